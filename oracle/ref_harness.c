/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C harness around the UNMODIFIED reference decoder/encoder compiled from
 * /root/reference (see oracle/Makefile, target _ref/libheaac_ref.so).  It drives the
 * reference's own AVCodec vtable (libavcodec/aacdec.c:2128-2142 `aac_decoder`,
 * libavcodec/aacenc.c `aac_encoder`) exactly as avcodec_decode_audio3 /
 * avcodec_encode_audio would (libavcodec/utils.c:536-548, 638-663), and taps the one
 * call site of the hot path, `spectral_to_sample(ac)` (aacdec.c:2078):
 *
 *   pre-hook : pack everything the synthesis back end consumes (coeffs, ICS/TNS side info,
 *              SBR and PS parameters) from AACContext into the C-ABI structs of
 *              include/heaac_b200.h.  This is the reference-side binding shown in
 *              a maintainer swapping the back end would add exactly this.
 *   post-hook: expose the reference's stage arrays (W, X_low, X_high, alpha, e_curr, gain,
 *              Y, X, ...) for stage-by-stage diffs.
 *
 * The two hook calls are injected at build time by a sed filter on the compiler's stdin
 * (oracle/Makefile); no reference source is copied into the repository.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>

#include "libavcodec/avcodec.h"
#include "libavcodec/aac.h"
#include "libavcodec/sbr.h"
#include "libavcodec/aacps.h"
#include "libavcodec/dsputil.h"

#include "heaac_b200.h"

extern AVCodec aac_decoder;
extern AVCodec aac_encoder;

/* ---- symbols the handful of compiled reference objects refer to but never call
 *      on the AAC path (video scaling, the generic parser glue, utils.c logging) ---- */
void av_log_missing_feature(void *avc, const char *feature, int want_sample)
{
    av_log(avc, AV_LOG_WARNING, "%s not implemented.\n", feature);
}
void ff_shrink22(uint8_t *d, int dw, const uint8_t *s, int sw, int w, int h) { abort(); }
void ff_shrink44(uint8_t *d, int dw, const uint8_t *s, int sw, int w, int h) { abort(); }
void ff_shrink88(uint8_t *d, int dw, const uint8_t *s, int sw, int w, int h) { abort(); }
void ff_img_copy_plane(uint8_t *d, int dw, const uint8_t *s, int sw, int bw, int h) { abort(); }
void ff_parse_close(AVCodecParserContext *s) { }
int  ff_aac_ac3_parse(AVCodecParserContext *s1, AVCodecContext *avctx,
                      const uint8_t **poutbuf, int *poutbuf_size,
                      const uint8_t *buf, int buf_size) { abort(); }

/* ---- decoder handle --------------------------------------------------------------- */
#define REF_MAX_ELEMS 8

typedef struct RefDec {
    AVCodecContext avctx;
    int      float_mode;
    int      n_elems;
    int      elem_type[REF_MAX_ELEMS];
    int      elem_id[REF_MAX_ELEMS];
    /* capture targets for the current frame (caller-owned, may be NULL) */
    float        *cap_coeffs;
    HeaacIcs     *cap_ics;
    HeaacTns     *cap_tns;
    HeaacSbrElem *cap_sbr;
    HeaacPsFrame *cap_ps;
    HeaacCce     *cap_cce;       /* one per TYPE_CCE element of the layout, in layout order */
    float        *cap_pcm_f32;   /* planar [ch][samples] */
    int           cap_pcm_f32_cap;
    int           last_samples;
    int           last_channels;
    int           hook_calls;
    /* drop-in mode: the product back end in place of spectral_to_sample (ref_dec_enable_dropin) */
    int           dropin;
    HeaacContext *be;
    int           be_sbr, be_ps, be_ds, be_ncc;
    long long     dropin_frames;
} RefDec;

static __thread RefDec *cur_dec;

/* RNE + saturate, what the SIMD float_to_int16 variants do (cvtps2dq + packssdw);
 * installing it makes aac_decode_init's scaling branch (aacdec.c:573-581) moot, so we
 * set the SIMD-branch constants by hand in ref_dec_open. */
static void tap_float_to_int16_interleave(int16_t *dst, const float **src, long len, int channels)
{
    RefDec *d = cur_dec;
    for (int c = 0; c < channels; c++) {
        for (long i = 0; i < len; i++) {
            long v = lrintf(src[c][i]);
            if (v < -32768) v = -32768;
            if (v >  32767) v =  32767;
            dst[i * channels + c] = (int16_t)v;
        }
        if (d && d->cap_pcm_f32 && (c + 1) * len <= d->cap_pcm_f32_cap)
            memcpy(d->cap_pcm_f32 + c * len, src[c], len * sizeof(float));
    }
}

static void pack_ics(HeaacIcs *o, const SingleChannelElement *sce)
{
    memset(o, 0, sizeof(*o));
    o->window_sequence = sce->ics.window_sequence[0];
    o->window_shape    = sce->ics.use_kb_window[0];
    o->max_sfb         = sce->ics.max_sfb;
    o->tns_present     = sce->tns.present;
    o->prev_valid = 1;
    o->prev_window_sequence = sce->ics.window_sequence[1];
    o->prev_window_shape    = sce->ics.use_kb_window[1];
}

static void pack_tns(HeaacTns *o, const SingleChannelElement *sce)
{
    memset(o, 0, sizeof(*o));
    if (!sce->tns.present)
        return;
    int is8 = sce->ics.window_sequence[0] == EIGHT_SHORT_SEQUENCE;
    for (int w = 0; w < sce->ics.num_windows; w++) {
        o->n_filt[w] = sce->tns.n_filt[w];
        for (int f = 0; f < sce->tns.n_filt[w]; f++) {
            int slot = is8 ? w : f;
            o->length[slot]    = sce->tns.length[w][f];
            o->order[slot]     = sce->tns.order[w][f];
            o->direction[slot] = sce->tns.direction[w][f];
            memcpy(o->coef[slot], sce->tns.coef[w][f], sizeof(o->coef[slot]));
        }
    }
}

static void pack_sbr(HeaacSbrElem *o, const SpectralBandReplication *s)
{
    memset(o, 0, sizeof(*o));
    o->start = s->start; o->reset = s->reset; o->bs_coupling = s->bs_coupling;
    o->bs_limiter_gains = s->bs_limiter_gains; o->bs_interpol_freq = s->bs_interpol_freq;
    o->bs_smoothing_mode = s->bs_smoothing_mode; o->num_patches = s->num_patches;
    o->n_q = s->n_q; o->n[0] = s->n[0]; o->n[1] = s->n[1]; o->n_lim = s->n_lim;
    o->k0 = s->k[0]; o->kx[0] = s->kx[0]; o->kx[1] = s->kx[1]; o->m[0] = s->m[0]; o->m[1] = s->m[1];
    for (int i = 0; i < 25; i++) o->f_tablelow[i]   = s->f_tablelow[i];
    for (int i = 0; i < 49; i++) o->f_tablehigh[i]  = s->f_tablehigh[i];
    for (int i = 0; i < 6;  i++) o->f_tablenoise[i] = s->f_tablenoise[i];
    for (int i = 0; i < 29; i++) o->f_tablelim[i]   = s->f_tablelim[i];
    for (int i = 0; i < 6;  i++) {
        o->patch_num_subbands[i]  = s->patch_num_subbands[i];
        o->patch_start_subband[i] = s->patch_start_subband[i];
    }
    for (int c = 0; c < 2; c++) {
        const SBRData *d = &s->data[c];
        HeaacSbrChan *oc = &o->ch[c];
        oc->bs_num_env = d->bs_num_env; oc->bs_num_noise = d->bs_num_noise;
        oc->bs_amp_res = d->bs_amp_res; oc->bs_add_harmonic_flag = d->bs_add_harmonic_flag;
        oc->e_a[0] = d->e_a[0]; oc->e_a[1] = d->e_a[1];
        oc->t_env_num_env_old = d->t_env_num_env_old;
        memcpy(oc->bs_freq_res, d->bs_freq_res, 7);
        memcpy(oc->t_env, d->t_env, 8);
        memcpy(oc->t_q, d->t_q, 3);
        memcpy(oc->bs_invf_mode, d->bs_invf_mode, 10);
        memcpy(oc->bs_add_harmonic, d->bs_add_harmonic, 48);
        /* env_facs / noise_facs still hold the integer-valued bitstream quantities here:
         * sbr_dequant runs inside ff_sbr_apply (aacsbr.c:1723-1725), i.e. after this hook. */
        for (int e = 0; e < 5; e++)
            for (int k = 0; k < 48; k++)
                oc->env_q[e][k] = (int16_t)lrintf(d->env_facs[e + 1][k]);
        for (int e = 0; e < 2; e++)
            for (int k = 0; k < 5; k++)
                oc->noise_q[e][k] = (int16_t)lrintf(d->noise_facs[e + 1][k]);
    }
}

static void pack_ps(HeaacPsFrame *o, const PSContext *p)
{
    memset(o, 0, sizeof(*o));
    o->start = p->start; o->num_env = p->num_env; o->num_env_old = p->num_env_old;
    o->iid_quant = p->iid_quant; o->nr_iid_par = p->nr_iid_par; o->nr_icc_par = p->nr_icc_par;
    o->nr_ipdopd_par = p->nr_ipdopd_par; o->icc_mode = p->icc_mode;
    o->enable_ipdopd = p->enable_ipdopd; o->is34bands = p->is34bands;
    o->is34bands_old = p->is34bands_old;
    for (int i = 0; i < 6; i++) o->border_position[i] = p->border_position[i];
    memcpy(o->iid_par, p->iid_par, sizeof(o->iid_par));
    memcpy(o->icc_par, p->icc_par, sizeof(o->icc_par));
    memcpy(o->ipd_par, p->ipd_par, sizeof(o->ipd_par));
    memcpy(o->opd_par, p->opd_par, sizeof(o->opd_par));
}

/* ChannelCoupling (aac.h:165-177) of every TYPE_CCE element of the layout -> HeaacCce: the (type, id_select, ch_select) entries
 * of apply_channel_coupling (aacdec.c:1870-1898) resolved into (core channel, gain) pairs -- the front-end half of the coupling
 * boundary (include/heaac_b200.h).  Core channels count in layout order. */
static void pack_cce_all(RefDec *d, AACContext *ac, HeaacCce *out)
{
    int first_ch[REF_MAX_ELEMS], ch = 0, k = 0;
    for (int e = 0; e < d->n_elems; e++) { first_ch[e] = ch; ch += d->elem_type[e] == TYPE_CPE ? 2 : 1; }
    for (int e = 0; e < d->n_elems; e++) {
        if (d->elem_type[e] != TYPE_CCE) continue;
        HeaacCce *o = out + k++;
        ChannelElement *cce = ac->che[TYPE_CCE][d->elem_id[e]];
        memset(o, 0, sizeof(*o));
        if (!cce) continue;
        const ChannelCoupling *coup = &cce->coup;
        o->coupling_point = coup->coupling_point;
        int index = 0, n = 0;
        for (int c = 0; c <= coup->num_coupled; c++) {
            int sel = coup->ch_select[c], te = -1;
            for (int t = 0; t < d->n_elems; t++)
                if (d->elem_type[t] == (int)coup->type[c] && d->elem_id[t] == coup->id_select[c]) te = t;
            if (te >= 0) {
                if (sel != 1 && n < HEAAC_CCE_MAX_TARGETS) { o->target_ch[n] = first_ch[te];     o->gain[n++] = coup->gain[index][0]; }
                if (sel != 2 && n < HEAAC_CCE_MAX_TARGETS) { o->target_ch[n] = first_ch[te] + 1; o->gain[n++] = coup->gain[index + (sel == 3)][0]; }
            }
            index += 1 + (sel == 3);
        }
        o->n_targets = n;
    }
}

/* ---- drop-in mode: the product library in place of spectral_to_sample, inside this process ------------------- */
static struct {
    void *lib;
    int  (*create)(const HeaacConfig *, HeaacContext **);
    void (*destroy)(HeaacContext *);
    int  (*synth)(HeaacContext *, const HeaacBatch *);
} g_be;

/* Hand the current frame to heaac_b200_synth and put its float PCM where spectral_to_sample would have left it
 * (che->ch[c].ret): the reference's own output stage (float_to_int16_interleave, aacdec.c:2096) runs on it afterwards.
 * Returns 1 on success. */
static int dropin_frame(RefDec *d, AACContext *ac)
{
    static __thread float coeffs[HEAAC_MAX_CHANNELS * 1024], pcm[HEAAC_MAX_CHANNELS * 2048];
    HeaacIcs ics[HEAAC_MAX_CHANNELS]; HeaacTns tns[HEAAC_MAX_CHANNELS];
    HeaacSbrElem sbr[REF_MAX_ELEMS]; HeaacPsFrame ps;
    const int want_sbr = ac->m4ac.sbr == 1, want_ps = ac->m4ac.ps == 1 && d->avctx.channels <= 2 && d->n_elems == 1 && d->elem_type[0] == TYPE_SCE;
    const int want_ds = want_sbr && ac->m4ac.ext_sample_rate < 2 * ac->m4ac.sample_rate;
    int ncc = 0;
    for (int e = 0; e < d->n_elems; e++) ncc += d->elem_type[e] == TYPE_CPE ? 2 : 1;
    if (!d->be || d->be_sbr != want_sbr || d->be_ps != want_ps || d->be_ds != want_ds) {
        /* (re)create the back end for what the front end has found out by now (implicit SBR / PS signalling is known
         * once the first frame's fill elements are parsed, i.e. before the first spectral_to_sample) */
        HeaacConfig cfg;
        if (d->be) g_be.destroy(d->be);
        d->be = NULL;
        memset(&cfg, 0, sizeof(cfg));
        cfg.abi_version = HEAAC_ABI_VERSION;
        cfg.n_streams = 1; cfg.n_elems = d->n_elems;
        for (int e = 0; e < d->n_elems; e++) cfg.elem_type[e] = (uint8_t)d->elem_type[e];
        cfg.sr_index = ac->m4ac.sampling_index;
        cfg.sbr = want_sbr; cfg.ps = want_ps; cfg.sbr_downsampled = want_ds;
        cfg.add_bias = ac->add_bias; cfg.sf_scale = ac->sf_scale;
        cfg.max_frames_per_call = 1; cfg.want_float_out = 1;
        if (g_be.create(&cfg, &d->be)) return 0;
        d->be_sbr = want_sbr; d->be_ps = want_ps; d->be_ds = want_ds; d->be_ncc = ncc;
    }
    memset(&ps, 0, sizeof(ps));
    int ch = 0, se = 0;
    for (int e = 0; e < d->n_elems; e++) {
        ChannelElement *che = ac->che[d->elem_type[e]][d->elem_id[e]];
        int nch = d->elem_type[e] == TYPE_CPE ? 2 : 1;
        if (!che) return 0;
        for (int c = 0; c < nch; c++, ch++) {
            memcpy(coeffs + 1024 * ch, che->ch[c].coeffs, 1024 * sizeof(float));
            pack_ics(ics + ch, &che->ch[c]);
            pack_tns(tns + ch, &che->ch[c]);
        }
        if (d->elem_type[e] != TYPE_LFE) {
            pack_sbr(sbr + se, &che->sbr);
            if (se == 0) pack_ps(&ps, &che->sbr.ps);
            se++;
        }
    }
    HeaacCce cce[REF_MAX_ELEMS];
    int n_cce = 0;
    for (int e = 0; e < d->n_elems; e++) n_cce += d->elem_type[e] == TYPE_CCE;
    if (n_cce) pack_cce_all(d, ac, cce);
    HeaacBatch b;
    memset(&b, 0, sizeof(b));
    b.n_frames = 1; b.coeffs = coeffs; b.ics = ics; b.tns = tns;
    b.sbr = want_sbr ? sbr : NULL; b.ps = want_ps ? &ps : NULL; b.pcm_f32 = pcm;
    b.cce = n_cce ? cce : NULL;
    if (g_be.synth(d->be, &b)) return 0;
    /* planar float PCM [out channel][len] -> the elements' ret buffers (with PS the SCE's second channel is R) */
    const int len = (want_sbr && !want_ds) ? 2048 : 1024;
    ch = 0;
    for (int e = 0; e < d->n_elems; e++) {
        ChannelElement *che = ac->che[d->elem_type[e]][d->elem_id[e]];
        int nout = want_ps ? 2 : (d->elem_type[e] == TYPE_CPE ? 2 : 1);
        if (d->elem_type[e] == TYPE_CCE) continue;          /* no output channel of its own (its signal is inside the targets now) */
        for (int c = 0; c < nout; c++, ch++)
            memcpy(che->ch[c].ret, pcm + (size_t)ch * len, len * sizeof(float));
    }
    d->dropin_frames++;
    return 1;
}

/* dlopen the product library and switch this decoder to drop-in mode.  Returns 0 on success. */
int ref_dec_enable_dropin(RefDec *d, const char *libpath)
{
    if (!g_be.lib) {
        void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
        if (!h) { av_log(NULL, AV_LOG_ERROR, "ref_harness: dlopen(%s): %s\n", libpath, dlerror()); return -1; }
        g_be.create  = (int (*)(const HeaacConfig *, HeaacContext **))dlsym(h, "heaac_b200_create");
        g_be.destroy = (void (*)(HeaacContext *))dlsym(h, "heaac_b200_destroy");
        g_be.synth   = (int (*)(HeaacContext *, const HeaacBatch *))dlsym(h, "heaac_b200_synth");
        if (!g_be.create || !g_be.destroy || !g_be.synth) { dlclose(h); return -2; }
        g_be.lib = h;
    }
    d->dropin = 1;
    return 0;
}
long long ref_dec_dropin_frames(RefDec *d) { return d->dropin_frames; }

/* called from aac_decode_frame in place of / right before spectral_to_sample(ac):
 *   if (!heaac_ref_pre_hook(ac)) spectral_to_sample(ac); */
int heaac_ref_pre_hook(AACContext *ac)
{
    RefDec *d = cur_dec;
    if (!d) return 0;
    d->hook_calls++;
    int ch = 0, se = 0;
    for (int e = 0; e < d->n_elems; e++) {
        ChannelElement *che = ac->che[d->elem_type[e]][d->elem_id[e]];
        int nch = d->elem_type[e] == TYPE_CPE ? 2 : 1;
        for (int c = 0; c < nch; c++, ch++) {
            if (!che) continue;
            if (d->cap_coeffs) memcpy(d->cap_coeffs + 1024 * ch, che->ch[c].coeffs, 1024 * sizeof(float));
            if (d->cap_ics) pack_ics(d->cap_ics + ch, &che->ch[c]);
            if (d->cap_tns) pack_tns(d->cap_tns + ch, &che->ch[c]);
        }
        if (d->elem_type[e] != TYPE_LFE) {
            if (che && d->cap_sbr) pack_sbr(d->cap_sbr + se, &che->sbr);
            if (che && d->cap_ps && se == 0) pack_ps(d->cap_ps, &che->sbr.ps);
            se++;
        }
    }
    if (d->cap_cce) pack_cce_all(d, ac, d->cap_cce);
    return d->dropin ? dropin_frame(d, ac) : 0;
}

void heaac_ref_post_hook(AACContext *ac) { }

/* ---- public harness API (ctypes) ------------------------------------------------ */
RefDec *ref_dec_open(const uint8_t *extradata, int extradata_size, int sample_rate, int channels,
                     int float_mode, int n_elems, const int *elem_type, const int *elem_id)
{
    RefDec *d = av_mallocz(sizeof(*d));
    if (!d) return NULL;
    d->avctx.codec = &aac_decoder;
    d->avctx.codec_type = AVMEDIA_TYPE_AUDIO;
    d->avctx.codec_id = CODEC_ID_AAC;
    d->avctx.sample_rate = sample_rate;
    d->avctx.channels = channels;
    d->avctx.priv_data = av_mallocz(aac_decoder.priv_data_size);
    if (extradata_size > 0) {
        d->avctx.extradata = av_mallocz(extradata_size + FF_INPUT_BUFFER_PADDING_SIZE);
        memcpy(d->avctx.extradata, extradata, extradata_size);
        d->avctx.extradata_size = extradata_size;
    }
    d->float_mode = float_mode;
    d->n_elems = n_elems;
    for (int i = 0; i < n_elems && i < REF_MAX_ELEMS; i++) {
        d->elem_type[i] = elem_type[i];
        d->elem_id[i]   = elem_id[i];
    }
    if (aac_decoder.init(&d->avctx) < 0) {
        av_free(d->avctx.priv_data); av_free(d);
        return NULL;
    }
    if (float_mode) {
        /* select the reference's own SIMD scaling branch, aacdec.c:577-581 */
        AACContext *ac = d->avctx.priv_data;
        ac->add_bias  = 0.0f;
        ac->sf_scale  = 1. / -1024.;
        ac->sf_offset = 60;
        ac->dsp.float_to_int16_interleave = tap_float_to_int16_interleave;
    }
    return d;
}

void ref_dec_close(RefDec *d)
{
    if (!d) return;
    if (d->be) g_be.destroy(d->be);
    aac_decoder.close(&d->avctx);
    av_free(d->avctx.priv_data);
    av_free(d->avctx.extradata);
    av_free(d);
}

void ref_dec_set_capture_cce(RefDec *d, HeaacCce *cce) { d->cap_cce = cce; }

void ref_dec_set_capture(RefDec *d, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                         HeaacSbrElem *sbr, HeaacPsFrame *ps, float *pcm_f32, int pcm_f32_cap)
{
    d->cap_coeffs = coeffs; d->cap_ics = ics; d->cap_tns = tns; d->cap_sbr = sbr; d->cap_ps = ps;
    d->cap_pcm_f32 = pcm_f32; d->cap_pcm_f32_cap = pcm_f32_cap;
}

/* One avcodec_decode_audio3-equivalent call.  `pkt` must carry
 * FF_INPUT_BUFFER_PADDING_SIZE zero bytes after `size`.  Returns bytes consumed or <0;
 * *out_bytes = bytes of s16 PCM written. */
int ref_dec_decode(RefDec *d, const uint8_t *pkt, int size, int16_t *out, int *out_bytes)
{
    AVPacket avpkt;
    memset(&avpkt, 0, sizeof(avpkt));
    avpkt.data = (uint8_t *)pkt;
    avpkt.size = size;
    cur_dec = d;
    int ret = aac_decoder.decode(&d->avctx, out, out_bytes, &avpkt);
    cur_dec = NULL;
    if (ret >= 0) {
        d->last_channels = d->avctx.channels;
        d->last_samples  = d->last_channels ? *out_bytes / 2 / d->last_channels : 0;
    }
    return ret;
}

/* Timing helper for the CPU baseline: decode `n` packets (concatenated in `blob`, each followed by
 * FF_INPUT_BUFFER_PADDING_SIZE zero bytes; offsets/sizes given) `repeat` times through the same context,
 * entirely in C.  Returns the number of output samples per channel produced, or <0 on error. */
long long ref_dec_decode_many(RefDec *d, const uint8_t *blob, const int *offsets, const int *sizes, int n, int repeat)
{
    static __thread int16_t out[8 * 2048 + 64] __attribute__((aligned(16)));
    long long samples = 0;
    for (int r = 0; r < repeat; r++)
        for (int i = 0; i < n; i++) {
            AVPacket avpkt;
            int out_bytes = sizeof(out);
            memset(&avpkt, 0, sizeof(avpkt));
            avpkt.data = (uint8_t *)blob + offsets[i];
            avpkt.size = sizes[i];
            cur_dec = d;
            int ret = aac_decoder.decode(&d->avctx, out, &out_bytes, &avpkt);
            cur_dec = NULL;
            if (ret < 0) return -1;
            samples += d->avctx.channels ? out_bytes / 2 / d->avctx.channels : 0;
        }
    return samples;
}

int ref_dec_channels(RefDec *d)    { return d->avctx.channels; }
int ref_dec_sample_rate(RefDec *d) { return d->avctx.sample_rate; }
int ref_dec_sbr(RefDec *d)         { return ((AACContext *)d->avctx.priv_data)->m4ac.sbr; }
int ref_dec_ps(RefDec *d)          { return ((AACContext *)d->avctx.priv_data)->m4ac.ps; }
int ref_dec_sr_index(RefDec *d)    { return ((AACContext *)d->avctx.priv_data)->m4ac.sampling_index; }
float ref_dec_add_bias(RefDec *d)  { return ((AACContext *)d->avctx.priv_data)->add_bias; }
float ref_dec_sf_scale(RefDec *d)  { return ((AACContext *)d->avctx.priv_data)->sf_scale; }

/* Stage arrays of element `e` after the last decode (post spectral_to_sample).
 * Returns the number of bytes copied, or <0 for an unknown name. */
int ref_dec_peek(RefDec *d, int e, int ch, const char *name, void *dst, int cap)
{
    AACContext *ac = d->avctx.priv_data;
    ChannelElement *che = ac->che[d->elem_type[e]][d->elem_id[e]];
    if (!che) return -2;
    SpectralBandReplication *s = &che->sbr;
    const void *p = NULL; int n = 0;
#define PK(nm, ptr, sz) if (!strcmp(name, nm)) { p = (ptr); n = (sz); }
    PK("coeffs",  che->ch[ch].coeffs, sizeof(che->ch[ch].coeffs))
    PK("saved",   che->ch[ch].saved, 512 * sizeof(float))
    PK("ret",     che->ch[ch].ret, sizeof(che->ch[ch].ret))
    PK("W",       s->data[ch].W, sizeof(s->data[ch].W))
    PK("Y",       s->data[ch].Y, sizeof(s->data[ch].Y))
    PK("g_temp",  s->data[ch].g_temp, sizeof(s->data[ch].g_temp))
    PK("q_temp",  s->data[ch].q_temp, sizeof(s->data[ch].q_temp))
    PK("env_facs", s->data[ch].env_facs, sizeof(s->data[ch].env_facs))
    PK("noise_facs", s->data[ch].noise_facs, sizeof(s->data[ch].noise_facs))
    PK("bw_array", s->data[ch].bw_array, sizeof(s->data[ch].bw_array))
    PK("s_indexmapped", s->data[ch].s_indexmapped, sizeof(s->data[ch].s_indexmapped))
    PK("ana_samples", s->data[ch].analysis_filterbank_samples, sizeof(s->data[ch].analysis_filterbank_samples))
    PK("X_low",   s->X_low, sizeof(s->X_low))
    PK("X_high",  s->X_high, sizeof(s->X_high))
    PK("X",       s->X[ch], sizeof(s->X[ch]))
    PK("alpha0",  s->alpha0, sizeof(s->alpha0))
    PK("alpha1",  s->alpha1, sizeof(s->alpha1))
    PK("e_origmapped", s->e_origmapped, sizeof(s->e_origmapped))
    PK("q_mapped", s->q_mapped, sizeof(s->q_mapped))
    PK("s_mapped", s->s_mapped, sizeof(s->s_mapped))
    PK("e_curr",  s->e_curr, sizeof(s->e_curr))
    PK("q_m",     s->q_m, sizeof(s->q_m))
    PK("s_m",     s->s_m, sizeof(s->s_m))
    PK("gain",    s->gain, sizeof(s->gain))
    PK("ps_delay", s->ps.delay, sizeof(s->ps.delay))
    PK("ps_ap_delay", s->ps.ap_delay, sizeof(s->ps.ap_delay))
    PK("ps_H11",  s->ps.H11, sizeof(s->ps.H11))
    PK("ps_H12",  s->ps.H12, sizeof(s->ps.H12))
    PK("ps_H21",  s->ps.H21, sizeof(s->ps.H21))
    PK("ps_H22",  s->ps.H22, sizeof(s->ps.H22))
#undef PK
    if (!p) return -1;
    if (n > cap) n = cap;
    memcpy(dst, p, n);
    return n;
}

/* ---- the reference's transform entry points, for direct known-answer checks ----- */
typedef struct RefMdct { FFTContext c; } RefMdct;
RefMdct *ref_mdct_init(int nbits, int inverse, double scale)
{
    RefMdct *m = av_mallocz(sizeof(*m));
    if (ff_mdct_init(&m->c, nbits, inverse, scale) < 0) { av_free(m); return NULL; }
    return m;
}
void ref_imdct_half(RefMdct *m, float *out, const float *in) { ff_imdct_half(&m->c, out, in); }
void ref_imdct_calc(RefMdct *m, float *out, const float *in) { ff_imdct_calc(&m->c, out, in); }
void ref_mdct_end(RefMdct *m) { ff_mdct_end(&m->c); av_free(m); }

void ref_kbd_window(float *w, float alpha, int n) { ff_kbd_window_init(w, alpha, n); }
void ref_vector_fmul_window(float *dst, const float *s0, const float *s1, const float *win, float bias, int len)
{
    ff_vector_fmul_window_c(dst, s0, s1, win, bias, len);
}
void ref_float_to_int16(int16_t *dst, const float *src, long len) { ff_float_to_int16_c(dst, src, len); }

/* ---- encoder (fixture generation only) ------------------------------------------- */
typedef struct RefEnc { AVCodecContext avctx; } RefEnc;

RefEnc *ref_enc_open(int sample_rate, int channels, int bit_rate)
{
    RefEnc *e = av_mallocz(sizeof(*e));
    e->avctx.codec = &aac_encoder;
    e->avctx.codec_type = AVMEDIA_TYPE_AUDIO;
    e->avctx.codec_id = CODEC_ID_AAC;
    e->avctx.sample_rate = sample_rate;
    e->avctx.channels = channels;
    e->avctx.bit_rate = bit_rate;
    e->avctx.profile = FF_PROFILE_UNKNOWN;
    e->avctx.flags |= CODEC_FLAG_BITEXACT;
    e->avctx.strict_std_compliance = -2;
    e->avctx.priv_data = av_mallocz(aac_encoder.priv_data_size);
    if (aac_encoder.init(&e->avctx) < 0) { av_free(e->avctx.priv_data); av_free(e); return NULL; }
    return e;
}
int ref_enc_extradata(RefEnc *e, uint8_t *dst, int cap)
{
    int n = e->avctx.extradata_size < cap ? e->avctx.extradata_size : cap;
    memcpy(dst, e->avctx.extradata, n);
    return n;
}
/* one 1024-sample frame of interleaved s16 in; returns packet bytes (0 on the first call) */
int ref_enc_encode(RefEnc *e, const int16_t *pcm, uint8_t *out, int out_cap)
{
    int r = aac_encoder.encode(&e->avctx, out, out_cap, (void *)pcm);
    e->avctx.frame_number++;
    return r;
}
int ref_enc_frame_bits(RefEnc *e) { return e->avctx.frame_bits; }
void ref_enc_close(RefEnc *e)
{
    if (!e) return;
    aac_encoder.close(&e->avctx);
    av_free(e->avctx.priv_data);
    av_free(e);
}

void ref_set_log_level(int level) { av_log_set_level(level); }
