/*
 * heaac_frontend.c -- host build of the front end (heaac_fe_core.h): table construction, AudioSpecificConfig
 * parsing and the heaac_fe_* entry points.  One HeaacFrontend per stream; different front ends may be driven
 * from different threads.
 *   AudioSpecificConfig          mpeg4audio.c:77-130, aacdec.c:462-500
 */
#include <pthread.h>
#include "heaac_fe_core.h"
#include "heaac_huff_data.h"
#include "heaac_tables_data.h"

/* ---------------------------------------------------------------- table construction (host only) */
static void vlc_build(FeTables *T, Vlc *v, const uint32_t *codes, const uint8_t *bits, int n)
{
    int sub_bits[1 << VLC_PRIM], sub_off[1 << VLC_PRIM], total = 1 << VLC_PRIM;
    memset(sub_bits, 0, sizeof(sub_bits));
    v->maxlen = 0;
    for (int i = 0; i < n; i++) {
        const int l = bits[i];
        if (l > v->maxlen) v->maxlen = l;
        if (l > VLC_PRIM) {
            const unsigned prefix = codes[i] >> (l - VLC_PRIM);
            if (l - VLC_PRIM > sub_bits[prefix]) sub_bits[prefix] = l - VLC_PRIM;
        }
    }
    for (int p = 0; p < (1 << VLC_PRIM); p++) { sub_off[p] = total; total += sub_bits[p] ? 1 << sub_bits[p] : 0; }
    if (T->pool_used + (uint32_t)total > FE_POOL_WORDS) abort();                 /* sized for the normative tables */
    v->off = T->pool_used;
    T->pool_used += (uint32_t)total;
    uint32_t *tab = T->pool + v->off;
    for (int p = 0; p < (1 << VLC_PRIM); p++)
        if (sub_bits[p]) tab[p] = ((uint32_t)sub_off[p] << 8) | 0x80u | (uint32_t)sub_bits[p];
    for (int i = 0; i < n; i++) {
        const int l = bits[i];
        if (!l) continue;
        if (l <= VLC_PRIM) {
            const unsigned base = codes[i] << (VLC_PRIM - l);
            for (unsigned k = 0; k < (1u << (VLC_PRIM - l)); k++) tab[base + k] = ((uint32_t)i << 8) | (uint32_t)l;
        } else {
            const int r = l - VLC_PRIM;
            const unsigned prefix = codes[i] >> r, rem = codes[i] & ((1u << r) - 1);
            const int sb = sub_bits[prefix];
            const unsigned base = sub_off[prefix] + (rem << (sb - r));
            for (unsigned k = 0; k < (1u << (sb - r)); k++) tab[base + k] = ((uint32_t)i << 8) | (uint32_t)l;
        }
    }
}

#define VB(v, name) vlc_build(T, &(v), name##_codes, name##_bits, (int)(sizeof(name##_bits)))
static void fe_tables_once(void)
{
    FeTables *T = (FeTables *)calloc(1, sizeof(FeTables));
    if (!T) abort();
    VB(T->vsf, heaac_sf);
    VB(T->vspec[0], heaac_spec1); VB(T->vspec[1], heaac_spec2); VB(T->vspec[2], heaac_spec3); VB(T->vspec[3], heaac_spec4);
    VB(T->vspec[4], heaac_spec5); VB(T->vspec[5], heaac_spec6); VB(T->vspec[6], heaac_spec7); VB(T->vspec[7], heaac_spec8);
    VB(T->vspec[8], heaac_spec9); VB(T->vspec[9], heaac_spec10); VB(T->vspec[10], heaac_spec11);
    VB(T->vsbr[T_ENV_15], heaac_sbr_t_huffman_env_1_5dB); VB(T->vsbr[F_ENV_15], heaac_sbr_f_huffman_env_1_5dB);
    VB(T->vsbr[T_ENV_BAL_15], heaac_sbr_t_huffman_env_bal_1_5dB); VB(T->vsbr[F_ENV_BAL_15], heaac_sbr_f_huffman_env_bal_1_5dB);
    VB(T->vsbr[T_ENV_30], heaac_sbr_t_huffman_env_3_0dB); VB(T->vsbr[F_ENV_30], heaac_sbr_f_huffman_env_3_0dB);
    VB(T->vsbr[T_ENV_BAL_30], heaac_sbr_t_huffman_env_bal_3_0dB); VB(T->vsbr[F_ENV_BAL_30], heaac_sbr_f_huffman_env_bal_3_0dB);
    VB(T->vsbr[T_NOISE_30], heaac_sbr_t_huffman_noise_3_0dB); VB(T->vsbr[T_NOISE_BAL_30], heaac_sbr_t_huffman_noise_bal_3_0dB);
    VB(T->vps[PS_IID_DF1], heaac_ps_huff_iid_df1); VB(T->vps[PS_IID_DT1], heaac_ps_huff_iid_dt1);
    VB(T->vps[PS_IID_DF0], heaac_ps_huff_iid_df0); VB(T->vps[PS_IID_DT0], heaac_ps_huff_iid_dt0);
    VB(T->vps[PS_ICC_DF], heaac_ps_huff_icc_df); VB(T->vps[PS_ICC_DT], heaac_ps_huff_icc_dt);
    VB(T->vps[PS_IPD_DF], heaac_ps_huff_ipd_df); VB(T->vps[PS_IPD_DT], heaac_ps_huff_ipd_dt);
    VB(T->vps[PS_OPD_DF], heaac_ps_huff_opd_df); VB(T->vps[PS_OPD_DT], heaac_ps_huff_opd_dt);
    for (int i = 0; i < 428; i++) T->pow2sf[i] = pow(2, (i - 200) / 4.);
    for (int i = 0; i < 8192; i++) T->pow43[i] = cbrtf(i) * i;
    for (int i = 0; i < 16; i++) T->cb_val[i] = (float)pow(i, 4.0 / 3.0);      /* == the correctly rounded literals */
    for (int bk = 0; bk < 11; bk++) {                                         /* index = sum (q_j + off) range^(dim-1-j) */
        const int bt = bk + 1, dim = bt < 5 ? 4 : 2;
        const int range = bt < 5 ? 3 : bt < 7 ? 9 : bt < 9 ? 8 : bt < 11 ? 13 : 17;
        const int lav = (bt == 1 || bt == 2) ? 1 : (bt == 5 || bt == 6) ? 4 : 0;
        for (int code = 0; code < spec_sizes[bk]; code++) {
            int c = code;
            for (int d = dim - 1; d >= 0; d--) {
                const int q = c % range - lav;
                c /= range;
                T->spec_val[bk][code][d] = q < 0 ? -T->cb_val[-q] : T->cb_val[q > 15 ? 0 : q];     /* 16 = escape marker */
            }
        }
    }
    memcpy(T->tns_map, heaac_tns_map, sizeof(T->tns_map));
    memcpy(T->swb1024, heaac_swb_offset_1024, sizeof(T->swb1024));
    memcpy(T->swb128, heaac_swb_offset_128, sizeof(T->swb128));
    memcpy(T->num_swb1024, heaac_num_swb_1024, sizeof(T->num_swb1024));
    memcpy(T->num_swb128, heaac_num_swb_128, sizeof(T->num_swb128));
    fe_tb_host = T;
}
static pthread_once_t fe_once = PTHREAD_ONCE_INIT;
static void fe_tables(void) { pthread_once(&fe_once, fe_tables_once); }

/* the built tables, for the device front end to upload (heaac_parse_gpu.cu) */
const void *heaac_fe_tables_blob(size_t *bytes)
{
    fe_tables();
    if (bytes) *bytes = sizeof(FeTables);
    return fe_tb_host;
}
size_t heaac_fe_state_size(void) { return sizeof(FE); }

/* ---------------------------------------------------------------- configuration (host only) */
static int get_aot(BR *b) { int t = br_get(b, 5); if (t == 31) t = 32 + br_get(b, 6); return t; }
static int get_rate(BR *b, int *idx) { *idx = br_get(b, 4); return *idx == 15 ? (int)br_get(b, 24) : sample_rates[*idx]; }

/* AudioSpecificConfig, mpeg4audio.c:77-130 + GASpecificConfig aacdec.c:425-455 */
static int parse_asc(FE *fe, const uint8_t *d, int size)
{
    BR b;
    br_init(&b, d, size);
    int idx, ext_idx;
    fe->object_type = get_aot(&b);
    fe->sample_rate = get_rate(&b, &idx);
    fe->sr_index = idx;
    int chan_config = br_get(&b, 4);
    fe->sbr = -1; fe->ps = -1; fe->ext_sample_rate = 0;
    if (fe->object_type == 5 || fe->object_type == 29) {          /* AOT_SBR / AOT_PS: explicit hierarchical signalling */
        if (fe->object_type == 29) fe->ps = 1;
        fe->sbr = 1;
        fe->ext_sample_rate = get_rate(&b, &ext_idx);
        fe->object_type = get_aot(&b);
    }
    if (fe->object_type != 2) return HEAAC_E_UNSUPPORTED;          /* AAC-LC only */
    if (idx > 12) return HEAAC_E_INVAL;
    const int specific = b.pos;                                    /* GASpecificConfig, aacdec.c:425-455 */
    if (br_get1(&b)) return HEAAC_E_UNSUPPORTED;                   /* frameLengthFlag: 960 */
    if (br_get1(&b)) br_skip(&b, 14);                              /* dependsOnCoreCoder */
    br_get1(&b);                                                   /* extensionFlag */
    if (!chan_config) {                                            /* the layout comes as a program config element */
        br_skip(&b, 4);                                            /* element_instance_tag */
        int r = fe_pce(fe, &b, 1);
        if (r) return r;
    }
    b.pos = specific;                                              /* the sync extension is looked for from here (mpeg4audio.c:105-125) */
    if (fe->sbr == -1 && br_left(&b) >= 16) {                      /* backward-compatible explicit signalling */
        while (br_left(&b) > 15) {
            if (br_show(&b, 11) == 0x2b7) {
                br_skip(&b, 11);
                int ext_aot = get_aot(&b);
                if (ext_aot == 5 && (fe->sbr = br_get1(&b)) == 1) fe->ext_sample_rate = get_rate(&b, &ext_idx);
                if (br_left(&b) > 11 && br_get(&b, 11) == 0x548) fe->ps = br_get1(&b);
                break;
            } else br_skip(&b, 1);
        }
    }
    if (fe->sbr == 1 && fe->ps == -1) fe->ps = 1;                  /* aacdec.c:478-479 */
    return chan_config ? set_layout(fe, chan_config) : 0;
}


int heaac_fe_open(FE **out, const uint8_t *extradata, int extradata_size, int sample_rate, int channels, int bias_mode)
{
    if (!out) return HEAAC_E_INVAL;
    fe_tables();
    FE *fe = (FE *)calloc(1, sizeof(FE));
    if (!fe) return HEAAC_E_NOMEM;
    fe->sf_offset = bias_mode ? 0 : 60;
    fe->random_state = 0x1f2e3d4c;                                 /* aacdec.c:567 */
    int r;
    if (extradata && extradata_size > 0) {
        r = parse_asc(fe, extradata, extradata_size);
    } else {                                                       /* ADTS-only stream: configured by the first header */
        fe->sbr = fe->ps = -1; fe->object_type = 2;
        fe->sample_rate = sample_rate;
        fe->sr_index = 15;
        for (int i = 0; i < 13; i++) if (sample_rates[i] == sample_rate) fe->sr_index = i;
        r = channels > 0 ? set_layout(fe, channels == 6 ? 6 : channels) : 0;
    }
    if (r) { free(fe); return r; }
    for (int e = 0; e < HEAAC_MAX_ELEMS; e++) {                    /* ff_aac_sbr_ctx_init, aacsbr.c:128-137 */
        fe->sbr_el[e].t.kx[0] = fe->sbr_el[e].t.kx[1] = 32;
        fe->sbr_el[e].ch[0].e_a[1] = fe->sbr_el[e].ch[1].e_a[1] = -1;
    }
    *out = fe;
    return 0;
}
void heaac_fe_close(FE *fe) { free(fe); }

/* One packet (an ADTS frame or a bare raw_data_block) -> the frame's C-ABI structures.
 * coeffs [ncc][1024]; ics/tns [ncc]; sbr [nse] (may be NULL for LC); ps (may be NULL).
 * Returns bytes consumed (>= 0) or a negative HEAAC_E_* code (the reference returns -1). */
int heaac_fe_parse(FE *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                   HeaacSbrElem *sbr, HeaacPsFrame *ps)
{
    return fe_parse_frame(fe, pkt, size, coeffs, ics, tns, sbr, ps);
}

/* The two-phase parse the device front end uses (fe_phase_a + fe_phase_b), run on the host: same results as
 * heaac_fe_parse by construction -- tests hold it to that. */
int heaac_fe_parse_two_phase(FE *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                             HeaacSbrElem *sbr, HeaacPsFrame *ps)
{
    if (!fe || !pkt || size <= 0 || !coeffs || !ics || !tns) return HEAAC_E_INVAL;
    FE *scratch = (FE *)malloc(sizeof(FE));
    FeTodo todo;
    if (!scratch) return HEAAC_E_NOMEM;
    fe_phase_a(fe, scratch, &todo, pkt, size, coeffs, ics, tns);
    free(scratch);
    return fe_phase_b(fe, &todo, pkt, size, coeffs, ics, tns, sbr, ps);
}

/* give an unconfigured front end (ADTS input, nothing parsed yet) the layout another one has found */
int heaac_fe_copy_layout(FE *dst, const FE *src)
{
    if (!dst || !src) return HEAAC_E_INVAL;
    dst->sample_rate = src->sample_rate; dst->sr_index = src->sr_index;
    if (src->by_tag) {                                             /* a layout from a program config element */
        dst->n_elems = src->n_elems; dst->ncc = src->ncc; dst->nse = src->nse; dst->n_cce = src->n_cce;
        dst->chan_config = 0; dst->by_tag = 1;
        memcpy(dst->elem_type, src->elem_type, sizeof(dst->elem_type));
        memcpy(dst->elem_tag, src->elem_tag, sizeof(dst->elem_tag));
        memcpy(dst->elem_ch0, src->elem_ch0, sizeof(dst->elem_ch0));
        memset(dst->cce, 0, sizeof(dst->cce));
        return 0;
    }
    return src->chan_config ? set_layout(dst, src->chan_config) : 0;
}

/* Batch use: give `dst` the layout and core rate `src` has found and FIX them -- the caller sizes every per-stream
 * output slot for that layout, so from now on a packet whose ADTS header announces another channel configuration or
 * sampling-rate index fails with HEAAC_E_INVAL before anything is written (fe_parse_frame). */
int heaac_fe_fix_layout(FE *dst, const FE *src)
{
    int r = heaac_fe_copy_layout(dst, src);
    if (r) return r;
    if (!dst->chan_config && !dst->by_tag) return HEAAC_E_INVAL;
    dst->layout_fixed = 1;
    return 0;
}

/* 1 when the stream asks for downsampled SBR: explicitly signalled extension rate below twice the core rate
 * (ff_sbr_apply: downsampled = m4ac.ext_sample_rate < sbr->sample_rate, aacsbr.c:1719) */
int heaac_fe_sbr_downsampled(const FE *fe)
{
    return fe && fe->sbr == 1 && fe->ext_sample_rate && fe->ext_sample_rate < 2 * fe->sample_rate;
}

/* the state block of an opened front end, for the device front end to replicate per stream */
const void *heaac_fe_state_blob(const FE *fe) { return fe; }

/* Stream properties known after open (explicit config) or after the first parsed frame (implicit SBR/PS). */
int heaac_fe_info(const FE *fe, int *ncc, int *nse, int *sbr, int *ps, int *sr_index, uint8_t elem_type[HEAAC_MAX_ELEMS])
{
    if (!fe) return HEAAC_E_INVAL;
    if (ncc) *ncc = fe->ncc;
    if (nse) *nse = fe->nse;
    if (sbr) *sbr = fe->sbr == 1;
    if (ps) *ps = fe->ps == 1 && fe->ncc == 1 && fe->sbr == 1;
    if (sr_index) *sr_index = fe->sr_index;
    if (elem_type) for (int e = 0; e < fe->n_elems; e++) elem_type[e] = (uint8_t)fe->elem_type[e];
    return fe->n_elems;
}

/* Layouts with coupling channel elements: the coupling side info of the frame heaac_fe_parse has just returned, one HeaacCce per
 * CCE in layout order (what HeaacBatch.cce takes).  Returns the number of CCEs of the layout. */
int heaac_fe_cce(const FE *fe, HeaacCce *out)
{
    if (!fe) return HEAAC_E_INVAL;
    if (out) memcpy(out, fe->cce, sizeof(HeaacCce) * (size_t)fe->n_cce);
    return fe->n_cce;
}
/* 1 when the layout came from a program config element: elements -- and output channels -- are then in the reference's
 * (instance tag, type) order and HeaacConfig.out_slot is the identity */
int heaac_fe_layout_from_pce(const FE *fe) { return fe && fe->by_tag; }

/* The device front end's handling of a chunk (phase A for every frame, then fe_chunk_b: the SBR payloads' first step for all
 * frames at once, the rest in frame order), run on the host with one lane: same results as n_frames calls of heaac_fe_parse by
 * construction -- tests hold it to that, including what a failed frame leaves behind (consumed[f] < 0, silent frame).
 * Arrays are frame-major for the layout `fe` has (coeffs [n_frames][ncc][1024], ics / tns [n_frames][ncc], sbr
 * [n_frames][n_sbr_elems] or NULL, ps [n_frames] or NULL); n_frames <= 32; the layout must be known (extradata). */
static __thread int chunk_stats[4];
/* how the calling thread's last heaac_fe_parse_chunk went: [0] passes of the SBR payloads' first step, [1] frames that went through
 * the one-pass parser, [2] of those: because a header changed how SBR data is read */
void heaac_fe_chunk_stats(int out[4]) { memcpy(out, chunk_stats, sizeof(chunk_stats)); }

int heaac_fe_parse_chunk(FE *fe, int n_frames, const uint8_t *const *pkt, const int *size, float *coeffs, HeaacIcs *ics,
                         HeaacTns *tns, HeaacSbrElem *sbr, HeaacPsFrame *ps, int *consumed)
{
    if (!fe || !pkt || !size || !coeffs || !ics || !tns || !consumed || n_frames < 1 || n_frames > 32 || !fe->ncc) return HEAAC_E_INVAL;
    const int ncc = fe->ncc, nse = fe->nse, P = nse > 0 ? nse : 1;
    size_t total = 0;
    uint32_t off[32];
    for (int f = 0; f < n_frames; f++) {
        if (size[f] <= 0 || !pkt[f]) return HEAAC_E_INVAL;
        off[f] = (uint32_t)total;
        total += ((size_t)size[f] + 15) & ~(size_t)7;
    }
    uint8_t *blob = (uint8_t *)calloc(1, total + 8);
    FE *scratch = (FE *)malloc(sizeof(FE));
    FeTodo *todo = (FeTodo *)calloc(n_frames, sizeof(FeTodo));
    FeSbrTape *tape = (FeSbrTape *)calloc((size_t)n_frames * P, sizeof(FeSbrTape));
    int r = HEAAC_E_NOMEM;
    if (blob && scratch && todo && tape) {
        for (int f = 0; f < n_frames; f++) {
            memcpy(blob + off[f], pkt[f], size[f]);
            fe_phase_a(fe, scratch, &todo[f], blob + off[f], size[f], coeffs + (size_t)f * ncc * 1024, ics + (size_t)f * ncc,
                       tns + (size_t)f * ncc);
        }
        FeChunkIO io;
        memset(&io, 0, sizeof(io));
        io.n_frames = n_frames; io.ncc = ncc; io.nse = nse; io.stride = 1;
        io.todo = todo; io.tape = tape; io.tapes_per_frame = P;
        io.blob = blob; io.off = off; io.size = size;
        io.coeffs = coeffs; io.ics = ics; io.tns = tns; io.sbr = sbr; io.ps = ps; io.consumed = consumed;
        memset(chunk_stats, 0, sizeof(chunk_stats));
        io.stats = chunk_stats;
        fe_chunk_b(fe, &io, 0, 1);
        r = 0;
    }
    free(blob); free(scratch); free(todo); free(tape);
    return r;
}
