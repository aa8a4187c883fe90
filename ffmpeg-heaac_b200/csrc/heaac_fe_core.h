/*
 * heaac_fe_core.h -- the front end proper: raw_data_block / ADTS parsing for AAC-LC, HE-AAC v1 and v2 into the
 * per-frame structures of the batched C-ABI (include/heaac_b200.h), written once for two targets:
 *   host   (heaac_frontend.c):   heaac_fe_* entry points, one HeaacFrontend per stream on host threads
 *   device (heaac_parse_gpu.cu): FE_DEVICE defined, one GPU thread per stream, packets in, structures written
 *                                straight into the back end's device staging buffers
 *
 * It restates, in own code, what the reference's parsers do up to the spectral_to_sample() call site
 * (libavcodec/aacdec.c:2078):
 *   ADTS header                  aac_parser.c:29-67, aacdec.c:1935-1971
 *   element loop                 aacdec.c:2002-2076
 *   ics_info / section / scalefactors / pulse / TNS / spectrum  aacdec.c:645-1245, 1332-1385
 *   CPE, M/S, intensity          aacdec.c:1390-1494
 *   fill element / implicit SBR  aacdec.c:1650-1690
 *   SBR header, grid, envelopes, noise, harmonics, extension   aacsbr.c:207-259, 609-1086
 *   PS data                      aacps.c:84-279
 * Dequantised coefficients are produced with the reference's arithmetic (in-book values from the codebook
 * vector literals, escapes through cbrtf(i)*i, 2^((sf-200)/4) scalefactor table) so that they equal the
 * reference's `coeffs[]` value for value.
 * Not supported (as in the scope table): PCE / channel_config 0, CCE, AAC-Main prediction, LTP, SSR, 960 frames.
 */
#ifndef HEAAC_FE_CORE_H
#define HEAAC_FE_CORE_H
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/heaac_b200.h"
#include "heaac_sbr_tables_core.h"

#define FE_MAX_CH 16
#define FE_MAX_CCE 4
#ifdef FE_DEVICE
#define FE_INL static __device__ __forceinline__
/* big routines off the hot paths of the device kernels -- the one-pass frame parser (phase B's fallback for frames phase A could not
 * do statelessly; phase A has its own inlined, specialised instance), ps_data, the table derivation: called, not inlined, so that
 * their registers and code do not weigh on the paths around them */
#define FE_BIG static __device__ __noinline__
#else
#define FE_INL static inline __attribute__((always_inline))
#define FE_BIG static
#endif
/* lanes of a warp that walk a structure together (second step of the SBR payload, packing); `nl` = the number of cooperating
 * lanes in scope: 1 on the host, and on the device where lane 0 goes through the one-pass parser alone (no barrier there: the
 * other lanes are not in that code) */
#ifdef FE_DEVICE
#define FE_SYNCWARP() do { if (nl > 1) __syncwarp(); } while (0)
#define FE_SYNCWARP_ALL() __syncwarp()
#define FE_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define FE_SYNCWARP() ((void)0)
#define FE_SYNCWARP_ALL() ((void)0)
#define FE_ANY(x) (x)
#endif

/* ---------------------------------------------------------------- bit reader */
/* MSB-first, with a 64-bit window held in registers: `cache` = bits [cpos, cpos + 64) of the packet, cpos a
 * multiple of 32.  A field of up to 25 bits is two shifts while it lies inside the window; otherwise the window
 * is reloaded at the 32-bit word that holds `pos` (so at least 33 bits follow).
 * Host: the reload is one unaligned 64-bit load while 8 bytes remain, a byte-wise zero-filling path at the tail,
 * so no input padding is required.  Device: packets sit 8-byte aligned in the staging blob followed by >= 8 zero
 * bytes, the reload is two aligned 32-bit loads.  Positions past the packet end read as zero on both. */
typedef struct { const uint8_t *buf; int nbits, pos, cpos; uint64_t cache; } BR;

FE_INL void br_refill(BR *b)
{
    const int w = b->pos >> 5;                                  /* 32-bit word that holds pos */
    b->cpos = w << 5;
    if (b->pos < 0 || b->pos >= b->nbits) { b->cache = 0; b->cpos = b->pos & ~31; return; }
#ifdef FE_DEVICE
    const unsigned *p = (const unsigned *)b->buf + w;
    b->cache = ((uint64_t)__byte_perm(p[0], 0, 0x0123) << 32) | __byte_perm(p[1], 0, 0x0123);
#else
    const int byte = w * 4, size = (b->nbits + 7) >> 3;
    if (byte + 8 <= size) {
        uint64_t v;
        memcpy(&v, b->buf + byte, 8);
        b->cache = __builtin_bswap64(v);
    } else {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v = (v << 8) | (byte + i < size ? b->buf[byte + i] : 0u);
        b->cache = v;
    }
#endif
}
FE_INL void br_init(BR *b, const uint8_t *buf, int size)
{
    b->buf = buf; b->nbits = size * 8; b->pos = 0;
    br_refill(b);
}
FE_INL unsigned br_show(BR *b, int n)                         /* 1 <= n <= 25 */
{
    unsigned off = (unsigned)(b->pos - b->cpos);
    if (off > (unsigned)(64 - n)) { br_refill(b); off = (unsigned)(b->pos - b->cpos); }
    return (unsigned)((b->cache << off) >> (64 - n));
}
FE_INL unsigned br_get(BR *b, int n) { if (!n) return 0; unsigned v = br_show(b, n); b->pos += n; return v; }
FE_INL unsigned br_get1(BR *b) { return br_get(b, 1); }
FE_INL void br_skip(BR *b, int n) { b->pos += n; }
FE_INL int br_left(const BR *b) { return b->nbits - b->pos; }

/* ---------------------------------------------------------------- tables built at start-up */
/* Two-level Huffman tables: VLC_PRIM bits index the primary table; codes longer than that go through one
 * sub-table per primary prefix.  Entry = (symbol or sub-table offset) << 8 | flag 0x80 (sub-table) | length (total
 * code length, or the sub-table's index width); 0 = no such code.  All tables live in one pool addressed by
 * offset, so the whole FeTables block can be copied to the device as is. */
#define VLC_PRIM 9
#define FE_POOL_WORDS (96 * 1024)
typedef struct { uint32_t off; int maxlen; } Vlc;

enum { T_ENV_15, F_ENV_15, T_ENV_BAL_15, F_ENV_BAL_15, T_ENV_30, F_ENV_30, T_ENV_BAL_30, F_ENV_BAL_30, T_NOISE_30, T_NOISE_BAL_30 };
enum { PS_IID_DF1, PS_IID_DT1, PS_IID_DF0, PS_IID_DT0, PS_ICC_DF, PS_ICC_DT, PS_IPD_DF, PS_IPD_DT, PS_OPD_DF, PS_OPD_DT };

typedef struct FeTables {
    Vlc vsf, vspec[11], vsbr[10], vps[10];
    uint32_t pool_used;
    float pow2sf[428];                 /* 2^((i-200)/4), aac_tablegen.h:33-38 */
    float pow43[8192];                 /* escapes: cbrtf(i)*i, cbrt_tablegen.h:35-44 */
    float cb_val[16];                  /* in-book values: the literals of codebook_vector*_vals, aactab.c:899-1008 */
    float spec_val[11][289][4] __attribute__((aligned(16)));   /* codeword -> dequantised values (signed books) / magnitudes */
    float tns_map[4][16];              /* tns_tmp2_map, aacdectab.h:43-69 */
    uint16_t swb1024[13][64], swb128[13][32];
    uint8_t num_swb1024[13], num_swb128[13];
    uint32_t pool[FE_POOL_WORDS];
} FeTables;

#ifdef FE_DEVICE
static __device__ const FeTables *fe_tb_dev;
#define TB fe_tb_dev
#else
static FeTables *fe_tb_host;
#define TB fe_tb_host
#endif

FE_TAB int spec_sizes[11] = { 81, 81, 81, 81, 81, 81, 64, 64, 169, 169, 289 };
FE_TAB int8_t sbr_lav[10] = { 60, 60, 24, 24, 31, 31, 12, 12, 31, 12 };      /* aacsbr.c:73-74 */
FE_TAB int8_t ps_off[10] = { 30, 30, 14, 14, 7, 7, 0, 0, 0, 0 };             /* aacpsdata.c:136-142 */

FE_INL int vlc_get_tab(BR *b, const uint32_t *tab, int maxlen)       /* all code lengths are <= 25 - ... <= 20 bits */
{
    const unsigned x = br_show(b, 25);
    uint32_t e = tab[x >> (25 - VLC_PRIM)];
    if (e & 0x80u) {
        const int sb = e & 31;
        e = tab[(e >> 8) + ((x >> (25 - VLC_PRIM - sb)) & ((1u << sb) - 1))];
    }
    if (!e) { b->pos += maxlen; return -1; }
    b->pos += e & 31;
    return (int)(e >> 8);
}
FE_INL int vlc_get(BR *b, const Vlc *v) { return vlc_get_tab(b, TB->pool + v->off, v->maxlen); }

/* 4 / 2 consecutive coefficients at once (band offsets are multiples of 4, pair positions of 2) */
FE_INL void fe_store4(float *p, float a, float b_, float c, float d)
{
#ifdef FE_DEVICE
    *(float4 *)p = make_float4(a, b_, c, d);
#else
    p[0] = a; p[1] = b_; p[2] = c; p[3] = d;
#endif
}
FE_INL void fe_store2(float *p, float a, float b_)
{
#ifdef FE_DEVICE
    *(float2 *)p = make_float2(a, b_);
#else
    p[0] = a; p[1] = b_;
#endif
}

FE_FN const uint16_t *heaac_fe_swb_offset(int sr_index, int is8, int *num_swb)
{
    *num_swb = is8 ? TB->num_swb128[sr_index] : TB->num_swb1024[sr_index];
    return is8 ? TB->swb128[sr_index] : TB->swb1024[sr_index];
}

FE_FN void fe_zero(float *p, int n)                             /* p 16-byte aligned and n a multiple of 4 wherever it matters */
{
    int i = 0;
    if (!((uintptr_t)p & 15)) for (; i + 4 <= n; i += 4) fe_store4(p + i, 0.0f, 0.0f, 0.0f, 0.0f);
    for (; i < n; i++) p[i] = 0.0f;
}
FE_INL void fe_copy32(void *dst, const void *src, int nwords)   /* both 4-byte aligned */
{
    uint32_t *d = (uint32_t *)dst;
    const uint32_t *q = (const uint32_t *)src;
    for (int i = 0; i < nwords; i++) d[i] = q[i];
}
FE_INL void fe_zero32(void *dst, int nwords)
{
    uint32_t *d = (uint32_t *)dst;
    for (int i = 0; i < nwords; i++) d[i] = 0;
}
FE_INL void fe_zero16(void *dst, int nhalf)                     /* 2-byte aligned */
{
    uint16_t *d = (uint16_t *)dst;
    for (int i = 0; i < nhalf; i++) d[i] = 0;
}

/* ---------------------------------------------------------------- state */
typedef struct FeChan {
    int ws[2], shape[2];                        /* window_sequence / use_kb_window [current, previous] */
    int max_sfb, num_groups, group_len[8], num_windows, num_swb;
    const uint16_t *swb;
    uint8_t band_type[128];
    uint8_t run_end[128];
    float sf[128];
    HeaacTns tns; int tns_present;
} FeChan;

typedef struct FeSbrCh {
    int frame_class, num_env, num_noise, amp_res, add_harmonic_flag;
    uint8_t freq_res[8], t_env[8], t_q[4], add_harmonic[48] /* 4-byte aligned */, df_env[5], df_noise[2], invf[2][5];
    uint8_t q_at;                               /* tapes: the border (index into t_env) where the second noise floor starts */
    int e_a[2], t_env_num_env_old;
    int16_t env[6][48], noise[3][5];            /* row 0 = last row of the previous frame (delta-time base) */
} FeSbrCh;

typedef struct FeSbr {
    int start, reset, amp_res_header, coupling;
    int start_freq, stop_freq, xover, freq_scale, alter_scale, noise_bands;
    int limiter_bands, limiter_gains, interpol_freq, smoothing_mode;
    HeaacSbrElem t;                             /* derived tables + kx/m */
    FeSbrCh ch[2];
} FeSbr;

typedef struct FePs {
    int start, enable_iid, iid_quant, nr_iid_par, nr_ipdopd_par, enable_icc, icc_mode, nr_icc_par, enable_ext;
    int frame_class, num_env_old, num_env, enable_ipdopd, is34, is34_old;
    int border[6];
    int8_t iid[5][34], icc[5][34], ipd[5][34], opd[5][34];
} FePs;

/* What one SBR payload says, before it meets the element's running state (see "SBR payload, in two steps") */
typedef struct FeSbrTape {
    int header;                                 /* a header came with the payload: hdr[] (SBH_*) */
    int hdr[11];
    int status;                                 /* TAPE_* */
    int coupling;
    int ext_pos, ext_left;                      /* extension blocks behind the data: bit position and size in bits */
    FeSbrCh ch[2];                              /* grid, flags, rows 1.. of env / noise (time-coded rows as differences) */
} FeSbrTape;

/* What the stateless first phase of the two-phase parse (fe_phase_a / fe_phase_b below) leaves for the second. */
#define FE_MAX_TODO 8
typedef struct FeTodo {
    int result;                                 /* phase A's return value: bytes consumed or a negative error */
    int serial;                                 /* the frame has to go through the one-pass parser (PNS, errors, overflow) */
    int n_sbr;
    struct { int pos, cnt, crc, prev_elem, prev_type; } sbr[FE_MAX_TODO];   /* SBR fill payloads: where and after what */
    uint8_t chan[FE_MAX_CH];                    /* 0 absent, 1 read its own ics_info, 2 shares the previous channel's */
} FeTodo;

struct HeaacFrontend {
    int sample_rate, sr_index, chan_config, object_type;
    int sbr, ps;                                /* -1 unknown (implicit signalling), 0 absent, 1 present */
    int ext_sample_rate, locked;
    int layout_fixed;                           /* batch use: the layout was fixed when the stream was opened; a header that disagrees is an error */
    int n_elems, elem_type[HEAAC_MAX_ELEMS];
    int elem_ch0[HEAAC_MAX_ELEMS];              /* first core channel of each element */
    int ncc, nse;
    int sf_offset;                              /* 60: s16-unit scaling, 0: C bias scaling (aacdec.c:573-581) */
    int random_state;
    int phase_a, saw_pns, last_common;          /* two-phase parse bookkeeping */
    FeTodo *todo;
    FeChan ch[FE_MAX_CH];
    FeSbr sbr_el[HEAAC_MAX_ELEMS];
    FePs psx;
    /* layouts from a program config element (host build only; the device front end takes channel_configuration 1-6) */
    int by_tag;                                 /* elements are addressed by (type, instance tag) -- che[type][id] of the reference */
    int elem_tag[HEAAC_MAX_ELEMS];
    int n_cce;
    HeaacCce cce[FE_MAX_CCE];                   /* coupling side info of the last parsed frame, one per CCE in layout order */
    FeSbrTape tape;                             /* the one-pass parser's tape */
};
typedef struct HeaacFrontend FE;

FE_TAB int sample_rates[16] = { 96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350, 0, 0, 0 };

/* ---------------------------------------------------------------- configuration */
FE_TAB int fe_lay[8][6] = {          /* element types in bitstream order (set_default_channel_config, aacdec.c) */
    { 0 }, { 1, HEAAC_TYPE_SCE }, { 1, HEAAC_TYPE_CPE }, { 2, HEAAC_TYPE_SCE, HEAAC_TYPE_CPE },
    { 3, HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_SCE }, { 3, HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE },
    { 4, HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE, HEAAC_TYPE_LFE }, { 0 } };
FE_FN int set_layout(FE *fe, int chan_config)
{
    if (chan_config < 1 || chan_config > 6) return HEAAC_E_UNSUPPORTED;
    fe->chan_config = chan_config;
    fe->n_elems = fe_lay[chan_config][0];
    fe->ncc = fe->nse = 0;
    fe->by_tag = 0; fe->n_cce = 0;
    for (int e = 0; e < fe->n_elems; e++) {
        fe->elem_type[e] = fe_lay[chan_config][1 + e];
        fe->elem_ch0[e] = fe->ncc;
        fe->ncc += fe->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
        fe->nse += fe->elem_type[e] != HEAAC_TYPE_LFE;
    }
    return 0;
}

/* program_config_element (ISO/IEC 14496-3 table 4.2; decode_pce, aacdec.c:280-340): reads the element and marks the listed
 * (type, tag) pairs in `present` (NULL: only steps over it -- a further PCE of a configured stream is not evaluated,
 * aacdec.c:2020-2030; that is all the device build needs).  Returns 0 or a negative HEAAC_E_* code. */
FE_BIG int fe_pce_read(BR *b, uint8_t (*present)[16])
{
    br_skip(b, 2);                                                 /* object_type */
    br_skip(b, 4);                                                 /* sampling_frequency_index (the container's wins, a mismatch is only a warning) */
    int count[6];                                                  /* front, side, back, lfe, assoc_data, cc */
    for (int i = 0; i < 6; i++) count[i] = br_get(b, i == 3 ? 2 : i == 4 ? 3 : 4);
    if (br_get1(b)) br_skip(b, 4);                                 /* mono_mixdown */
    if (br_get1(b)) br_skip(b, 4);                                 /* stereo_mixdown */
    if (br_get1(b)) br_skip(b, 3);                                 /* matrix_mixdown */
    for (int grp = 0; grp < 6; grp++)
        for (int i = 0; i < count[grp]; i++) {
            if (grp == 4) { br_skip(b, 4); continue; }             /* associated data elements */
            const int flag = grp == 3 ? 0 : br_get1(b), tag = br_get(b, 4);
            const int type = grp == 3 ? HEAAC_TYPE_LFE : grp == 5 ? HEAAC_TYPE_CCE : flag ? HEAAC_TYPE_CPE : HEAAC_TYPE_SCE;
            if (present) present[type][tag] = 1;                   /* (cc_element_is_ind_sw is not kept: every CCE says so itself) */
        }
    b->pos = (b->pos + 7) & ~7;
    const int comment = 8 * br_get(b, 8);
    if (br_left(b) < comment) return HEAAC_E_INVAL;
    br_skip(b, comment);
    return 0;
}

#ifndef FE_DEVICE
/* The reference marks che_pos[type][tag] for every element a PCE lists and then walks tags 0..15 x types SCE, CPE, CCE, LFE
 * to allocate elements and number the output channels (output_configure, aacdec.c:246-262): the layout built here is that
 * walk -- elements ordered by (tag, type), output channels in element order, coupling channels without an output of their own.
 * apply != 0: make it the stream's layout; 0: only step over it. */
FE_FN int fe_pce(FE *fe, BR *b, int apply)
{
    uint8_t present[4][16];
    memset(present, 0, sizeof(present));
    int r = fe_pce_read(b, apply ? present : NULL);
    if (r) return r;
    if (!apply) return 0;
    int n = 0, ncc = 0, nse = 0, n_cce = 0;
    for (int tag = 0; tag < 16; tag++)
        for (int type = 0; type < 4; type++) {
            if (!present[type][tag]) continue;
            if (n >= HEAAC_MAX_ELEMS || ncc + 1 + (type == HEAAC_TYPE_CPE) > FE_MAX_CH) return HEAAC_E_UNSUPPORTED;
            if (type == HEAAC_TYPE_CCE && n_cce++ >= FE_MAX_CCE) return HEAAC_E_UNSUPPORTED;
            fe->elem_type[n] = type; fe->elem_tag[n] = tag; fe->elem_ch0[n] = ncc;
            ncc += type == HEAAC_TYPE_CPE ? 2 : 1;
            nse += type == HEAAC_TYPE_SCE || type == HEAAC_TYPE_CPE;
            n++;
        }
    if (!n) return HEAAC_E_INVAL;
    fe->n_elems = n; fe->ncc = ncc; fe->nse = nse; fe->n_cce = n_cce;
    fe->chan_config = 0; fe->by_tag = 1;
    memset(fe->cce, 0, sizeof(fe->cce));
    return 0;
}

#endif


/* ---------------------------------------------------------------- ICS */
FE_FN int ics_info(FE *fe, FeChan *c, BR *b)                     /* aacdec.c:645-710 */
{
    if (br_get1(b)) return HEAAC_E_INVAL;
    c->ws[1] = c->ws[0]; c->ws[0] = br_get(b, 2);
    c->shape[1] = c->shape[0]; c->shape[0] = br_get1(b);
    c->num_groups = 1; c->group_len[0] = 1;
    if (c->ws[0] == HEAAC_EIGHT_SHORT) {
        c->max_sfb = br_get(b, 4);
        for (int i = 0; i < 7; i++) {
            if (br_get1(b)) c->group_len[c->num_groups - 1]++;
            else { c->num_groups++; c->group_len[c->num_groups - 1] = 1; }
        }
        c->num_windows = 8;
        c->swb = heaac_fe_swb_offset(fe->sr_index, 1, &c->num_swb);
    } else {
        c->max_sfb = br_get(b, 6);
        c->num_windows = 1;
        c->swb = heaac_fe_swb_offset(fe->sr_index, 0, &c->num_swb);
        if (br_get1(b)) return HEAAC_E_UNSUPPORTED;                /* predictor_data_present: not AAC-LC */
    }
    if (c->max_sfb > c->num_swb) return HEAAC_E_INVAL;
    return 0;
}

FE_FN int band_types(FeChan *c, BR *b)                           /* aacdec.c:720-756 */
{
    int idx = 0;
    const int bits = c->ws[0] == HEAAC_EIGHT_SHORT ? 3 : 5;
    for (int g = 0; g < c->num_groups; g++) {
        int k = 0;
        while (k < c->max_sfb) {
            int end = k, incr, bt = br_get(b, 4);
            if (bt == 12) return HEAAC_E_INVAL;
            while ((incr = br_get(b, bits)) == (1 << bits) - 1) end += incr;
            end += incr;
            if (br_left(b) < 0 || end > c->max_sfb) return HEAAC_E_INVAL;
            for (; k < end; k++) { c->band_type[idx] = bt; c->run_end[idx++] = end; }
        }
    }
    return 0;
}

/* Scale factor data (ISO/IEC 14496-3 4.6.2.3; what decode_scalefactors yields, aacdec.c:768-822).  Every coded band belongs to one
 * of three KINDS -- spectral, noise (PNS), intensity position -- and each kind keeps one running value over the whole channel,
 * advanced by a Huffman-coded difference per band (the first noise band carries a 9-bit PCM word instead).  A kind is described
 * by how its running value is range-checked and where it points into the one 2^(x/4) table (the table in the function). */
FE_FN int scalefactors(FE *fe, FeChan *c, BR *b, int global_gain)
{
    enum { K_SPECTRAL, K_NOISE, K_INTENSITY };
    const int sfo = fe->sf_offset + (c->ws[0] == HEAAC_EIGHT_SHORT ? 12 : 0);
    /* kind:          spectral          noise             intensity
     * running value  global_gain       global_gain - 90  0
     * range check    value             value             value + 100          (each within 0..255)
     * table index    sfo + value       sfo + 100 + value 200 - value
     * sign           -                 -                 +                                                    */
    int v_spectral = global_gain, v_noise = global_gain - 90, v_intensity = 0;      /* (scalars: they stay in registers on the device) */
    int noise_pcm = 1;
    const int n = c->num_groups * c->max_sfb;
    for (int idx = 0; idx < n; idx++) {
        const int bt = c->band_type[idx];
        if (bt == 0) { c->sf[idx] = 0.f; continue; }
        const int k = bt == 13 ? K_NOISE : bt >= 14 ? K_INTENSITY : K_SPECTRAL;
        int value = k == K_NOISE ? v_noise : k == K_INTENSITY ? v_intensity : v_spectral;
        if (k == K_NOISE && noise_pcm) { value += br_get(b, 9) - 256; noise_pcm = 0; }
        else value += vlc_get(b, &TB->vsf) - 60;
        if ((unsigned)(value + (k == K_INTENSITY ? 100 : 0)) > 255U) return HEAAC_E_INVAL;
        const float gain = TB->pow2sf[k == K_INTENSITY ? 200 - value : sfo + value + (k == K_NOISE ? 100 : 0)];
        c->sf[idx] = k == K_INTENSITY ? gain : -gain;
        if (k == K_NOISE) v_noise = value; else if (k == K_INTENSITY) v_intensity = value; else v_spectral = value;
    }
    return 0;
}

FE_FN int tns_data(FeChan *c, BR *b)                             /* aacdec.c:854-887 */
{
    const int is8 = c->ws[0] == HEAAC_EIGHT_SHORT;
    const int max_order = is8 ? 7 : 12;
    fe_zero32(&c->tns, (int)(sizeof(c->tns) / 4));
    for (int w = 0; w < c->num_windows; w++) {
        int nf = br_get(b, 2 - is8);
        c->tns.n_filt[w] = nf;
        if (!nf) continue;
        int coef_res = br_get1(b);
        for (int f = 0; f < nf; f++) {
            int slot = is8 ? w : f;
            c->tns.length[slot] = br_get(b, 6 - 2 * is8);
            int order = br_get(b, 5 - 2 * is8);
            if (order > max_order) return HEAAC_E_INVAL;
            c->tns.order[slot] = order;
            if (order) {
                c->tns.direction[slot] = br_get1(b);
                int compress = br_get1(b), len = coef_res + 3 - compress, map = 2 * compress + coef_res;
                for (int i = 0; i < order; i++) c->tns.coef[slot][i] = TB->tns_map[map][br_get(b, len)];
            }
        }
    }
    return 0;
}

/* spectral data + dequantisation, aacdec.c:988-1245.  Value of a coefficient: sign * |q|^(4/3) * sf with the
 * products evaluated exactly as the reference does (table value times scalefactor). */
FE_FN int spectrum(FE *fe, FeChan *c, BR *b, float *coef, int pulse_present, const int *ppos, const int *pamp, int npulse)
{
    const int cw = 1024 / c->num_windows;
    const uint16_t *off = c->swb;
    float *base = coef;
    int idx = 0;
    for (int g = 0; g < c->num_windows; g++)
        fe_zero(coef + g * 128 + off[c->max_sfb], cw - off[c->max_sfb]);
    for (int g = 0; g < c->num_groups; g++) {
        int glen = c->group_len[g];
        for (int i = 0; i < c->max_sfb; i++, idx++) {
            const int bt = c->band_type[idx];
            float *cfo = coef + off[i];
            const int len = off[i + 1] - off[i];
            const float sf = c->sf[idx];
            if (bt == 0 || bt >= 14) {
                for (int w = 0; w < glen; w++, cfo += 128) fe_zero(cfo, len);
            } else if (bt == 13) {                                 /* PNS, aacdec.c:1021-1036 */
                fe->saw_pns = 1;
                for (int w = 0; w < glen; w++, cfo += 128) {
                    float energy = 0.0f;
                    for (int k = 0; k < len; k++) {
                        fe->random_state = (int)((unsigned)fe->random_state * 1664525u + 1013904223u);   /* lcg_random, aacdec.c:508-511 */
                        cfo[k] = fe->random_state;
                    }
                    for (int k = 0; k < len; k++) energy += cfo[k] * cfo[k];
                    float scale = sf / sqrtf(energy);
                    for (int k = 0; k < len; k++) cfo[k] = cfo[k] * scale;
                }
            } else {
                const uint32_t *tab = TB->pool + TB->vspec[bt - 1].off;
                const int maxlen = TB->vspec[bt - 1].maxlen;
                const float (*val)[4] = TB->spec_val[bt - 1];
                const int size = spec_sizes[bt - 1];
                if (bt == 1 || bt == 2) {                          /* signed quads */
                    for (int w = 0; w < glen; w++, cfo += 128)
                        for (int k = 0; k < len; k += 4) {
                            const int code = vlc_get_tab(b, tab, maxlen);
                            if ((unsigned)code >= (unsigned)size) return HEAAC_E_INVAL;
                            const float *v = val[code];
                            fe_store4(cfo + k, v[0] * sf, v[1] * sf, v[2] * sf, v[3] * sf);
                        }
                } else if (bt == 5 || bt == 6) {                   /* signed pairs */
                    for (int w = 0; w < glen; w++, cfo += 128)
                        for (int k = 0; k < len; k += 2) {
                            const int code = vlc_get_tab(b, tab, maxlen);
                            if ((unsigned)code >= (unsigned)size) return HEAAC_E_INVAL;
                            fe_store2(cfo + k, val[code][0] * sf, val[code][1] * sf);
                        }
                } else if (bt < 5) {                               /* unsigned quads + one sign bit per non-zero value */
                    for (int w = 0; w < glen; w++, cfo += 128)
                        for (int k = 0; k < len; k += 4) {
                            const int code = vlc_get_tab(b, tab, maxlen);
                            if ((unsigned)code >= (unsigned)size) return HEAAC_E_INVAL;
                            const float *v = val[code];
                            unsigned sg = br_show(b, 4) << 28;     /* next sign bit in the MSB */
                            int used = 0;
                            float m[4];
                            for (int d = 0; d < 4; d++) {
                                m[d] = v[d];
                                if (m[d] != 0.0f) { if (sg & 0x80000000u) m[d] = -m[d]; sg <<= 1; used++; }
                            }
                            b->pos += used;
                            fe_store4(cfo + k, m[0] * sf, m[1] * sf, m[2] * sf, m[3] * sf);
                        }
                } else if (bt < 11) {                              /* unsigned pairs + sign bits */
                    for (int w = 0; w < glen; w++, cfo += 128)
                        for (int k = 0; k < len; k += 2) {
                            const int code = vlc_get_tab(b, tab, maxlen);
                            if ((unsigned)code >= (unsigned)size) return HEAAC_E_INVAL;
                            unsigned sg = br_show(b, 2) << 30;
                            int used = 0;
                            float m0 = val[code][0], m1 = val[code][1];
                            if (m0 != 0.0f) { if (sg & 0x80000000u) m0 = -m0; sg <<= 1; used++; }
                            if (m1 != 0.0f) { if (sg & 0x80000000u) m1 = -m1; used++; }
                            b->pos += used;
                            fe_store2(cfo + k, m0 * sf, m1 * sf);
                        }
                } else {                                           /* book 11: pairs, sign bits, escapes (aacdec.c:1177-1209) */
                    for (int w = 0; w < glen; w++, cfo += 128)
                        for (int k = 0; k < len; k += 2) {
                            const int code = vlc_get_tab(b, tab, maxlen);
                            if ((unsigned)code >= (unsigned)size) return HEAAC_E_INVAL;
                            const int q0 = code / 17, q1 = code % 17;
                            const int n0 = q0 ? (int)br_get1(b) : 0, n1 = q1 ? (int)br_get1(b) : 0;
                            float m0 = val[code][0], m1 = val[code][1];
                            if (q0 == 16) {
                                int n = 0;
                                while (br_get1(b)) { if (++n > 8) return HEAAC_E_INVAL; }
                                n += 4;
                                m0 = TB->pow43[(1 << n) + br_get(b, n)];
                            }
                            if (q1 == 16) {
                                int n = 0;
                                while (br_get1(b)) { if (++n > 8) return HEAAC_E_INVAL; }
                                n += 4;
                                m1 = TB->pow43[(1 << n) + br_get(b, n)];
                            }
                            fe_store2(cfo + k, (n0 ? -m0 : m0) * sf, (n1 ? -m1 : m1) * sf);
                        }
                }
            }
        }
        coef += glen << 7;
    }
    if (pulse_present) {                                           /* aacdec.c:1219-1234 */
        idx = 0;
        for (int i = 0; i < npulse; i++) {
            float co = base[ppos[i]];
            while (off[idx + 1] <= ppos[i]) idx++;
            /* (a pulse beyond max_sfb would be scaled by whatever an EARLIER frame left in band_type / sf in the reference,
             *  aacdec.c:1226; such streams are not conformant -- here the pulse is dropped, on every path alike) */
            if (idx < c->max_sfb && c->band_type[idx] != 13 && c->sf[idx]) {
                float ico = -pamp[i];
                if (co) {
                    co /= c->sf[idx];
                    ico = co / sqrtf(sqrtf(fabsf(co))) + (co > 0 ? -ico : ico);
                }
                base[ppos[i]] = cbrtf(fabsf(ico)) * ico * c->sf[idx];
            }
        }
    }
    return 0;
}

FE_FN int decode_ics(FE *fe, FeChan *c, BR *b, float *coef, int common_window)       /* aacdec.c:1332-1385 */
{
    int r, ppos[4], pamp[4], npulse = 0, pulse_present;
    int global_gain = br_get(b, 8);
    if (!common_window && (r = ics_info(fe, c, b))) return r;
    if ((r = band_types(c, b))) return r;
    if ((r = scalefactors(fe, c, b, global_gain))) return r;
    if ((pulse_present = br_get1(b))) {                            /* aacdec.c:829-847 */
        if (c->ws[0] == HEAAC_EIGHT_SHORT) return HEAAC_E_INVAL;
        npulse = br_get(b, 2) + 1;
        int swb = br_get(b, 6);
        if (swb >= c->num_swb) return HEAAC_E_INVAL;
        ppos[0] = c->swb[swb] + br_get(b, 5);
        if (ppos[0] > 1023) return HEAAC_E_INVAL;
        pamp[0] = br_get(b, 4);
        for (int i = 1; i < npulse; i++) {
            ppos[i] = br_get(b, 5) + ppos[i - 1];
            if (ppos[i] > 1023) return HEAAC_E_INVAL;
            pamp[i] = br_get(b, 4);
        }
    }
    if ((c->tns_present = br_get1(b)) && (r = tns_data(c, b))) return r;
    if (br_get1(b)) return HEAAC_E_UNSUPPORTED;                    /* gain control (SSR) */
    return spectrum(fe, c, b, coef, pulse_present, ppos, pamp, npulse);
}

FE_FN int decode_cpe(FE *fe, FeChan *c0, FeChan *c1, BR *b, float *co0, float *co1)  /* aacdec.c:1460-1494 */
{
    uint32_t ms_mask_w[32];                                        /* 128 flags, cleared / set by words */
    uint8_t *ms_mask = (uint8_t *)ms_mask_w;
    int r, ms_present = 0;
    int common = br_get1(b);
    fe->last_common = common;
    fe_zero32(ms_mask_w, 32);
    if (common) {
        if ((r = ics_info(fe, c0, b))) return r;
        int keep = c1->shape[0];
        c1->ws[0] = c0->ws[0]; c1->ws[1] = c0->ws[1]; c1->shape[0] = c0->shape[0]; c1->shape[1] = keep;
        c1->max_sfb = c0->max_sfb; c1->num_groups = c0->num_groups; memcpy(c1->group_len, c0->group_len, sizeof(c0->group_len));
        c1->num_windows = c0->num_windows; c1->num_swb = c0->num_swb; c1->swb = c0->swb;
        ms_present = br_get(b, 2);
        if (ms_present == 3) return HEAAC_E_INVAL;
        if (ms_present == 1) for (int i = 0; i < c0->num_groups * c0->max_sfb; i++) ms_mask[i] = br_get1(b);
        else if (ms_present == 2) for (int i = 0; i < 32; i++) ms_mask_w[i] = 0x01010101u;
    }
    if ((r = decode_ics(fe, c0, b, co0, common))) return r;
    if ((r = decode_ics(fe, c1, b, co1, common))) return r;
    if (common && ms_present) {                                    /* apply_mid_side_stereo, aacdec.c:1390-1411 */
        float *a = co0, *d = co1;
        int idx = 0;
        for (int g = 0; g < c0->num_groups; g++) {
            for (int i = 0; i < c0->max_sfb; i++, idx++) {
                if (ms_mask[idx] && c0->band_type[idx] < 13 && c1->band_type[idx] < 13)
                    for (int w = 0; w < c0->group_len[g]; w++)
                        for (int k = c0->swb[i]; k < c0->swb[i + 1]; k += 4) {        /* band widths are multiples of 4 */
                            float *pa = a + w * 128 + k, *pd = d + w * 128 + k;
                            const float a0 = pa[0], a1 = pa[1], a2 = pa[2], a3 = pa[3];
                            const float d0 = pd[0], d1 = pd[1], d2 = pd[2], d3 = pd[3];
                            fe_store4(pa, a0 + d0, a1 + d1, a2 + d2, a3 + d3);
                            fe_store4(pd, a0 - d0, a1 - d1, a2 - d2, a3 - d3);
                        }
            }
            a += c0->group_len[g] * 128; d += c0->group_len[g] * 128;
        }
    }
    {                                                              /* apply_intensity_stereo, aacdec.c:1420-1452 */
        float *a = co0, *d = co1;
        int idx = 0;
        for (int g = 0; g < c1->num_groups; g++) {
            for (int i = 0; i < c1->max_sfb;) {
                if (c1->band_type[idx] == 14 || c1->band_type[idx] == 15) {
                    int end = c1->run_end[idx];
                    for (; i < end; i++, idx++) {
                        int cs = -1 + 2 * (c1->band_type[idx] - 14);
                        if (ms_present) cs *= 1 - 2 * ms_mask[idx];
                        float scale = cs * c1->sf[idx];
                        for (int w = 0; w < c1->group_len[g]; w++)
                            for (int k = c1->swb[i]; k < c1->swb[i + 1]; k++) d[w * 128 + k] = scale * a[w * 128 + k];
                    }
                } else {
                    int end = c1->run_end[idx];
                    idx += end - i; i = end;
                }
            }
            a += c1->group_len[g] * 128; d += c1->group_len[g] * 128;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- SBR */

#ifndef FE_DEVICE
/* coupling_channel_element (decode_cce, aacdec.c:1503-1568) for an INDEPENDENTLY SWITCHED element: one gain per list, the
 * first list's being 1.  The element / channel selection that apply_channel_coupling (aacdec.c:1870-1898) makes when it meets
 * a target is resolved here into the ABI's list of (target channel, gain) operations.  Anything else than AFTER_IMDCT
 * coupling is outside the ABI (HEAAC_E_UNSUPPORTED). */
FE_FN int decode_cce(FE *fe, int e, BR *b, float *coeffs)
{
    int k = 0;
    for (int i = 0; i < e; i++) k += fe->elem_type[i] == HEAAC_TYPE_CCE;
    HeaacCce *o = &fe->cce[k];
    int sel_type[8], sel_tag[8], sel_ch[8], lists = 0;
    const int ind_sw = br_get1(b), n_coupled = br_get(b, 3) + 1;
    for (int c = 0; c < n_coupled; c++) {
        sel_type[c] = br_get1(b) ? HEAAC_TYPE_CPE : HEAAC_TYPE_SCE;
        sel_tag[c] = br_get(b, 4);
        sel_ch[c] = sel_type[c] == HEAAC_TYPE_CPE ? (int)br_get(b, 2) : 2;
        lists += 1 + (sel_ch[c] == 3);
    }
    br_skip(b, 1);                                                 /* cc_domain: only meaningful for dependent coupling */
    if (!ind_sw) return HEAAC_E_UNSUPPORTED;
    br_skip(b, 1);                                                 /* gain_element_sign: per-band gains only */
    const float scale = (float)pow(2., pow(2., (int)br_get(b, 2) - 3));
    int r = decode_ics(fe, &fe->ch[fe->elem_ch0[e]], b, coeffs + 1024 * fe->elem_ch0[e], 0);
    if (r) return r;
    float gain[16];
    gain[0] = 1.f;
    for (int c = 1; c < lists; c++)
        gain[c] = (float)pow(scale, -(vlc_get(b, &TB->vsf) - 60));  /* common_gain_element_present is implied */
    int n = 0, list = 0;
    for (int c = 0; c < n_coupled; c++) {
        int te = -1;
        for (int t = 0; t < fe->n_elems; t++)
            if (fe->elem_type[t] == sel_type[c] && fe->elem_tag[t] == sel_tag[c]) te = t;
        if (te >= 0) {
            if (sel_ch[c] != 1 && n < HEAAC_CCE_MAX_TARGETS) { o->target_ch[n] = (uint8_t)fe->elem_ch0[te];       o->gain[n++] = gain[list]; }
            if (sel_ch[c] != 2 && n < HEAAC_CCE_MAX_TARGETS) { o->target_ch[n] = (uint8_t)(fe->elem_ch0[te] + 1); o->gain[n++] = gain[list + (sel_ch[c] == 3)]; }
        }
        list += 1 + (sel_ch[c] == 3);
    }
    for (int i = n; i < HEAAC_CCE_MAX_TARGETS; i++) { o->target_ch[i] = 0; o->gain[i] = 0.f; }
    o->coupling_point = 3; o->n_targets = (uint8_t)n;
    return 0;
}
#endif


FE_BIG int sbr_tables(FE *fe, FeSbr *s)                           /* sbr_reset, aacsbr.c:1023-1034 */
{
    HeaacSbrElem t = s->t;
    int r = fe_sbr_tables(fe->sample_rate, s->start_freq, s->stop_freq, s->xover, s->freq_scale, s->alter_scale,
                                  s->noise_bands, s->limiter_bands, &t);
    if (r) return r;
    s->t = t;
    return 0;
}

/* ---------------------------------------------------------------- SBR payload, in two steps
 * An SBR payload is read in two steps that meet at a FeSbrTape:
 *   sbr_*_bits   what the payload SAYS, read with nothing from the previous frame: the header fields, the time grid, the
 *                difference-coded envelope / noise rows with the differences along frequency already summed up and the
 *                differences along time left as differences.  Needs only the band counts of the element's tables (which
 *                change with a header that changes them, and with nothing else);
 *   sbr_*_apply  folds a tape into the element's running state: previous-frame bookkeeping of the grid, time differences
 *                added to the row before, last row carried over.  No bits are read; the rows are walked by `nl` cooperating
 *                lanes (host: lane 0 of 1; device: the 32 lanes of the stream's warp).
 * The one-pass parser runs both per payload.  The device front end runs the first step for all frames of a chunk at once,
 * one frame per lane, and the second in frame order (fe_chunk_b). */
enum { SBH_AMP_RES, SBH_START, SBH_STOP, SBH_XOVER, SBH_FREQ_SCALE, SBH_ALTER_SCALE, SBH_NOISE_BANDS,
       SBH_LIM_BANDS, SBH_LIM_GAINS, SBH_INTERPOL, SBH_SMOOTHING, SBH_COUNT };
enum { TAPE_DATA = 0,       /* ch[] hold the frame's data */
       TAPE_NONE = 1,       /* the element has not seen a header yet: the payload carries nothing to read */
       TAPE_BAD = 2,        /* invalid grid */
       TAPE_STALE = 3,      /* a header that changes how the data is read: the payload has to go through the one-pass parser */
       TAPE_SKIP = 4 };     /* a payload the frame parser steps over (nothing before it to attach it to, behind an LFE) */

FE_FN void sbr_header_bits(BR *b, int *h)                         /* read_sbr_header, aacsbr.c:207-259 */
{
    h[SBH_AMP_RES] = br_get1(b);
    h[SBH_START] = br_get(b, 4); h[SBH_STOP] = br_get(b, 4); h[SBH_XOVER] = br_get(b, 3);
    br_skip(b, 2);
    const int extra1 = br_get1(b), extra2 = br_get1(b);
    if (extra1) { h[SBH_FREQ_SCALE] = br_get(b, 2); h[SBH_ALTER_SCALE] = br_get1(b); h[SBH_NOISE_BANDS] = br_get(b, 2); }
    else { h[SBH_FREQ_SCALE] = 2; h[SBH_ALTER_SCALE] = 1; h[SBH_NOISE_BANDS] = 2; }
    if (extra2) { h[SBH_LIM_BANDS] = br_get(b, 2); h[SBH_LIM_GAINS] = br_get(b, 2); h[SBH_INTERPOL] = br_get1(b); h[SBH_SMOOTHING] = br_get1(b); }
    else { h[SBH_LIM_BANDS] = 2; h[SBH_LIM_GAINS] = 2; h[SBH_INTERPOL] = 1; h[SBH_SMOOTHING] = 1; }
}

/* 1 if the element, as it stands, reads its data the way a stream with header h does (same amplitude resolution, same tables) */
FE_FN int sbr_header_same(const FeSbr *s, const int *h)
{
    return s->start && s->amp_res_header == h[SBH_AMP_RES] && s->start_freq == h[SBH_START] && s->stop_freq == h[SBH_STOP] &&
           s->xover == h[SBH_XOVER] && s->freq_scale == h[SBH_FREQ_SCALE] && s->alter_scale == h[SBH_ALTER_SCALE] &&
           s->noise_bands == h[SBH_NOISE_BANDS];
}

FE_FN void sbr_header_apply(FE *fe, FeSbr *s, const int *h)
{
    if (s->start_freq != h[SBH_START] || s->stop_freq != h[SBH_STOP] || s->xover != h[SBH_XOVER] ||
        s->freq_scale != h[SBH_FREQ_SCALE] || s->alter_scale != h[SBH_ALTER_SCALE] || s->noise_bands != h[SBH_NOISE_BANDS]) s->reset = 1;
    const int old_lim = s->limiter_bands;
    s->start = 1;
    s->amp_res_header = h[SBH_AMP_RES];
    s->start_freq = h[SBH_START]; s->stop_freq = h[SBH_STOP]; s->xover = h[SBH_XOVER];
    s->freq_scale = h[SBH_FREQ_SCALE]; s->alter_scale = h[SBH_ALTER_SCALE]; s->noise_bands = h[SBH_NOISE_BANDS];
    s->limiter_bands = h[SBH_LIM_BANDS]; s->limiter_gains = h[SBH_LIM_GAINS];
    s->interpol_freq = h[SBH_INTERPOL]; s->smoothing_mode = h[SBH_SMOOTHING];
    if (s->limiter_bands != old_lim && !s->reset) {               /* sbr_make_f_tablelim only */
        HeaacSbrElem t = s->t;
        if (!fe_sbr_tables(fe->sample_rate, s->start_freq, s->stop_freq, s->xover, s->freq_scale, s->alter_scale,
                                   s->noise_bands, s->limiter_bands, &t)) {
            memcpy(s->t.f_tablelim, t.f_tablelim, sizeof(t.f_tablelim));
            s->t.n_lim = t.n_lim;
        }
    }
}

/* The time grid of a channel (ISO/IEC 14496-3 4.6.18.3.3; what read_sbr_grid yields, aacsbr.c:609-749).  The four frame classes
 * differ only in which ends of the frame are VARIABLE: class bit 1 = the leading border, bit 0 = the trailing one.  A variable end
 * sends its absolute border and a count of relative borders that step inwards from it by 2 r + 2 slots each; a fixed end sits
 * at slot 0 / 16.  With both ends fixed the envelopes divide the frame evenly instead.
 * Fills the grid fields of a tape channel (freq_res[1..], t_env, t_q, e_a[1]); what the grid takes over from the previous frame
 * (freq_res[0], e_a[0], t_env_num_env_old) is sbr_chan_apply's. */
FE_FN int sbr_grid_bits(FeSbrCh *c, BR *b, int amp_res_header)
{
    c->amp_res = amp_res_header;
    const int cls = c->frame_class = br_get(b, 2);
    const int var_lead = cls >> 1, var_trail = cls & 1;
    unsigned ptr = 0;
    if (!cls) {
        const int n = 1 << br_get(b, 2);
        c->num_env = n;
        if (n == 1) c->amp_res = 0;
        if (n > 4) return -1;
        const int step = (16 + (n >> 1)) / n;
        for (int e = 0; e < n; e++) c->t_env[e] = (uint8_t)(e * step);
        c->t_env[n] = 16;
        const int res = br_get1(b);
        for (int e = 1; e <= n; e++) c->freq_res[e] = (uint8_t)res;
    } else {
        const int first = var_lead ? br_get(b, 2) : 0;
        const int last = 16 + (var_trail ? br_get(b, 2) : 0);
        const int n_lead = var_lead ? br_get(b, 2) : 0, n_trail = var_trail ? br_get(b, 2) : 0;
        const int n = n_lead + n_trail + 1;
        c->num_env = n;
        if (n > 5) return -1;
        c->t_env[0] = (uint8_t)first;
        c->t_env[n] = (uint8_t)last;
        for (int e = 1; e <= n_lead; e++) c->t_env[e] = (uint8_t)(c->t_env[e - 1] + 2 * br_get(b, 2) + 2);
        for (int e = n - 1; e >= n - n_trail; e--) c->t_env[e] = (uint8_t)(c->t_env[e + 1] - 2 * br_get(b, 2) - 2);
        ptr = br_get(b, n > 3 ? 3 : n > 1 ? 2 : 1);                  /* enough bits for 0 .. n */
        /* one resolution flag per envelope; a frame that is variable at the end only sends them last envelope first */
        if (cls == 1) for (int e = n; e >= 1; e--) c->freq_res[e] = br_get1(b);
        else          for (int e = 1; e <= n; e++) c->freq_res[e] = br_get1(b);
    }
    const int n = c->num_env;
    if (ptr > (unsigned)n + 1) return -1;
    for (int e = 1; e <= n; e++) if (c->t_env[e - 1] > c->t_env[e]) return -1;
    /* noise floors: one for a single envelope, else two, split at the border the pointer selects (the middle one without a pointer) */
    c->num_noise = n > 1 ? 2 : 1;
    c->t_q[0] = c->t_env[0]; c->t_q[c->num_noise] = c->t_env[n];
    if (n > 1) {
        unsigned at;
        if (!cls) at = n >> 1;
        else if (var_trail) {
            /* counted back from the end: max(pointer - 1, 1), in UNSIGNED arithmetic as in the reference (aacsbr.c:729): pointer 0
             * wraps and lands one past the last border */
            const unsigned back = ptr - 1u;
            at = (unsigned)n - (back > 1u ? back : 1u);
        } else
            at = ptr == 0 ? 1u : ptr == 1 ? (unsigned)n - 1u : ptr - 1u;
        /* `at` can be n + 1 (pointer 0): the reference then reads the entry behind the frame's last border, i.e. whatever an
         * earlier frame left there (aacsbr.c:729-733) -- so the border is looked up when the grid meets the element's state */
        c->q_at = (uint8_t)at;
    }
    /* the envelope that starts at a transient (l_A) */
    c->e_a[1] = -1;
    if (var_trail && ptr) c->e_a[1] = n + 1 - (int)ptr;
    else if (cls == 2 && ptr > 1) c->e_a[1] = (int)ptr - 1;
    return 0;
}

FE_FN void sbr_dtdf_bits(FeSbrCh *c, BR *b)
{
    for (int i = 0; i < c->num_env; i++) c->df_env[i] = br_get1(b);
    for (int i = 0; i < c->num_noise; i++) c->df_noise[i] = br_get1(b);
}

/* One matrix of delta-coded SBR parameters -- envelope scalefactors or noise floors (ISO/IEC 14496-3 4.6.18.3.5; what
 * read_sbr_envelope / read_sbr_noise yield, aacsbr.c:787-890).  Row r + 1 is coded either along FREQUENCY (a start value of
 * `start_bits`, then Huffman-coded differences between neighbouring bands: summed up in the first step) or along TIME
 * (differences to row r: left as differences in the first step).  Every value is scaled by `scale` (2 for the balance channel of
 * a coupled pair).  A payload holds two to four such matrices back to back; sbr_data_bits lists them in a plan and reads them
 * ONE VALUE PER STEP (sbr_plan_step), so that lanes reading different frames side by side stay in lockstep. */
typedef struct SbrRows {
    int16_t *m;
    const uint8_t *along_time, *res;            /* per row; res[r + 1]: 0 = low, 1 = high resolution band table */
    int stride, rows, n_low, n_high, tab_time, tab_freq, start_bits, scale;
} SbrRows;
FE_TAB uint8_t sbr_res_none[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };       /* noise floors all sit on the noise band table */

/* The second step for such a matrix: rows along time are added to the row before them.  The two rows may sit on different band
 * tables (res[]: 0 = low, 1 = high resolution; res[0] = the last row of the previous frame): a high band then refers to the
 * low band that contains it, a low band to the high band it starts with (`odd`: the high table has an odd number of bands, so
 * the pairing is shifted by one).  Row 0 receives a copy of the last row: the base of the next frame's first time difference. */
FE_FN void sbr_rows_apply(int16_t *m, const int16_t *tape, int stride, int rows, const uint8_t *along_time, const uint8_t *res,
                          int n_low, int n_high, int odd, int lane, int nl)
{
    for (int r = 0; r < rows; r++) {
        int16_t *cur = m + (r + 1) * stride;
        const int16_t *prev = m + r * stride, *in = tape + (r + 1) * stride;
        const int hi = res[r + 1], n = hi ? n_high : n_low;
        const int time = along_time[r], same = res[r + 1] == res[r];
        for (int j = lane; j < n; j += nl) {
            int v = in[j];
            if (time) {
                const int ref = same ? j : hi ? (j + odd) >> 1 : (j ? 2 * j - odd : 0);
                v = (int16_t)(prev[ref] + v);
            }
            cur[j] = (int16_t)v;
        }
        FE_SYNCWARP();
    }
    for (int j = lane; j < stride; j += nl) m[j] = m[rows * stride + j];
    FE_SYNCWARP();
}

FE_FN void sbr_envelope_rows(SbrRows *q, const FeSbr *s, FeSbrCh *c, int ch, int coupling)
{
    const int balance = coupling && ch, fine = !c->amp_res;              /* 1.5 dB steps need one more start bit and their own tables */
    q->m = &c->env[0][0]; q->stride = 48; q->rows = c->num_env; q->along_time = c->df_env; q->res = c->freq_res;
    q->n_low = s->t.n[0]; q->n_high = s->t.n[1];
    q->tab_time = balance ? (fine ? T_ENV_BAL_15 : T_ENV_BAL_30) : (fine ? T_ENV_15 : T_ENV_30);
    q->tab_freq = balance ? (fine ? F_ENV_BAL_15 : F_ENV_BAL_30) : (fine ? F_ENV_15 : F_ENV_30);
    q->start_bits = (balance ? 5 : 6) + fine; q->scale = balance + 1;
}

FE_FN void sbr_noise_rows(SbrRows *q, const FeSbr *s, FeSbrCh *c, int ch, int coupling)
{
    const int balance = coupling && ch;
    q->m = &c->noise[0][0]; q->stride = 5; q->rows = c->num_noise; q->along_time = c->df_noise; q->res = sbr_res_none;
    q->n_low = q->n_high = s->t.n_q;
    q->tab_time = balance ? T_NOISE_BAL_30 : T_NOISE_30; q->tab_freq = balance ? F_ENV_BAL_30 : F_ENV_30;
    q->start_bits = 5; q->scale = balance + 1;
}

/* ---------------------------------------------------------------- PS (ff_ps_read_data, aacps.c:150-279) */
/* parameter bands per iid / icc mode (modes 3..5: the same counts with the finer iid quantiser) and of the phase parameters */
FE_TAB int8_t ps_bands_of_mode[6] = { 10, 20, 34, 10, 20, 34 }, ps_phase_bands_of_mode[6] = { 5, 11, 17, 5, 11, 17 };

/* One envelope's worth of one Parametric Stereo parameter (iid, icc, ipd or opd; ISO/IEC 14496-3 8.6.4.6.2).  A parameter is
 * described by its Huffman tables (frequency / time differences), the offset of their symbol range and how a value is kept in
 * range: iid and icc are range-CHECKED (a violation invalidates the frame), the phase parameters WRAP modulo 8. */
typedef struct PsParam { int8_t (*rows)[34]; int count, tab_freq, tab_time, lo, hi, wrap; } PsParam;

FE_FN int ps_param_env(const FePs *p, BR *b, const PsParam *q, int e)
{
    const int along_time = br_get1(b);
    const int table = along_time ? q->tab_time : q->tab_freq;
    const int off = q->wrap ? 0 : ps_off[table];
    int base_row = e ? e - 1 : p->num_env_old - 1;                  /* time differences refer to the previous envelope (of the previous frame) */
    if (base_row < 0) base_row = 0;
    int run = 0;
    for (int k = 0; k < q->count; k++) {
        int v = (along_time ? q->rows[base_row][k] : run) + vlc_get(b, &TB->vps[table]) - off;
        if (q->wrap) v &= 7;
        run = v;
        q->rows[e][k] = (int8_t)v;
        if (!q->wrap && (q->rows[e][k] < q->lo || q->rows[e][k] > q->hi)) return -1;     /* (the stored 8-bit value is what counts) */
    }
    return 0;
}

/* ps_data() (ISO/IEC 14496-3 8.6.4.6.2; what ff_ps_read_data yields, aacps.c:150-279): optional header, envelope borders, then
 * the enabled parameters envelope by envelope, the phase parameters inside an extension container. */
FE_BIG int ps_read(FePs *p, BR *host, int bits_left)
{
    BR bb = *host, *b = &bb;
    const int start_pos = host->pos;
    const int header = br_get1(b);
    if (header) {
        if ((p->enable_iid = br_get1(b))) {
            const int mode = br_get(b, 3);
            if (mode > 5) goto err;
            p->nr_iid_par = ps_bands_of_mode[mode]; p->iid_quant = mode > 2; p->nr_ipdopd_par = ps_phase_bands_of_mode[mode];
        }
        if ((p->enable_icc = br_get1(b))) {
            p->icc_mode = br_get(b, 3);
            if (p->icc_mode > 5) goto err;
            p->nr_icc_par = ps_bands_of_mode[p->icc_mode];
        }
        p->enable_ext = br_get1(b);
    }
    /* envelopes: 0 / 1 / 2 / 4 evenly spaced (fixed class) or 1..4 with explicit borders */
    p->frame_class = br_get1(b);
    p->num_env_old = p->num_env;
    {
        const int code = br_get(b, 2);
        p->num_env = p->frame_class ? code + 1 : (code == 3 ? 4 : code);
    }
    p->border[0] = -1;
    for (int e = 1; e <= p->num_env; e++)
        p->border[e] = p->frame_class ? (int)br_get(b, 5) : e * (32 / p->num_env) - 1;
    {
        const PsParam iid = { p->iid, p->nr_iid_par, p->iid_quant ? PS_IID_DF1 : PS_IID_DF0, p->iid_quant ? PS_IID_DT1 : PS_IID_DT0,
                              -(7 + 8 * p->iid_quant), 7 + 8 * p->iid_quant, 0 };
        const PsParam icc = { p->icc, p->nr_icc_par, PS_ICC_DF, PS_ICC_DT, 0, 7, 0 };
        if (p->enable_iid) { for (int e = 0; e < p->num_env; e++) if (ps_param_env(p, b, &iid, e)) goto err; }
        else fe_zero16(p->iid, (int)(sizeof(p->iid) / 2));
        if (p->enable_icc) { for (int e = 0; e < p->num_env; e++) if (ps_param_env(p, b, &icc, e)) goto err; }
        else fe_zero16(p->icc, (int)(sizeof(p->icc) / 2));
    }
    if (p->enable_ext) {
        /* extension container: a byte count, then elements of a 2-bit id each; id 0 carries the phase parameters (aacps.c:120-139) */
        int left = br_get(b, 4);
        if (left == 15) left += br_get(b, 8);
        left *= 8;
        while (left > 7) {
            const int id = br_get(b, 2), at = b->pos;
            if (id == 0) {
                const PsParam ipd = { p->ipd, p->nr_ipdopd_par, PS_IPD_DF, PS_IPD_DT, 0, 7, 1 };
                const PsParam opd = { p->opd, p->nr_ipdopd_par, PS_OPD_DF, PS_OPD_DT, 0, 7, 1 };
                if ((p->enable_ipdopd = br_get1(b)))
                    for (int e = 0; e < p->num_env; e++) { ps_param_env(p, b, &ipd, e); ps_param_env(p, b, &opd, e); }
                br_skip(b, 1);
            }
            left -= 2 + (b->pos - at);
        }
        if (left < 0) goto err;
        br_skip(b, left);
    }
    if (!p->num_env || p->border[p->num_env] < 31) {
        /* the last envelope must end at the frame's last slot: if it does not (or there is none), one more envelope that repeats
         * the last parameters is appended (aacps.c:235-252) */
        const int last = p->num_env ? p->num_env - 1 : p->num_env_old - 1;
        if (last >= 0 && last != p->num_env) {
            if (p->enable_iid) memcpy(p->iid[p->num_env], p->iid[last], 34);
            if (p->enable_icc) memcpy(p->icc[p->num_env], p->icc[last], 34);
            if (p->enable_ipdopd) { memcpy(p->ipd[p->num_env], p->ipd[last], 34); memcpy(p->opd[p->num_env], p->opd[last], 34); }
        }
        p->num_env++;
        p->border[p->num_env] = 31;
    }
    p->is34_old = p->is34;
    if (p->enable_iid || p->enable_icc)
        p->is34 = (p->enable_iid && p->nr_iid_par == 34) || (p->enable_icc && p->nr_icc_par == 34);
    if (!p->enable_ipdopd) fe_zero32(p->ipd, (int)((sizeof(p->ipd) + sizeof(p->opd)) / 4));     /* contiguous, 4-byte aligned */
    if (header) p->start = 1;
    {
        const int consumed = b->pos - start_pos;
        if (consumed <= bits_left) { host->pos += consumed; return consumed; }
    }
err:
    p->start = 0;
    host->pos = start_pos + bits_left;
    return bits_left;
}

/* ---------------------------------------------------------------- SBR payload */
/* sbr_data (read_sbr_data, aacsbr.c:982-1021), first step: the payload behind the header into the tape.  `s` is only read: the
 * band counts of its tables and the header's amplitude resolution.  Returns 0, or -1 for an invalid grid.
 * wide: the 32 lanes of a warp are in here together, each with the payload of another frame (valid = 0: a lane that has none and
 * only keeps the others company).  The envelope / noise values -- nearly all of the bits -- are then read in lockstep, one value
 * per lane and step, until no lane has any left; grids and flags before, harmonics after are read lane by lane. */
FE_INL int sbr_data_bits(const FeSbr *s, BR *b, int is_cpe, FeSbrTape *t, int wide, int valid)
{
    SbrRows plan[4];
    int n_mat = 0, bad = 0;
    if (valid) {
        FeSbrCh *c0 = &t->ch[0], *c1 = &t->ch[1];
        const int n_q = s->t.n_q;
        t->coupling = 0;
        if (!is_cpe) {
            if (br_get1(b)) br_skip(b, 4);
            if (sbr_grid_bits(c0, b, s->amp_res_header)) bad = 1;
            else {
                sbr_dtdf_bits(c0, b);
                for (int i = 0; i < n_q; i++) c0->invf[0][i] = br_get(b, 2);
                sbr_envelope_rows(&plan[0], s, c0, 0, 0); sbr_noise_rows(&plan[1], s, c0, 0, 0);
                n_mat = 2;
            }
        } else {
            if (br_get1(b)) br_skip(b, 8);
            if ((t->coupling = br_get1(b))) {
                if (sbr_grid_bits(c0, b, s->amp_res_header)) bad = 1;
                else {
                    /* the second channel shares the grid (copy_sbr_grid, aacsbr.c:751-766) */
                    c1->frame_class = c0->frame_class; c1->num_env = c0->num_env; c1->num_noise = c0->num_noise; c1->amp_res = c0->amp_res;
                    c1->e_a[1] = c0->e_a[1];
                    memcpy(c1->freq_res, c0->freq_res, sizeof(c0->freq_res));
                    memcpy(c1->t_env, c0->t_env, sizeof(c0->t_env));
                    memcpy(c1->t_q, c0->t_q, sizeof(c0->t_q));
                    sbr_dtdf_bits(c0, b); sbr_dtdf_bits(c1, b);
                    for (int i = 0; i < n_q; i++) c0->invf[0][i] = br_get(b, 2);
                    sbr_envelope_rows(&plan[0], s, c0, 0, 1); sbr_noise_rows(&plan[1], s, c0, 0, 1);
                    sbr_envelope_rows(&plan[2], s, c1, 1, 1); sbr_noise_rows(&plan[3], s, c1, 1, 1);
                    n_mat = 4;
                }
            } else {
                if (sbr_grid_bits(c0, b, s->amp_res_header) || sbr_grid_bits(c1, b, s->amp_res_header)) bad = 1;
                else {
                    sbr_dtdf_bits(c0, b); sbr_dtdf_bits(c1, b);
                    for (int i = 0; i < n_q; i++) c0->invf[0][i] = br_get(b, 2);
                    for (int i = 0; i < n_q; i++) c1->invf[0][i] = br_get(b, 2);
                    sbr_envelope_rows(&plan[0], s, c0, 0, 0); sbr_envelope_rows(&plan[1], s, c1, 1, 0);
                    sbr_noise_rows(&plan[2], s, c0, 0, 0); sbr_noise_rows(&plan[3], s, c1, 1, 0);
                    n_mat = 4;
                }
            }
        }
    }
    /* the matrices, one value per step: (mi, r, j) = matrix, row, band of the next value; acc = the running sum of a row
     * coded along frequency */
    int mi = 0, r = 0, j = 0, acc = 0;
    for (;;) {
        while (mi < n_mat) {                                       /* step over finished rows and matrices (and empty ones) */
            const SbrRows *q = &plan[mi];
            if (r >= q->rows) { mi++; r = 0; j = 0; continue; }
            if (j >= (q->res[r + 1] ? q->n_high : q->n_low)) { r++; j = 0; continue; }
            break;
        }
        const int more = mi < n_mat;
        if (!(wide ? FE_ANY(more) : more)) break;
        if (more) {
            const SbrRows *q = &plan[mi];
            const int time = q->along_time[r];
            int v;
            if (!time && j == 0) v = q->scale * (int)br_get(b, q->start_bits);
            else {
                const int tab = time ? q->tab_time : q->tab_freq;
                v = q->scale * (vlc_get(b, &TB->vsbr[tab]) - sbr_lav[tab]);
                if (!time) v = (int16_t)(acc + v);
            }
            acc = v;
            q->m[(r + 1) * q->stride + j] = (int16_t)v;
            j++;
        }
        if (wide) FE_SYNCWARP_ALL();
    }
    if (!valid || bad) return -bad;
    const int n_high = s->t.n[1];
    for (int c = 0; c <= is_cpe; c++)
        if ((t->ch[c].add_harmonic_flag = br_get1(b))) for (int i = 0; i < n_high; i++) t->ch[c].add_harmonic[i] = br_get1(b);
    t->ext_pos = t->ext_left = 0;
    if (br_get1(b)) {                                              /* bs_extended_data: noted, read by sbr_ext_apply */
        int left = br_get(b, 4);
        if (left == 15) left += br_get(b, 8);
        t->ext_left = left << 3;
        t->ext_pos = b->pos;
    }
    return 0;
}

/* second step for one channel: the tape's grid and rows into the running state (read_sbr_grid's and copy_sbr_grid's
 * previous-frame bookkeeping, read_sbr_invf's mode history, the time differences).  twin: the second channel of a coupled
 * pair takes ALL its inverse filtering modes from the first channel's state (aacsbr.c:1006-1007). */
FE_FN void sbr_chan_apply(const FeSbr *s, FeSbrCh *c, const FeSbrCh *t, const FeSbrCh *twin, int lane, int nl)
{
    const int num_env_old = c->num_env;
    const int res_old = c->freq_res[num_env_old], t_old = c->t_env[num_env_old], e_a_old = -(c->e_a[1] != num_env_old);
    const int n_q = s->t.n_q, n_high = s->t.n[1];
    FE_SYNCWARP();                                                 /* every lane has the old values */
    if (lane == 0) {
        c->frame_class = t->frame_class; c->num_env = t->num_env; c->num_noise = t->num_noise; c->amp_res = t->amp_res;
        c->add_harmonic_flag = t->add_harmonic_flag;
        c->e_a[0] = e_a_old; c->e_a[1] = t->e_a[1];
        c->t_env_num_env_old = t_old;
        c->freq_res[0] = (uint8_t)res_old;
    }
    const int n_env = t->num_env, n_noise = t->num_noise;            /* only what the payload holds: the tape's other entries are stale */
    for (int i = lane; i < 8; i += nl) {
        if (twin) {                                                /* copy_sbr_grid, aacsbr.c:751-766 */
            if (i >= 1 && i <= 6) c->freq_res[i] = twin->freq_res[i];
            c->t_env[i] = twin->t_env[i];
            if (i < 3) c->t_q[i] = twin->t_q[i];
        } else {
            if (i && i <= n_env) c->freq_res[i] = t->freq_res[i];
            if (i <= n_env) c->t_env[i] = t->t_env[i];
            if (i == 0 || i == n_noise) c->t_q[i] = t->t_q[i];
        }
        if (i < n_env) c->df_env[i] = t->df_env[i];
        if (i < n_noise) c->df_noise[i] = t->df_noise[i];
        if (i < 5) {
            c->invf[1][i] = c->invf[0][i];
            if (twin) c->invf[0][i] = twin->invf[0][i];
            else if (i < n_q) c->invf[0][i] = t->invf[0][i];
        }
    }
    if (t->add_harmonic_flag) for (int i = lane; i < n_high; i += nl) c->add_harmonic[i] = t->add_harmonic[i];
    if (!twin && n_noise > 1) {                                    /* the second noise floor's border, from the borders as they stand now */
        FE_SYNCWARP();
        if (lane == 0) c->t_q[1] = c->t_env[t->q_at < 8 ? t->q_at : 7];
    }
    uint8_t res[8];
    res[0] = (uint8_t)res_old;
    for (int e = 1; e < 8; e++) res[e] = t->freq_res[e];
    const uint8_t one_table[4] = { 0, 0, 0, 0 };
    sbr_rows_apply(&c->env[0][0], &t->env[0][0], 48, t->num_env, t->df_env, res, s->t.n[0], n_high, n_high & 1, lane, nl);
    sbr_rows_apply(&c->noise[0][0], &t->noise[0][0], 5, t->num_noise, t->df_noise, one_table, n_q, n_q, 0, lane, nl);
}

FE_FN void sbr_data_apply(FeSbr *s, const FeSbrTape *t, int is_cpe, int lane, int nl)
{
    if (lane == 0 && is_cpe) s->coupling = t->coupling;
    sbr_chan_apply(s, &s->ch[0], &t->ch[0], NULL, lane, nl);
    if (is_cpe) sbr_chan_apply(s, &s->ch[1], &t->ch[1], t->coupling ? &s->ch[0] : NULL, lane, nl);
}

/* the extension blocks behind the SBR data (aacsbr.c:1009-1019 region): Parametric Stereo is the one that is read */
FE_FN void sbr_ext_apply(FE *fe, BR *b, const FeSbrTape *t)
{
    int left = t->ext_left;
    if (left <= 7 || fe->ps != 1) return;                          /* nothing but bits to step over */
    b->pos = t->ext_pos;
    while (left > 7) {
        left -= 2;
        int id = br_get(b, 2);
        if (id == 2) left -= ps_read(&fe->psx, b, left);          /* EXTENSION_ID_PS */
        else { br_skip(b, left); left = 0; }
    }
}

/* ff_decode_sbr_extension (aacsbr.c:1044-1086), one pass: both steps for one payload */
FE_BIG void sbr_extension(FE *fe, FeSbr *s, BR *host, int crc, int cnt, int is_cpe)
{
    BR bb = *host, *b = &bb;
    FeSbrTape *t = &fe->tape;
    host->pos += cnt * 8 - 4;
    s->reset = 0;
    if (!fe->ext_sample_rate) fe->ext_sample_rate = 2 * fe->sample_rate;
    if (crc) br_skip(b, 10);
    s->t.kx[0] = s->t.kx[1];
    s->t.m[0] = s->t.m[1];
    if (br_get1(b)) { sbr_header_bits(b, t->hdr); sbr_header_apply(fe, s, t->hdr); }
    if (s->reset && sbr_tables(fe, s)) s->start = 0;               /* "pure upsampling mode" */
    if (!s->start) return;
    if (sbr_data_bits(s, b, is_cpe, t, 0, 1)) { s->start = 0; return; }
    sbr_data_apply(s, t, is_cpe, 0, 1);
    sbr_ext_apply(fe, b, t);
}

/* The first step alone, for a payload whose frame is not the stream's current one (the device front end: all frames of a chunk
 * at once).  `s` is the element as the chunk's earlier frames will have left it PROVIDED none of them changes how data is read:
 * a header that does makes the tape TAPE_STALE, and the caller sends the payload through sbr_extension when its turn comes. */
FE_FN void sbr_tape_bits(const FeSbr *s, BR *host, int crc, int is_cpe, FeSbrTape *t, int wide, int valid)
{
    BR bb, *b = &bb;
    int go = 0;                                                    /* this lane has data to read */
    if (valid) {
        bb = *host;
        if (crc) br_skip(b, 10);
        if ((t->header = br_get1(b))) sbr_header_bits(b, t->hdr);
        if (t->header && !sbr_header_same(s, t->hdr)) t->status = TAPE_STALE;
        else if (!s->start) t->status = TAPE_NONE;
        else go = 1;
    }
    if (!wide && !go) return;
    const int r = sbr_data_bits(s, b, is_cpe, t, wide, go);
    if (go) t->status = r ? TAPE_BAD : TAPE_DATA;
}

/* ... and the second step for such a tape, by nl cooperating lanes (scalars by lane 0) */
FE_FN void sbr_tape_apply(FE *fe, FeSbr *s, const uint8_t *pkt, int size, int is_cpe, const FeSbrTape *t, int lane, int nl)
{
    if (lane == 0) {
        s->reset = 0;
        if (!fe->ext_sample_rate) fe->ext_sample_rate = 2 * fe->sample_rate;
        s->t.kx[0] = s->t.kx[1];
        s->t.m[0] = s->t.m[1];
        if (t->header) sbr_header_apply(fe, s, t->hdr);           /* same tables: no reset */
    }
    FE_SYNCWARP();
    if (!s->start || t->status == TAPE_NONE) return;
    if (t->status == TAPE_BAD) { FE_SYNCWARP(); if (lane == 0) s->start = 0; return; }
    sbr_data_apply(s, t, is_cpe, lane, nl);
    if (lane == 0) {
        BR br;
        br_init(&br, pkt, size);
        sbr_ext_apply(fe, &br, t);
    }
}

/* ---------------------------------------------------------------- frame */
/* HeaacSbrElem: 140 bytes of header + tables, then ch[2] of 592 bytes each (env_q at +88); HeaacPsFrame: 20 bytes of
 * header, then the four parameter arrays back to back.  Everything is 4-byte aligned, so the copies go by words. */
/* The packers take (lane, nl): the word copies are strided over nl cooperating lanes, the scalar fields are written by
 * lane 0.  The host passes (0, 1); the device front end calls them with all 32 lanes of the stream's warp. */
FE_INL void fe_copy32_l(void *dst, const void *src, int nwords, int lane, int nl)
{
    uint32_t *d = (uint32_t *)dst;
    const uint32_t *q = (const uint32_t *)src;
    for (int i = lane; i < nwords; i += nl) d[i] = q[i];
}
FE_FN void pack_sbr_l(const FeSbr *s, HeaacSbrElem *o, int lane, int nl)
{
    fe_copy32_l(&o->n[0], &s->t.n[0], (140 - 8) / 4, lane, nl);      /* tables; the first 8 bytes are the flags below */
    if (lane == 0) {
        o->start = s->start; o->reset = s->reset; o->bs_coupling = s->coupling;
        o->bs_limiter_gains = s->limiter_gains; o->bs_interpol_freq = s->interpol_freq; o->bs_smoothing_mode = s->smoothing_mode;
        o->num_patches = s->t.num_patches; o->n_q = s->t.n_q;
    }
    for (int c = 0; c < 2; c++) {
        const FeSbrCh *d = &s->ch[c];
        HeaacSbrChan *oc = &o->ch[c];
        if (d->add_harmonic_flag) fe_copy32_l(oc->bs_add_harmonic, d->add_harmonic, 48 / 4, lane, nl);
        else for (int i = lane; i < 48 / 4; i += nl) ((uint32_t *)oc->bs_add_harmonic)[i] = 0;
        fe_copy32_l(oc->env_q, d->env[1], 5 * 48 * 2 / 4, lane, nl);
        /* the 40 bytes before bs_add_harmonic, one byte per lane and turn */
        for (int i = lane; i < 40; i += nl) {
            int v = 0;
            if (i >= 28) { if (i < 38) v = d->invf[(i - 28) / 5][(i - 28) % 5]; }
            else if (i >= 24) { if (i < 27) v = d->t_q[i - 24]; }
            else if (i >= 16) v = d->t_env[i - 16];
            else if (i >= 8) { if (i < 15) v = d->freq_res[i - 8]; }
            else v = i == 0 ? d->num_env : i == 1 ? d->num_noise : i == 2 ? d->amp_res : i == 3 ? d->add_harmonic_flag :
                     i == 4 ? d->e_a[0] : i == 5 ? d->e_a[1] : i == 6 ? d->t_env_num_env_old : 0;
            ((uint8_t *)oc)[i] = (uint8_t)v;
        }
        for (int i = lane; i < 12; i += nl) ((int16_t *)oc->noise_q)[i] = i < 10 ? d->noise[1 + i / 5][i % 5] : 0;   /* + reserved2 */
    }
}
FE_FN void pack_ps_l(const FePs *p, HeaacPsFrame *o, int lane, int nl)
{
    fe_copy32_l(o->iid_par, p->iid, 4 * 5 * 34 / 4, lane, nl);        /* iid, icc, ipd, opd are contiguous in both structures */
    if (lane == 0) {
        fe_zero32(o, 20 / 4);
        o->start = p->start; o->num_env = p->num_env; o->num_env_old = p->num_env_old; o->iid_quant = p->iid_quant;
        o->nr_iid_par = p->nr_iid_par; o->nr_icc_par = p->nr_icc_par; o->nr_ipdopd_par = p->nr_ipdopd_par; o->icc_mode = p->icc_mode;
        o->enable_ipdopd = p->enable_ipdopd; o->is34bands = p->is34; o->is34bands_old = p->is34_old;
        for (int i = 0; i < 6; i++) o->border_position[i] = (int8_t)p->border[i];
    }
}
/* all SBR elements + PS of the frame the stream's state currently describes */
FE_FN void fe_pack_l(const FE *fe, HeaacSbrElem *sbr, HeaacPsFrame *ps, int lane, int nl)
{
    if (sbr) for (int e = 0, se = 0; e < fe->n_elems; e++) if (fe->elem_type[e] != HEAAC_TYPE_LFE) pack_sbr_l(&fe->sbr_el[se], &sbr[se], lane, nl), se++;
    if (ps) pack_ps_l(&fe->psx, ps, lane, nl);
}

/* An SBR fill payload met after element `prev_elem` (decode_extension_payload, aacdec.c:1650-1690): implicit
 * signalling, then ff_decode_sbr_extension.  `b` stands right after the 4-bit extension type; leaves it after
 * the payload. */
FE_FN void fe_sbr_payload(FE *fe, BR *b, int crc, int cnt, int prev_elem, int prev_type)
{
    if (prev_elem < 0 || fe->sbr == 0 || (fe->sbr == -1 && fe->locked)) {
        br_skip(b, 8 * cnt - 4);
        return;
    }
    if (fe->ps == -1 && !fe->locked && fe->ncc == 1) { fe->sbr = 1; fe->ps = 1; }
    else fe->sbr = 1;
    if (prev_type == HEAAC_TYPE_LFE) {     /* read_sbr_data rejects the element type (aacsbr.c:997-1002): the LFE stays without SBR data */
        br_skip(b, 8 * cnt - 4);
        return;
    }
    int se = 0;
    for (int i = 0; i < prev_elem; i++) se += fe->elem_type[i] != HEAAC_TYPE_LFE;
    sbr_extension(fe, &fe->sbr_el[se], b, crc, cnt, prev_type == 1);
}

/* One packet (an ADTS frame or a bare raw_data_block) -> the frame's C-ABI structures.
 * coeffs [ncc][1024]; ics/tns [ncc]; sbr [nse] (may be NULL for LC); ps (may be NULL).
 * Returns bytes consumed (>= 0) or a negative HEAAC_E_* code (the reference returns -1). */
/* pa: compile-time constant at both call sites -- 1 = the stateless first phase of the two-phase parse (SBR payloads are only
 * noted: the whole SBR / PS parser drops out of that instance, and of the kernel built on it), 0 = the one-pass parser */
FE_INL int fe_parse_frame_x(FE *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                            HeaacSbrElem *sbr, HeaacPsFrame *ps, const int pa)
{
    if (!fe || !pkt || size <= 0 || !coeffs || !ics) return HEAAC_E_INVAL;
    BR br, *b = &br;
    br_init(b, pkt, size);
    if (br_show(b, 12) == 0xfff) {                                 /* ADTS header, aac_parser.c:29-67 */
        br_skip(b, 12);
        br_skip(b, 1); br_skip(b, 2);
        int crc_absent = br_get1(b);
        br_skip(b, 2);                                            /* profile */
        int sr = br_get(b, 4);
        if (!sample_rates[sr]) return HEAAC_E_INVAL;
        br_skip(b, 1);
        int ch = br_get(b, 3);
        br_skip(b, 4);
        int fsize = br_get(b, 13);
        if (fsize < 7) return HEAAC_E_INVAL;
        br_skip(b, 11);
        int rdb = br_get(b, 2);
        if (rdb) return HEAAC_E_UNSUPPORTED;
        if (!crc_absent) br_skip(b, 16);
        if (ch && (fe->locked || fe->layout_fixed)) {
            /* Output slots (and the batch's staging buffers) are sized for the layout the stream was opened with: a
             * header announcing another channel configuration or core rate is rejected before anything is written.
             * (The reference reconfigures its output instead, aacdec.c:1994-2006; a batch cannot.) */
            if (fe->chan_config && (ch != fe->chan_config || sr != fe->sr_index)) return HEAAC_E_INVAL;
        } else if (ch) {
            int r = set_layout(fe, ch); if (r) return r; fe->sample_rate = sample_rates[sr]; fe->sr_index = sr;
        }
#ifndef FE_DEVICE
        else if (!fe->locked && !fe->layout_fixed) {              /* channel_configuration 0: the layout follows as a PCE */
            fe->sample_rate = sample_rates[sr]; fe->sr_index = sr;
        }
#endif
    }
#ifdef FE_DEVICE
    if (!fe->n_elems) return HEAAC_E_INVAL;
#endif
    int seen[4] = { 0, 0, 0, 0 };
    int prev_elem = -1, prev_type = 7, type;
    while ((type = br_get(b, 3)) != 7) {
        int id = br_get(b, 4), e = -1, r = 0;
#ifndef FE_DEVICE
        if (fe->by_tag && type <= 3) {                             /* the element with this tag (che[type][id], aacdec.c:118-128) */
            for (int i = 0; i < fe->n_elems; i++)
                if (fe->elem_type[i] == type && fe->elem_tag[i] == id) { e = i; break; }
            if (e < 0) return HEAAC_E_INVAL;
        } else
#endif
        if (type == 0 || type == 1 || type == 3) {                 /* k-th element of this type in the layout */
            int k = 0;
            for (int i = 0; i < fe->n_elems; i++)
                if (fe->elem_type[i] == type && k++ == seen[type]) { e = i; break; }
            if (e < 0) return HEAAC_E_INVAL;
            seen[type]++;
        }
        switch (type) {
        case 0: case 3:
            r = decode_ics(fe, &fe->ch[fe->elem_ch0[e]], b, coeffs + 1024 * fe->elem_ch0[e], 0);
            if (pa) fe->todo->chan[fe->elem_ch0[e]] = 1;
            break;
        case 1:
            r = decode_cpe(fe, &fe->ch[fe->elem_ch0[e]], &fe->ch[fe->elem_ch0[e] + 1], b,
                           coeffs + 1024 * fe->elem_ch0[e], coeffs + 1024 * (fe->elem_ch0[e] + 1));
            if (pa) { fe->todo->chan[fe->elem_ch0[e]] = 1; fe->todo->chan[fe->elem_ch0[e] + 1] = fe->last_common ? 2 : 1; }
            break;
        case 4: {                                                  /* DSE, aacdec.c:1190-1200 region */
            int align = br_get1(b), cnt = br_get(b, 8);
            if (cnt == 255) cnt += br_get(b, 8);
            if (align) b->pos = (b->pos + 7) & ~7;
            br_skip(b, 8 * cnt);
            break; }
        case 6: {                                                  /* FIL */
            int cnt = id;
            if (cnt == 15) cnt += br_get(b, 8) - 1;
            if (br_left(b) < 8 * cnt) return HEAAC_E_INVAL;
            while (cnt > 0) {                                      /* decode_extension_payload, aacdec.c:1650-1690 */
                int ext = br_get(b, 4), res = cnt;
                if (ext == 0xd || ext == 0xe) {
                    if (pa) {                                      /* stateless phase: note it for phase B, skip the bits */
                        FeTodo *td = fe->todo;
                        if (td->n_sbr < FE_MAX_TODO) {
                            td->sbr[td->n_sbr].pos = b->pos; td->sbr[td->n_sbr].cnt = cnt; td->sbr[td->n_sbr].crc = ext == 0xe;
                            td->sbr[td->n_sbr].prev_elem = prev_elem; td->sbr[td->n_sbr].prev_type = prev_type;
                            td->n_sbr++;
                        } else td->serial = 1;
                        br_skip(b, 8 * cnt - 4);
                    } else
                        fe_sbr_payload(fe, b, ext == 0xe, cnt, prev_elem, prev_type);
                } else br_skip(b, 8 * cnt - 4);
                cnt -= res;
            }
            break; }
#ifndef FE_DEVICE
        case 2:
            if (e < 0) return HEAAC_E_UNSUPPORTED;                 /* a coupling channel needs a layout that lists it (a PCE) */
            r = decode_cce(fe, e, b, coeffs);
            if (pa) fe->todo->chan[fe->elem_ch0[e]] = 1;
            break;
        case 5:                                                    /* in band: configures a stream nothing else has configured */
            r = fe_pce(fe, b, !fe->n_elems && !fe->locked && !fe->layout_fixed);
            break;
#else
        case 5: r = fe_pce_read(b, NULL); break;                   /* the stream is configured: stepped over */
#endif
        default: return HEAAC_E_UNSUPPORTED;                       /* CCE outside a PCE layout (device build: any CCE) */
        }
        if (r) return r;
        if (type <= 3) { prev_elem = e; prev_type = type; } else if (type != 6) { prev_elem = -1; prev_type = type; }
        if (br_left(b) < 3) return HEAAC_E_INVAL;
    }
#ifndef FE_DEVICE
    if (!fe->n_elems) return HEAAC_E_INVAL;
#endif
    fe->locked = 1;
    for (int c = 0; c < fe->ncc; c++) {
        const FeChan *fc = &fe->ch[c];
        fe_zero32(&ics[c], 2);
        ics[c].window_sequence = fc->ws[0]; ics[c].window_shape = fc->shape[0]; ics[c].max_sfb = fc->max_sfb;
        ics[c].tns_present = fc->tns_present;
        ics[c].prev_valid = 1; ics[c].prev_window_sequence = fc->ws[1]; ics[c].prev_window_shape = fc->shape[1];
        if (tns) {                                                 /* read by the back end only when tns_present */
            if (fc->tns_present) fe_copy32(&tns[c], &fc->tns, (int)(sizeof(HeaacTns) / 4));
            else fe_zero32(tns[c].n_filt, 2);
        }
    }
    if (!pa) fe_pack_l(fe, sbr, ps, 0, 1);
    int consumed = (b->pos + 7) >> 3;
    return consumed > size ? size : consumed;
}

FE_BIG int fe_parse_frame(FE *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                         HeaacSbrElem *sbr, HeaacPsFrame *ps)
{
    return fe_parse_frame_x(fe, pkt, size, coeffs, ics, tns, sbr, ps, 0);
}



/* ---------------------------------------------------------------- two-phase parse
 * The device front end spreads the frames of a chunk over the lanes of a warp.  Entropy decoding of the core
 * elements -- nearly all of the work -- needs nothing from the previous frame, so phase A runs fe_parse_frame on a
 * frame-local scratch state (head + channel blocks only) and merely notes where the SBR payloads are.  Phase B then
 * walks the frames in order on the stream's persistent state: previous-window bookkeeping (ics_info and the
 * common_window copy of decode_cpe, aacdec.c:645-710, 1464-1470), implicit signalling, the SBR / PS payloads (delta
 * coded against the previous frame) and the packing.  Frames that touch sequential state inside the core elements
 * (PNS draws from the stream's random generator) or that failed are simply redone by the one-pass parser. */
FE_FN size_t fe_scratch_bytes(int ncc) { return offsetof(FE, ch) + (size_t)ncc * sizeof(FeChan); }

FE_FN void fe_phase_a(const FE *stream, FE *scratch, FeTodo *todo, const uint8_t *pkt, int size, float *coeffs,
                      HeaacIcs *ics, HeaacTns *tns)
{
    fe_copy32(scratch, stream, (int)(offsetof(FE, ch) / 4));
    scratch->phase_a = 1; scratch->saw_pns = 0; scratch->todo = todo;
    todo->serial = 0; todo->n_sbr = 0;
    for (int c = 0; c < FE_MAX_CH; c++) todo->chan[c] = 0;
#ifndef FE_DEVICE
    scratch->by_tag = 0;                                           /* (beyond the head that was copied) */
    if (stream->by_tag) { todo->serial = 1; todo->result = 0; return; }    /* program config element layouts: one-pass parser only */
#endif
    todo->result = fe_parse_frame_x(scratch, pkt, size, coeffs, ics, tns, NULL, NULL, 1);
    if (todo->result < 0 || scratch->saw_pns || scratch->ncc != stream->ncc) todo->serial = 1;
    else
        for (int c = 0; c < stream->ncc; c++) if (!todo->chan[c]) fe_zero(coeffs + 1024 * c, 1024);
}

/* previous-window bookkeeping of the frame phase A has parsed (ics_info and the common_window copy of decode_cpe) */
FE_FN void fe_windows_b(FE *fe, const FeTodo *todo, HeaacIcs *ics)
{
    for (int c = 0; c < fe->ncc; c++) {
        FeChan *st = &fe->ch[c];
        if (todo->chan[c] == 1) {
            st->ws[1] = st->ws[0]; st->ws[0] = ics[c].window_sequence;
            st->shape[1] = st->shape[0]; st->shape[0] = ics[c].window_shape;
        } else if (todo->chan[c] == 2) {
            const int keep = st->shape[0];
            st->ws[0] = st[-1].ws[0]; st->ws[1] = st[-1].ws[1]; st->shape[0] = st[-1].shape[0]; st->shape[1] = keep;
        } else {
            fe_zero32(&ics[c], 2);
            ics[c].window_sequence = (uint8_t)st->ws[0]; ics[c].window_shape = (uint8_t)st->shape[0];
        }
        ics[c].prev_valid = 1; ics[c].prev_window_sequence = (uint8_t)st->ws[1]; ics[c].prev_window_shape = (uint8_t)st->shape[1];
    }
}

FE_FN int fe_phase_b(FE *fe, const FeTodo *todo, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                     HeaacSbrElem *sbr, HeaacPsFrame *ps)
{
    if (todo->serial) return fe_parse_frame(fe, pkt, size, coeffs, ics, tns, sbr, ps);
    fe_windows_b(fe, todo, ics);
    if (todo->n_sbr) {
        BR br;
        br_init(&br, pkt, size);
        for (int k = 0; k < todo->n_sbr; k++) {
            br.pos = todo->sbr[k].pos;
            fe_sbr_payload(fe, &br, todo->sbr[k].crc, todo->sbr[k].cnt, todo->sbr[k].prev_elem, todo->sbr[k].prev_type);
        }
    }
    fe->locked = 1;
    fe_pack_l(fe, sbr, ps, 0, 1);
    return todo->result;
}

/* ---------------------------------------------------------------- phase B for a whole chunk of one stream
 * The SBR payloads of a chunk's frames are read in two steps (see "SBR payload, in two steps"): the first step for ALL frames
 * that are still to do at once -- one frame per lane -- against the element states as they stand; then the frames are finished
 * in order: window bookkeeping, the second step of their payloads (the lanes walk the rows together), Parametric Stereo,
 * packing.  A frame that has to go through the one-pass parser (fe_phase_b: flagged by phase A, or with a header that changes
 * how SBR data is read) is done by lane 0 alone; if it changed what the tapes of the later frames were read with, those are
 * read again.  Frame f's packet and outputs sit at index f * stride of the arrays (device: stride = streams of the batch). */
typedef struct FeChunkIO {
    int n_frames, ncc, nse;
    size_t stride;
    const FeTodo *todo;                         /* [n_frames], this stream's */
    FeSbrTape *tape; int tapes_per_frame;       /* [n_frames][tapes_per_frame], this stream's */
    const uint8_t *blob; const uint32_t *off; const int *size;
    float *coeffs; HeaacIcs *ics; HeaacTns *tns; HeaacSbrElem *sbr; HeaacPsFrame *ps; int *consumed;
    int *stats;                                 /* optional (host): [0] first-step passes, [1] frames through the one-pass parser, [2] of those: stale tapes */
} FeChunkIO;

#ifdef FE_DEVICE
#define FE_BCAST(v) __shfl_sync(0xffffffffu, (v), 0)
#else
#define FE_BCAST(v) (v)
#endif

/* everything of an element's state the first step reads */
FE_FN uint64_t fe_sbr_reads(const FeSbr *s)
{
    return (uint64_t)(s->start & 1) | (uint64_t)(s->amp_res_header & 1) << 1 | (uint64_t)(s->start_freq & 15) << 2 |
           (uint64_t)(s->stop_freq & 15) << 6 | (uint64_t)(s->xover & 7) << 10 | (uint64_t)(s->freq_scale & 3) << 13 |
           (uint64_t)(s->alter_scale & 1) << 15 | (uint64_t)(s->noise_bands & 3) << 16 |
           (uint64_t)(s->t.n[0] & 255) << 18 | (uint64_t)(s->t.n[1] & 255) << 26 | (uint64_t)(s->t.n_q & 255) << 34;
}
FE_FN int fe_sbr_elem_of(const FE *fe, int elem)
{
    int se = 0;
    for (int i = 0; i < elem; i++) se += fe->elem_type[i] != HEAAC_TYPE_LFE;
    return se;
}

FE_FN void fe_chunk_b(FE *fe, const FeChunkIO *io, int lane, int nl)
{
    const int n = io->n_frames, P = io->tapes_per_frame, ncc = io->ncc, nse = io->nse;
    int k = 0;
    while (k < n) {
        const int taped = fe->sbr == 1;
        if (io->stats && taped) io->stats[0]++;
        if (taped)
            for (int f0 = k; f0 < n; f0 += nl) {                   /* first step, one frame per lane; the lanes stay together */
                const int f = f0 + lane;
                const FeTodo *td = io->todo + (f < n ? f : k);
                const int use = f < n && !td->serial && td->n_sbr <= P;
                BR br;
                if (use) br_init(&br, io->blob + io->off[(size_t)f * io->stride], io->size[(size_t)f * io->stride]);
                for (int p = 0; p < P; p++) {
                    FeSbrTape *t = io->tape + (size_t)(f < n ? f : k) * P + p;
                    int valid = use && p < td->n_sbr, se = 0, pt = 0, crc = 0;
                    if (valid) {
                        const int pe = td->sbr[p].prev_elem;
                        pt = td->sbr[p].prev_type; crc = td->sbr[p].crc;
                        if (pe < 0 || pt == HEAAC_TYPE_LFE) { t->status = TAPE_SKIP; valid = 0; }
                        else { br.pos = td->sbr[p].pos; se = fe_sbr_elem_of(fe, pe); }
                    }
                    if (valid || nl > 1) sbr_tape_bits(&fe->sbr_el[se], &br, crc, pt == 1, t, nl > 1, valid);
                }
            }
        FE_SYNCWARP();
        int again = 0;
        for (; k < n && !again; k++) {                             /* the frames in order */
            const FeTodo *td = io->todo + k;
            const size_t idx = (size_t)k * io->stride;
            const uint8_t *pkt = io->blob + io->off[idx];
            const int size = io->size[idx];
            float *co = io->coeffs + idx * ncc * 1024;
            HeaacIcs *ic = io->ics + idx * ncc;
            HeaacTns *tn = io->tns + idx * ncc;
            HeaacSbrElem *sb = io->sbr ? io->sbr + idx * nse : NULL;
            HeaacPsFrame *pp = io->ps ? io->ps + idx : NULL;
            int serial = !taped || td->serial || td->n_sbr > P, r;
            if (!serial)
                for (int p = 0; p < td->n_sbr; p++) serial |= io->tape[(size_t)k * P + p].status == TAPE_STALE;
            if (io->stats && serial) { io->stats[1]++; io->stats[2] += taped && !td->serial && td->n_sbr <= P; }
            if (serial) {
                /* can this frame change what the first step reads?  only while SBR payloads are read at all */
                const int live = k + 1 < n && (fe->sbr == 1 || (fe->sbr == -1 && !fe->locked));
                const int sbr_before = fe->sbr;
                uint64_t before[HEAAC_MAX_ELEMS];
                if (live) for (int e = 0; e < nse; e++) before[e] = fe_sbr_reads(&fe->sbr_el[e]);
                FE_SYNCWARP();
                r = 0;
                if (lane == 0) {
                    r = fe_phase_b(fe, td, pkt, size, co, ic, tn, NULL, NULL);
                    if (r >= 0 && fe->ncc != ncc) r = HEAAC_E_INVAL;   /* layout differs from the batch's */
                    io->consumed[idx] = r;
                }
                FE_SYNCWARP();
                r = FE_BCAST(r);
                if (live) {
                    again = fe->sbr != sbr_before;
                    for (int e = 0; e < nse; e++) again |= before[e] != fe_sbr_reads(&fe->sbr_el[e]);
                }
            } else {
                if (lane == 0) fe_windows_b(fe, td, ic);
                for (int p = 0; p < td->n_sbr; p++) {
                    const FeSbrTape *t = io->tape + (size_t)k * P + p;
                    if (t->status == TAPE_SKIP) continue;
                    if (lane == 0 && fe->ps == -1 && !fe->locked && fe->ncc == 1) fe->ps = 1;   /* implicit signalling, aacdec.c:1659-1676 */
                    sbr_tape_apply(fe, &fe->sbr_el[fe_sbr_elem_of(fe, td->sbr[p].prev_elem)], pkt, size, td->sbr[p].prev_type == 1,
                                   t, lane, nl);
                    FE_SYNCWARP();
                }
                r = fe->ncc != ncc ? HEAAC_E_INVAL : td->result;
                if (lane == 0) { fe->locked = 1; io->consumed[idx] = r; }
            }
            FE_SYNCWARP();                                         /* lane 0's state writes before the other lanes' reads */
            if (r >= 0) {
                fe_pack_l(fe, sb, pp, lane, nl);
            } else {                                               /* silent frame; the batch goes on */
                for (int i = lane; i < ncc * 1024; i += nl) co[i] = 0.0f;
                for (int i = lane; i < (int)(sizeof(HeaacIcs) * ncc / 4); i += nl) ((uint32_t *)ic)[i] = 0;
                for (int i = lane; i < (int)(sizeof(HeaacTns) * ncc / 4); i += nl) ((uint32_t *)tn)[i] = 0;
                if (sb) for (int i = lane; i < (int)(sizeof(HeaacSbrElem) * nse / 4); i += nl) ((uint32_t *)sb)[i] = 0;
                if (pp) for (int i = lane; i < (int)(sizeof(HeaacPsFrame) / 4); i += nl) ((uint32_t *)pp)[i] = 0;
            }
            FE_SYNCWARP();                                         /* the pack reads state the next frame overwrites */
        }
    }
}

#endif
