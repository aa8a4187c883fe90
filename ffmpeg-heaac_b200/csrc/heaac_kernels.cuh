/*
 * heaac_kernels.cuh -- CTA-level routines of the HE-AAC synthesis path (sm_100a).
 *
 * Each routine is executed by one CTA that owns one channel (or one PS stream) for one
 * frame.  Work inside a routine is expressed as strided loops over `tid/nt` separated by
 * HEAAC_SYNC(), with every intermediate of the frame staged in shared memory; HBM is touched
 * only for the frame's inputs, the per-channel state block and the outputs.
 *
 * Reference functions restated here (all from /root/reference/libavcodec):
 *   core_channel   : apply_tns aacdec.c:1698-1736 (+ lpc.h:61-103), imdct_and_windowing
 *                    aacdec.c:1741-1806, ff_imdct_half_c mdct.c:124-159,
 *                    ff_vector_fmul_window_c dsputil.c:3832-3845
 *   sbr_channel    : sbr_dequant aacsbr.c:1089-1128, sbr_qmf_analysis 1136-1169, sbr_lf_gen
 *                    1337-1357, sbr_hf_inverse_filter 1232-1313, sbr_chirp 1316-1334, sbr_hf_gen
 *                    1360-1409, sbr_mapping 1451-1496, sbr_env_estimate 1499-1546, sbr_gain_calc
 *                    1552-1605, sbr_hf_assemble 1608-1714, sbr_x_gen 1412-1446
 *   ps_stream      : ff_ps_apply aacps.c:973-992 and everything it calls (283-971)
 *   synth_channel  : sbr_qmf_synthesis aacsbr.c:1175-1230, float_to_int16 dsputil.c:3972-4001
 */
#ifndef HEAAC_KERNELS_CUH
#define HEAAC_KERNELS_CUH

#include "heaac_dev.cuh"
#include "heaac_regfft.cuh"

#ifdef HEAAC_HOST_EMU
#define HEAAC_CONST_TABLE static const
#else
#define HEAAC_CONST_TABLE __device__ const
#endif
/* double literals narrowed to float exactly as the reference's `static const float` tables are */
HEAAC_CONST_TABLE float heaac_h_smooth[5] = { (float)0.33333333333333, (float)0.30150283239582, (float)0.21816949906249,
                                              (float)0.11516383427084, (float)0.03183050093751 };   /* aacsbr.c:1616-1622 */
HEAAC_CONST_TABLE float heaac_limgain[4] = { (float)0.70795, (float)1.0, (float)1.41254, (float)10000000000.0 };  /* aacsbr.c:1557 */

#define HEAAC_MIN(a, b) ((a) > (b) ? (b) : (a))     /* FFMIN semantics incl. NaN behaviour */
#define HEAAC_MAX(a, b) ((a) > (b) ? (a) : (b))     /* FFMAX */

/* ------------------------------------------------------------------------------------
 * float -> int16 (dsputil.c:3972-3982 for the C "bias" convention; cvtps2dq-style RNE +
 * saturation for the s16-unit convention; both give the same s16)
 * ---------------------------------------------------------------------------------- */
HEAAC_DEV int16_t heaac_to_s16(float x, int bias_mode)
{
    if (bias_mode) {
        int32_t bits;
#ifdef HEAAC_HOST_EMU
        memcpy(&bits, &x, 4);
#else
        bits = __float_as_int(x);
#endif
        long long tmp = bits;
        if (tmp & 0xf0000) tmp = (0x43c0ffffLL - tmp) >> 31;
        return (int16_t)(tmp - 0x8000);
    }
#ifdef HEAAC_HOST_EMU
    float r = heaac_rint(x);
    r = r < -32768.0f ? -32768.0f : (r > 32767.0f ? 32767.0f : r);
    return (int16_t)(int)r;
#else
    short r;                                   /* round-to-nearest-even + saturate in one instruction */
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return r;
#endif
}

/* ------------------------------------------------------------------------------------
 * Batched in-place complex FFT in shared memory, X[m] = sum_k x[k] e^{+2 pi i km/n}
 * (the un-normalised inverse convention of ff_fft_calc with inverse = 1, fft.c:81-176).
 * Input must already be in bit-reversed order.  `nbatch` transforms of 2^log2n points lie
 * back to back.  Radix-2, one butterfly per work item, log2n barriers.
 * ---------------------------------------------------------------------------------- */
/* element i of the FFT arrays lives at i + (i >> 3): with 8 consecutive points per work item in the first pass and
 * strides of 8 and 64 in the later ones, the skew keeps every pass (nearly) free of bank conflicts */
#define HEAAC_ZP(i) ((i) + ((i) >> 3))
#define HEAAC_Z_FLOATS (512 + 64)

/* One radix-2 butterfly of the network: (a, b) <- (a + w b, a - w b), w = exp(2 pi i j / 2^s) from the 512-point table. */
#define HEAAC_BFLY(ar, ai, br, bi, wr, wi) do { \
        const float tr_ = (br) * (wr) - (bi) * (wi), ti_ = (br) * (wi) + (bi) * (wr); \
        (br) = (ar) - tr_; (bi) = (ai) - ti_; (ar) = (ar) + tr_; (ai) = (ai) + ti_; } while (0)

template <int NT>
HEAAC_DEV void heaac_fft_smem(float *re, float *im, int log2n, int nbatch,
                              const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    /* The same radix-2 decimation-in-time network as before (same butterflies, same twiddles, same operation order
     * inside each butterfly), but three stages at a time: a work item loads the 8 points that stages 3p+1..3p+3
     * connect, runs their 12 butterflies in registers and stores them back -- 3 passes and barriers for 512 points
     * (2 for 64) instead of 9 (6), and a third of the shared-memory traffic. */
    constexpr int nt = NT;
    const int n = 1 << log2n;
    const int items = nbatch * (n >> 3);
    const int lead = log2n % 3;                             /* sizes that are not a power of 8: plain stages first */
    for (int s = 1; s <= lead; s++) {
        const int half = 1 << (s - 1), hn = n >> 1;
        HEAAC_NOUNROLL
        for (int w = tid; w < nbatch * hn; w += nt) {
            const int bt = w / hn, r = w - bt * hn;
            const int j = r & (half - 1), grp = r >> (s - 1);
            const int i0 = (bt << log2n) + (grp << s) + j, i1 = i0 + half;
            const float wr = T->fft_wre[j * (512 >> s)], wi = T->fft_wim[j * (512 >> s)];
            float ar = re[HEAAC_ZP(i0)], ai = im[HEAAC_ZP(i0)], br = re[HEAAC_ZP(i1)], bi = im[HEAAC_ZP(i1)];
            HEAAC_BFLY(ar, ai, br, bi, wr, wi);
            re[HEAAC_ZP(i0)] = ar; im[HEAAC_ZP(i0)] = ai; re[HEAAC_ZP(i1)] = br; im[HEAAC_ZP(i1)] = bi;
        }
        HEAAC_SYNC();
    }
    for (int p3 = lead; p3 < log2n; p3 += 3) {
        const int h0 = 1 << p3;                             /* butterfly span of the first stage of this pass */
        HEAAC_NOUNROLL
        for (int w = tid; w < items; w += nt) {
            const int bt = w / (n >> 3), r = w - bt * (n >> 3);
            const int j = r & (h0 - 1), blk = r >> p3;
            const int base = (bt << log2n) + (blk << (p3 + 3)) + j;
            float xr[8], xi[8];
            HEAAC_UNROLL
            for (int m = 0; m < 8; m++) { const int e = HEAAC_ZP(base + m * h0); xr[m] = re[e]; xi[m] = im[e]; }
            {                                               /* stage 3p+1: pairs (m, m+1), twiddle index j */
                const int tw = j * (512 >> (p3 + 1));
                const float wr = T->fft_wre[tw], wi = T->fft_wim[tw];
                HEAAC_UNROLL
                for (int m = 0; m < 8; m += 2) HEAAC_BFLY(xr[m], xi[m], xr[m + 1], xi[m + 1], wr, wi);
            }
            {                                               /* stage 3p+2: pairs (m, m+2), twiddle index j + (m & 1) h0 */
                const int ts = 512 >> (p3 + 2);
                const float w0r = T->fft_wre[j * ts], w0i = T->fft_wim[j * ts];
                const float w1r = T->fft_wre[(j + h0) * ts], w1i = T->fft_wim[(j + h0) * ts];
                HEAAC_BFLY(xr[0], xi[0], xr[2], xi[2], w0r, w0i);
                HEAAC_BFLY(xr[1], xi[1], xr[3], xi[3], w1r, w1i);
                HEAAC_BFLY(xr[4], xi[4], xr[6], xi[6], w0r, w0i);
                HEAAC_BFLY(xr[5], xi[5], xr[7], xi[7], w1r, w1i);
            }
            {                                               /* stage 3p+3: pairs (m, m+4), twiddle index j + m h0 */
                const int ts = 512 >> (p3 + 3);
                HEAAC_UNROLL
                for (int m = 0; m < 4; m++) {
                    const float wr = T->fft_wre[(j + m * h0) * ts], wi = T->fft_wim[(j + m * h0) * ts];
                    HEAAC_BFLY(xr[m], xi[m], xr[m + 4], xi[m + 4], wr, wi);
                }
            }
            HEAAC_UNROLL
            for (int m = 0; m < 8; m++) { const int e = HEAAC_ZP(base + m * h0); re[e] = xr[m]; im[e] = xi[m]; }
        }
        HEAAC_SYNC();
    }
}

HEAAC_DEV int heaac_bitrev(int k, int bits)
{
#ifdef HEAAC_HOST_EMU
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((k >> b) & 1) << (bits - 1 - b);
    return r;
#else
    return (int)(__brev((unsigned)k) >> (32 - bits));
#endif
}

/* Batched half-IMDCT of size N = 4 * n4 (mdct.c:124-159): `in` holds nbatch * N/2 inputs,
 * `out` receives nbatch * N/2 outputs; in and out may alias (input is consumed by the
 * pre-rotation before anything is written).  zr/zi: nbatch * n4 floats of scratch each. */
template <int NT>
HEAAC_DEV void heaac_imdct_half_smem(float *out, const float *in, float *zr, float *zi,
                                     int log2n4, int nbatch,
                                     const float *HEAAC_RESTRICT tcos, const float *HEAAC_RESTRICT tsin,
                                     const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    constexpr int nt = NT;
    const int n4 = 1 << log2n4, n2 = n4 << 1, n8 = n4 >> 1;
    HEAAC_NOUNROLL
    for (int w = tid; w < nbatch * n4; w += nt) {
        int b = w >> log2n4, k = w & (n4 - 1);
        const float *x = in + b * n2;
        float a = x[n2 - 1 - 2 * k], c = x[2 * k];
        int rk = heaac_bitrev(k, log2n4);
        zr[HEAAC_ZP(b * n4 + rk)] = a * tcos[k] - c * tsin[k];
        zi[HEAAC_ZP(b * n4 + rk)] = a * tsin[k] + c * tcos[k];
    }
    HEAAC_SYNC();
    heaac_fft_smem<NT>(zr, zi, log2n4, nbatch, T, tid);
    heaac_cp_async_wait<0>();        /* what the caller fetched ahead for the windowing is visible behind the barrier below */
    HEAAC_NOUNROLL
    for (int w = tid; w < nbatch * n8; w += nt) {
        int b = w / n8, k = w - b * n8;
        int p = n8 - k - 1, q = n8 + k;
        float *o = out + b * n2;
        const float zrp = zr[HEAAC_ZP(b * n4 + p)], zip = zi[HEAAC_ZP(b * n4 + p)];
        const float zrq = zr[HEAAC_ZP(b * n4 + q)], ziq = zi[HEAAC_ZP(b * n4 + q)];
        float r0 = zip * tsin[p] - zrp * tcos[p];
        float i1 = zip * tcos[p] + zrp * tsin[p];
        float r1 = ziq * tsin[q] - zrq * tcos[q];
        float i0 = ziq * tcos[q] + zrq * tsin[q];
        o[2 * p] = r0; o[2 * p + 1] = i0;
        o[2 * q] = r1; o[2 * q + 1] = i1;
    }
    HEAAC_SYNC();
}

/* ====================================================================================
 * AAC-LC core: TNS + IMDCT + windowing + overlap-add for one channel-frame
 * ================================================================================== */
struct HeaacCoreSmem {
    float zr[HEAAC_Z_FLOATS], zi[HEAAC_Z_FLOATS];
    float temp[128];
    float lpc[8][HEAAC_TNS_MAX_ORDER];
    int   tns_start[8], tns_size[8], tns_inc[8], tns_order[8];
};

template <int NT>
HEAAC_DEV void heaac_vfw(float *dst, const float *s0, const float *s1, const float *HEAAC_RESTRICT win,
                         float bias, int len,  int tid)
{
    constexpr int nt = NT;
    HEAAC_NOUNROLL
    for (int i = tid; i < len; i += nt) {
        int j = 2 * len - 1 - i;
        float a = s0[i], b = s1[len - 1 - i], wi = win[i], wj = win[j];
        dst[i] = a * wj - b * wi + bias;
        dst[j] = a * wi + b * wj + bias;
    }
}

/* One of the (at most 8) independent TNS filters of a channel-frame (aacdec.c:1701-1728): slot = window for short
 * blocks, filter index (< 3) for long ones.  Returns 0 if the slot holds no filter; else the coefficient range as
 * (first index, count, step) and, if want_lpc, the direct-form coefficients (compute_lpc_coefs(normalize = 0),
 * lpc.h:75-96). */
HEAAC_DEV int heaac_tns_chain(const HeaacIcs *ics, const HeaacTns *HEAAC_RESTRICT tns, int sr_index,
                              const HeaacDevTables *HEAAC_RESTRICT T, int slot, int *start_o, int *size_o, int *inc_o,
                              int *order_o, float *lpc, int want_lpc)
{
    const int is8 = ics->window_sequence == HEAAC_EIGHT_SHORT;
    const int w = is8 ? slot : 0, f = is8 ? 0 : slot;
    const int nf = tns->n_filt[w];
    if (f >= nf || (!is8 && slot >= 3)) return 0;
    const uint16_t *swb = is8 ? T->swb_offset_128[sr_index] : T->swb_offset_1024[sr_index];
    const int num_swb = is8 ? T->num_swb_128[sr_index] : T->num_swb_1024[sr_index];
    const int tmb = is8 ? T->tns_max_bands_128[sr_index] : T->tns_max_bands_1024[sr_index];
    const int mmm = HEAAC_MIN(tmb, (int)ics->max_sfb);
    int bottom = num_swb, top = num_swb;
    for (int g = 0; g <= f; g++) {                   /* walk down to this filter's band range */
        const int sl = is8 ? w : g;
        top = bottom;
        bottom = HEAAC_MAX(0, top - (int)tns->length[sl]);
    }
    /* side info of device-resident batches reaches the kernels unchecked: an order beyond the coefficient array
     * (aac.h:44, never produced by a parser) is cut, not followed out of bounds */
    const int order = HEAAC_MIN((int)tns->order[slot], HEAAC_TNS_MAX_ORDER);
    if (!order) return 0;
    int start = swb[HEAAC_MIN(bottom, mmm)];
    const int end = swb[HEAAC_MIN(top, mmm)];
    const int size = end - start;
    int inc = 1;
    if (size <= 0) return 0;
    if (tns->direction[slot]) { inc = -1; start = end - 1; }
    if (want_lpc)
        for (int i = 0; i < order; i++) {
            const float r = -tns->coef[slot][i];
            lpc[i] = r;
            for (int j = 0; j < (i + 1) >> 1; j++) {
                const float a = lpc[j], b = lpc[i - 1 - j];
                lpc[j] = a + r * b;
                lpc[i - 1 - j] = b + r * a;
            }
        }
    *start_o = start + w * 128; *size_o = size; *inc_o = inc; *order_o = order;
    return 1;
}

#ifndef HEAAC_HOST_EMU
/* The all-pole filter of one chain with the last ORD outputs and the coefficients in registers:
 *   y[m] = (((x[m] - y[m-1] a[0]) - y[m-2] a[1]) - ...)      in exactly this order (apply_tns, aacdec.c:1729-1733);
 * taps beyond the filter order have a = 0 and taps before the start of the range see y = 0, which leaves the sum
 * unchanged bit for bit (x - 0 = x), so one unrolled form per order class serves every order. */
template <int ORD>
__device__ __forceinline__ void heaac_tns_run(const float *HEAAC_RESTRICT in, float *out, int start, int size, int inc,
                                              int order, const float *lpc)
{
    float a[ORD], h[ORD];
    HEAAC_UNROLL
    for (int i = 0; i < ORD; i++) { a[i] = i < order ? lpc[i] : 0.0f; h[i] = 0.0f; }
    for (int m = 0; m < size; m++, start += inc) {
        float acc = in[start];
        HEAAC_UNROLL
        for (int i = 0; i < ORD; i++) acc -= h[i] * a[i];
        out[start] = acc;
        HEAAC_UNROLL
        for (int i = ORD - 1; i > 0; i--) h[i] = h[i - 1];
        h[0] = acc;
    }
}
#endif

/* `c`   : 1024 coefficients of this channel-frame, already staged in shared memory; overwritten (TNS in
 *         place, then the half-IMDCT output "buf")
 * `out` : 1024 floats in shared or global memory (time-domain output, bias included) */
template <int NT>
HEAAC_DEV void heaac_core_channel(HeaacCoreSmem *sm, float *c,
                                  const HeaacIcs *HEAAC_RESTRICT ics_p, const HeaacTns *HEAAC_RESTRICT tns,
                                  const float *HEAAC_RESTRICT tns_done,
                                  HeaacChanState *st, float *out, float bias, int sr_index,
                                  const HeaacDevTables *HEAAC_RESTRICT T,  int tid, const float *saved_rd = nullptr)
{
    /* saved_rd: where to read the overlap-add history from (a copy the caller fetched ahead of time; it may be `out` itself:
     * every output sample is computed by the thread that reads the history value at the same index); NULL: the state block */
    constexpr int nt = NT;
    HeaacIcs ics = *ics_p;
    ics.window_sequence &= 3; ics.window_shape &= 1;         /* unchecked side info stays inside the case table */
    const int ws0 = ics.window_sequence;
    const int ws1 = (ics.prev_valid ? (int)ics.prev_window_sequence : st->prev_ws) & 3;
    const int shape_prev = (ics.prev_valid ? (int)ics.prev_window_shape : st->prev_shape) & 1;
    const int is8 = ws0 == HEAAC_EIGHT_SHORT;

    /* ---- TNS, aacdec.c:1698-1736 ---- */
    const int use_tns = ics.tns_present && tns;
    if (use_tns && tns_done) {
        /* the filters already ran in k_tns (one thread per chain, registers only): take over the filtered ranges */
        for (int slot = 0; slot < 8; slot++) {
            int start, size, inc, order;
            float lpc_unused[HEAAC_TNS_MAX_ORDER];
            if (!heaac_tns_chain(&ics, tns, sr_index, T, slot, &start, &size, &inc, &order, lpc_unused, 0)) continue;
            const int lo = inc > 0 ? start : start - (size - 1);
            HEAAC_NOUNROLL
            for (int i = tid; i < size; i += nt) c[lo + i] = tns_done[lo + i];
        }
        HEAAC_SYNC();
    }
#ifdef HEAAC_HOST_EMU                                    /* (the kernels always run the pre-pass; the in-place form serves the logic tests) */
    else if (use_tns) {
        for (int slot = tid; slot < 8; slot += nt) {
            int start, size, inc, order;
            sm->tns_size[slot] = 0;
            if (!heaac_tns_chain(&ics, tns, sr_index, T, slot, &start, &size, &inc, &order, sm->lpc[slot], 1)) continue;
            sm->tns_start[slot] = start; sm->tns_size[slot] = size; sm->tns_inc[slot] = inc; sm->tns_order[slot] = order;
        }
        HEAAC_SYNC();
        /* all-pole filters over disjoint coefficient ranges: one serial chain per work item */
        for (int slot = tid; slot < 8; slot += nt) {
            int size = sm->tns_size[slot];
            if (size <= 0) continue;
            int start = sm->tns_start[slot], inc = sm->tns_inc[slot], order = sm->tns_order[slot];
            const float *lpc = sm->lpc[slot];
            for (int m = 0; m < size; m++, start += inc) {
                float acc = c[start];
                int lim = HEAAC_MIN(m, order);
                for (int i = 1; i <= lim; i++)
                    acc -= c[start - i * inc] * lpc[i - 1];
                c[start] = acc;
            }
        }
        HEAAC_SYNC();
    }
#endif
    else {
        HEAAC_SYNC();
    }

    /* ---- half IMDCT: 1 x N=2048 or 8 x N=256, aacdec.c:1755-1763 ---- */
    if (is8) {
        heaac_imdct_half_smem<NT>(c, c, sm->zr, sm->zi, 6, 8, T->tcos_256, T->tsin_256, T, tid);
        HEAAC_SYNC();        /* (the 64-point FFTs take one pass less than the 512-point one: every path through this routine crosses
                              * the same number of barriers, so that channels with different window sequences may share a CTA) */
    } else
        heaac_imdct_half_smem<NT>(c, c, sm->zr, sm->zi, 9, 1, T->tcos_2048, T->tsin_2048, T, tid);

    /* ---- windowing + overlap-add, aacdec.c:1771-1805 ---- */
    const float *buf = c;
    float *saved = st->saved;
    const float *saved_r = saved_rd ? saved_rd : st->saved;
    const float *sw      = ics.window_shape ? T->kbd_128 : T->sine_128;
    const float *lw_prev = shape_prev ? T->kbd_1024 : T->sine_1024;
    const float *sw_prev = shape_prev ? T->kbd_128 : T->sine_128;
    const int prev_long = ws1 == HEAAC_ONLY_LONG || ws1 == HEAAC_LONG_STOP;

    if (prev_long && (ws0 == HEAAC_ONLY_LONG || ws0 == HEAAC_LONG_START)) {
        heaac_vfw<NT>(out, saved_r, buf, lw_prev, bias, 512, tid);
    } else {
        HEAAC_NOUNROLL
        for (int i = tid; i < 448; i += nt) out[i] = saved_r[i] + bias;
        heaac_vfw<NT>(out + 448, saved_r + 448, buf, sw_prev, bias, 64, tid);
        if (is8) {
            for (int j = 1; j < 4; j++)
                heaac_vfw<NT>(out + 448 + 128 * j, buf + 128 * (j - 1) + 64, buf + 128 * j, sw, bias, 64, tid);
            heaac_vfw<NT>(sm->temp, buf + 3 * 128 + 64, buf + 4 * 128, sw, bias, 64, tid);
        } else {
            HEAAC_NOUNROLL
            for (int i = 576 + tid; i < 1024; i += nt) out[i] = buf[i - 512] + bias;
        }
    }
    HEAAC_SYNC();        /* all reads of the old `saved` are done; temp[] is complete */
    if (is8) {
        HEAAC_NOUNROLL
        for (int i = tid; i < 64; i += nt) {
            out[960 + i] = sm->temp[i];
            saved[i] = sm->temp[64 + i] - bias;
        }
        for (int j = 0; j < 3; j++)
            heaac_vfw<NT>(saved + 64 + 128 * j, buf + (4 + j) * 128 + 64, buf + (5 + j) * 128, sw, 0.0f, 64, tid);
        HEAAC_NOUNROLL
        for (int i = tid; i < 64; i += nt) saved[448 + i] = buf[7 * 128 + 64 + i];
    } else if (ws0 == HEAAC_LONG_START) {
        HEAAC_NOUNROLL
        for (int i = tid; i < 448; i += nt) saved[i] = buf[512 + i];
        HEAAC_NOUNROLL
        for (int i = tid; i < 64; i += nt)  saved[448 + i] = buf[7 * 128 + 64 + i];
    } else {
        HEAAC_NOUNROLL
        for (int i = tid; i < 512; i += nt) saved[i] = buf[512 + i];
    }
    if (tid == 0) { st->prev_ws = ws0; st->prev_shape = ics.window_shape; }
    HEAAC_SYNC();
}

/* ====================================================================================
 * Independent channel coupling + the float -> s16 epilogue for layouts with coupling channel elements
 * (apply_channel_coupling / apply_independent_coupling, aacdec.c:1849-1898): every output channel of one stream-frame.
 *   f32  : this stream's decoded channels, planar [ncc][1024], bias included (what imdct_and_windowing leaves in ret)
 *   cce  : this stream-frame's coupling side info, [n_cce]
 * ================================================================================== */
template <int NT>
HEAAC_DEV void heaac_couple_stream(const HeaacGeom *g, const float *f32, const HeaacCce *HEAAC_RESTRICT cce,
                                   int16_t *pcm, float *pcm_f32, int tid)
{
    constexpr int nt = NT;
    const float bias = g->add_bias;
    const int bias_mode = bias != 0.0f;
    for (int i = tid; i < 1024; i += nt)
        for (int c = 0; c < g->ncc; c++) {
            const int slot = g->out_slot[c];
            if (slot == 0xff) continue;                      /* a coupling channel itself: no output */
            float v = f32[c * 1024 + i];
            for (int k = 0; k < g->n_cce; k++) {
                if (cce[k].coupling_point != 3) continue;
                const float src = heaac_fsub(f32[g->cce_ch[k] * 1024 + i], bias);
                const int nt_ = HEAAC_MIN((int)cce[k].n_targets, HEAAC_CCE_MAX_TARGETS);
                for (int t = 0; t < nt_; t++)
                    if (cce[k].target_ch[t] == c) v = heaac_fadd(v, heaac_fmul(cce[k].gain[t], src));
            }
            if (pcm)     pcm[(size_t)i * g->nout + slot] = heaac_to_s16(v, bias_mode);
            if (pcm_f32) pcm_f32[(size_t)slot * 1024 + i] = v;
        }
}

#include "heaac_sbr.cuh"

/* ====================================================================================
 * Parametric Stereo for one stream-frame (ff_ps_apply, aacps.c:973-992)
 * ================================================================================== */
/* slot dimension padded to 33: the time-serial phases run with lanes = bands, and an unpadded row (64 floats) would put
 * all lanes of a warp on the same bank */
#define HEAAC_PS_NS 33
struct HeaacPsSmem {
    float L[91][HEAAC_PS_NS][2];    /* Lbuf: hybrid-domain mono signal, then left output */
    float R[91][HEAAC_PS_NS][2];    /* Rbuf: decorrelated signal, then right output */
    float in[5][44][2];             /* in_buf: 6 history + 38 new slots of QMF bands 0..4 */
    float power[34][HEAAC_PS_NS];   /* per parameter band and slot: the power, then (in place) the transient gain */
    HeaacPsFrame p;
};

HEAAC_DEV void heaac_ps_hyb_cx(const float (*in)[2], const float (*f)[2], float *re, float *im)
{
    float sr = f[6][0] * in[6][0], si = f[6][0] * in[6][1];          /* aacps.c:311,344 */
    for (int j = 0; j < 6; j++) {
        float a_re = in[j][0], a_im = in[j][1], b_re = in[12 - j][0], b_im = in[12 - j][1];
        sr += f[j][0] * (a_re + b_re) - f[j][1] * (a_im - b_im);
        si += f[j][0] * (a_im + b_im) + f[j][1] * (a_re - b_re);
    }
    *re = sr; *im = si;
}

/* index remapping between 10/20/34 parameter bands (aacps.c:461-605) */
HEAAC_DEV int heaac_ps_map_idx(const int8_t *par, int num_par, int is34, int full, int b)
{
    /* returns the remapped index for target band b; caller guarantees b is a band the reference writes */
    if (is34) {
        if (num_par == 20 || num_par == 11) {
            if (b == 4) return (par[2] + par[3]) / 2;
            if (b == 1) return (par[0] + par[1]) / 2;
            const int8_t src[34] = { 0, 0, 1, 2, 0, 3, 4, 4, 5, 5, 6, 7, 8, 8, 9, 9, 10, 11, 12, 13,
                                     14, 14, 15, 15, 16, 16, 17, 17, 18, 18, 18, 18, 19, 19 };
            return par[src[b]];
        } else if (num_par == 10 || num_par == 5) {
            const int8_t src[34] = { 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6,
                                     7, 7, 7, 7, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9 };
            if (!full && b == 16) return 0;
            return par[src[b]];
        }
        return par[b];
    }
    if (num_par == 34 || num_par == 17) {
        switch (b) {
        case 0:  return (2 * par[0] + par[1]) / 3;
        case 1:  return (par[1] + 2 * par[2]) / 3;
        case 2:  return (2 * par[3] + par[4]) / 3;
        case 3:  return (par[4] + 2 * par[5]) / 3;
        case 4:  return (par[6] + par[7]) / 2;
        case 5:  return (par[8] + par[9]) / 2;
        case 6:  return par[10];
        case 7:  return par[11];
        case 8:  return (par[12] + par[13]) / 2;
        case 9:  return (par[14] + par[15]) / 2;
        case 10: return par[16];
        case 11: return par[17];
        case 12: return par[18];
        case 13: return par[19];
        case 14: return (par[20] + par[21]) / 2;
        case 15: return (par[22] + par[23]) / 2;
        case 16: return (par[24] + par[25]) / 2;
        case 17: return (par[26] + par[27]) / 2;
        case 18: return (par[28] + par[29] + par[30] + par[31]) / 4;
        default: return (par[32] + par[33]) / 2;
        }
    } else if (num_par == 10 || num_par == 5) {
        if (!full && b == 10) return 0;
        return par[b >> 1];
    }
    return par[b];
}

/* map_val_20_to_34 / map_val_34_to_20 on one row of 34 floats (aacps.c:501-523, 607-643) */
HEAAC_DEV void heaac_ps_map_val(float *v, int to34)
{
    float o[34];
    if (to34) {
        const int8_t src[34] = { 0, 0, 1, 2, 0, 3, 4, 4, 5, 5, 6, 7, 8, 8, 9, 9, 10, 11, 12, 13,
                                 14, 14, 15, 15, 16, 16, 17, 17, 18, 18, 18, 18, 19, 19 };
        for (int b = 0; b < 34; b++) o[b] = v[src[b]];
        o[4] = (v[2] + v[3]) * 0.5f;
        o[1] = (v[0] + v[1]) * 0.5f;
    } else {
        for (int b = 0; b < 34; b++) o[b] = v[b];
        o[0] = (2 * v[0] + v[1]) * 0.33333333f;   o[1] = (v[1] + 2 * v[2]) * 0.33333333f;
        o[2] = (2 * v[3] + v[4]) * 0.33333333f;   o[3] = (v[4] + 2 * v[5]) * 0.33333333f;
        o[4] = (v[6] + v[7]) * 0.5f;              o[5] = (v[8] + v[9]) * 0.5f;
        o[6] = v[10];                             o[7] = v[11];
        o[8] = (v[12] + v[13]) * 0.5f;            o[9] = (v[14] + v[15]) * 0.5f;
        o[10] = v[16]; o[11] = v[17]; o[12] = v[18]; o[13] = v[19];
        o[14] = (v[20] + v[21]) * 0.5f;           o[15] = (v[22] + v[23]) * 0.5f;
        o[16] = (v[24] + v[25]) * 0.5f;           o[17] = (v[26] + v[27]) * 0.5f;
        o[18] = (v[28] + v[29] + v[30] + v[31]) * 0.25f;
        o[19] = (v[32] + v[33]) * 0.5f;
    }
    for (int b = 0; b < 34; b++) v[b] = o[b];
}

/* XL / XR: the element's band-major X buffers in global scratch (XL in/out, XR out).
 * top_in = kx + m of the SBR element (aacsbr.c:1753). */
template <int NT>
HEAAC_DEV void heaac_ps_stream(HeaacPsSmem *sm, const HeaacPsFrame *HEAAC_RESTRICT frame, HeaacPsState *st,
                               float *XL, float *XR, int top_in,
                               const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    constexpr int nt = NT;
    const float *Lre = XL, *Lim = XL + HEAAC_X_PLANE;
    /* Everything the first phases need comes from global memory (the frame's parameters, the QMF bands k_chain has just
     * written, the hybrid filter history): issue it all at once, one barrier, instead of three dependent round trips. */
    const int start_g = frame->start, is34_g = frame->is34bands;
    {
        const uint32_t *src = (const uint32_t *)frame;
        uint32_t *dst = (uint32_t *)&sm->p;
        for (int i = tid; i < (int)(sizeof(HeaacPsFrame) / 4); i += nt) dst[i] = src[i];
    }
    if (start_g) {
        for (int w = tid; w < 5 * 44; w += nt) {              /* in_buf: 6 history slots + this frame, QMF bands 0..4 */
            int i = w / 44, j = w - i * 44;
            if (j < 6) { sm->in[i][j][0] = st->in_hist[i][j][0]; sm->in[i][j][1] = st->in_hist[i][j][1]; }
            else       { sm->in[i][j][0] = Lre[i * HEAAC_XS + (j - 6)]; sm->in[i][j][1] = Lim[i * HEAAC_XS + (j - 6)]; }
        }
        const int nb = is34_g ? 59 : 61, q0 = is34_g ? 5 : 3, k0 = is34_g ? 32 : 10;   /* QMF bands that pass through */
        for (int w = tid; w < nb * 32; w += nt) {
            int i = w >> 5, n = w & 31;
            sm->L[i + k0][n][0] = Lre[(i + q0) * HEAAC_XS + n]; sm->L[i + k0][n][1] = Lim[(i + q0) * HEAAC_XS + n];
        }
    }
    HEAAC_SYNC();
    const HeaacPsFrame *p = &sm->p;
    if (!p->start) {                                      /* aacsbr.c:1755: R = copy of L */
        for (int i = tid; i < HEAAC_X_FLOATS; i += nt) XR[i] = XL[i];
        HEAAC_SYNC();
        return;
    }
    const int is34 = p->is34bands;
    const int nr_bands = is34 ? 91 : 71, nr_par = is34 ? 34 : 20, nr_ap = is34 ? 50 : 30;
    const int short_delay_band = is34 ? 62 : 42, decay_cutoff = is34 ? 32 : 10;
    const int8_t *k_to_i = is34 ? T->k_to_i_34 : T->k_to_i_20;
    const int top = top_in + nr_bands - 64;
    const int switched = is34 != p->is34bands_old;        /* aacps.c:665-671: all decorrelator state is cleared */

    /* ---- hybrid analysis, aacps.c:359-395 (in_buf and the pass-through bands are already in shared memory) ---- */
    for (int w = tid; w < 5 * 6; w += nt) {
        int i = w / 6, j = w - i * 6;
        st->in_hist[i][j][0] = sm->in[i][32 + j][0]; st->in_hist[i][j][1] = sm->in[i][32 + j][1];
    }
    if (is34) {
        for (int w = tid; w < 32 * 32; w += nt) {         /* 12 + 8 + 4 + 4 + 4 sub-subbands x 32 slots */
            int q = w >> 5, n = w & 31;
            int band, sub; const float (*f)[7][2];
            if (q < 12)      { band = 0; sub = q;      f = T->f34_0_12; }
            else if (q < 20) { band = 1; sub = q - 12; f = T->f34_1_8; }
            else             { band = 2 + ((q - 20) >> 2); sub = (q - 20) & 3; f = T->f34_2_4; }
            heaac_ps_hyb_cx(&sm->in[band][n], f[sub], &sm->L[q][n][0], &sm->L[q][n][1]);
        }
    } else {
        for (int n = tid; n < 32; n += nt) {              /* hybrid6_cx + 2 x hybrid2_re */
            float t[8][2];
            for (int q = 0; q < 8; q++) heaac_ps_hyb_cx(&sm->in[0][n], T->f20_0_8[q], &t[q][0], &t[q][1]);
            for (int c = 0; c < 2; c++) {
                sm->L[0][n][c] = t[6][c]; sm->L[1][n][c] = t[7][c]; sm->L[2][n][c] = t[0][c]; sm->L[3][n][c] = t[1][c];
                sm->L[4][n][c] = t[2][c] + t[5][c];
                sm->L[5][n][c] = t[3][c] + t[4][c];
            }
            for (int b = 1; b <= 2; b++) {                /* aacps.c:283-300, reverse = (b == 1) */
                const float (*in)[2] = &sm->in[b][n];
                const float *f = T->g1_Q2;
                float re_in = f[6] * in[6][0], re_op = 0.0f, im_in = f[6] * in[6][1], im_op = 0.0f;
                for (int j = 0; j < 6; j += 2) {
                    re_op += f[j + 1] * (in[j + 1][0] + in[12 - j - 1][0]);
                    im_op += f[j + 1] * (in[j + 1][1] + in[12 - j - 1][1]);
                }
                int rev = b == 1, base = 6 + 2 * (b - 1);
                sm->L[base + rev][n][0]  = re_in + re_op; sm->L[base + rev][n][1]  = im_in + im_op;
                sm->L[base + !rev][n][0] = re_in - re_op; sm->L[base + !rev][n][1] = im_in - im_op;
            }
        }
    }
    HEAAC_SYNC();

    /* ---- decorrelation, aacps.c:645-754 ---- */
    for (int w = tid; w < nr_par * 32; w += nt) {
        int i = w >> 5, n = w & 31;
        float pw = 0.0f;                                  /* sum over the bands of parameter band i, ascending k (aacps.c:690-697) */
        const uint8_t *kl = T->ps_inv_k[is34];
        for (int q = T->ps_inv_off[is34][i]; q < T->ps_inv_off[is34][i + 1]; q++) {
            const int k = kl[q];
            pw += sm->L[k][n][0] * sm->L[k][n][0] + sm->L[k][n][1] * sm->L[k][n][1];
        }
        sm->power[i][n] = pw;
    }
    HEAAC_SYNC();
    /* ---- one region without barriers: the transient detector (a serial recursion per parameter band), the decorrelator (a
     *      serial recursion per hybrid band) and the mixing matrices (serial over envelopes per parameter band) are independent
     *      of each other -- the transient gain only scales the decorrelator's output, which is done where that output is
     *      read -- so with 256 threads they run side by side on thread groups T, H and A ---- */
    constexpr bool PS_SPLIT = nt >= 256;
    const bool inT = !PS_SPLIT || tid < 64, inH = !PS_SPLIT || (tid >= 64 && tid < 128), inA = !PS_SPLIT || tid >= 128;
    const int tT = tid, nT = PS_SPLIT ? 64 : nt, tH = PS_SPLIT ? tid - 64 : tid, nH = PS_SPLIT ? 64 : nt;
    const int tA = PS_SPLIT ? tid - 128 : tid, nA = PS_SPLIT ? nt - 128 : nt;
    if (inT) {
    for (int i = tT; i < nr_par; i += nT) {               /* transient detection: serial in time */
        float pdn = switched ? 0.0f : st->peak_decay_nrg[i];
        float psm = switched ? 0.0f : st->power_smooth[i];
        float pds = switched ? 0.0f : st->peak_decay_diff_smooth[i];
        for (int n = 0; n < 32; n++) {
            float pw = sm->power[i][n];
            float decayed = 0.76592833836465f * pdn;
            pdn = HEAAC_MAX(decayed, pw);
            psm += 0.25f * (pw - psm);
            pds += 0.25f * (pdn - pw - pds);
            float denom = 1.5f * pds;
            sm->power[i][n] = (denom > psm) ? heaac_fdiv(psm, denom) : 1.0f;
        }
        st->peak_decay_nrg[i] = pdn; st->power_smooth[i] = psm; st->peak_decay_diff_smooth[i] = pds;
    }
    if (switched)
        for (int i = nr_par + tT; i < 34; i += nT) {
            st->peak_decay_nrg[i] = 0.0f; st->power_smooth[i] = 0.0f; st->peak_decay_diff_smooth[i] = 0.0f;
        }
    }
    if (inA) {
    for (int k = tA; k < nr_bands; k += nA) {             /* one hybrid band per work item, serial in time */
        const int cleared = switched || k >= top;         /* aacps.c:980-983 + 665-671 */
        const float (*Lk)[2] = sm->L[k];
        float (*Rk)[2] = sm->R[k];
        if (k < nr_ap) {
            float g = 1.f - 0.05f * (k - decay_cutoff);
            g = g < 0.f ? 0.f : (g > 1.f ? 1.f : g);
            /* everything the recursion carries lives in registers: the three links' last five outputs (ap[m][4] newest), their
             * coefficients, and the two input samples the 2-slot delay still needs from the previous frame (delay[k][12..13]) */
            const float ag0 = 0.65143905753106f * g, ag1 = 0.56471812200776f * g, ag2 = 0.48954165955695f * g;
            float ap[3][5][2];
            HEAAC_UNROLL
            for (int m = 0; m < 3; m++) {
                HEAAC_UNROLL
                for (int j = 0; j < 5; j++) {
                    ap[m][j][0] = cleared ? 0.0f : st->ap_delay[k][m][j][0];
                    ap[m][j][1] = cleared ? 0.0f : st->ap_delay[k][m][j][1];
                }
            }
            float fd[3][2];
            HEAAC_UNROLL
            for (int m = 0; m < 3; m++) { fd[m][0] = T->Q_fract_allpass[is34][k][m][0]; fd[m][1] = T->Q_fract_allpass[is34][k][m][1]; }
            const float pf_re = T->phi_fract[is34][k][0], pf_im = T->phi_fract[is34][k][1];
            float d2_re = cleared ? 0.0f : st->delay[k][12][0], d2_im = cleared ? 0.0f : st->delay[k][12][1];     /* two slots back */
            float d1_re = cleared ? 0.0f : st->delay[k][13][0], d1_im = cleared ? 0.0f : st->delay[k][13][1];     /* one slot back */
            for (int n = 0; n < 32; n++) {
                const float d_re = d2_re, d_im = d2_im;       /* delay[k][n + 12] */
                d2_re = d1_re; d2_im = d1_im; d1_re = Lk[n][0]; d1_im = Lk[n][1];
                float in_re = d_re * pf_re - d_im * pf_im;
                float in_im = d_re * pf_im + d_im * pf_re;
                HEAAC_UNROLL
                for (int m = 0; m < 3; m++) {
                    /* ap_delay[k][m][n + 5 - link_delay], link_delay = 3 + m: ring of the last 5 values, ap[m][4] newest */
                    const float agm = m == 0 ? ag0 : m == 1 ? ag1 : ag2;
                    float a_re = agm * in_re, a_im = agm * in_im;
                    float ld_re = ap[m][2 - m][0], ld_im = ap[m][2 - m][1];
                    float nw_re = in_re, nw_im = in_im;
                    in_re = ld_re * fd[m][0] - ld_im * fd[m][1] - a_re;
                    in_im = ld_re * fd[m][1] + ld_im * fd[m][0] - a_im;
                    nw_re += agm * in_re;
                    nw_im += agm * in_im;
                    HEAAC_UNROLL
                    for (int j = 0; j < 4; j++) { ap[m][j][0] = ap[m][j + 1][0]; ap[m][j][1] = ap[m][j + 1][1]; }
                    ap[m][4][0] = nw_re; ap[m][4][1] = nw_im;
                }
                Rk[n][0] = in_re;                             /* (the transient gain is applied where R is read: the mixing) */
                Rk[n][1] = in_im;
            }
            HEAAC_UNROLL
            for (int m = 0; m < 3; m++) {
                HEAAC_UNROLL
                for (int j = 0; j < 5; j++) { st->ap_delay[k][m][j][0] = ap[m][j][0]; st->ap_delay[k][m][j][1] = ap[m][j][1]; }
            }
        } else {
            const int d = k < short_delay_band ? 14 : 1;
            for (int n = 0; n < 32; n++) {
                const int pos = n + 14 - d;               /* index into [previous frame's delay line (14) | this frame (32)] */
                float d_re, d_im;
                if (pos < 14) { d_re = cleared ? 0.0f : st->delay[k][pos][0]; d_im = cleared ? 0.0f : st->delay[k][pos][1]; }
                else          { d_re = Lk[pos - 14][0]; d_im = Lk[pos - 14][1]; }
                Rk[n][0] = d_re;
                Rk[n][1] = d_im;
            }
        }
        for (int j = 0; j < 14; j++) { st->delay[k][j][0] = Lk[18 + j][0]; st->delay[k][j][1] = Lk[18 + j][1]; }
    }
    /* bands beyond nr_bands keep their state unless everything was cleared */
    if (switched) {
        for (int w = tA; w < (91 - nr_bands) * 14; w += nA) {
            int k = nr_bands + w / 14, j = w % 14;
            st->delay[k][j][0] = 0.0f; st->delay[k][j][1] = 0.0f;
        }
        for (int w = tA; w < (50 - nr_ap) * 15; w += nA) {
            int k = nr_ap + w / 15, r = w % 15;
            st->ap_delay[k][r / 5][r % 5][0] = 0.0f; st->ap_delay[k][r / 5][r % 5][1] = 0.0f;
        }
    }
    /* ap_delay of all-pass bands at or above `top` that this frame did not run: cleared (aacps.c:982-983) */
    /* (bands < nr_ap are always run above, with cleared history when k >= top) */

    }
    if (inH) {
    /* ---- stereo_processing, aacps.c:794-971: H matrices.  Work items own what they touch, so no barrier is needed inside:
     *      one item per parameter band for the hand-over of the previous frame's last envelope, the index remapping and the
     *      matrices of all envelopes ---- */
    {
        const float (*H_LUT)[8][4] = p->icc_mode < 3 ? T->HA : T->HB;
        for (int b = tH; b < 34; b += nH) {               /* serial over envelopes: the phase smoother has memory */
            /* the matrices live in the stream's state block (all six rows persist: stale rows are re-read by later frames,
             * aacps.c:816-823); row 0 becomes the previous frame's last envelope -- column b by the item that owns column b,
             * before it overwrites the rows of this frame's envelopes */
            for (int r = 0; r < 8; r++) st->H[r >> 1][r & 1][0][b] = st->H[r >> 1][r & 1][p->num_env_old][b];
            if (switched && b < 17) { st->ipd_hist[b] = 0; st->opd_hist[b] = 0; }       /* ipdopd_reset, aacps.c:141-148 */
            if (b >= nr_par) continue;
            int oh = st->opd_hist[b < 34 ? b : 0], ih = st->ipd_hist[b < 34 ? b : 0];
            for (int e = 0; e < p->num_env; e++) {
                const int iid = heaac_ps_map_idx(p->iid_par[e], p->nr_iid_par, is34, 1, b);
                const int icc = heaac_ps_map_idx(p->icc_par[e], p->nr_icc_par, is34, 1, b);
                /* indices are range-checked by the parser (aacps.c:116-117); clamp so that a corrupt
                 * parameter block can never index outside the table */
                int li = (int8_t)iid + 7 + 23 * p->iid_quant, lc = (int8_t)icc;
                li = li < 0 ? 0 : (li > 45 ? 45 : li);
                lc = lc < 0 ? 0 : (lc > 7 ? 7 : lc);
                const float *lut = H_LUT[li][lc];
                float h11 = lut[0], h12 = lut[1], h21 = lut[2], h22 = lut[3];
                if (p->enable_ipdopd && b < p->nr_ipdopd_par) {
                    /* (20-band processing with 17 phase parameters: bands 11..16 have no remapped value -- the reference reads
                     * uninitialised memory there, aacps.c:845-848; zero here) */
                    const int has_pd = b < (is34 ? 17 : 11);
                    const int ipd = has_pd ? (int8_t)heaac_ps_map_idx(p->ipd_par[e], p->nr_ipdopd_par, is34, 0, b) : 0;
                    const int opd = has_pd ? (int8_t)heaac_ps_map_idx(p->opd_par[e], p->nr_ipdopd_par, is34, 0, b) : 0;
                    int opd_idx = (oh * 8 + (opd & 7)) & 511;
                    int ipd_idx = (ih * 8 + (ipd & 7)) & 511;
                    float opd_re = T->pd_re_smooth[opd_idx], opd_im = T->pd_im_smooth[opd_idx];
                    float ipd_re = T->pd_re_smooth[ipd_idx], ipd_im = T->pd_im_smooth[ipd_idx];
                    oh = opd_idx & 0x3F; ih = ipd_idx & 0x3F;
                    float adj_re = opd_re * ipd_re + opd_im * ipd_im;
                    float adj_im = opd_im * ipd_re - opd_re * ipd_im;
                    st->H[0][1][e + 1][b] = h11 * opd_im;  h11 = h11 * opd_re;
                    st->H[1][1][e + 1][b] = h12 * adj_im;  h12 = h12 * adj_re;
                    st->H[2][1][e + 1][b] = h21 * opd_im;  h21 = h21 * opd_re;
                    st->H[3][1][e + 1][b] = h22 * adj_im;  h22 = h22 * adj_re;
                }
                st->H[0][0][e + 1][b] = h11; st->H[1][0][e + 1][b] = h12;
                st->H[2][0][e + 1][b] = h21; st->H[3][0][e + 1][b] = h22;
            }
            st->opd_hist[b] = (int8_t)oh; st->ipd_hist[b] = (int8_t)ih;
        }
    }
    }
    HEAAC_SYNC();
    if (switched) {                                       /* 20 <-> 34 bands: the previous frame's matrices are remapped (aacps.c:824-843) */
        for (int w = tid; w < 8; w += nt) heaac_ps_map_val(st->H[w >> 1][w & 1][0], is34);
        HEAAC_SYNC();
    }
    /* ---- mixing: one (hybrid band, envelope) per work item -- the interpolation of the matrices restarts from H[e] at every
     *      envelope border (aacps.c:905-920), so envelopes are independent; serial over the envelope's slots (the interpolated
     *      coefficients are running sums) ---- */
    for (int w = tid; w < nr_bands * p->num_env; w += nt) {
        const int e0 = w / nr_bands, k = w - e0 * nr_bands;
        const int b = k_to_i[k];
        const int neg = p->enable_ipdopd && ((is34 && k <= 13 && k >= 9) || (!is34 && k <= 1));
        for (int e = e0; e <= e0; e++) {
            int start = p->border_position[e], stop = p->border_position[e + 1];
            float width = heaac_fdiv(1.f, (float)(stop - start));
            float hr[4], hi[4], hrs[4], his[4];
            for (int h = 0; h < 4; h++) {
                hr[h] = st->H[h][0][e][b];
                hrs[h] = (st->H[h][0][e + 1][b] - hr[h]) * width;
                hi[h] = 0.0f; his[h] = 0.0f;
                if (p->enable_ipdopd) {
                    hi[h] = neg ? -st->H[h][1][e][b] : st->H[h][1][e][b];
                    his[h] = (st->H[h][1][e + 1][b] - hi[h]) * width;
                }
            }
            for (int n = start + 1; n <= stop; n++) {
                const float tg = sm->power[b][n];             /* transient gain of the decorrelated signal, aacps.c:747-752 */
                float l_re = sm->L[k][n][0], l_im = sm->L[k][n][1], r_re = tg * sm->R[k][n][0], r_im = tg * sm->R[k][n][1];
                for (int h = 0; h < 4; h++) hr[h] += hrs[h];
                if (p->enable_ipdopd) {
                    for (int h = 0; h < 4; h++) hi[h] += his[h];
                    sm->L[k][n][0] = hr[0] * l_re + hr[2] * r_re - hi[0] * l_im - hi[2] * r_im;
                    sm->L[k][n][1] = hr[0] * l_im + hr[2] * r_im + hi[0] * l_re + hi[2] * r_re;
                    sm->R[k][n][0] = hr[1] * l_re + hr[3] * r_re - hi[1] * l_im - hi[3] * r_im;
                    sm->R[k][n][1] = hr[1] * l_im + hr[3] * r_im + hi[1] * l_re + hi[3] * r_re;
                } else {
                    sm->L[k][n][0] = hr[0] * l_re + hr[2] * r_re;
                    sm->L[k][n][1] = hr[0] * l_im + hr[2] * r_im;
                    sm->R[k][n][0] = hr[1] * l_re + hr[3] * r_re;
                    sm->R[k][n][1] = hr[1] * l_im + hr[3] * r_im;
                }
            }
        }
    }
    HEAAC_SYNC();

    /* ---- hybrid synthesis x 2, aacps.c:397-445: back to 64 QMF bands, slots 0..31 ---- */
    for (int w = tid; w < 2 * 32 * 64; w += nt) {
        int side = w >> 11, q = (w >> 5) & 63, n = w & 31;      /* slot fastest */
        const float (*in)[HEAAC_PS_NS][2] = side ? sm->R : sm->L;
        float *X = side ? XR : XL;
        float re, im;
        if (is34) {
            if (q < 5) {
                int lo = q == 0 ? 0 : q == 1 ? 12 : 20 + 4 * (q - 2);
                int cnt = q == 0 ? 12 : q == 1 ? 8 : 4;
                re = 0.0f; im = 0.0f;
                for (int i = 0; i < cnt; i++) { re += in[lo + i][n][0]; im += in[lo + i][n][1]; }
            } else { re = in[q + 27][n][0]; im = in[q + 27][n][1]; }
        } else {
            if (q == 0) {
                re = in[0][n][0] + in[1][n][0] + in[2][n][0] + in[3][n][0] + in[4][n][0] + in[5][n][0];
                im = in[0][n][1] + in[1][n][1] + in[2][n][1] + in[3][n][1] + in[4][n][1] + in[5][n][1];
            } else if (q == 1) { re = in[6][n][0] + in[7][n][0]; im = in[6][n][1] + in[7][n][1]; }
            else if (q == 2)   { re = in[8][n][0] + in[9][n][0]; im = in[8][n][1] + in[9][n][1]; }
            else               { re = in[q + 7][n][0]; im = in[q + 7][n][1]; }
        }
        X[q * HEAAC_XS + n] = re;
        X[HEAAC_X_PLANE + q * HEAAC_XS + n] = im;
    }
    HEAAC_SYNC();
}

/* ====================================================================================
 * QMF synthesis (sbr_qmf_synthesis, aacsbr.c:1175-1230) + float->s16 for the 1 or 2 output channels of
 * one stream-frame.  64 threads per channel:
 *   phase 1  thread = (plane re/im, slot i): one half-IMDCT (N = 128, scale 1/64) in registers, inputs read
 *            coalesced from the band-major X buffer (lanes = slots)
 *   phase 2  the im-plane threads combine b0 (parked in the slot's own v block) and b1 (registers) into the
 *            slot's 128-sample v block in place (aacsbr.c:1210-1213)
 *   phase 3  thread = (sample n, channel): the 10-tap polyphase sum for all 32 slots with the window
 *            column w[64 j + n] held in registers; s16 stores of the two channels interleave into
 *            contiguous bytes
 * v blocks use a row stride of 129 floats: phase 2 (lanes = slots) and phase 3 (lanes = samples) are both
 * free of bank conflicts.
 * ================================================================================== */
#define HEAAC_V_STRIDE 129
#define HEAAC_SYN_CH_FLOATS (41 * HEAAC_V_STRIDE + 7)    /* = 16 (mod 32): the 2nd channel starts 16 banks away */
struct HeaacSynthSmem {
    float ch[2][HEAAC_SYN_CH_FLOATS];   /* per channel: v[41][129], block 0 = newest, blocks 32..40 = history */
};

template <int NT>
HEAAC_DEV void heaac_synth_stream(HeaacSynthSmem *sm, const float *HEAAC_RESTRICT X, HeaacSynthState *st, int nch,
                                  int nout, int out_slot0, int out_slot1, int16_t *pcm, float *pcm_f32, float scale, float bias,
                                  const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    constexpr int nt = NT;
    const int nthr = 64 * nch;                             /* nch = channels of this element, nout = of the stream */
    for (int w = tid; w < nch * 1152; w += nt) {           /* history -> blocks 32..40, asynchronously: first needed in phase 3 */
        int c = w / 1152, r = w - c * 1152;
        heaac_cp_async4(&sm->ch[c][(32 + (r >> 7)) * HEAAC_V_STRIDE + (r & 127)], &st[c].v[r]);
    }
    heaac_cp_async_commit();
    /* ---- phase 1: 64 transforms per channel, all at once: the threads of the real plane park b0 in the slot's own
     *      v block, those of the imaginary plane keep b1 in registers across the barrier; phase 2: v = (-b0 + b1, b0 + b1) ---- */
    for (int t0 = 0; t0 < nthr; t0 += nt) {             /* trip count uniform over the CTA: the barrier inside is safe */
        const int t = t0 + tid;
        const bool act = t < nthr;
        const int c = t >> 6, plane = (t >> 5) & 1, i = t & 31;
        const float *x = X + (size_t)c * HEAAC_X_FLOATS + plane * HEAAC_X_PLANE + i;
        float *v = sm->ch[act ? c : 0] + (31 - i) * HEAAC_V_STRIDE;
        float b1[64];
        if (act && plane == 0) {
            heaac_imdct128_reg<HEAAC_MDCT_SYN>(
                [&](int j) { return HEAAC_LDX(x + j * HEAAC_XS); },
                [&](int j, float b) { v[j] = b; });
        } else if (act) {
            heaac_imdct128_reg<HEAAC_MDCT_SYN>(
                [&](int j) { float xv = HEAAC_LDX(x + j * HEAAC_XS); return (j & 1) ? -xv : xv; },     /* aacsbr.c:1205-1207 */
                [&](int j, float b) { b1[j] = b; });
        }
        HEAAC_SYNC();
        if (act && plane == 1) {
            HEAAC_UNROLL
            for (int n = 0; n < 32; n++) {              /* v holds b0[0..63]; becomes v[0..127] in place */
                const float p1 = v[63 - n], p2 = v[n], q1 = b1[n], q2 = b1[63 - n];
                v[n]       = -p1 + q1;
                v[127 - n] =  p1 + q1;
                v[63 - n]  = -p2 + q2;
                v[64 + n]  =  p2 + q2;
            }
        }
    }
    heaac_cp_async_wait<0>();
    HEAAC_SYNC();
    /* ---- phase 3: windowing, out[n] = sum_j v[off_j + n] w[64 j + n] (aacsbr.c:1215-1227) ---- */
    const int scale_and_bias = scale != 1.0f || bias != 0.0f;
    const int bias_mode = bias != 0.0f;
    for (int t = tid; t < nthr; t += nt) {
        const int c = nch == 2 ? (t & 1) : 0, n = nch == 2 ? (t >> 1) : t;
        float w[10];
        HEAAC_UNROLL
        for (int j = 0; j < 10; j++) w[j] = T->qmf_us[64 * j + n];
        const float *vb = sm->ch[c] + n;
        const int slot = c ? out_slot1 : out_slot0;
        for (int i = 0; i < 32; i++) {
            const float *v = vb + (31 - i) * HEAAC_V_STRIDE;
            float acc = v[0] * w[0] + 0.0f;
            HEAAC_UNROLL
            for (int j = 1; j < 10; j++)
                acc = v[j * HEAAC_V_STRIDE + ((j & 1) ? 64 : 0)] * w[j] + acc;
            if (scale_and_bias) acc = acc * scale + bias;
            const int o = i * 64 + n;
            if (pcm)     pcm[(size_t)o * nout + slot] = heaac_to_s16(acc, bias_mode);
            if (pcm_f32) pcm_f32[(size_t)slot * 2048 + o] = acc;
        }
    }
    for (int w = tid; w < nch * 1152; w += nt) {           /* newest 9 blocks become the next frame's history */
        int c = w / 1152, r = w - c * 1152;
        st[c].v[r] = sm->ch[c][(r >> 7) * HEAAC_V_STRIDE + (r & 127)];
    }
    HEAAC_SYNC();
}


/* Downsampled synthesis bank (div = 1, aacsbr.c:1194-1203): one N=128 half-IMDCT per slot on (-Re X[0..31], Im X[31..0]),
 * 64 new v values per slot, 32 output samples per slot through the decimated prototype (sbr_qmf_window_ds[n] =
 * us[2n]); history = 9 blocks of 64.  Same shared-memory rows as the 64-band bank, half of each row used. */
template <int NT>
HEAAC_DEV void heaac_synth_stream_ds(HeaacSynthSmem *sm, const float *HEAAC_RESTRICT X, HeaacSynthState *st, int nch,
                                     int nout, int out_slot0, int out_slot1, int16_t *pcm, float *pcm_f32, float scale,
                                     float bias, const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    constexpr int nt = NT;
    const int nthr = 32 * nch;
    for (int w = tid; w < nch * 576; w += nt) {
        int c = w / 576, r = w - c * 576;
        sm->ch[c][(32 + (r >> 6)) * HEAAC_V_STRIDE + (r & 63)] = st[c].v[r];
    }
    for (int t = tid; t < nthr; t += nt) {
        const int c = t >> 5, i = t & 31;
        const float *x0 = X + (size_t)c * HEAAC_X_FLOATS + i, *x1 = x0 + HEAAC_X_PLANE;
        float *v = sm->ch[c] + (31 - i) * HEAAC_V_STRIDE;
        heaac_imdct128_reg<HEAAC_MDCT_SYN>(
            [&](int j) { return j < 32 ? -HEAAC_LDX(x0 + j * HEAAC_XS) : HEAAC_LDX(x1 + (63 - j) * HEAAC_XS); },
            [&](int j, float b) { if (j & 1) v[(63 - j) >> 1] = b; else v[63 - ((62 - j) >> 1)] = -b; });
    }
    HEAAC_SYNC();
    const int scale_and_bias = scale != 1.0f || bias != 0.0f;
    const int bias_mode = bias != 0.0f;
    for (int t = tid; t < nthr; t += nt) {
        const int c = nch == 2 ? (t & 1) : 0, n = nch == 2 ? (t >> 1) : t;
        float w[10];
        HEAAC_UNROLL
        for (int j = 0; j < 10; j++) w[j] = T->qmf_us[2 * (32 * j + n)];
        const float *vb = sm->ch[c] + n;
        const int slot = c ? out_slot1 : out_slot0;
        for (int i = 0; i < 32; i++) {
            const float *v = vb + (31 - i) * HEAAC_V_STRIDE;
            float acc = v[0] * w[0] + 0.0f;
            HEAAC_UNROLL
            for (int j = 1; j < 10; j++)
                acc = v[j * HEAAC_V_STRIDE + ((j & 1) ? 32 : 0)] * w[j] + acc;
            if (scale_and_bias) acc = acc * scale + bias;
            const int o = i * 32 + n;
            if (pcm)     pcm[(size_t)o * nout + slot] = heaac_to_s16(acc, bias_mode);
            if (pcm_f32) pcm_f32[(size_t)slot * 1024 + o] = acc;
        }
    }
    for (int w = tid; w < nch * 576; w += nt) {            /* newest 9 blocks of 64 become the next frame's history */
        int c = w / 576, r = w - c * 576;
        st[c].v[r] = sm->ch[c][(r >> 6) * HEAAC_V_STRIDE + (r & 63)];
    }
    HEAAC_SYNC();
}

#endif /* HEAAC_KERNELS_CUH */
