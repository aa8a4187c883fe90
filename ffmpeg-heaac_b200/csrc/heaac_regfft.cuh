/*
 * heaac_regfft.cuh -- register-resident small transforms for the QMF banks.
 *
 * One thread computes one complete half-IMDCT of size N = 128 (ff_imdct_half_c, mdct.c:124-159, with the
 * N/4 = 32-point complex FFT of fft.c inside) entirely in registers: every loop below has compile-time
 * bounds and is fully unrolled, so all array indices and all twiddle factors (heaac_twiddles_gen.h) are
 * compile-time constants -- the twiddles become FFMA/FMUL immediates and no shared memory, shuffles or
 * barriers are needed inside a transform.  32 QMF time slots map to the 32 lanes of a warp.
 *
 * The arithmetic is the oracle's radix-2 decimation-in-time butterfly network; butterflies whose twiddle
 * is exactly 1 or i skip the multiplication (the reference's table holds cos(pi/2) = 6.1e-17 there, so the
 * result can differ from the table-driven form by one rounding in rare cases).
 */
#ifndef HEAAC_REGFFT_CUH
#define HEAAC_REGFFT_CUH

#include "heaac_dev.cuh"
#include "heaac_twiddles_gen.h"

#ifdef HEAAC_HOST_EMU
#define HEAAC_UNROLL
#define HEAAC_NOUNROLL
#define HEAAC_UNROLL4
#else
#define HEAAC_UNROLL _Pragma("unroll")
/* work-item loops stay rolled: the kernels are bound by instruction fetch, not by loop overhead */
#define HEAAC_NOUNROLL _Pragma("unroll 1")
#define HEAAC_UNROLL4 _Pragma("unroll 4")
#endif

HEAAC_HD constexpr int heaac_brev5(int k)
{
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

enum { HEAAC_MDCT_SYN = 0, HEAAC_MDCT_ANA = 1 };

template <int KIND> HEAAC_HD constexpr float heaac_tcos128(int k) { return KIND == HEAAC_MDCT_SYN ? heaac_tcos_syn(k) : heaac_tcos_ana(k); }
template <int KIND> HEAAC_HD constexpr float heaac_tsin128(int k) { return KIND == HEAAC_MDCT_SYN ? heaac_tsin_syn(k) : heaac_tsin_ana(k); }

/* one radix-2 DIT stage (butterfly span 2^(S-1)) of the 32-point FFT; all bounds are constant expressions */
template <int S>
HEAAC_DEV void heaac_fft32_stage(float (&re)[32], float (&im)[32])
{
    constexpr int half = 1 << (S - 1);
    HEAAC_UNROLL
    for (int g = 0; g < 32; g += 2 * half) {
        HEAAC_UNROLL
        for (int j = 0; j < half; j++) {
            const int tw = j * (32 >> S);                  /* twiddle exponent in units of 2 pi / 32 */
            const int a = g + j, b = a + half;
            float tr, ti;
            if (tw == 0)      { tr = re[b];  ti = im[b]; }
            else if (tw == 8) { tr = -im[b]; ti = re[b]; }
            else {
                const float wr = heaac_w32_re(tw), wi = heaac_w32_im(tw);
                tr = re[b] * wr - im[b] * wi;
                ti = re[b] * wi + im[b] * wr;
            }
            const float ar = re[a], ai = im[a];
            re[b] = ar - tr; im[b] = ai - ti;
            re[a] = ar + tr; im[a] = ai + ti;
        }
    }
}

/* in-place 32-point FFT, X[m] = sum_k x[k] e^{+2 pi i km/32}; input in bit-reversed order */
HEAAC_DEV void heaac_fft32_reg(float (&re)[32], float (&im)[32])
{
    heaac_fft32_stage<1>(re, im);
    heaac_fft32_stage<2>(re, im);
    heaac_fft32_stage<3>(re, im);
    heaac_fft32_stage<4>(re, im);
    heaac_fft32_stage<5>(re, im);
}

/* Half-IMDCT, N = 128: `ld(j)` returns input j (0..63), `st(j, v)` receives output j (0..63). */
template <int KIND, class Load, class Store>
HEAAC_DEV void heaac_imdct128_reg(Load ld, Store st)
{
    float zr[32], zi[32];
    HEAAC_UNROLL
    for (int k = 0; k < 32; k++) {                         /* pre-rotation into bit-reversed slots */
        const float a = ld(63 - 2 * k), c = ld(2 * k);
        const float tc = heaac_tcos128<KIND>(k), ts = heaac_tsin128<KIND>(k);
        zr[heaac_brev5(k)] = a * tc - c * ts;
        zi[heaac_brev5(k)] = a * ts + c * tc;
    }
    heaac_fft32_reg(zr, zi);
    HEAAC_UNROLL
    for (int k = 0; k < 16; k++) {                         /* post-rotation + reordering */
        const int p = 15 - k, q = 16 + k;
        const float tcp = heaac_tcos128<KIND>(p), tsp = heaac_tsin128<KIND>(p);
        const float tcq = heaac_tcos128<KIND>(q), tsq = heaac_tsin128<KIND>(q);
        const float r0 = zi[p] * tsp - zr[p] * tcp;
        const float i1 = zi[p] * tcp + zr[p] * tsp;
        const float r1 = zi[q] * tsq - zr[q] * tcq;
        const float i0 = zi[q] * tcq + zr[q] * tsq;
        st(2 * p, r0); st(2 * p + 1, i0);
        st(2 * q, r1); st(2 * q + 1, i1);
    }
}

/* ------------------------------------------------------------------------------------
 * The same half-IMDCT (N = 128) by FOUR cooperating threads of one warp (a "quad"): a quarter of the instruction stream
 * per thread -- the one-thread form is ~1300 straight-line instructions, which one warp per transform batch executes
 * alone at instruction-fetch speed while the rest of the CTA waits.  Same butterfly network, same operations:
 *
 *   part A, thread t: the 8 points at bit-reversed positions 8t .. 8t+7: pre-rotation, DIT stages 1-3 (butterfly spans
 *           1, 2, 4 stay inside the 8 points; their twiddles are the same for every t: compile-time) -> exchange buffer
 *   part B, thread t: the 8 points 4j + t, j = 0..7: stages 4-5 (spans 8 and 16 = j +- 2, j +- 4; twiddles depend on t and
 *           come from the FFT table), post-rotation -> 16 outputs
 *
 * The quads of a warp are laid out as lane = 8 t + q (q = 0..7: eight transforms per warp); the exchange buffer of such a
 * group holds element e of transform q at [(e >> 3) * 72 + (e & 7) * 8 + q] (per plane: HEAAC_QX_FLOATS floats), which is
 * free of bank conflicts for both access patterns.  Only the four threads of a quad exchange data, so a warp-level
 * barrier between A and B is enough.
 *
 * ld(hi, lo, odd): input hi + lo (hi: compile-time multiple of 8, lo = 0..7 at run time, odd = parity of the index);
 * st(hi, lo, v): output hi + lo.  tcos / tsin: the MDCT twiddles of the instance; wre / wim: e^{+2 pi i j / 32} at [16 j].
 * ---------------------------------------------------------------------------------- */
#define HEAAC_QX_FLOATS 288

HEAAC_HD constexpr int heaac_brev3(int k) { return ((k & 1) << 2) | (k & 2) | ((k & 4) >> 2); }

template <class Load>
HEAAC_DEV void heaac_imdct128_quad_a(Load ld, const float *HEAAC_RESTRICT tcos, const float *HEAAC_RESTRICT tsin, int t,
                                     float *xre, float *xim)
{
    const int bt = ((t & 1) << 1) | (t >> 1);              /* positions 8t + j hold k = brev5(8t + j) = 4 brev3(j) + brev2(t) */
    const float *tc = tcos + bt, *ts = tsin + bt;
    float zr[8], zi[8];
    HEAAC_UNROLL
    for (int j = 0; j < 8; j++) {
        const int bj = heaac_brev3(j);
        const float a = ld(56 - 8 * bj, 7 - 2 * bt, 1), c = ld(8 * bj, 2 * bt, 0);      /* in[63 - 2k], in[2k] */
        const float tcv = tc[4 * bj], tsv = ts[4 * bj];
        zr[j] = a * tcv - c * tsv;
        zi[j] = a * tsv + c * tcv;
    }
    HEAAC_UNROLL
    for (int s = 1; s <= 3; s++) {
        const int half = 1 << (s - 1);
        HEAAC_UNROLL
        for (int g = 0; g < 8; g += 2 * half) {
            HEAAC_UNROLL
            for (int j = 0; j < half; j++) {
                const int tw = j * (32 >> s);
                const int a = g + j, b = a + half;
                float tr, ti;
                if (tw == 0)      { tr = zr[b];  ti = zi[b]; }
                else if (tw == 8) { tr = -zi[b]; ti = zr[b]; }
                else {
                    const float wr = tw == 4 ? heaac_w32_re(4) : heaac_w32_re(12), wi = tw == 4 ? heaac_w32_im(4) : heaac_w32_im(12);
                    tr = zr[b] * wr - zi[b] * wi;
                    ti = zr[b] * wi + zi[b] * wr;
                }
                const float ar = zr[a], ai = zi[a];
                zr[b] = ar - tr; zi[b] = ai - ti;
                zr[a] = ar + tr; zi[a] = ai + ti;
            }
        }
    }
    HEAAC_UNROLL
    for (int j = 0; j < 8; j++) { xre[t * 72 + j * 8] = zr[j]; xim[t * 72 + j * 8] = zi[j]; }
}

template <class Store>
HEAAC_DEV void heaac_imdct128_quad_b(Store st, const float *HEAAC_RESTRICT tcos, const float *HEAAC_RESTRICT tsin,
                                     const float *HEAAC_RESTRICT wre, const float *HEAAC_RESTRICT wim, int t,
                                     const float *xre, const float *xim)
{
    float zr[8], zi[8];
    HEAAC_UNROLL
    for (int j = 0; j < 8; j++) {                          /* point 4j + t */
        zr[j] = xre[(j >> 1) * 72 + (4 * (j & 1) + t) * 8];
        zi[j] = xim[(j >> 1) * 72 + (4 * (j & 1) + t) * 8];
    }
    HEAAC_UNROLL
    for (int s = 4; s <= 5; s++) {
        const int hj = s == 4 ? 2 : 4;                     /* butterfly span in units of j */
        HEAAC_UNROLL
        for (int j = 0; j < 8; j++) {
            if (j & hj) continue;
            /* twiddle exponent of point i = 4j + t: (i mod 8) * 2 in stage 4, i mod 16 in stage 5 */
            const int tw = s == 4 ? (4 * (j & 1) + t) * 2 : 4 * (j & 3) + t;
            const float wr = wre[16 * tw], wi = wim[16 * tw];
            const int a = j, b = j + hj;
            const float tr = zr[b] * wr - zi[b] * wi;
            const float ti = zr[b] * wi + zi[b] * wr;
            const float ar = zr[a], ai = zi[a];
            zr[b] = ar - tr; zi[b] = ai - ti;
            zr[a] = ar + tr; zi[a] = ai + ti;
        }
    }
    const float *tc = tcos + t, *ts = tsin + t;
    HEAAC_UNROLL
    for (int j = 0; j < 8; j++) {                          /* out[2i] = zi ts - zr tc, out[63 - 2i] = zi tc + zr ts, i = 4j + t */
        const float tcv = tc[4 * j], tsv = ts[4 * j];
        st(8 * j, 2 * t, zi[j] * tsv - zr[j] * tcv);
        st(56 - 8 * j, 7 - 2 * t, zi[j] * tcv + zr[j] * tsv);
    }
}

#endif
