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
#else
#define HEAAC_UNROLL _Pragma("unroll")
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

#endif
