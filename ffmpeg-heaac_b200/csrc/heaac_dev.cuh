/*
 * heaac_dev.cuh -- device-side data layout of the B200 HE-AAC synthesis back end.
 *
 * Everything a stream needs between frames lives in HBM in the structs below
 * (SURVEY 8a "state": the live parts of SingleChannelElement aac.h:204-217,
 * SBRData sbr.h:59-107 and PSContext aacps.h:64-74).  One CTA owns one struct for the
 * duration of a kernel, so each block is read and written with coalesced, vectorised
 * accesses.
 *
 * The CTA-level routines in heaac_kernels.cuh are written as strided loops over work
 * items separated by HEAAC_SYNC().  With HEAAC_HOST_EMU defined they compile as plain C++
 * (one "thread") -- that build exists ONLY for tests/emu (logic checks on machines without a
 * GPU); the product library never contains it and has no CPU path.
 */
#ifndef HEAAC_DEV_CUH
#define HEAAC_DEV_CUH

#include <stdint.h>
#include "../../include/heaac_b200.h"

#ifdef HEAAC_HOST_EMU
#define HEAAC_DEV   static inline
#define HEAAC_SYNC() ((void)0)
#define HEAAC_SYNCWARP() ((void)0)
#define HEAAC_SYNC_OR(x) ((x) != 0)
#define HEAAC_LDX(p) (*(p))
#define HEAAC_RESTRICT
#include <math.h>
#include <string.h>
static inline float heaac_fmul(float a, float b) { return a * b; }
static inline float heaac_fadd(float a, float b) { return a + b; }
static inline float heaac_fsub(float a, float b) { return a - b; }
static inline float heaac_fdiv(float a, float b) { return a / b; }
static inline float heaac_fsqrt(float a) { return sqrtf(a); }
static inline float heaac_rint(float a) { return rintf(a); }
#elif defined(__CUDACC__)
#define HEAAC_DEV   __device__ __forceinline__
#define HEAAC_SYNC() __syncthreads()
#define HEAAC_SYNCWARP() __syncwarp()
#define HEAAC_SYNC_OR(x) __syncthreads_or(x)      /* barrier + "does the predicate hold for any thread of the CTA" */
/* the inter-kernel X buffer is written by the kernel before: read it through L2, never through the non-coherent path */
#define HEAAC_LDX(p) __ldcg(p)
#define HEAAC_RESTRICT __restrict__
/* IEEE round-to-nearest, never contracted into FMA: used where the reference's result is
 * order-sensitive (aacsbr.c:1259 "does not seem numerically stable") */
__device__ __forceinline__ float heaac_fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float heaac_fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float heaac_fsub(float a, float b) { return __fsub_rn(a, b); }
/* (division and square root expand to 10-20 instructions plus a slow path each and are used at some thirty places of the
 * chain kernel, none of them hot: called, not inlined, they keep the kernel's instruction footprint down) */
static __device__ __noinline__ float heaac_fdiv(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __noinline__ float heaac_fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float heaac_rint(float a) { return rintf(a); }
#endif

/* ---- asynchronous global -> shared copies (LDGSTS): issue early, wait right before first use ---- */
#ifdef HEAAC_HOST_EMU
static inline void heaac_cp_async4(float *dst, const float *src) { *dst = *src; }
static inline void heaac_cp_async16(float *dst, const float *src) { memcpy(dst, src, 16); }
static inline void heaac_cp_async_commit() {}
template <int N> static inline void heaac_cp_async_wait() {}
#elif defined(__CUDACC__)
__device__ __forceinline__ void heaac_cp_async4(float *dst, const float *src)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void heaac_cp_async16(float *dst, const float *src)     /* both 16-byte aligned */
{
    unsigned s = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void heaac_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void heaac_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#endif

/* ---- read-only tables, built on the host (heaac_tables.cpp) and copied to HBM once ---- */
struct HeaacDevTables {
    float sine_1024[1024], sine_128[128], kbd_1024[1024], kbd_128[128];   /* aacdec.c:593-596 */
    /* MDCT twiddles (mdct.c:94-100): N=2048 and N=256 scale 1; N=128 scale 1/64 (synthesis)
     * and N=128 scale -2 (analysis) */
    float tcos_2048[512], tsin_2048[512];
    float tcos_256[64],   tsin_256[64];
    float tcos_syn[32],   tsin_syn[32];
    float tcos_ana[32],   tsin_ana[32];
    /* FFT twiddles e^{+2 pi i j/512}, j < 256 (smaller sizes use a stride) */
    float fft_wre[256], fft_wim[256];
    uint16_t bitrev512[512];
    uint8_t  bitrev32[32];
    float qmf_us[640];              /* synthesis prototype (aacsbrdata.h:270-353 + aacsbr.c:117-120) */
    float qmf_ds[320];              /* analysis prototype, us[2n] (aacsbr.c:122-123) */
    float noise[1024];              /* sbr_noise_table[512][2] */
    uint16_t swb_offset_1024[13][64];
    uint16_t swb_offset_128[13][32];
    uint8_t  num_swb_1024[13], num_swb_128[13], tns_max_bands_1024[13], tns_max_bands_128[13];
    /* Parametric Stereo (aacps_tablegen.h:87-209, aacpsdata.c:145-163) */
    float pd_re_smooth[512], pd_im_smooth[512];
    float HA[46][8][4], HB[46][8][4];
    float f20_0_8[8][7][2], f34_0_12[12][7][2], f34_1_8[8][7][2], f34_2_4[4][7][2];
    float Q_fract_allpass[2][50][3][2], phi_fract[2][50][2];
    float g1_Q2[8];
    int8_t k_to_i_20[71 + 1], k_to_i_34[91 + 1];
    /* inverse of k_to_i: the hybrid bands of parameter band i, in ascending order, are ps_inv_k[mode][ps_inv_off[mode][i] ..
     * ps_inv_off[mode][i + 1]) (mode 0: 20-band, 1: 34-band) */
    uint8_t ps_inv_off[2][36], ps_inv_k[2][92];
};

/* ---- per core channel state --------------------------------------------------------
 * Members that are moved as blocks start at multiples of 16 bytes and the struct size is a multiple of 16
 * (cp.async.bulk needs 16-byte aligned addresses and sizes). */
struct HeaacChanState {
    float saved[512];               /* overlap-add history, sce->saved[0..511] */
    /* SBR (only touched when cfg.sbr) */
    float ana_hist[288];            /* analysis_filterbank_samples[0..287] */
    float w_tail[2 * 32 * 8];       /* W[1][24..31] of the previous frame, [plane][band][slot - 24] */
    float y_tail[2 * 64 * 8];       /* Y[1][32..37], [plane][band][slot - 32] (6 of 8 slots used) */
    float g_hist[10 * 48];          /* g_temp rows 32..41 (the only rows ever read as history) */
    float q_hist[10 * 48];          /* q_temp rows 32..41 */
    float bw_array[8];
    int32_t prev_ws, prev_shape;    /* ics.window_sequence[1], ics.use_kb_window[1] */
    uint32_t f_indexnoise, f_indexsine;
    uint8_t s_index0[48];           /* s_indexmapped[0] */
};
static_assert(sizeof(HeaacChanState) % 16 == 0, "HeaacChanState must keep 16-byte granularity");

/* ---- per OUTPUT channel state (QMF synthesis history) ------------------------------- */
struct HeaacSynthState {
    float v[1152];                  /* synthesis_filterbank_samples, newest 128-block first */
};

/* ---- per PS stream state (aacps.h:64-74, live parts) -------------------------------- */
struct HeaacPsState {
    float in_hist[5][6][2];         /* in_buf[5][0..5] */
    float delay[91][14][2];         /* delay[k][0..13] after the frame shift */
    float ap_delay[50][3][5][2];    /* ap_delay[k][m][0..4] */
    float peak_decay_nrg[34], power_smooth[34], peak_decay_diff_smooth[34];
    float H[4][2][6][34];           /* H11,H12,H21,H22 [re/im][envelope row 0..5][band]: all rows, stale ones are re-read (aacps.c:816-823) */
    int8_t opd_hist[34], ipd_hist[34];
    uint32_t pad[3];
};

/* Inter-kernel QMF-domain buffer of one channel: X[plane re/im][band 0..63][slot], slot fastest so that
 * kernels whose lanes are time slots read and write it coalesced (the reference's sbr->X is
 * [2][38][64], sbr.h:162). */
#define HEAAC_XS        40                   /* slot stride (38 used) */
#define HEAAC_X_PLANE   (64 * HEAAC_XS)
#define HEAAC_X_FLOATS  (2 * HEAAC_X_PLANE)

/* compact per-launch geometry */
struct HeaacGeom {
    int n_streams, ncc, nout, nse, out_len;
    int ds;                          /* downsampled SBR synthesis (32 bands) */
    int sr_index, sbr, ps;
    float add_bias, sf_scale;
    int n_elems;
    uint8_t elem_type[HEAAC_MAX_ELEMS];
    uint8_t ch_elem[HEAAC_MAX_CHANNELS];     /* core channel -> element index */
    uint8_t ch_in_elem[HEAAC_MAX_CHANNELS];  /* core channel -> 0/1 inside its element */
    uint8_t ch_sbr_elem[HEAAC_MAX_CHANNELS]; /* core channel -> SBR element index (0xff for LFE) */
    uint8_t elem_is_cpe[HEAAC_MAX_CHANNELS]; /* per core channel: its element is a CPE */
    uint8_t out_slot[HEAAC_MAX_CHANNELS];    /* core (or synthesised) channel -> position in the output */
    uint8_t elem_first_out[HEAAC_MAX_ELEMS]; /* element -> its first output channel */
    uint8_t elem_nout[HEAAC_MAX_ELEMS];      /* element -> number of output channels (2 with PS) */
    int max_elem_ch;
    int n_cce;                               /* coupling channel elements: coded channels without an output channel */
    uint8_t cce_ch[HEAAC_MAX_ELEMS];         /* their core channel indices, in element order */
};

#endif
