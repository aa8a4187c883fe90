/*
 * heaac_b200_experimental.h -- entry points of EXPERIMENTS that ship in libheaac_b200.so but are not part of the decode path.
 */
#ifndef HEAAC_B200_EXPERIMENTAL_H
#define HEAAC_B200_EXPERIMENTAL_H
#ifdef __cplusplus
extern "C" {
#endif

/* The transform part of the 64-band synthesis QMF (sbr_qmf_synthesis, aacsbr.c:1199-1213) as a 3 x TF32 split GEMM on
 * tcgen05.mma (csrc/heaac_tc_experiment.cu; driver and error analysis: tools/tc_qmf_experiment.py).
 *   a_hi_img / a_lo_img : the 128 x 128 transform matrix split into tf32 hi / lo parts, each as the 64 KB UMMA K-major
 *                         no-swizzle shared-memory image (host pointers)
 *   X, V                : ncol columns of 128 floats each, host pointers (ncol a multiple of 64)
 *   ms_tc / ms_copy     : average kernel time over `iters` launches of the GEMM kernel / of a copy kernel with the same traffic
 * Returns 0, or a negative HEAAC_E_* code. */
int heaac_b200_exp_tc_synth(const float *a_hi_img, const float *a_lo_img, const float *X, float *V, int ncol, int iters,
                            float *ms_tc, float *ms_copy);

#ifdef __cplusplus
}
#endif
#endif
