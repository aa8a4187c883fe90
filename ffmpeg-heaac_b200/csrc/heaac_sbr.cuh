/*
 * heaac_sbr.cuh -- small helpers shared by the SBR routines (heaac_chain.cuh) and the Parametric Stereo / synthesis
 * routines (heaac_kernels.cuh).
 */
#ifndef HEAAC_SBR_CUH
#define HEAAC_SBR_CUH

#include "heaac_dev.cuh"
#include "heaac_regfft.cuh"

/* (row, lane) loop nests: rows are dealt to warps, the fastest index to lanes, so that everything that
 * depends only on the row is computed once per warp iteration.  With nt = 1 (host emulation) they
 * degenerate to plain nested loops. */
#define HEAAC_WARP(tid, nt)    ((nt) >= 32 ? (tid) >> 5 : 0)
#define HEAAC_NWARPS(nt)       ((nt) >= 32 ? (nt) >> 5 : 1)
#define HEAAC_LANE(tid, nt)    ((nt) >= 32 ? (tid) & 31 : 0)
#define HEAAC_LSTRIDE(nt)      ((nt) >= 32 ? 32 : 1)

/* exact 2^(k/2): exp2f of a multiple of 0.5 equals ldexp(1 or sqrt2, n) (checked against glibc) */
HEAAC_DEV float heaac_exp2_half(float x)
{
    float fl = floorf(x);
    float m = (x != fl) ? 1.41421356237309504880f : 1.0f;
    return ldexpf(m, (int)fl);
}

#endif
