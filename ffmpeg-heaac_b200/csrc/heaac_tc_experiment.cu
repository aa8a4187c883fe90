/*
 * heaac_tc_experiment.cu -- EXPERIMENT, not on the product path: the transform part of the 64-band synthesis bank
 * (sbr_qmf_synthesis, aacsbr.c:1199-1213: two N = 128 half-IMDCTs + combine) as a GEMM on the 5th-generation tensor cores.
 *
 *   V[128 x N] = M[128 x 128] . X[128 x N]        N = 32 columns per channel-frame, M fixed (tools/tc_qmf_experiment.py)
 *
 * with the 3 x TF32 split the north star names:  D = A_hi B_hi + A_hi B_lo + A_lo B_hi  (fp32 accumulation in TMEM).
 * One CTA per SM, persistent over tiles of 64 columns:
 *   - A_hi / A_lo (2 x 64 KB) stay in shared memory for the life of the CTA, in the UMMA K-major no-swizzle layout (core
 *     matrices of 8 rows x 16 bytes, LBO = 128 B along K, SBO = 4096 B along M); the host passes them as ready-made images;
 *   - per tile the 128 threads read 64 columns of X, split every value into tf32 hi + tf32 lo and store both in the same layout
 *     (2 x 32 KB), fence.proxy.async, barrier;
 *   - one thread issues 3 x 16 tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 64, K = 8) and commits to an mbarrier;
 *   - all four warps read the accumulator back with tcgen05.ld (32 lanes x 32 bit per warp, 64 columns) and store V.
 * Nothing is overlapped (no double buffering of B or of the accumulator): the measured time is an upper bound of what a
 * tuned kernel would need, the tensor-pipe share is read from ncu.  heaac_b200_exp_tc_synth also times a plain copy kernel
 * with the same global traffic as the lower bound.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TC_TILE_N 64
#define TC_A_BYTES (128 * 128 * 4)
#define TC_B_BYTES (TC_TILE_N * 128 * 4)
#define TC_SMEM (2 * TC_A_BYTES + 2 * TC_B_BYTES + 64)

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor): start address >> 4 in bits [0,14), leading
 * byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version = 1 in [46,48), layout type 0 (no swizzle) in [61,64) */
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ float tc_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(128, 1)
k_tc_synth(const float4 *__restrict__ a_hi_img, const float4 *__restrict__ a_lo_img, const float *__restrict__ X, float *__restrict__ V, int ntiles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float4 *sAhi = reinterpret_cast<float4 *>(smem);
    float4 *sAlo = reinterpret_cast<float4 *>(smem + TC_A_BYTES);
    unsigned char *sBhi = smem + 2 * TC_A_BYTES, *sBlo = sBhi + TC_B_BYTES;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + 2 * TC_A_BYTES + 2 * TC_B_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < TC_A_BYTES / 16; i += 128) { sAhi[i] = a_hi_img[i]; sAlo[i] = a_lo_img[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                  /* one warp allocates 64 TMEM columns (128 lanes x 64 x fp32 accumulator) */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(tc_smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    /* instruction descriptor (UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
     * N >> 3 at bit 17, M >> 4 at bit 24 */
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_TILE_N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        /* ---- B operand: 64 columns x 128 values, split, into the K-major core-matrix layout ---- */
        const float4 *xt = reinterpret_cast<const float4 *>(X + (size_t)tile * TC_TILE_N * 128);
        for (int idx = tid; idx < TC_TILE_N * 32; idx += 128) {
            const int r = idx & 7, kc = (idx >> 3) & 31, jb = idx >> 8;
            const float4 v = xt[(jb * 8 + r) * 32 + kc];
            float4 h, l;
            h.x = tc_tf32(v.x); h.y = tc_tf32(v.y); h.z = tc_tf32(v.z); h.w = tc_tf32(v.w);
            l.x = tc_tf32(v.x - h.x); l.y = tc_tf32(v.y - h.y); l.z = tc_tf32(v.z - h.z); l.w = tc_tf32(v.w - h.w);
            const int off = jb * 4096 + kc * 128 + r * 16;
            *reinterpret_cast<float4 *>(sBhi + off) = h;
            *reinterpret_cast<float4 *>(sBlo + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     /* generic-proxy stores -> visible to the tensor core */
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t aH = tc_smem_u32(sAhi), aL = tc_smem_u32(sAlo), bH = tc_smem_u32(sBhi), bL = tc_smem_u32(sBlo);
            #pragma unroll 1
            for (int pass = 0; pass < 3; pass++) {
                const uint32_t a0 = pass == 2 ? aL : aH, b0 = pass == 1 ? bL : bH;
                #pragma unroll 1
                for (int ks = 0; ks < 16; ks++) {                     /* K = 8 per instruction: two core matrices = 256 B */
                    const uint64_t ad = tc_desc(a0 + ks * 256, 128, 4096), bd = tc_desc(b0 + ks * 256, 128, 4096);
                    const uint32_t acc = (pass | ks) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(mbar)) : "memory");
        }
        {   /* everyone waits for the accumulator */
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(tc_smem_u32(mbar)), "r"(phase) : "memory");
            } while (!ok);
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        /* ---- epilogue: thread = TMEM lane = v index n; 64 columns = the tile's 64 time slots ---- */
        float *vt = V + (size_t)tile * TC_TILE_N * 128 + tid;
        #pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + half * 32;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            #pragma unroll
            for (int j = 0; j < 32; j++) vt[(size_t)(half * 32 + j) * 128] = __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                              /* B and the accumulator are free for the next tile */
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem_d) : "memory");
}

/* the same global traffic without any arithmetic: the floor of this formulation */
__global__ void __launch_bounds__(128)
k_tc_copy(const float4 *__restrict__ X, float4 *__restrict__ V, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * 128 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 128) V[i] = X[i];
}

extern "C" int heaac_b200_exp_tc_synth(const float *a_hi_img, const float *a_lo_img, const float *X, float *V, int ncol, int iters,
                                       float *ms_tc, float *ms_copy)
{
    if (!a_hi_img || !a_lo_img || !X || !V || ncol <= 0 || ncol % TC_TILE_N || iters <= 0) return -22;
#define TCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "heaac_b200_exp_tc_synth: %s: %s\n", #call, cudaGetErrorString(e_)); return -5; } } while (0)
    float *dA, *dB, *dX, *dV;
    const size_t xb = (size_t)ncol * 128 * sizeof(float);
    TCK(cudaMalloc(&dA, TC_A_BYTES)); TCK(cudaMalloc(&dB, TC_A_BYTES)); TCK(cudaMalloc(&dX, xb)); TCK(cudaMalloc(&dV, xb));
    TCK(cudaMemcpy(dA, a_hi_img, TC_A_BYTES, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(dB, a_lo_img, TC_A_BYTES, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(dX, X, xb, cudaMemcpyHostToDevice));
    TCK(cudaMemset(dV, 0, xb));
    TCK(cudaFuncSetAttribute(k_tc_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    int dev = 0, sms = 0;
    TCK(cudaGetDevice(&dev));
    TCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ntiles = ncol / TC_TILE_N, grid = ntiles < sms ? ntiles : sms;
    cudaEvent_t e0, e1;
    TCK(cudaEventCreate(&e0)); TCK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) k_tc_synth<<<grid, 128, TC_SMEM>>>((const float4 *)dA, (const float4 *)dB, dX, dV, ntiles);
    TCK(cudaDeviceSynchronize());
    TCK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) k_tc_synth<<<grid, 128, TC_SMEM>>>((const float4 *)dA, (const float4 *)dB, dX, dV, ntiles);
    TCK(cudaEventRecord(e1));
    TCK(cudaEventSynchronize(e1));
    TCK(cudaEventElapsedTime(ms_tc, e0, e1));
    *ms_tc /= iters;
    TCK(cudaMemcpy(V, dV, xb, cudaMemcpyDeviceToHost));
    for (int w = 0; w < 3; w++) k_tc_copy<<<sms * 8, 128>>>((const float4 *)dX, (float4 *)dV, xb / 16);
    TCK(cudaDeviceSynchronize());
    TCK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) k_tc_copy<<<sms * 8, 128>>>((const float4 *)dX, (float4 *)dV, xb / 16);
    TCK(cudaEventRecord(e1));
    TCK(cudaEventSynchronize(e1));
    TCK(cudaEventElapsedTime(ms_copy, e0, e1));
    *ms_copy /= iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dX); cudaFree(dV);
#undef TCK
    return 0;
}
