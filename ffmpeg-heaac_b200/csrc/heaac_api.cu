/*
 * heaac_api.cu -- C-ABI (include/heaac_b200.h) and kernel launches of the B200 HE-AAC
 * synthesis back end.  sm_100a only; there is no CPU path: every entry point fails with
 * HEAAC_E_NODEV when no CUDA device is usable.
 *
 * Launches per batch:
 *   AAC-LC          per frame   k_tns (pre-pass, only with TNS) + k_core (IMDCT + window + overlap-add + s16)
 *   HE-AAC v1       per CALL    k_chain: core transform -> QMF analysis -> HF generation / adjustment -> QMF synthesis
 *                               -> s16 for every frame of the call, all intermediates in shared memory (heaac_chain.cuh)
 *   HE-AAC v2 / downsampled SBR   per frame   k_chain (stops at X) + k_ps (hybrid analysis .. synthesis) + k_synth
 * All per-stream state stays in HBM between launches (heaac_dev.cuh).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "heaac_chain.cuh"
#include "heaac_context.cuh"
#ifndef HEAAC_X_GROUP
#define HEAAC_X_GROUP 8            /* frames per chain launch on the paths that go through the inter-kernel X buffer */
#endif

extern "C" void heaac_build_tables(HeaacDevTables *t);

#define HEAAC_THREADS 256
#ifndef HEAAC_SBR_THREADS
#define HEAAC_SBR_THREADS 256
#endif
#ifndef HEAAC_CORE_MIN_BLOCKS
#define HEAAC_CORE_MIN_BLOCKS 4
#endif

static void prof_begin(HeaacContext *c, int kind)
{
    if (!c->prof_on || c->prof_n >= c->prof_cap) return;
    c->prof_kind[c->prof_n] = (unsigned char)kind;
    cudaEventRecord(c->prof_ev[2 * c->prof_n], c->stream);
}
static void prof_end(HeaacContext *c)
{
    if (!c->prof_on || c->prof_n >= c->prof_cap) return;
    cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], c->stream);
    c->prof_n++;
}

/* ------------------------------------------------------------------------------------ */
/* TNS pre-pass: one thread per (channel, chain slot).  The all-pole filters are serial in the coefficient index and
 * only 1..8 of them exist per channel, so inside k_core they would hold a whole CTA for the length of the slowest
 * chain; here every chain of the tick runs at once.  Reads the caller's coefficients, writes the filtered ranges to
 * `done` (same indexing); k_core overlays them. */
__global__ void __launch_bounds__(128)
k_tns(int n_channels, int sr_index, const float *__restrict__ coeffs, const HeaacIcs *__restrict__ ics,
      const HeaacTns *__restrict__ tns, float *done, const HeaacDevTables *__restrict__ T)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int gc = t >> 3, slot = t & 7;
    if (gc >= n_channels) return;
    const HeaacIcs ic = ics[gc];
    if (!ic.tns_present) return;
    int start, size, inc, order;
    float lpc[HEAAC_TNS_MAX_ORDER];
    if (!heaac_tns_chain(&ic, tns + gc, sr_index, T, slot, &start, &size, &inc, &order, lpc, 1)) return;
    const float *in = coeffs + (size_t)gc * 1024;
    float *out = done + (size_t)gc * 1024;
    if (order <= 4)       heaac_tns_run<4>(in, out, start, size, inc, order, lpc);
    else if (order <= 8)  heaac_tns_run<8>(in, out, start, size, inc, order, lpc);
    else if (order <= 12) heaac_tns_run<12>(in, out, start, size, inc, order, lpc);
    else                  heaac_tns_run<HEAAC_TNS_MAX_ORDER>(in, out, start, size, inc, order, lpc);
}

/* AAC-LC channel-frames (no SBR).  Persistent: the CTAs of the grid claim channels from a device counter (so CTAs that
 * are scheduled late, e.g. behind a concurrently running parse kernel, do not hold the kernel up) and stage the NEXT
 * claimed channel's coefficients into the other half of a double buffer with cp.async while working on the current one.
 * work[0] = next unclaimed channel, work[1] = CTAs finished; the last CTA out resets both for the next launch. */
__global__ void __launch_bounds__(HEAAC_THREADS, HEAAC_CORE_MIN_BLOCKS)
k_core(HeaacGeom g, int n_channels, int n_frames, const float *__restrict__ coeffs, const HeaacIcs *__restrict__ ics,
       const HeaacTns *__restrict__ tns, HeaacChanState *chan,
       int16_t *pcm, float *pcm_f32, const HeaacDevTables *__restrict__ T, unsigned *work,
       const float *__restrict__ tns_done)
{
    /* A CTA claims a channel and takes it through all n_frames frames of the call (frame-major arrays: frame f of channel gc
     * at index f * n_channels + gc); the coefficients of the next frame -- or of the next channel's first frame -- are fetched
     * while the current one is transformed. */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_claim;
    HeaacCoreSmem *sm = reinterpret_cast<HeaacCoreSmem *>(smem_raw);
    float *stage = reinterpret_cast<float *>(smem_raw + sizeof(HeaacCoreSmem));   /* [2][1024] */
    float *out = stage + 2048;                       /* [1024] */
    const size_t pcm_frame = (size_t)g.n_streams * 1024 * g.nout;
    if (threadIdx.x == 0) s_claim = (int)atomicAdd(&work[0], 1u);
    __syncthreads();
    int gc = s_claim;
    if (gc < n_channels)
        heaac_cp_async16(stage + 4 * threadIdx.x, coeffs + (size_t)gc * 1024 + 4 * threadIdx.x);
    heaac_cp_async_commit();
    __syncthreads();                                 /* s_claim has been read by everyone */
    for (int it = 0; gc < n_channels; ) {
        const int s = gc / g.ncc, c = gc - s * g.ncc;
        int gn = gc;
        for (int f = 0; f < n_frames; f++, it++) {
            const size_t fc = (size_t)f * n_channels + gc;
            const bool last = f == n_frames - 1;
            float *cur = stage + (it & 1) * 1024, *nxt = stage + ((it + 1) & 1) * 1024;
            if (last && threadIdx.x == 0) s_claim = (int)atomicAdd(&work[0], 1u);
            heaac_cp_async_wait<0>();                /* this frame's coefficients have landed */
            __syncthreads();
            if (last) gn = s_claim;
            if (!last || gn < n_channels) {
                const size_t nc = last ? (size_t)gn : fc + n_channels;
                heaac_cp_async16(nxt + 4 * threadIdx.x, coeffs + nc * 1024 + 4 * threadIdx.x);
            }
            heaac_cp_async_commit();
            heaac_core_channel<HEAAC_THREADS>(sm, cur, ics + fc, tns ? tns + fc : nullptr,
                                              tns_done ? tns_done + fc * 1024 : nullptr,
                                              chan + gc, out, g.add_bias, g.sr_index, T, threadIdx.x);
            const int bias_mode = g.add_bias != 0.0f;
            if (pcm) {
                int16_t *dst = pcm + (size_t)f * pcm_frame + (size_t)s * 1024 * g.nout + g.out_slot[c];
                for (int i = threadIdx.x; i < 1024; i += HEAAC_THREADS) dst[(size_t)i * g.nout] = heaac_to_s16(out[i], bias_mode);
            }
            if (pcm_f32) {
                float *dst = pcm_f32 + (size_t)f * pcm_frame + ((size_t)s * g.nout + g.out_slot[c]) * 1024;
                for (int i = threadIdx.x; i < 1024; i += HEAAC_THREADS) dst[i] = out[i];
            }
            __syncthreads();                         /* out / stage / s_claim are reused by the next frame */
        }
        gc = gn;
    }
    if (threadIdx.x == 0 && atomicAdd(&work[1], 1u) == gridDim.x - 1) { work[0] = 0; work[1] = 0; }
}

/* ------------------------------------------------------------------------------------
 * k_chain: the whole SBR channel pipeline of a batch in ONE launch.  One CTA owns one coded channel and walks through the
 * frames of the call: core transform (IMDCT + window + overlap-add) -> 32-band QMF analysis -> HF generation / adjustment
 * -> 64-band QMF synthesis -> s16, with every intermediate (the time samples, W, X_low, X_high, Y, X, the v buffer) in
 * shared memory (heaac_chain.cuh): HBM sees the frame's coefficients and side info, the channel's state and the PCM,
 * nothing else.  The coefficients of frame f + 1 are fetched with a bulk asynchronous copy (cp.async.bulk + mbarrier,
 * issued by one thread) while frame f is being synthesised; the synthesis history stays in shared memory between frames.
 * emit = 1 (Parametric Stereo, downsampled SBR): the chain stops at X, which goes to the inter-kernel buffer;
 * emit = 2 (keep_intermediates): X is written out as well, and the chain goes on.
 * `time_in` (ff_sbr_apply shim): the element's time samples are given, there is no core transform.
 * ---------------------------------------------------------------------------------- */
#ifndef HEAAC_CHAIN_THREADS
#define HEAAC_CHAIN_THREADS 128              /* threads per channel */
#endif
#ifndef HEAAC_CHAIN_CH
#define HEAAC_CHAIN_CH 1                     /* channels per CTA (measured: 1 and 2 give the same throughput) */
#endif
#ifndef HEAAC_CHAIN_MIN_BLOCKS
#define HEAAC_CHAIN_MIN_BLOCKS 5             /* CTAs per SM (shared memory allows 5 channels) */
#endif

/* CH channels per CTA, HEAAC_CHAIN_THREADS threads each, in lockstep (every barrier is CTA-wide and reached by all threads
 * whatever the channels' side info says): the kernel is bound by instruction fetch -- a CTA runs through nearly all of its code
 * once per frame -- and the channels of a CTA share every fetched line.  `first_ch`: global index of the grid's first channel. */
template <int CH>
__global__ void __launch_bounds__(HEAAC_CHAIN_THREADS * CH, CH == HEAAC_CHAIN_CH ? HEAAC_CHAIN_MIN_BLOCKS : 4)
k_chain(HeaacGeom g, int first_ch, int n_frames, int emit, const float *__restrict__ coeffs, const float *__restrict__ time_in,
        const HeaacIcs *__restrict__ ics, const HeaacTns *__restrict__ tns, const float *__restrict__ tns_done,
        const HeaacSbrElem *__restrict__ sbr, HeaacChanState *chan, HeaacSynthState *synth, float *X,
        int16_t *pcm, float *pcm_f32, const HeaacDevTables *__restrict__ T)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int half = threadIdx.x / HEAAC_CHAIN_THREADS, tid = threadIdx.x - half * HEAAC_CHAIN_THREADS;
    HeaacChainSmem *sm = reinterpret_cast<HeaacChainSmem *>(smem_raw) + half;
    const int gc = first_ch + blockIdx.x * CH + half;
    const int s = gc / g.ncc, c = gc - s * g.ncc;
    const int se = g.ch_sbr_elem[c];
    HeaacChainArgs a;
    a.n_frames = n_frames; a.emit = emit == 3 ? 1 : emit;      /* 3 = 1 with one X per frame of the launch */
    a.x_frame = emit == 3 ? (size_t)g.n_streams * g.nout * HEAAC_X_FLOATS : 0;
    a.frame_ch = (size_t)g.n_streams * g.ncc;
    a.frame_elems = (size_t)g.n_streams * g.nse;
    a.pcm_frame = (size_t)g.n_streams * 2048 * g.nout;
    a.coeffs = coeffs ? coeffs + (size_t)gc * 1024 : nullptr;
    a.time_in = time_in ? time_in + (size_t)gc * 1024 : nullptr;
    a.ics = ics ? ics + gc : nullptr;
    a.tns = tns ? tns + gc : nullptr;
    a.tns_done = tns_done ? tns_done + (size_t)gc * 1024 : nullptr;
    a.sbr = se == 0xff ? nullptr : sbr + (size_t)s * g.nse + se;      /* an LFE carries no SBR data: upsampled only */
    a.st = chan + gc;
    a.sy = synth + (size_t)s * g.nout + c;
    a.X = X ? X + ((size_t)s * g.nout + c) * HEAAC_X_FLOATS : nullptr;
    a.pcm = pcm ? pcm + (size_t)s * 2048 * g.nout : nullptr;
    a.pcm_f32 = pcm_f32 ? pcm_f32 + (size_t)s * 2048 * g.nout : nullptr;
    a.ch_in_elem = g.ch_in_elem[c]; a.is_cpe = g.elem_is_cpe[c]; a.nout = g.nout; a.out_slot = g.out_slot[c]; a.sr_index = g.sr_index;
    a.qmf_scale = -1024.0f * g.sf_scale; a.add_bias = g.add_bias;
    heaac_chain_channel<HEAAC_CHAIN_THREADS>(sm, a, T, tid);
}

/* all n_ch channels (every channel of an SBR configuration carries SBR): CTAs of HEAAC_CHAIN_CH channels, then the rest one by one */
static void launch_chain(HeaacContext *c, unsigned n_ch, int n_frames, int emit, const float *coeffs, const float *time_in,
                         const HeaacIcs *ics, const HeaacTns *tns, const float *tns_done, const HeaacSbrElem *sbr, HeaacChanState *chan,
                         HeaacSynthState *synth, float *X, int16_t *pcm, float *pcm_f32)
{
    const HeaacGeom g = c->geom;
    const unsigned groups = n_ch / HEAAC_CHAIN_CH, rest = n_ch - groups * HEAAC_CHAIN_CH;
    if (groups)
        k_chain<HEAAC_CHAIN_CH><<<groups, HEAAC_CHAIN_THREADS * HEAAC_CHAIN_CH, c->smem_chain * HEAAC_CHAIN_CH, c->stream>>>(
            g, 0, n_frames, emit, coeffs, time_in, ics, tns, tns_done, sbr, chan, synth, X, pcm, pcm_f32, c->d_tables);
    if (rest)
        k_chain<1><<<rest, HEAAC_CHAIN_THREADS, c->smem_chain, c->stream>>>(
            g, (int)(groups * HEAAC_CHAIN_CH), n_frames, emit, coeffs, time_in, ics, tns, tns_done, sbr, chan, synth, X, pcm, pcm_f32, c->d_tables);
}

__global__ void __launch_bounds__(HEAAC_THREADS)
k_ps(HeaacGeom g, const HeaacPsFrame *__restrict__ ps, const HeaacSbrElem *__restrict__ sbr,
     HeaacPsState *pstate, float *X, const HeaacDevTables *__restrict__ T)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeaacPsSmem *sm = reinterpret_cast<HeaacPsSmem *>(smem_raw);
    const int s = blockIdx.x;
    float *XL = X + (size_t)s * 2 * HEAAC_X_FLOATS, *XR = XL + HEAAC_X_FLOATS;
    const HeaacSbrElem *e = sbr + (size_t)s * g.nse;
    heaac_ps_stream<HEAAC_THREADS>(sm, ps + s, pstate + s, XL, XR, (int)e->kx[1] + (int)e->m[1], T, threadIdx.x);
}

#ifndef HEAAC_SYNTH_MIN_BLOCKS
#define HEAAC_SYNTH_MIN_BLOCKS 4
#endif
/* QMF synthesis from the inter-kernel X buffer: the two output channels of a Parametric Stereo stream, and downsampled SBR */
__global__ void __launch_bounds__(128, HEAAC_SYNTH_MIN_BLOCKS)
k_synth(HeaacGeom g, const float *X, HeaacSynthState *synth,
        int16_t *pcm, float *pcm_f32, const HeaacDevTables *__restrict__ T)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeaacSynthSmem *sm = reinterpret_cast<HeaacSynthSmem *>(smem_raw);
    /* one CTA per (stream, element): its 1 or 2 output channels */
    const int s = blockIdx.x / g.n_elems, e = blockIdx.x - s * g.n_elems;
    const int c0 = g.elem_first_out[e], nch = g.elem_nout[e];
    const float qmf_scale = -1024.0f * g.sf_scale;
    const size_t go = (size_t)s * g.nout + c0;
    if (g.ds)
        heaac_synth_stream_ds<128>(sm, X + go * HEAAC_X_FLOATS, synth + go, nch, g.nout, g.out_slot[c0], g.out_slot[c0 + (nch > 1)],
                                   pcm ? pcm + (size_t)s * 1024 * g.nout : nullptr,
                                   pcm_f32 ? pcm_f32 + (size_t)s * g.nout * 1024 : nullptr,
                                   qmf_scale, g.add_bias, T, threadIdx.x);
    else
        heaac_synth_stream<128>(sm, X + go * HEAAC_X_FLOATS, synth + go, nch, g.nout, g.out_slot[c0], g.out_slot[c0 + (nch > 1)],
                                pcm ? pcm + (size_t)s * 2048 * g.nout : nullptr,
                                pcm_f32 ? pcm_f32 + (size_t)s * g.nout * 2048 : nullptr,
                                qmf_scale, g.add_bias, T, threadIdx.x);
}

/* independent channel coupling + s16 epilogue, one CTA per stream (layouts with coupling channel elements only) */
__global__ void __launch_bounds__(HEAAC_THREADS)
k_couple(HeaacGeom g, const float *__restrict__ f32, const HeaacCce *__restrict__ cce, int16_t *pcm, float *pcm_f32)
{
    const int s = blockIdx.x;
    heaac_couple_stream<HEAAC_THREADS>(&g, f32 + (size_t)s * g.ncc * 1024, cce + (size_t)s * g.n_cce,
                                       pcm ? pcm + (size_t)s * 1024 * g.nout : nullptr,
                                       pcm_f32 ? pcm_f32 + (size_t)s * g.nout * 1024 : nullptr, threadIdx.x);
}

/* ------------------------------------------------------------------------------------ */
int heaac_cuda_err(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    fprintf(stderr, "heaac_b200: %s: %s\n", what, cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation) return HEAAC_E_NOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return HEAAC_E_NODEV;
    return HEAAC_E_LAUNCH;
}

static void build_geom(const HeaacConfig *cfg, HeaacGeom *g)
{
    memset(g, 0, sizeof(*g));
    g->n_streams = cfg->n_streams;
    g->n_elems = cfg->n_elems;
    int ch = 0, se = 0;
    for (int e = 0; e < cfg->n_elems; e++) {
        g->elem_type[e] = cfg->elem_type[e];
        int nch = cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
        for (int c = 0; c < nch; c++, ch++) {
            g->ch_elem[ch] = (uint8_t)e;
            g->ch_in_elem[ch] = (uint8_t)c;
            g->ch_sbr_elem[ch] = cfg->elem_type[e] == HEAAC_TYPE_LFE ? 0xff : (uint8_t)se;
            g->elem_is_cpe[ch] = nch == 2;
        }
        if (cfg->elem_type[e] != HEAAC_TYPE_LFE) se++;
    }
    /* k_synth works per element (an LFE included: it is upsampled like the others): its first output channel and channel count
     * (PS: mono in, 2 out) */
    {
        int o = 0;
        g->max_elem_ch = 1;
        for (int e = 0; e < cfg->n_elems; e++) {
            int nch = cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
            g->elem_first_out[e] = (uint8_t)o;
            g->elem_nout[e] = (uint8_t)(cfg->ps ? 2 : nch);
            if (nch > g->max_elem_ch) g->max_elem_ch = nch;
            o += nch;
        }
    }
    g->ncc = ch;
    g->nse = se;
    for (int e = 0, c0 = 0; e < cfg->n_elems; c0 += cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1, e++)
        if (cfg->elem_type[e] == HEAAC_TYPE_CCE) g->cce_ch[g->n_cce++] = (uint8_t)c0;
    g->nout = cfg->ps ? 2 : ch - g->n_cce;
    g->ds = cfg->sbr && cfg->sbr_downsampled;
    g->out_len = cfg->sbr && !g->ds ? 2048 : 1024;
    g->sr_index = cfg->sr_index;
    g->sbr = cfg->sbr != 0;
    g->ps = cfg->ps != 0;
    g->add_bias = cfg->add_bias;
    g->sf_scale = cfg->sf_scale;
    {
        int any = 0;
        for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) any |= cfg->out_slot[i];
        for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) g->out_slot[i] = (any && !cfg->ps) ? cfg->out_slot[i] : (uint8_t)i;
        if (g->n_cce) {                              /* coupling channels have no output; the others keep their order */
            int o = 0;
            for (int i = 0; i < ch; i++) {
                const bool is_cce = cfg->elem_type[g->ch_elem[i]] == HEAAC_TYPE_CCE;
                g->out_slot[i] = is_cce ? 0xff : (any ? cfg->out_slot[i] : (uint8_t)o);
                o += !is_cce;
            }
        }
    }
}

static int init_state(HeaacContext *c)
{
    const HeaacGeom *g = &c->geom;
    size_t nch = (size_t)g->n_streams * g->ncc;
    CK(cudaMemsetAsync(c->d_chan, 0, nch * sizeof(HeaacChanState), c->stream));
    if (g->sbr) {
        CK(cudaMemsetAsync(c->d_synth, 0, (size_t)g->n_streams * g->nout * sizeof(HeaacSynthState), c->stream));
        if (c->d_X) CK(cudaMemsetAsync(c->d_X, 0, c->x_frames * (size_t)g->n_streams * g->nout * HEAAC_X_FLOATS * sizeof(float), c->stream));
    }
    if (g->ps) CK(cudaMemsetAsync(c->d_ps, 0, (size_t)g->n_streams * sizeof(HeaacPsState), c->stream));
    /* everything the reference initialises to non-zero (kx' = 32, e_a[1] = -1; aacsbr.c:128-137) lives in
     * the parser-owned parameter structs, so the all-zero block IS the post-init state. */
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int heaac_b200_create(const HeaacConfig *cfg, HeaacContext **out)
{
    if (!cfg || !out) return HEAAC_E_INVAL;
    *out = nullptr;
    if (cfg->abi_version != HEAAC_ABI_VERSION) return HEAAC_E_INVAL;
    if (cfg->n_streams <= 0 || cfg->n_elems <= 0 || cfg->n_elems > HEAAC_MAX_ELEMS) return HEAAC_E_INVAL;
    if (cfg->sr_index < 0 || cfg->sr_index > 12) return HEAAC_E_INVAL;
    if (cfg->max_frames_per_call <= 0) return HEAAC_E_INVAL;
    int nch = 0, has_lfe = 0, n_cce = 0;
    for (int e = 0; e < cfg->n_elems; e++) {
        int t = cfg->elem_type[e];
        if (t != HEAAC_TYPE_SCE && t != HEAAC_TYPE_CPE && t != HEAAC_TYPE_LFE && t != HEAAC_TYPE_CCE) return HEAAC_E_INVAL;
        nch += t == HEAAC_TYPE_CPE ? 2 : 1;
        has_lfe |= t == HEAAC_TYPE_LFE;
        n_cce += t == HEAAC_TYPE_CCE;
    }
    (void)has_lfe;
    /* coupling channel elements: independent coupling of AAC-LC layouts (with SBR the coupled signal would have to go through
     * ff_sbr_apply as well, aacdec.c:1923-1928: not built) */
    if (n_cce && (cfg->sbr || cfg->ps)) return HEAAC_E_UNSUPPORTED;
    if (nch > HEAAC_MAX_CHANNELS) return HEAAC_E_INVAL;
    if (cfg->ps && !(cfg->sbr && cfg->n_elems == 1 && cfg->elem_type[0] == HEAAC_TYPE_SCE)) return HEAAC_E_INVAL;
    {
        int any = 0, seen = 0;
        for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) any |= cfg->out_slot[i];
        if (any && !cfg->ps) {
            for (int i = 0, e = 0, left = 0; i < nch; i++) {
                if (!left) left = cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
                const bool is_cce = cfg->elem_type[e] == HEAAC_TYPE_CCE;
                if (!--left) e++;
                if (is_cce) continue;
                if (cfg->out_slot[i] >= nch - n_cce || (seen >> cfg->out_slot[i]) & 1) return HEAAC_E_INVAL;
                seen |= 1 << cfg->out_slot[i];
            }
        }
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        fprintf(stderr, "heaac_b200: no usable CUDA device (this library has no CPU path)\n");
        return HEAAC_E_NODEV;
    }
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) {
        fprintf(stderr, "heaac_b200: device %d is sm_%d%d; this build targets sm_100a only\n", cfg->device, prop.major, prop.minor);
        return HEAAC_E_NODEV;
    }

    HeaacContext *c = new (std::nothrow) HeaacContext();
    if (!c) return HEAAC_E_NOMEM;
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    build_geom(cfg, &c->geom);
    const HeaacGeom *g = &c->geom;
    int r = 0;
#define CKX(call) do { r = heaac_cuda_err((call), #call); if (r) { heaac_b200_destroy(c); return r; } } while (0)
    CKX(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKX(cudaEventCreate(&c->ev0));
    CKX(cudaEventCreate(&c->ev1));
    {
        HeaacDevTables *h = (HeaacDevTables *)malloc(sizeof(HeaacDevTables));
        if (!h) { heaac_b200_destroy(c); return HEAAC_E_NOMEM; }
        heaac_build_tables(h);
        r = heaac_cuda_err(cudaMalloc(&c->d_tables, sizeof(HeaacDevTables)), "cudaMalloc tables");
        if (!r) r = heaac_cuda_err(cudaMemcpy(c->d_tables, h, sizeof(HeaacDevTables), cudaMemcpyHostToDevice), "upload tables");
        free(h);
        if (r) { heaac_b200_destroy(c); return r; }
    }
    size_t S = g->n_streams, T = cfg->max_frames_per_call;
    CKX(cudaMalloc(&c->d_chan, S * g->ncc * sizeof(HeaacChanState)));
    if (!g->sbr) {                                   /* AAC-LC: one frame tick of TNS scratch; with SBR it is sized on first use */
        CKX(cudaMalloc(&c->d_tns_done, S * g->ncc * 1024 * sizeof(float)));
        c->tns_done_frames = 1;
    }
    CKX(cudaMalloc(&c->d_work, 2 * sizeof(unsigned)));
    CKX(cudaMemset(c->d_work, 0, 2 * sizeof(unsigned)));
    if (g->sbr) {
        CKX(cudaMalloc(&c->d_synth, S * g->nout * sizeof(HeaacSynthState)));
        if (g->ps || g->ds || cfg->keep_intermediates) { CKX(cudaMalloc(&c->d_X, S * g->nout * HEAAC_X_FLOATS * sizeof(float))); c->x_frames = 1; }
    }
    if (g->ps) CKX(cudaMalloc(&c->d_ps, S * sizeof(HeaacPsState)));
    if (g->n_cce) CKX(cudaMalloc(&c->d_couple, S * g->ncc * 1024 * sizeof(float)));
    c->chunk = (int)((T + 3) / 4);
    if (c->chunk > 8) c->chunk = 8;
    if (cfg->chunk_frames > 0) c->chunk = cfg->chunk_frames;
    if (c->chunk > 32) c->chunk = 32;
    if ((size_t)c->chunk > T) c->chunk = (int)T;
    if (c->chunk < 1) c->chunk = 1;
    {
        const size_t Tc = c->chunk;
        CKX(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        CKX(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 8; i++) CKX(cudaEventCreateWithFlags(&c->ev_ticket[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            CKX(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CKX(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            CKX(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
            CKX(cudaMalloc(&c->s_coeffs[i], Tc * S * g->ncc * 1024 * sizeof(float)));
            CKX(cudaMalloc(&c->s_ics[i], Tc * S * g->ncc * sizeof(HeaacIcs)));
            CKX(cudaMalloc(&c->s_tns[i], Tc * S * g->ncc * sizeof(HeaacTns)));
            if (g->sbr) CKX(cudaMalloc(&c->s_sbr[i], Tc * S * g->nse * sizeof(HeaacSbrElem)));
            if (g->ps)  CKX(cudaMalloc(&c->s_ps[i], Tc * S * sizeof(HeaacPsFrame)));
            if (g->n_cce) CKX(cudaMalloc(&c->s_cce[i], Tc * S * g->n_cce * sizeof(HeaacCce)));
            CKX(cudaMalloc(&c->s_pcm[i], Tc * S * g->out_len * g->nout * sizeof(int16_t)));
            if (cfg->want_float_out) CKX(cudaMalloc(&c->s_pcm_f32[i], Tc * S * g->out_len * g->nout * sizeof(float)));
        }
    }

    c->smem_core = sizeof(HeaacCoreSmem) + (2048 + 1024) * sizeof(float);
    c->smem_chain = sizeof(HeaacChainSmem);
    c->smem_ps = sizeof(HeaacPsSmem);
    c->smem_synth = sizeof(HeaacSynthSmem);
    CKX(cudaFuncSetAttribute(k_core, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_core));
    CKX(cudaFuncSetAttribute(k_chain<HEAAC_CHAIN_CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_chain * HEAAC_CHAIN_CH));
    CKX(cudaFuncSetAttribute(k_chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_chain));
    CKX(cudaFuncSetAttribute(k_ps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_ps));
    CKX(cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_synth));
    {
        int per_sm = 0;
        CKX(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_core, HEAAC_THREADS, c->smem_core));
        c->core_grid = prop.multiProcessorCount * (per_sm > 0 ? per_sm : 1);
    }
    r = init_state(c);
    if (r) { heaac_b200_destroy(c); return r; }
#undef CKX
    *out = c;
    return 0;
}

extern "C" void heaac_b200_destroy(HeaacContext *c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tables); cudaFree(c->d_chan); cudaFree(c->d_synth); cudaFree(c->d_ps);
    cudaFree(c->d_X); cudaFree(c->d_work); cudaFree(c->d_tns_done); cudaFree(c->d_couple);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->s_coeffs[i]); cudaFree(c->s_ics[i]); cudaFree(c->s_tns[i]); cudaFree(c->s_sbr[i]); cudaFree(c->s_ps[i]); cudaFree(c->s_cce[i]);
        cudaFree(c->s_pcm[i]); cudaFree(c->s_pcm_f32[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    for (int i = 0; i < 8; i++) if (c->ev_ticket[i]) cudaEventDestroy(c->ev_ticket[i]);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (int i = 0; i < 2 * c->prof_cap; i++) cudaEventDestroy(c->prof_ev[i]);
    free(c->prof_ev); free(c->prof_kind);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int heaac_b200_reset_state(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    return init_state(c);
}

extern "C" int heaac_b200_n_core_channels(const HeaacContext *c) { return c ? c->geom.ncc : HEAAC_E_INVAL; }
extern "C" int heaac_b200_n_out_channels(const HeaacContext *c)  { return c ? c->geom.nout : HEAAC_E_INVAL; }
extern "C" int heaac_b200_frame_out_len(const HeaacContext *c)   { return c ? c->geom.out_len : HEAAC_E_INVAL; }
extern "C" int heaac_b200_n_sbr_elems(const HeaacContext *c)     { return c ? c->geom.nse : HEAAC_E_INVAL; }
extern "C" long long heaac_b200_launch_count(const HeaacContext *c) { return c ? c->launches : 0; }
extern "C" uintptr_t heaac_b200_stream(const HeaacContext *c) { return c ? (uintptr_t)c->stream : 0; }

/* launches for frames [0, n) of a device-resident batch on c->stream */
int heaac_launch_frames(HeaacContext *c, int n_frames, const float *coeffs, const HeaacIcs *ics, const HeaacTns *tns,
                         const HeaacSbrElem *sbr, const HeaacPsFrame *ps, int16_t *pcm, float *pcm_f32, const HeaacCce *cce)
{
    const HeaacGeom g = c->geom;
    const size_t S = g.n_streams;
    const size_t pcm_frame = S * g.out_len * g.nout;
    const unsigned n_ch = (unsigned)(S * g.ncc);
    if (g.sbr) {
        const float *tns_done = nullptr;
        if (tns) {
            /* TNS pre-pass for every frame of the call at once (the filters of different frames are independent) */
            if (c->tns_done_frames < (size_t)n_frames) {
                CK(cudaStreamSynchronize(c->stream));
                cudaFree(c->d_tns_done); c->d_tns_done = nullptr; c->tns_done_frames = 0;
                CK(cudaMalloc(&c->d_tns_done, (size_t)n_frames * n_ch * 1024 * sizeof(float)));
                c->tns_done_frames = n_frames;
            }
            const size_t n_all = (size_t)n_frames * n_ch;
            k_tns<<<(unsigned)((n_all * 8 + 127) / 128), 128, 0, c->stream>>>((int)n_all, g.sr_index, coeffs, ics, tns, c->d_tns_done, c->d_tables);
            c->launches++;
            tns_done = c->d_tns_done;
        }
        const int with_x = c->cfg.keep_intermediates != 0;
        if (!g.ps && !g.ds) {                        /* HE-AAC v1: the whole call is one launch */
            prof_begin(c, HEAAC_K_CHAIN);
            launch_chain(c, n_ch, n_frames, with_x ? 2 : 0, coeffs, nullptr, ics, tns, tns_done, sbr, c->d_chan, c->d_synth, c->d_X, pcm, pcm_f32);
            prof_end(c);
            c->launches++;
            return heaac_cuda_err(cudaGetLastError(), "kernel launch");
        }
        /* Parametric Stereo / downsampled SBR: X goes through the inter-kernel buffer.  The chain runs for a group of frames per
         * launch (channel state and history stay in shared memory, the next frame's inputs are fetched under the current one) and
         * leaves one X per frame; Parametric Stereo and the synthesis bank follow frame by frame */
        const size_t x_frame = S * g.nout * HEAAC_X_FLOATS;
        const int group = n_frames < HEAAC_X_GROUP ? n_frames : HEAAC_X_GROUP;
        if (c->x_frames < (size_t)group) {
            CK(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_X); c->d_X = nullptr; c->x_frames = 0;
            CK(cudaMalloc(&c->d_X, (size_t)group * x_frame * sizeof(float)));
            CK(cudaMemsetAsync(c->d_X, 0, (size_t)group * x_frame * sizeof(float), c->stream));
            c->x_frames = group;
        }
        for (int f0 = 0; f0 < n_frames; f0 += group) {
            const int nf = n_frames - f0 < group ? n_frames - f0 : group;
            const size_t fo = (size_t)f0 * n_ch;
            prof_begin(c, HEAAC_K_CHAIN);
            launch_chain(c, n_ch, nf, 3, coeffs + fo * 1024, nullptr, ics + fo, tns ? tns + fo : nullptr, tns_done ? tns_done + fo * 1024 : nullptr,
                         sbr + (size_t)f0 * S * g.nse, c->d_chan, c->d_synth, c->d_X, nullptr, nullptr);
            prof_end(c);
            c->launches++;
            for (int f = f0; f < f0 + nf; f++) {
                const HeaacSbrElem *sb = sbr + (size_t)f * S * g.nse;
                float *X = c->d_X + (size_t)(f - f0) * x_frame;
                int16_t *pc = pcm ? pcm + (size_t)f * pcm_frame : nullptr;
                float *pf = pcm_f32 ? pcm_f32 + (size_t)f * pcm_frame : nullptr;
                if (g.ps) {
                    prof_begin(c, HEAAC_K_PS);
                    k_ps<<<(unsigned)S, HEAAC_THREADS, c->smem_ps, c->stream>>>(g, ps + (size_t)f * S, sb, c->d_ps, X, c->d_tables);
                    prof_end(c);
                    c->launches++;
                }
                prof_begin(c, HEAAC_K_SYNTH);
                k_synth<<<(unsigned)(S * g.n_elems), 128, c->smem_synth, c->stream>>>(g, X, c->d_synth, pc, pf, c->d_tables);
                prof_end(c);
                c->launches++;
            }
        }
        return heaac_cuda_err(cudaGetLastError(), "kernel launch");
    }
    if (tns) {
        /* AAC-LC: the TNS pre-pass for every frame of the call at once, like on the SBR paths (the filters are serial chains: the
         * more of them in flight, the better their latency is hidden) */
        if (c->tns_done_frames < (size_t)n_frames) {
            CK(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_tns_done); c->d_tns_done = nullptr; c->tns_done_frames = 0;
            CK(cudaMalloc(&c->d_tns_done, (size_t)n_frames * n_ch * 1024 * sizeof(float)));
            c->tns_done_frames = n_frames;
        }
        const size_t n_all = (size_t)n_frames * n_ch;
        prof_begin(c, HEAAC_K_CORE);
        k_tns<<<(unsigned)((n_all * 8 + 127) / 128), 128, 0, c->stream>>>((int)n_all, g.sr_index, coeffs, ics, tns, c->d_tns_done, c->d_tables);
        prof_end(c);
        c->launches++;
    }
    const unsigned grid = n_ch < (unsigned)c->core_grid ? n_ch : (unsigned)c->core_grid;
    if (!g.n_cce) {                                  /* AAC-LC: one launch for the whole call, every CTA takes its channels through all frames */
        prof_begin(c, HEAAC_K_CORE);
        k_core<<<grid, HEAAC_THREADS, c->smem_core, c->stream>>>(g, (int)n_ch, n_frames, coeffs, ics, tns, c->d_chan, pcm, pcm_f32,
                                                                   c->d_tables, c->d_work, tns ? c->d_tns_done : nullptr);
        prof_end(c);
        c->launches++;
        return heaac_cuda_err(cudaGetLastError(), "kernel launch");
    }
    for (int f = 0; f < n_frames; f++) {
        /* layouts with coupling channel elements, frame by frame: every coded channel goes to the planar float buffer (its own slot
         * each), k_couple adds the coupling channels to their targets and converts */
        const size_t fo = (size_t)f * n_ch;
        HeaacGeom gp = g;
        gp.nout = g.ncc;
        for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) gp.out_slot[i] = (uint8_t)i;
        prof_begin(c, HEAAC_K_CORE);
        k_core<<<grid, HEAAC_THREADS, c->smem_core, c->stream>>>(gp, (int)n_ch, 1, coeffs + fo * 1024, ics + fo, tns ? tns + fo : nullptr, c->d_chan,
                                                                   nullptr, c->d_couple, c->d_tables, c->d_work,
                                                                   tns ? c->d_tns_done + fo * 1024 : nullptr);
        k_couple<<<(unsigned)S, HEAAC_THREADS, 0, c->stream>>>(g, c->d_couple, cce + (size_t)f * S * g.n_cce,
                                                                 pcm ? pcm + (size_t)f * pcm_frame : nullptr, pcm_f32 ? pcm_f32 + (size_t)f * pcm_frame : nullptr);
        prof_end(c);
        c->launches += 2;
    }
    return heaac_cuda_err(cudaGetLastError(), "kernel launch");
}

/* Enqueue a batch.  Host-pointer batches go through the double-buffered staging sets; `seq` counts staged
 * chunks over the life of the context so that submissions may follow one another without a host sync. */
extern "C" int heaac_b200_synth_submit(HeaacContext *c, const HeaacBatch *b)
{
    if (!c || !b || b->n_frames <= 0 || !b->coeffs || !b->ics) return HEAAC_E_INVAL;
    const HeaacGeom *g = &c->geom;
    if (g->sbr && !b->sbr) return HEAAC_E_INVAL;
    if (g->ps && !b->ps) return HEAAC_E_INVAL;
    if (g->n_cce && !b->cce) return HEAAC_E_INVAL;
    if (!b->pcm && !b->pcm_f32) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    if (b->on_device)
        return heaac_launch_frames(c, b->n_frames, b->coeffs, b->ics, b->tns, b->sbr, b->ps, b->pcm, b->pcm_f32, b->cce);

    if (b->pcm_f32 && !c->s_pcm_f32[0]) return HEAAC_E_INVAL;   /* context created without want_float_out */
    const size_t S = g->n_streams;
    const int Tc = c->chunk;
    for (int f0 = 0; f0 < b->n_frames; f0 += Tc, c->seq++) {
        const int buf = (int)(c->seq & 1);
        const int n = b->n_frames - f0 < Tc ? b->n_frames - f0 : Tc;
        const size_t nch = (size_t)n * S * g->ncc, och = (size_t)f0 * S * g->ncc;
        /* H2D on s_in: staging set `buf` is free once the kernels of the chunk two back are done */
        if (c->seq >= 2) CK(cudaStreamWaitEvent(c->s_in, c->ev_done[buf], 0));
        CK(cudaMemcpyAsync(c->s_coeffs[buf], b->coeffs + och * 1024, nch * 1024 * sizeof(float), cudaMemcpyHostToDevice, c->s_in));
        CK(cudaMemcpyAsync(c->s_ics[buf], b->ics + och, nch * sizeof(HeaacIcs), cudaMemcpyHostToDevice, c->s_in));
        if (b->tns) CK(cudaMemcpyAsync(c->s_tns[buf], b->tns + och, nch * sizeof(HeaacTns), cudaMemcpyHostToDevice, c->s_in));
        if (g->sbr) CK(cudaMemcpyAsync(c->s_sbr[buf], b->sbr + (size_t)f0 * S * g->nse, (size_t)n * S * g->nse * sizeof(HeaacSbrElem),
                                       cudaMemcpyHostToDevice, c->s_in));
        if (g->ps)  CK(cudaMemcpyAsync(c->s_ps[buf], b->ps + (size_t)f0 * S, (size_t)n * S * sizeof(HeaacPsFrame),
                                       cudaMemcpyHostToDevice, c->s_in));
        if (g->n_cce) CK(cudaMemcpyAsync(c->s_cce[buf], b->cce + (size_t)f0 * S * g->n_cce, (size_t)n * S * g->n_cce * sizeof(HeaacCce),
                                         cudaMemcpyHostToDevice, c->s_in));
        CK(cudaEventRecord(c->ev_in[buf], c->s_in));
        /* kernels on the context stream: need the inputs, and the PCM staging set drained by the D2H two chunks back */
        CK(cudaStreamWaitEvent(c->stream, c->ev_in[buf], 0));
        if (c->seq >= 2) CK(cudaStreamWaitEvent(c->stream, c->ev_out[buf], 0));
        int r = heaac_launch_frames(c, n, c->s_coeffs[buf], c->s_ics[buf], b->tns ? c->s_tns[buf] : nullptr, c->s_sbr[buf], c->s_ps[buf],
                              c->s_pcm[buf], b->pcm_f32 ? c->s_pcm_f32[buf] : nullptr, c->s_cce[buf]);
        if (r) return r;
        CK(cudaEventRecord(c->ev_done[buf], c->stream));
        /* D2H on s_out */
        const size_t pcm_n = (size_t)n * S * g->out_len * g->nout, pcm_o = (size_t)f0 * S * g->out_len * g->nout;
        CK(cudaStreamWaitEvent(c->s_out, c->ev_done[buf], 0));
        if (b->pcm) CK(cudaMemcpyAsync(b->pcm + pcm_o, c->s_pcm[buf], pcm_n * sizeof(int16_t), cudaMemcpyDeviceToHost, c->s_out));
        if (b->pcm_f32) CK(cudaMemcpyAsync(b->pcm_f32 + pcm_o, c->s_pcm_f32[buf], pcm_n * sizeof(float), cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaEventRecord(c->ev_out[buf], c->s_out));
    }
    return 0;
}

/* all host INPUT buffers of the batches submitted so far have been read (they may be overwritten) */
extern "C" int heaac_b200_inputs_consumed(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->s_in));
    return 0;
}

/* A ticket marks "all host-pointer work submitted so far": ticket_wait returns once the PCM of everything submitted
 * before the ticket was taken is in host memory, without waiting for later submissions.  Up to 8 outstanding. */
extern "C" long long heaac_b200_ticket(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    /* the event slot of the ticket issued 8 tickets ago is about to be reused: finish that ticket's work first (back
     * pressure on the submitter), so that "its event was reused" really does mean "its PCM is in host memory" */
    if (c->n_tickets >= 8) CK(cudaEventSynchronize(c->ev_ticket[c->n_tickets & 7]));
    CK(cudaEventRecord(c->ev_ticket[c->n_tickets & 7], c->s_out));
    return c->n_tickets++;
}

extern "C" int heaac_b200_ticket_wait(HeaacContext *c, long long ticket)
{
    if (!c || ticket < 0 || ticket >= c->n_tickets) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    /* a reused slot holds a later event on the same stream (s_out): waiting for it covers the older ticket too, and
     * heaac_b200_ticket() has already waited for the older work before reusing the slot */
    CK(cudaEventSynchronize(c->ev_ticket[ticket & 7]));
    return heaac_cuda_err(cudaGetLastError(), "ticket_wait");
}

/* everything submitted so far is complete: PCM is in the caller's buffers */
extern "C" int heaac_b200_wait(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->s_in));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->s_out));
    return heaac_cuda_err(cudaGetLastError(), "wait");
}

extern "C" int heaac_b200_synth(HeaacContext *c, const HeaacBatch *b)
{
    int r = heaac_b200_synth_submit(c, b);
    if (r) return r;
    if (b->on_device) return 0;                 /* device-resident batches stay asynchronous on the context's stream */
    return heaac_b200_wait(c);
}

/* page-locked host memory: what the H2D / D2H copies of the host-pointer path overlap best with */
extern "C" int heaac_b200_host_alloc(HeaacContext *c, size_t bytes, void **hptr)
{
    if (!c || !hptr) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    cudaError_t e = cudaHostAlloc(hptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return HEAAC_E_NOMEM; }
    return 0;
}

extern "C" int heaac_b200_host_free(HeaacContext *c, void *hptr)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaFreeHost(hptr));
    return 0;
}

extern "C" int heaac_b200_chunk_frames(const HeaacContext *c) { return c ? c->chunk : HEAAC_E_INVAL; }

/* aac_channel_layout_map (aacdectab.h:74-82): output order of the reference for channel configurations 1..7 as
 * (element type, instance) pairs; output_configure (aacdec.c:214-272) assigns output channels in this order. */
extern "C" int heaac_b200_reference_out_slots(const uint8_t *elem_type, int n_elems, uint8_t out_slot[HEAAC_MAX_CHANNELS])
{
    static const uint8_t cfg_elems[7][5] = {
        { HEAAC_TYPE_SCE }, { HEAAC_TYPE_CPE }, { HEAAC_TYPE_SCE, HEAAC_TYPE_CPE }, { HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_SCE },
        { HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE }, { HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE, HEAAC_TYPE_LFE },
        { HEAAC_TYPE_SCE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE, HEAAC_TYPE_CPE, HEAAC_TYPE_LFE } };
    static const uint8_t cfg_n[7] = { 1, 1, 2, 3, 3, 4, 5 };
    static const uint8_t order[7][5][2] = {                      /* {type, instance} in output order */
        { { HEAAC_TYPE_SCE, 0 } }, { { HEAAC_TYPE_CPE, 0 } }, { { HEAAC_TYPE_CPE, 0 }, { HEAAC_TYPE_SCE, 0 } },
        { { HEAAC_TYPE_CPE, 0 }, { HEAAC_TYPE_SCE, 0 }, { HEAAC_TYPE_SCE, 1 } },
        { { HEAAC_TYPE_CPE, 0 }, { HEAAC_TYPE_SCE, 0 }, { HEAAC_TYPE_CPE, 1 } },
        { { HEAAC_TYPE_CPE, 0 }, { HEAAC_TYPE_SCE, 0 }, { HEAAC_TYPE_LFE, 0 }, { HEAAC_TYPE_CPE, 1 } },
        { { HEAAC_TYPE_CPE, 0 }, { HEAAC_TYPE_SCE, 0 }, { HEAAC_TYPE_LFE, 0 }, { HEAAC_TYPE_CPE, 2 }, { HEAAC_TYPE_CPE, 1 } } };
    if (!elem_type || !out_slot || n_elems <= 0 || n_elems > HEAAC_MAX_ELEMS) return HEAAC_E_INVAL;
    int ncc = 0;
    for (int e = 0; e < n_elems; e++) ncc += elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
    if (ncc > HEAAC_MAX_CHANNELS) return HEAAC_E_INVAL;
    memset(out_slot, 0, HEAAC_MAX_CHANNELS);
    int cc = -1;
    for (int k = 0; k < 7 && cc < 0; k++)
        if (cfg_n[k] == n_elems && !memcmp(cfg_elems[k], elem_type, n_elems)) cc = k;
    if (cc < 0) { for (int c = 0; c < ncc; c++) out_slot[c] = (uint8_t)c; return ncc; }     /* identity */
    int ch = 0, seen[8] = { 0 };
    for (int e = 0; e < n_elems; e++) {
        const int type = elem_type[e], inst = seen[type & 7]++;
        int pos = 0;
        for (int k = 0; k < n_elems; k++) {
            if (order[cc][k][0] == type && order[cc][k][1] == inst) break;
            pos += order[cc][k][0] == HEAAC_TYPE_CPE ? 2 : 1;
        }
        out_slot[ch++] = (uint8_t)pos;
        if (type == HEAAC_TYPE_CPE) out_slot[ch++] = (uint8_t)(pos + 1);
    }
    return ncc;
}

extern "C" int heaac_b200_sync(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    return heaac_cuda_err(cudaGetLastError(), "sync");
}

extern "C" int heaac_b200_dev_alloc(HeaacContext *c, size_t bytes, void **dptr)
{
    if (!c || !dptr) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMalloc(dptr, bytes));
    return 0;
}
extern "C" int heaac_b200_dev_free(HeaacContext *c, void *dptr)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaFree(dptr));
    return 0;
}
extern "C" int heaac_b200_dev_upload(HeaacContext *c, void *dptr, const void *hptr, size_t bytes)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int heaac_b200_dev_download(HeaacContext *c, void *hptr, const void *dptr, size_t bytes)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int heaac_b200_timer_start(HeaacContext *c)
{
    if (!c) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return 0;
}
extern "C" int heaac_b200_timer_stop(HeaacContext *c, float *ms)
{
    if (!c || !ms) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

/* Per-kernel timing with CUDA events on the context's stream.  enable(capacity) starts recording up to
 * `capacity` launches; read() synchronises and returns total milliseconds and launch counts per kernel
 * kind (0 core, 1 sbr, 2 ps, 3 synth), then clears the record. */
extern "C" int heaac_b200_profile_enable(HeaacContext *c, int capacity)
{
    if (!c || capacity < 0) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    if (capacity > c->prof_cap) {
        cudaEvent_t *ev = (cudaEvent_t *)realloc(c->prof_ev, sizeof(cudaEvent_t) * 2 * capacity);
        unsigned char *kd = (unsigned char *)realloc(c->prof_kind, capacity);
        if (!ev || !kd) return HEAAC_E_NOMEM;
        c->prof_ev = ev; c->prof_kind = kd;
        for (int i = 2 * c->prof_cap; i < 2 * capacity; i++) CK(cudaEventCreate(&c->prof_ev[i]));
        c->prof_cap = capacity;
    }
    c->prof_n = 0;
    c->prof_on = capacity > 0;
    return 0;
}
extern "C" int heaac_b200_profile_read(HeaacContext *c, float ms[4], long long counts[4])
{
    if (!c || !ms || !counts) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < HEAAC_K_COUNT; k++) { ms[k] = 0.0f; counts[k] = 0; }
    for (int i = 0; i < c->prof_n; i++) {
        float t = 0.0f;
        CK(cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        ms[c->prof_kind[i]] += t;
        counts[c->prof_kind[i]]++;
    }
    c->prof_n = 0;
    return 0;
}

extern "C" int heaac_b200_debug_read(HeaacContext *c, const char *name, void *dst, size_t bytes)
{
    if (!c || !name || !dst) return HEAAC_E_INVAL;
    const HeaacGeom *g = &c->geom;
    const void *src = nullptr; size_t avail = 0;
    if (!strcmp(name, "X") && c->d_X) { src = c->d_X; avail = (size_t)g->n_streams * g->nout * HEAAC_X_FLOATS * sizeof(float); }
    else if (!strcmp(name, "chan_state")) { src = c->d_chan; avail = (size_t)g->n_streams * g->ncc * sizeof(HeaacChanState); }
    else if (!strcmp(name, "synth_state") && c->d_synth) { src = c->d_synth; avail = (size_t)g->n_streams * g->nout * sizeof(HeaacSynthState); }
    else if (!strcmp(name, "ps_state") && c->d_ps) { src = c->d_ps; avail = (size_t)g->n_streams * sizeof(HeaacPsState); }
    else if (!strcmp(name, "tables")) { src = c->d_tables; avail = sizeof(HeaacDevTables); }
    if (!src) return HEAAC_E_INVAL;
    if (bytes > avail) bytes = avail;
    return heaac_b200_dev_download(c, dst, src, bytes) ? HEAAC_E_LAUNCH : (int)(bytes > 0x7fffffff ? 0x7fffffff : bytes);
}

/* ---- ff_sbr_apply / ff_ps_apply shims (batch of one) ---------------------------------------
 * The reference's two stage entry points with plain arrays instead of its private structs: everything they keep
 * between calls (QMF delay lines, W / Y tails, gain history, PS delay lines ...) lives in a one-stream context.
 * For parity checks and incremental bring-up; the batched path does not go through them. */
struct HeaacSbrShim {
    HeaacContext *ctx;
    float *d_time;                   /* [nch][1024] core output handed in */
    HeaacSbrElem *d_elem;
    HeaacPsFrame *d_psf;
    float *d_out;                    /* [nout][2048] */
    int nch, nout, ps;
};

extern "C" int heaac_b200_sbr_init(HeaacSbrShim **out, int id_aac, int core_sr_index, int ps, float sf_scale, float add_bias)
{
    if (!out || (id_aac != HEAAC_TYPE_SCE && id_aac != HEAAC_TYPE_CPE) || (ps && id_aac != HEAAC_TYPE_SCE)) return HEAAC_E_INVAL;
    HeaacSbrShim *s = (HeaacSbrShim *)calloc(1, sizeof(*s));
    if (!s) return HEAAC_E_NOMEM;
    HeaacConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = HEAAC_ABI_VERSION; cfg.n_streams = 1; cfg.n_elems = 1; cfg.elem_type[0] = (uint8_t)id_aac;
    cfg.sr_index = core_sr_index; cfg.sbr = 1; cfg.ps = ps ? 1 : 0; cfg.add_bias = add_bias; cfg.sf_scale = sf_scale;
    cfg.max_frames_per_call = 1; cfg.want_float_out = 1;
    int r = heaac_b200_create(&cfg, &s->ctx);
    if (r) { free(s); return r; }
    s->nch = id_aac == HEAAC_TYPE_CPE ? 2 : 1; s->nout = ps ? 2 : s->nch; s->ps = ps ? 1 : 0;
    if (cudaMalloc(&s->d_time, sizeof(float) * 2 * 1024) != cudaSuccess || cudaMalloc(&s->d_elem, sizeof(HeaacSbrElem)) != cudaSuccess ||
        cudaMalloc(&s->d_psf, sizeof(HeaacPsFrame)) != cudaSuccess || cudaMalloc(&s->d_out, sizeof(float) * 2 * 2048) != cudaSuccess) {
        cudaGetLastError();
        heaac_b200_destroy(s->ctx); cudaFree(s->d_time); cudaFree(s->d_elem); cudaFree(s->d_psf); cudaFree(s->d_out); free(s);
        return HEAAC_E_NOMEM;
    }
    *out = s;
    return 0;
}

extern "C" void heaac_b200_sbr_close(HeaacSbrShim *s)
{
    if (!s) return;
    cudaFree(s->d_time); cudaFree(s->d_elem); cudaFree(s->d_psf); cudaFree(s->d_out);
    heaac_b200_destroy(s->ctx);
    free(s);
}

/* == ff_sbr_apply(ac, sbr, id_aac, L, R), aacsbr.c:1716-1771.  L (and R for a CPE) hold the 1024 core samples of the
 * frame on entry (the element's ret[] after imdct_and_windowing) and the 2048 output samples on exit; with PS, R is
 * output only.  `elem` is the element's SBR side info for this frame, `psf` the PS side info (PS streams). */
extern "C" int heaac_b200_sbr_apply(HeaacSbrShim *s, const HeaacSbrElem *elem, const HeaacPsFrame *psf, float *L, float *R)
{
    if (!s || !elem || !L || (s->nout == 2 && !R) || (s->ps && !psf)) return HEAAC_E_INVAL;
    HeaacContext *c = s->ctx;
    const HeaacGeom g = c->geom;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(s->d_time, L, sizeof(float) * 1024, cudaMemcpyHostToDevice, c->stream));
    if (s->nch == 2) CK(cudaMemcpyAsync(s->d_time + 1024, R, sizeof(float) * 1024, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(s->d_elem, elem, sizeof(HeaacSbrElem), cudaMemcpyHostToDevice, c->stream));
    if (s->ps) CK(cudaMemcpyAsync(s->d_psf, psf, sizeof(HeaacPsFrame), cudaMemcpyHostToDevice, c->stream));
    /* the chain from the element's time samples (no core transform): QMF analysis .. X (.. synthesis without PS) */
    launch_chain(c, (unsigned)s->nch, 1, s->ps ? 1 : 0, nullptr, s->d_time, nullptr, nullptr, nullptr, s->d_elem, c->d_chan, c->d_synth, c->d_X,
                 nullptr, s->ps ? nullptr : s->d_out);
    if (s->ps) {
        k_ps<<<1, HEAAC_THREADS, c->smem_ps, c->stream>>>(g, s->d_psf, s->d_elem, c->d_ps, c->d_X, c->d_tables);
        k_synth<<<1, 128, c->smem_synth, c->stream>>>(g, c->d_X, c->d_synth, nullptr, s->d_out, c->d_tables);
    }
    c->launches += 1 + 2 * s->ps;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(L, s->d_out, sizeof(float) * 2048, cudaMemcpyDeviceToHost, c->stream));
    if (s->nout == 2) CK(cudaMemcpyAsync(R, s->d_out + 2048, sizeof(float) * 2048, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

/* == ff_ps_apply(avctx, ps, L, R, top), aacps.c:973-992: L[2][38][64] (re / im planes, slot, QMF band) is processed in
 * place for slots 0..31 (slots 32..37 are look-ahead), R[2][38][64] receives the second channel.  Uses the PS state
 * of a shim created with ps = 1. */
extern "C" int heaac_b200_ps_apply(HeaacSbrShim *s, const HeaacPsFrame *psf, float *L, float *R, int top)
{
    if (!s || !s->ps || !psf || !L || !R || top < 0 || top > 64) return HEAAC_E_INVAL;
    HeaacContext *c = s->ctx;
    const HeaacGeom g = c->geom;
    CK(cudaSetDevice(c->cfg.device));
    float *h = (float *)calloc(2 * HEAAC_X_FLOATS, sizeof(float));
    if (!h) return HEAAC_E_NOMEM;
    for (int pl = 0; pl < 2; pl++)                         /* reference [plane][slot][band] -> device [plane][band][slot] */
        for (int i = 0; i < 38; i++)
            for (int k = 0; k < 64; k++) h[pl * HEAAC_X_PLANE + k * HEAAC_XS + i] = L[(pl * 38 + i) * 64 + k];
    HeaacSbrElem e;
    memset(&e, 0, sizeof(e));
    e.kx[1] = (uint8_t)top;                               /* k_ps derives `top` as kx + m */
    int r = heaac_cuda_err(cudaMemcpyAsync(c->d_X, h, sizeof(float) * HEAAC_X_FLOATS, cudaMemcpyHostToDevice, c->stream), "upload X");
    if (!r) r = heaac_cuda_err(cudaMemcpyAsync(s->d_elem, &e, sizeof(e), cudaMemcpyHostToDevice, c->stream), "upload elem");
    if (!r) r = heaac_cuda_err(cudaMemcpyAsync(s->d_psf, psf, sizeof(HeaacPsFrame), cudaMemcpyHostToDevice, c->stream), "upload ps");
    if (!r) {
        k_ps<<<1, HEAAC_THREADS, c->smem_ps, c->stream>>>(g, s->d_psf, s->d_elem, c->d_ps, c->d_X, c->d_tables);
        c->launches++;
        r = heaac_cuda_err(cudaGetLastError(), "k_ps");
    }
    if (!r) r = heaac_cuda_err(cudaMemcpyAsync(h, c->d_X, sizeof(float) * 2 * HEAAC_X_FLOATS, cudaMemcpyDeviceToHost, c->stream), "download X");
    if (!r) r = heaac_cuda_err(cudaStreamSynchronize(c->stream), "sync");
    if (!r)
        for (int ch = 0; ch < 2; ch++) {
            float *dst = ch ? R : L;
            for (int pl = 0; pl < 2; pl++)
                for (int i = 0; i < 32; i++)
                    for (int k = 0; k < 64; k++)
                        dst[(pl * 38 + i) * 64 + k] = h[ch * HEAAC_X_FLOATS + pl * HEAAC_X_PLANE + k * HEAAC_XS + i];
        }
    free(h);
    return r;
}

/* ---- ff_mdct_init / ff_imdct_half / ff_mdct_end shims (batch of one) -------------------- */
struct HeaacMdct {
    int nbits; float *d_tcos, *d_tsin, *d_in, *d_out; HeaacDevTables *d_tables; cudaStream_t stream;
};

__global__ void k_imdct_half_one(float *out, const float *in, const float *tcos, const float *tsin,
                                 int log2n4, const HeaacDevTables *T)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *buf = reinterpret_cast<float *>(smem_raw);
    const int n4 = 1 << log2n4, n2 = n4 * 2;
    float *zr = buf + n2, *zi = zr + n4 + (n4 >> 3);           /* skewed layout: HEAAC_ZP */
    for (int i = threadIdx.x; i < n2; i += blockDim.x) buf[i] = in[i];
    __syncthreads();
    heaac_imdct_half_smem<HEAAC_THREADS>(buf, buf, zr, zi, log2n4, 1, tcos, tsin, T, threadIdx.x);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) out[i] = buf[i];
}

extern "C" int heaac_b200_mdct_init(HeaacMdct **out, int nbits, int inverse, double scale)
{
    if (!out || !inverse) return HEAAC_E_INVAL;
    if (nbits < 5 || nbits > 11) return HEAAC_E_UNSUPPORTED;    /* FFT twiddle table covers up to 512 points */
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return HEAAC_E_NODEV;
    HeaacMdct *m = (HeaacMdct *)calloc(1, sizeof(*m));
    if (!m) return HEAAC_E_NOMEM;
    m->nbits = nbits;
    int n = 1 << nbits, n4 = n >> 2;
    float *tc = (float *)malloc(sizeof(float) * n4 * 2), *ts = tc + n4;
    double theta = 1.0 / 8.0 + (scale < 0 ? n4 : 0), sc = sqrt(fabs(scale));
    for (int i = 0; i < n4; i++) {
        double alpha = 2 * 3.14159265358979323846 * (i + theta) / n;
        tc[i] = -cos(alpha) * sc; ts[i] = -sin(alpha) * sc;
    }
    HeaacDevTables *h = (HeaacDevTables *)malloc(sizeof(HeaacDevTables));
    heaac_build_tables(h);
    int r = 0;
    if (!r) r = heaac_cuda_err(cudaMalloc(&m->d_tables, sizeof(HeaacDevTables)), "cudaMalloc");
    if (!r) r = heaac_cuda_err(cudaMemcpy(m->d_tables, h, sizeof(HeaacDevTables), cudaMemcpyHostToDevice), "cudaMemcpy");
    if (!r) r = heaac_cuda_err(cudaMalloc(&m->d_tcos, sizeof(float) * n4 * 2), "cudaMalloc");
    if (!r) r = heaac_cuda_err(cudaMemcpy(m->d_tcos, tc, sizeof(float) * n4 * 2, cudaMemcpyHostToDevice), "cudaMemcpy");
    if (!r) r = heaac_cuda_err(cudaMalloc(&m->d_in, sizeof(float) * n), "cudaMalloc");
    free(h); free(tc);
    if (r) { heaac_b200_mdct_end(m); return r; }
    m->d_tsin = m->d_tcos + n4;
    m->d_out = m->d_in + n / 2;
    *out = m;
    return 0;
}

extern "C" void heaac_b200_imdct_half(HeaacMdct *m, float *output, const float *input)
{
    if (!m) return;
    int n = 1 << m->nbits, n2 = n >> 1, n4 = n >> 2;
    cudaMemcpy(m->d_in, input, sizeof(float) * n2, cudaMemcpyHostToDevice);
    k_imdct_half_one<<<1, 256, sizeof(float) * (n2 + 2 * (n4 + (n4 >> 3)))>>>(m->d_out, m->d_in, m->d_tcos, m->d_tsin,
                                                                 m->nbits - 2, m->d_tables);
    cudaMemcpy(output, m->d_out, sizeof(float) * n2, cudaMemcpyDeviceToHost);
}

extern "C" void heaac_b200_mdct_end(HeaacMdct *m)
{
    if (!m) return;
    cudaFree(m->d_tables); cudaFree(m->d_tcos); cudaFree(m->d_in);
    free(m);
}

extern "C" const char *heaac_b200_version(void) { return "heaac_b200 0.1 (sm_100a)"; }
extern "C" const char *heaac_b200_strerror(int err)
{
    switch (err) {
    case HEAAC_OK: return "ok";
    case HEAAC_E_INVAL: return "invalid argument";
    case HEAAC_E_NOMEM: return "out of memory";
    case HEAAC_E_NODEV: return "no usable sm_100 CUDA device (no CPU path exists)";
    case HEAAC_E_LAUNCH: return "CUDA launch or runtime failure";
    case HEAAC_E_UNSUPPORTED: return "not supported by this build";
    default: return "unknown error";
    }
}

/* struct sizes, so that language bindings can assert their mirrors (abi.py) */
extern "C" int heaac_b200_sizeof(const char *name)
{
    if (!strcmp(name, "HeaacIcs")) return (int)sizeof(HeaacIcs);
    if (!strcmp(name, "HeaacTns")) return (int)sizeof(HeaacTns);
    if (!strcmp(name, "HeaacSbrChan")) return (int)sizeof(HeaacSbrChan);
    if (!strcmp(name, "HeaacSbrElem")) return (int)sizeof(HeaacSbrElem);
    if (!strcmp(name, "HeaacPsFrame")) return (int)sizeof(HeaacPsFrame);
    if (!strcmp(name, "HeaacCce")) return (int)sizeof(HeaacCce);
    if (!strcmp(name, "HeaacConfig")) return (int)sizeof(HeaacConfig);
    if (!strcmp(name, "HeaacBatch")) return (int)sizeof(HeaacBatch);
    if (!strcmp(name, "HeaacDecoderConfig")) return (int)sizeof(HeaacDecoderConfig);
    if (!strcmp(name, "HeaacChanState")) return (int)sizeof(HeaacChanState);
    if (!strcmp(name, "HeaacSynthState")) return (int)sizeof(HeaacSynthState);
    if (!strcmp(name, "HeaacPsState")) return (int)sizeof(HeaacPsState);
    if (!strcmp(name, "HeaacDevTables")) return (int)sizeof(HeaacDevTables);
    return -1;
}
