/*
 * heaac_parse_gpu.cu -- the front end on the GPU: packets in, PCM out, nothing but bitstreams over PCIe.
 *
 * The same parser source as the host front end (heaac_fe_core.h, compiled here with FE_DEVICE) runs with one warp
 * per stream: entropy decoding is serial inside a frame but the streams of a batch are independent, and so are the
 * core elements of different frames of one stream.  The parse kernels write coeffs / ICS / TNS / SBR / PS
 * structures straight into the back end's device staging sets, so the host-to-device traffic of a batch shrinks
 * from the dequantised spectra (4 KiB per channel-frame) to the packets themselves (~0.2 KiB).
 *
 * Per-stream parser state (struct HeaacFrontend) lives in HBM between calls.  Chunks are double buffered exactly
 * like the host-pointer path of heaac_b200_synth_submit: packets H2D + k_parse_a on s_in, k_parse_b on its own
 * stream, synthesis kernels on the context's stream, PCM D2H on s_out.
 */
#define FE_DEVICE 1
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "heaac_context.cuh"
#ifndef HEAAC_PARSE_AFTER_SYNTH
#define HEAAC_PARSE_AFTER_SYNTH 1
#endif
#include "heaac_fe_core.h"

/* pipeline trace (event times of every chunk on stderr): a debug build only, -DHEAAC_DEBUG_TIMING (make EXTRA=-DHEAAC_DEBUG_TIMING) */
#ifdef HEAAC_DEBUG_TIMING
#define HEAAC_DBG_TIMING 1
#else
#define HEAAC_DBG_TIMING 0
#endif

extern "C" const void *heaac_fe_tables_blob(size_t *bytes);
extern "C" size_t heaac_fe_state_size(void);
extern "C" const void *heaac_fe_state_blob(const HeaacFrontend *fe);

struct HeaacGpuParser {
    HeaacContext *ctx;
    FE *d_fe;                        /* [S] parser state */
    FE *d_template;                  /* state right after open, for reset */
    FeTables *d_tb;
    uint8_t *d_scratch;              /* [S][32] frame-local parser states of phase A (head + channel blocks) */
    size_t scratch_stride;
    FeTodo *d_todo[2];               /* [S][32] per staging set: phase B of chunk k reads one while phase A of k+1 fills the other */
    int gather_threads;              /* host threads that copy a chunk's packets into the page-locked blob (chunks of 8 MB and more) */
    FeSbrTape *d_tape;               /* [S][32][tapes_per_frame]: the SBR payloads' first step (k_parse_b; its launches are serialised on s_b) */
    int tapes_per_frame;
    cudaStream_t s_b;                /* phase B's stream */
    cudaEvent_t ev_a[2];             /* phase A of the chunk in staging set i is done */
    /* per staging set: packets of one chunk, 8-byte aligned and zero padded */
    uint8_t *h_blob[2], *d_blob[2];
    uint32_t *h_off[2], *d_off[2];   /* [chunk][S] byte offset of each packet in the blob */
    int *h_size[2], *d_size[2];
    int *d_consumed[2];
    int *h_consumed[8];              /* page-locked, [max_frames_per_call][S] per outstanding call: keeps the D2H asynchronous */
    int *user_consumed[8];           /* where each outstanding call's consumed[] goes once it is waited for */
    int user_n[8];
    size_t blob_cap;
    cudaEvent_t ev_h2d[2];
    cudaEvent_t ev_k0[2], ev_k1[2];  /* around k_parse, for parser_stats */
    int timed[2];
    double k_ms;
    long long k_timed;
    long long launches;
    /* HEAAC_DEBUG_TIMING: device timeline of the last submissions */
    cudaEvent_t dbg_base, dbg_ev[64][4];
    int dbg_n, dbg_on;
};

/* Two kernels per chunk, one warp per stream in both (see heaac_fe_core.h, "two-phase parse"):
 *   k_parse_a  lane t parses the core elements of the chunk's frame t on a frame-local scratch state, independent of
 *              the other frames and of the stream's history;
 *   k_parse_b  lane 0 walks the frames in order on the stream's persistent state (window history, SBR / PS payloads);
 *              the whole warp packs the result.
 * They run on two CUDA streams, so phase B of chunk k overlaps phase A of chunk k+1. */
#ifndef HEAAC_PARSE_THREADS
#define HEAAC_PARSE_THREADS 128
#endif
#ifndef HEAAC_PARSE_MIN_BLOCKS
#define HEAAC_PARSE_MIN_BLOCKS (384 / HEAAC_PARSE_THREADS)
#endif
__global__ void __launch_bounds__(HEAAC_PARSE_THREADS, HEAAC_PARSE_MIN_BLOCKS)
k_parse_a(const FE *fes, uint8_t *scratch, size_t scratch_stride, FeTodo *todo, const uint8_t *blob, const uint32_t *off,
          const int *size, int n_frames, int S, int ncc, float *coeffs, HeaacIcs *ics, HeaacTns *tns)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= S || lane >= n_frames) return;
    const size_t idx = (size_t)lane * S + s;
    fe_phase_a(fes + s, (FE *)(scratch + ((size_t)s * 32 + lane) * scratch_stride), todo + (size_t)s * 32 + lane,
               blob + off[idx], size[idx], coeffs + idx * ncc * 1024, ics + idx * ncc, tns + idx * ncc);
}

__global__ void __launch_bounds__(HEAAC_PARSE_THREADS, HEAAC_PARSE_MIN_BLOCKS)
k_parse_b(FE *fes, const FeTodo *todo, FeSbrTape *tape, int tapes_per_frame, const uint8_t *blob, const uint32_t *off,
          const int *size, int n_frames, int S, int ncc, int nse, int has_sbr, int has_ps, float *coeffs, HeaacIcs *ics,
          HeaacTns *tns, HeaacSbrElem *sbr, HeaacPsFrame *ps, int *consumed)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= S) return;
    FeChunkIO io;
    io.n_frames = n_frames; io.ncc = ncc; io.nse = nse; io.stride = (size_t)S;
    io.todo = todo + (size_t)s * 32;
    io.tape = tape + (size_t)s * 32 * tapes_per_frame; io.tapes_per_frame = tapes_per_frame;
    io.blob = blob; io.off = off + s; io.size = size + s;
    io.coeffs = coeffs + (size_t)s * ncc * 1024; io.ics = ics + (size_t)s * ncc; io.tns = tns + (size_t)s * ncc;
    io.sbr = has_sbr ? sbr + (size_t)s * nse : nullptr; io.ps = has_ps ? ps + s : nullptr;
    io.consumed = consumed + s; io.stats = nullptr;
    fe_chunk_b(fes + s, &io, lane, 32);
}

static void parser_free(HeaacGpuParser *p)
{
    if (!p) return;
    for (int b = 0; b < 2; b++) {
        if (p->h_blob[b]) cudaFreeHost(p->h_blob[b]);
        if (p->h_off[b]) cudaFreeHost(p->h_off[b]);
        if (p->h_size[b]) cudaFreeHost(p->h_size[b]);
        if (p->d_blob[b]) cudaFree(p->d_blob[b]);
        if (p->d_off[b]) cudaFree(p->d_off[b]);
        if (p->d_size[b]) cudaFree(p->d_size[b]);
        if (p->d_consumed[b]) cudaFree(p->d_consumed[b]);
        if (p->ev_h2d[b]) cudaEventDestroy(p->ev_h2d[b]);
        if (p->ev_k0[b]) cudaEventDestroy(p->ev_k0[b]);
        if (p->ev_k1[b]) cudaEventDestroy(p->ev_k1[b]);
    }
    for (int i = 0; i < 8; i++) if (p->h_consumed[i]) cudaFreeHost(p->h_consumed[i]);
    if (p->d_fe) cudaFree(p->d_fe);
    if (p->d_template) cudaFree(p->d_template);
    if (p->d_scratch) cudaFree(p->d_scratch);
    if (p->d_tape) cudaFree(p->d_tape);
    for (int i = 0; i < 2; i++) { if (p->d_todo[i]) cudaFree(p->d_todo[i]); if (p->ev_a[i]) cudaEventDestroy(p->ev_a[i]); }
    if (p->s_b) cudaStreamDestroy(p->s_b);
    delete p;
}

extern "C" void heaac_b200_parser_debug_timeline(HeaacGpuParser *p);
extern "C" void heaac_b200_parser_destroy(HeaacGpuParser *p)
{
    if (!p) return;
    heaac_b200_parser_debug_timeline(p);
    cudaSetDevice(p->ctx->cfg.device);
    cudaStreamSynchronize(p->ctx->s_in);
    cudaStreamSynchronize(p->s_b);
    cudaStreamSynchronize(p->ctx->stream);
    cudaStreamSynchronize(p->ctx->s_out);
    parser_free(p);
}

extern "C" int heaac_b200_parser_reset(HeaacGpuParser *p)
{
    if (!p) return HEAAC_E_INVAL;
    HeaacContext *c = p->ctx;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->s_in));
    for (int s = 0; s < c->geom.n_streams; s++)
        CK(cudaMemcpyAsync(p->d_fe + s, p->d_template, sizeof(FE), cudaMemcpyDeviceToDevice, c->s_in));
    CK(cudaStreamSynchronize(c->s_in));
    return 0;
}

/* `fe` = a host front end opened for this stream configuration and not yet used: its state block is the initial
 * state of every stream's device parser. */
extern "C" int heaac_b200_parser_create(HeaacContext *c, const HeaacFrontend *fe, size_t max_packet_bytes, HeaacGpuParser **out)
{
    if (!c || !fe || !out) return HEAAC_E_INVAL;
    if (heaac_fe_state_size() != sizeof(FE)) return HEAAC_E_INVAL;            /* host and device builds must agree */
    CK(cudaSetDevice(c->cfg.device));
    HeaacGpuParser *p = new (std::nothrow) HeaacGpuParser();
    if (!p) return HEAAC_E_NOMEM;
    memset(p, 0, sizeof(*p));
    p->ctx = c;
    const size_t S = c->geom.n_streams, n = (size_t)c->chunk * S;
    if (!max_packet_bytes) max_packet_bytes = 6144 / 8 * c->geom.ncc + 64;    /* 6144 bits per channel, aacenc.c:595 */
    p->blob_cap = n * ((max_packet_bytes + 15) & ~(size_t)7);
    int r = 0;
#define CKP(call) do { if (!r) r = heaac_cuda_err((call), #call); } while (0)
    size_t tb_bytes = 0;
    const void *tb = heaac_fe_tables_blob(&tb_bytes);
    if (tb_bytes != sizeof(FeTables)) r = HEAAC_E_INVAL;
    {
        /* One copy of the parser tables per device for the life of the process: the kernels reach it through a
         * __device__ pointer, so it must not go away while any parser on that device is alive. */
        static std::mutex mu;
        static FeTables *per_device[64];
        std::lock_guard<std::mutex> lk(mu);
        const int dev = c->cfg.device;
        if (dev < 0 || dev >= 64) r = HEAAC_E_INVAL;
        if (!r && !per_device[dev]) {
            FeTables *d = nullptr;
            CKP(cudaMalloc(&d, sizeof(FeTables)));
            CKP(cudaMemcpy(d, tb, sizeof(FeTables), cudaMemcpyHostToDevice));
            CKP(cudaMemcpyToSymbol(fe_tb_dev, &d, sizeof(d)));
            if (!r) per_device[dev] = d;
            else if (d) cudaFree(d);
        }
        if (!r) p->d_tb = per_device[dev];
    }
    CKP(cudaMalloc(&p->d_fe, S * sizeof(FE)));
    CKP(cudaMalloc(&p->d_template, sizeof(FE)));
    p->scratch_stride = (offsetof(FE, ch) + (size_t)c->geom.ncc * sizeof(FeChan) + 15) & ~(size_t)15;
    CKP(cudaMalloc(&p->d_scratch, S * 32 * p->scratch_stride));
    for (int i = 0; i < 2; i++) {
        CKP(cudaMalloc(&p->d_todo[i], S * 32 * sizeof(FeTodo)));
        CKP(cudaEventCreateWithFlags(&p->ev_a[i], cudaEventDisableTiming));
    }
    {
        const unsigned hw = std::thread::hardware_concurrency();
        p->gather_threads = hw >= 16 ? 4 : hw >= 8 ? 2 : 1;
    }
    p->tapes_per_frame = c->geom.nse > 0 ? c->geom.nse : 1;
    CKP(cudaMalloc(&p->d_tape, (c->geom.sbr ? S * 32 * p->tapes_per_frame : 1) * sizeof(FeSbrTape)));
    CKP(cudaStreamCreateWithFlags(&p->s_b, cudaStreamNonBlocking));
    CKP(cudaMemcpy(p->d_template, heaac_fe_state_blob(fe), sizeof(FE), cudaMemcpyHostToDevice));
    for (int i = 0; i < 8; i++) CKP(cudaHostAlloc(&p->h_consumed[i], (size_t)c->cfg.max_frames_per_call * S * sizeof(int), cudaHostAllocDefault));
    for (int b = 0; b < 2 && !r; b++) {
        CKP(cudaHostAlloc(&p->h_blob[b], p->blob_cap, cudaHostAllocDefault));
        CKP(cudaHostAlloc(&p->h_off[b], n * sizeof(uint32_t), cudaHostAllocDefault));
        CKP(cudaHostAlloc(&p->h_size[b], n * sizeof(int), cudaHostAllocDefault));
        CKP(cudaMalloc(&p->d_blob[b], p->blob_cap));
        CKP(cudaMalloc(&p->d_off[b], n * sizeof(uint32_t)));
        CKP(cudaMalloc(&p->d_size[b], n * sizeof(int)));
        CKP(cudaMalloc(&p->d_consumed[b], n * sizeof(int)));
        CKP(cudaEventCreateWithFlags(&p->ev_h2d[b], cudaEventDisableTiming));
        CKP(cudaEventCreate(&p->ev_k0[b]));
        CKP(cudaEventCreate(&p->ev_k1[b]));
    }
#undef CKP
    if (r) { parser_free(p); return r; }
    if (HEAAC_DBG_TIMING) {
        p->dbg_on = 1;
        cudaEventCreate(&p->dbg_base);
        for (int i = 0; i < 64; i++) for (int k = 0; k < 4; k++) cudaEventCreate(&p->dbg_ev[i][k]);
    }
    *out = p;
    r = heaac_b200_parser_reset(p);
    if (r) { parser_free(p); *out = nullptr; }
    return r;
}

extern "C" long long heaac_b200_parser_launch_count(const HeaacGpuParser *p) { return p ? p->launches : 0; }

/* average k_parse duration (CUDA events on its stream) over the launches harvested so far */
extern "C" int heaac_b200_parser_stats(const HeaacGpuParser *p, double *ms_per_launch, long long *timed_launches)
{
    if (!p) return HEAAC_E_INVAL;
    if (ms_per_launch) *ms_per_launch = p->k_timed ? p->k_ms / (double)p->k_timed : 0.0;
    if (timed_launches) *timed_launches = p->k_timed;
    return 0;
}

/* n_frames x n_streams packets -> PCM, asynchronously.  pkt / size / consumed are frame-major [n_frames][n_streams]
 * host arrays; pcm / pcm_f32 are host buffers as in heaac_b200_synth.  The packets are copied before the call
 * returns; pcm / pcm_f32 / consumed are filled by the time heaac_b200_parser_wait(ticket) returns.
 * Returns the ticket (>= 0) or a negative error. */
extern "C" long long heaac_b200_parser_submit(HeaacGpuParser *p, int n_frames, const uint8_t *const *pkt, const int *size,
                                              int16_t *pcm, float *pcm_f32, int *consumed)
{
    if (!p || n_frames <= 0 || !pkt || !size || (!pcm && !pcm_f32)) return HEAAC_E_INVAL;
    HeaacContext *c = p->ctx;
    const HeaacGeom *g = &c->geom;
    if (pcm_f32 && !c->s_pcm_f32[0]) return HEAAC_E_INVAL;
    if (n_frames > c->cfg.max_frames_per_call) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    const size_t S = g->n_streams;
    const int Tc = c->chunk;
    const int slot = (int)(c->n_tickets & 7);
    if (c->n_tickets >= 8 && p->user_consumed[slot]) {
        /* the call that used this slot 8 submissions ago was never waited for: complete it now (its consumed[] goes to
         * its caller's array) before h_consumed[slot] / user_consumed[slot] are reused */
        CK(cudaEventSynchronize(c->ev_ticket[slot]));
        memcpy(p->user_consumed[slot], p->h_consumed[slot], (size_t)p->user_n[slot] * sizeof(int));
        p->user_consumed[slot] = nullptr;
    }
    const int dbg = HEAAC_DBG_TIMING;
    double t_sync = 0, t_gather = 0, t_submit = 0, t_wait = 0;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    for (int f0 = 0; f0 < n_frames; f0 += Tc, c->seq++) {
        const int buf = (int)(c->seq & 1);
        const int n = n_frames - f0 < Tc ? n_frames - f0 : Tc;
        const size_t np = (size_t)n * S;
        const double t0 = now();
        /* the pinned packet set `buf` was last read by the H2D recorded in ev_h2d[buf] */
        if (c->seq >= 2) CK(cudaEventSynchronize(p->ev_h2d[buf]));
        if (p->timed[buf] && cudaEventQuery(p->ev_k1[buf]) == cudaSuccess) {        /* harvest the timing of two chunks ago */
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p->ev_k0[buf], p->ev_k1[buf]) == cudaSuccess && p->launches > 8) { p->k_ms += ms; p->k_timed++; }
            p->timed[buf] = 0;
        }
        const double t1 = now();
        size_t used = 0;
        for (size_t i = 0; i < np; i++) {                          /* layout: 8-byte aligned, >= 8 zero bytes after each */
            const int sz = size[(size_t)f0 * S + i];
            const size_t room = ((size_t)(sz > 0 ? sz : 0) + 15) & ~(size_t)7;
            if (sz < 0 || used + room > p->blob_cap) return HEAAC_E_INVAL;
            p->h_off[buf][i] = (uint32_t)used;
            p->h_size[buf][i] = sz;
            used += room;
        }
        {   /* gather into the page-locked blob; large chunks (AAC-LC 5.1: ~80 MB) by several host threads */
            uint8_t *dst = p->h_blob[buf];
            const uint32_t *off = p->h_off[buf];
            const uint8_t *const *src = pkt + (size_t)f0 * S;
            const int *sz = size + (size_t)f0 * S;
            auto copy_range = [=](size_t i0, size_t i1) {
                for (size_t i = i0; i < i1; i++) {
                    const size_t room = ((size_t)sz[i] + 15) & ~(size_t)7;
                    memcpy(dst + off[i], src[i], sz[i]);
                    memset(dst + off[i] + sz[i], 0, room - sz[i]);
                }
            };
            const int nt = used >= ((size_t)8 << 20) ? p->gather_threads : 1;
            if (nt <= 1) copy_range(0, np);
            else {
                std::vector<std::thread> th;
                for (int k = 1; k < nt; k++) th.emplace_back(copy_range, np * k / nt, np * (k + 1) / nt);
                copy_range(0, np / nt);
                for (auto &t : th) t.join();
            }
        }
        const double t2 = now();
        if (c->seq >= 2) CK(cudaStreamWaitEvent(c->s_in, c->ev_done[buf], 0));     /* staging set free again */
        CK(cudaMemcpyAsync(p->d_blob[buf], p->h_blob[buf], used, cudaMemcpyHostToDevice, c->s_in));
        CK(cudaMemcpyAsync(p->d_off[buf], p->h_off[buf], np * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_in));
        CK(cudaMemcpyAsync(p->d_size[buf], p->h_size[buf], np * sizeof(int), cudaMemcpyHostToDevice, c->s_in));
        CK(cudaEventRecord(p->ev_h2d[buf], c->s_in));
        const int di = p->dbg_on && p->dbg_n < 64 ? p->dbg_n++ : -1;
        if (di == 0) CK(cudaEventRecord(p->dbg_base, c->s_in));
        if (di >= 0) CK(cudaEventRecord(p->dbg_ev[di][0], c->s_in));
        if (c->seq >= 2) CK(cudaStreamWaitEvent(c->s_in, c->ev_out[buf], 0));      /* d_consumed[buf] drained */
#if HEAAC_PARSE_AFTER_SYNTH
        /* the front end of this chunk starts when the synthesis of the chunk before has left the SMs: side by side the two
         * latency-bound kernels only slow each other down (measured) */
        if (c->seq >= 1) CK(cudaStreamWaitEvent(c->s_in, c->ev_done[buf ^ 1], 0));
#endif
        {
            const int threads = HEAAC_PARSE_THREADS, warps_per_cta = threads / 32;
            const int grid = (int)((S + warps_per_cta - 1) / warps_per_cta);
            if (!p->timed[buf]) CK(cudaEventRecord(p->ev_k0[buf], c->s_in));
            k_parse_a<<<grid, threads, 0, c->s_in>>>(p->d_fe, p->d_scratch, p->scratch_stride, p->d_todo[buf], p->d_blob[buf],
                                                     p->d_off[buf], p->d_size[buf], n, (int)S, g->ncc, c->s_coeffs[buf],
                                                     c->s_ics[buf], c->s_tns[buf]);
            CK(cudaGetLastError());
            CK(cudaEventRecord(p->ev_a[buf], c->s_in));
            CK(cudaStreamWaitEvent(p->s_b, p->ev_a[buf], 0));
            k_parse_b<<<grid, threads, 0, p->s_b>>>(p->d_fe, p->d_todo[buf], p->d_tape, g->sbr ? p->tapes_per_frame : 0,
                                                     p->d_blob[buf], p->d_off[buf], p->d_size[buf], n,
                                                     (int)S, g->ncc, g->nse, g->sbr, g->ps, c->s_coeffs[buf], c->s_ics[buf],
                                                     c->s_tns[buf], c->s_sbr[buf], c->s_ps[buf], p->d_consumed[buf]);
            CK(cudaGetLastError());
            if (!p->timed[buf]) { CK(cudaEventRecord(p->ev_k1[buf], p->s_b)); p->timed[buf] = 1; }
            p->launches += 2;
        }
        CK(cudaEventRecord(c->ev_in[buf], p->s_b));
        if (di >= 0) CK(cudaEventRecord(p->dbg_ev[di][1], c->s_in));
        CK(cudaStreamWaitEvent(c->stream, c->ev_in[buf], 0));
        if (c->seq >= 2) CK(cudaStreamWaitEvent(c->stream, c->ev_out[buf], 0));
        int r = heaac_launch_frames(c, n, c->s_coeffs[buf], c->s_ics[buf], c->s_tns[buf], c->s_sbr[buf], c->s_ps[buf],
                                    c->s_pcm[buf], pcm_f32 ? c->s_pcm_f32[buf] : nullptr);
        if (r) return r;
        CK(cudaEventRecord(c->ev_done[buf], c->stream));
        if (di >= 0) CK(cudaEventRecord(p->dbg_ev[di][2], c->stream));
        const size_t pcm_n = np * g->out_len * g->nout, pcm_o = (size_t)f0 * S * g->out_len * g->nout;
        CK(cudaStreamWaitEvent(c->s_out, c->ev_done[buf], 0));
        if (pcm) CK(cudaMemcpyAsync(pcm + pcm_o, c->s_pcm[buf], pcm_n * sizeof(int16_t), cudaMemcpyDeviceToHost, c->s_out));
        if (pcm_f32) CK(cudaMemcpyAsync(pcm_f32 + pcm_o, c->s_pcm_f32[buf], pcm_n * sizeof(float), cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaMemcpyAsync(p->h_consumed[slot] + (size_t)f0 * S, p->d_consumed[buf], np * sizeof(int), cudaMemcpyDeviceToHost, c->s_out));
        CK(cudaEventRecord(c->ev_out[buf], c->s_out));
        if (di >= 0) CK(cudaEventRecord(p->dbg_ev[di][3], c->s_out));
        const double t3 = now();
        t_sync += t1 - t0; t_gather += t2 - t1; t_submit += t3 - t2;
    }
    p->user_consumed[slot] = consumed; p->user_n[slot] = (int)((size_t)n_frames * S);
    if (dbg) fprintf(stderr, "parser_submit: sync %.2f gather %.2f submit %.2f ms\n", t_sync, t_gather, t_submit);
    (void)t_wait;
    return heaac_b200_ticket(c);
}

extern "C" int heaac_b200_parser_wait(HeaacGpuParser *p, long long ticket)
{
    if (!p) return HEAAC_E_INVAL;
    int r = heaac_b200_ticket_wait(p->ctx, ticket);
    if (r) return r;
    const int slot = (int)(ticket & 7);
    if (p->ctx->n_tickets - ticket <= 8 && p->user_consumed[slot]) {      /* (an older ticket's consumed[] was delivered when its slot was reused) */
        memcpy(p->user_consumed[slot], p->h_consumed[slot], (size_t)p->user_n[slot] * sizeof(int));
        p->user_consumed[slot] = nullptr;
    }
    return 0;
}

extern "C" void heaac_b200_parser_debug_timeline(HeaacGpuParser *p)
{
    if (!p || !p->dbg_on) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < p->dbg_n; i++) {
        float t[4] = { 0, 0, 0, 0 };
        for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], p->dbg_base, p->dbg_ev[i][k]);
        fprintf(stderr, "chunk %2d: h2d done %7.2f  parse done %7.2f  synth done %7.2f  d2h done %7.2f ms\n", i, t[0], t[1], t[2], t[3]);
    }
    p->dbg_n = 0;
}

extern "C" int heaac_b200_parser_decode(HeaacGpuParser *p, int n_frames, const uint8_t *const *pkt, const int *size,
                                        int16_t *pcm, float *pcm_f32, int *consumed)
{
    long long t = heaac_b200_parser_submit(p, n_frames, pkt, size, pcm, pcm_f32, consumed);
    return t < 0 ? (int)t : heaac_b200_parser_wait(p, t);
}

/* ---- test hook: the device build of the sbr_reset table chain on arbitrary headers -------------------
 * params[i] = { start_freq, stop_freq, xover_band, freq_scale, alter_scale, noise_bands, limiter_bands, 0 };
 * out[i] / rc[i] = what fe_sbr_tables gives for it on the GPU.  Host pointers. */
__global__ void k_sbr_tables_probe(int core_rate, int n, const uint8_t *params, HeaacSbrElem *out, int *rc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = params + 8 * (size_t)i;
    HeaacSbrElem t;
    memset(&t, 0, sizeof(t));
    rc[i] = fe_sbr_tables(core_rate, p[0], p[1], p[2], p[3], p[4], p[5], p[6], &t);
    out[i] = t;
}

extern "C" int heaac_b200_sbr_tables_device(HeaacContext *c, int core_rate, int n, const uint8_t *params,
                                            HeaacSbrElem *out, int *rc)
{
    if (!c || n <= 0 || !params || !out || !rc) return HEAAC_E_INVAL;
    CK(cudaSetDevice(c->cfg.device));
    uint8_t *d_p = nullptr; HeaacSbrElem *d_o = nullptr; int *d_r = nullptr;
    int r = 0;
#define CKT(call) do { if (!r) r = heaac_cuda_err((call), #call); } while (0)
    CKT(cudaMalloc(&d_p, 8 * (size_t)n));
    CKT(cudaMalloc(&d_o, sizeof(HeaacSbrElem) * (size_t)n));
    CKT(cudaMalloc(&d_r, sizeof(int) * (size_t)n));
    CKT(cudaMemcpy(d_p, params, 8 * (size_t)n, cudaMemcpyHostToDevice));
    if (!r) k_sbr_tables_probe<<<(n + 127) / 128, 128, 0, c->stream>>>(core_rate, n, d_p, d_o, d_r);
    CKT(cudaGetLastError());
    CKT(cudaStreamSynchronize(c->stream));
    CKT(cudaMemcpy(out, d_o, sizeof(HeaacSbrElem) * (size_t)n, cudaMemcpyDeviceToHost));
    CKT(cudaMemcpy(rc, d_r, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
#undef CKT
    cudaFree(d_p); cudaFree(d_o); cudaFree(d_r);
    return r;
}
