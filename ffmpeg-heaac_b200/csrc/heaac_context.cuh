/*
 * heaac_context.cuh -- the back-end context shared by the translation units of the library
 * (heaac_api.cu: kernels + C-ABI, heaac_parse_gpu.cu: device front end).  Internal; not part of the ABI.
 */
#ifndef HEAAC_CONTEXT_CUH
#define HEAAC_CONTEXT_CUH
#include <cuda_runtime.h>
#include "heaac_dev.cuh"

/* ------------------------------------------------------------------------------------ */
struct HeaacContext {
    HeaacConfig cfg;
    HeaacGeom geom;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    HeaacDevTables *d_tables;
    HeaacChanState *d_chan;          /* [S][ncc] */
    HeaacSynthState *d_synth;        /* [S][nout] (SBR only) */
    HeaacPsState *d_ps;              /* [S] (PS only) */
    float *d_tns_done;               /* TNS-filtered coefficient ranges (k_tns -> k_core / k_chain), frame-major like coeffs */
    size_t tns_done_frames;          /* frames d_tns_done holds: 1 for AAC-LC (per frame tick), the call's frames with SBR */
    unsigned *d_work;                /* k_core's channel counter + finished-CTA counter */
    float *d_couple;                 /* [S][ncc][1024] decoded channels of one frame tick, planar (layouts with coupling channel elements) */
    HeaacCce *s_cce[2];              /* staging, like s_ics */
    float *d_X;                      /* [x_frames][S][nout][2][64][40]: X between k_chain, k_ps and k_synth (PS / downsampled SBR / keep_intermediates) */
    size_t x_frames;                 /* frames d_X holds (1, or the chain's frames per launch on the PS / downsampled paths) */
    /* staging for host-pointer batches, sized for max_frames_per_call */
    /* Host-pointer batches are streamed through two sets of device staging buffers of `chunk` frames each:
     * H2D of chunk k+1, the kernels of chunk k and D2H of chunk k-1 overlap on three streams. */
    int chunk;
    unsigned long long seq;            /* staged chunks so far (selects the staging set) */
    float *s_coeffs[2]; HeaacIcs *s_ics[2]; HeaacTns *s_tns[2]; HeaacSbrElem *s_sbr[2]; HeaacPsFrame *s_ps[2];
    int16_t *s_pcm[2]; float *s_pcm_f32[2];
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[2], ev_done[2], ev_out[2];
    cudaEvent_t ev_ticket[8];          /* completion marks on s_out, heaac_b200_ticket_* */
    long long n_tickets;
    long long launches;
    size_t smem_core, smem_chain, smem_ps, smem_synth;
    int core_grid;                   /* persistent k_core: resident CTAs on the whole GPU */
    /* optional per-kernel event timing (heaac_b200_profile_*) */
    int prof_on, prof_n, prof_cap;
    cudaEvent_t *prof_ev;            /* 2 events per recorded launch */
    unsigned char *prof_kind;        /* HEAAC_K_* per recorded launch */
};

enum { HEAAC_K_CORE = 0, HEAAC_K_CHAIN = 1, HEAAC_K_PS = 2, HEAAC_K_SYNTH = 3, HEAAC_K_COUNT = 4 };     /* (k_couple is timed with k_core) */

/* launches for frames [0, n) of a device-resident batch on c->stream (heaac_api.cu) */
int heaac_launch_frames(HeaacContext *c, int n_frames, const float *coeffs, const HeaacIcs *ics, const HeaacTns *tns,
                        const HeaacSbrElem *sbr, const HeaacPsFrame *ps, int16_t *pcm, float *pcm_f32, const HeaacCce *cce = nullptr);
int heaac_cuda_err(cudaError_t e, const char *what);
#define CK(call) do { int _r = heaac_cuda_err((call), #call); if (_r) return _r; } while (0)

#endif
