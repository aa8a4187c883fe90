/*
 * heaac_b200.h -- C-ABI of the B200-native HE-AAC synthesis back end.
 *
 * This library replaces ONE path of the reference decoder (aconverse/ffmpeg-heaac,
 * libavcodec 52.78): everything from spectral_to_sample() downwards
 *   libavcodec/aacdec.c:1903-1933  spectral_to_sample      (driver)
 *   libavcodec/aacdec.c:1698-1736  apply_tns
 *   libavcodec/aacdec.c:1741-1806  imdct_and_windowing
 *   libavcodec/aacsbr.c:1716-1771  ff_sbr_apply            (QMF analysis, HF generation/adjustment, QMF synthesis)
 *   libavcodec/aacps.c:973-992     ff_ps_apply             (hybrid analysis, decorrelation, stereo processing)
 *   libavcodec/dsputil.c:3972-4001 float_to_int16_interleave
 * for a BATCH of independent streams, on one B200.  Bitstream parsing stays on the host;
 * the structs below are exactly the per-frame outputs of the reference's parsers
 * (decode_ics / read_sbr_data / ff_ps_read_data) that the synthesis path consumes.
 *
 * Plain C, plain pointers and sizes; no CUDA, torch or C++ types appear in any signature.
 * All functions return 0 (or a non-negative count) on success and a negative HEAAC_E_* code
 * on failure -- the reference's own `int` convention (aacdec.c: -1 / AVERROR(x)).
 * One handle per GPU; calls on one handle must be serialised by the caller; different
 * handles are independent (same threading contract as AVCodecContext, utils.c:467-473).
 */
#ifndef HEAAC_B200_H
#define HEAAC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEAAC_ABI_VERSION 2

/* ---- error codes ---------------------------------------------------------------- */
#define HEAAC_OK            0
#define HEAAC_E_INVAL      (-22)   /* bad argument (EINVAL)                        */
#define HEAAC_E_NOMEM      (-12)   /* host or device allocation failed (ENOMEM)    */
#define HEAAC_E_NODEV      (-19)   /* no usable sm_100 device / CUDA runtime error */
#define HEAAC_E_LAUNCH     (-5)    /* kernel launch or execution failed (EIO)      */
#define HEAAC_E_UNSUPPORTED (-38)  /* valid request this build does not cover      */

/* ---- element / window enums (values identical to aac.h:46-71) ------------------- */
enum { HEAAC_TYPE_SCE = 0, HEAAC_TYPE_CPE = 1, HEAAC_TYPE_CCE = 2, HEAAC_TYPE_LFE = 3 };
enum { HEAAC_ONLY_LONG = 0, HEAAC_LONG_START = 1, HEAAC_EIGHT_SHORT = 2, HEAAC_LONG_STOP = 3 };

#define HEAAC_MAX_ELEMS     8      /* channel elements per stream in one layout      */
#define HEAAC_MAX_CHANNELS  16
#define HEAAC_FRAME_LEN     1024   /* core frame length (960 not supported, as in the reference) */
#define HEAAC_TNS_MAX_ORDER 20     /* aac.h:44 */

/* ---- per-frame side info: core (AAC-LC) channel ----------------------------------
 * Mirrors the fields of IndividualChannelStream (aac.h:135-150) and
 * TemporalNoiseShaping (aac.h:155-162) that apply_tns / imdct_and_windowing read.
 * The previous frame's window_sequence / window_shape are device state. */
typedef struct HeaacIcs {
    uint8_t window_sequence;   /* ics.window_sequence[0]                         */
    uint8_t window_shape;      /* ics.use_kb_window[0]: 1 = KBD, 0 = sine        */
    uint8_t max_sfb;           /* ics.max_sfb                                    */
    uint8_t tns_present;       /* tns.present                                    */
    /* Optional: the parser's own record of the previous frame (ics.window_sequence[1], ics.use_kb_window[1]).
     * When prev_valid == 0 the back end uses what it saw in the previous frame of this channel; a front end that
     * mirrors the reference's bookkeeping exactly (a CPE with common_window copies channel 0's previous sequence
     * to channel 1, aacdec.c:1467-1469) sets prev_valid = 1. */
    uint8_t prev_valid;
    uint8_t prev_window_sequence;
    uint8_t prev_window_shape;
    uint8_t reserved;
} HeaacIcs;

/* TNS filters in bitstream order.  Long blocks: window 0 holds up to 3 filters in slots
 * 0..2.  EIGHT_SHORT: window w holds at most one filter, in slot w (aacdec.c:854-887). */
typedef struct HeaacTns {
    uint8_t n_filt[8];                        /* tns.n_filt[w]                     */
    uint8_t length[8];                        /* tns.length[w][filt]               */
    uint8_t order[8];                         /* tns.order[w][filt]                */
    uint8_t direction[8];                     /* tns.direction[w][filt]            */
    float   coef[8][HEAAC_TNS_MAX_ORDER];     /* tns.coef[w][filt][]: reflection coefficients (tns_tmp2_map values) */
} HeaacTns;

/* ---- per-frame side info: SBR ------------------------------------------------------
 * Mirrors SBRData (sbr.h:59-107) bitstream fields as left by read_sbr_data
 * (aacsbr.c:982-1021) BEFORE sbr_dequant: env_q / noise_q are the integer-valued
 * env_facs[1..] / noise_facs[1..] rows. */
typedef struct HeaacSbrChan {
    uint8_t bs_num_env;
    uint8_t bs_num_noise;
    uint8_t bs_amp_res;
    uint8_t bs_add_harmonic_flag;
    int8_t  e_a[2];                 /* l_APrev, l_A (aacsbr.c:741-746)               */
    uint8_t t_env_num_env_old;
    uint8_t reserved0;
    uint8_t bs_freq_res[8];         /* [0] = last of previous frame, [e+1] per envelope */
    uint8_t t_env[8];
    uint8_t t_q[4];
    uint8_t bs_invf_mode[2][5];     /* [0] current, [1] previous frame               */
    uint8_t reserved1[2];
    uint8_t bs_add_harmonic[48];
    int16_t env_q[5][48];           /* env_facs[e+1][k], quantised                   */
    int16_t noise_q[2][5];          /* noise_facs[e+1][k], quantised                 */
    uint8_t reserved2[4];
} HeaacSbrChan;

/* Mirrors the SpectralBandReplication fields (sbr.h:112-156) that ff_sbr_apply reads:
 * header flags and the frequency tables derived by sbr_reset (aacsbr.c:1023-1034). */
typedef struct HeaacSbrElem {
    uint8_t start;                  /* sbr->start: 0 = pure upsampling mode          */
    uint8_t reset;                  /* sbr->reset                                    */
    uint8_t bs_coupling;
    uint8_t bs_limiter_gains;
    uint8_t bs_interpol_freq;
    uint8_t bs_smoothing_mode;
    uint8_t num_patches;
    uint8_t n_q;
    uint8_t n[2];                   /* N_low, N_high                                 */
    uint8_t n_lim;
    uint8_t k0;                     /* sbr->k[0]                                     */
    uint8_t kx[2];                  /* kx' (previous frame), kx                      */
    uint8_t m[2];                   /* M', M                                         */
    uint8_t f_tablelow[25];
    uint8_t f_tablehigh[49];
    uint8_t f_tablenoise[6];
    uint8_t f_tablelim[29];
    uint8_t patch_num_subbands[6];
    uint8_t patch_start_subband[6];
    uint8_t reserved[3];
    HeaacSbrChan ch[2];
} HeaacSbrElem;

/* ---- per-frame side info: Parametric Stereo ---------------------------------------
 * Mirrors the PSContext bitstream fields (aacps.h:40-63) after ff_ps_read_data. */
typedef struct HeaacPsFrame {
    uint8_t start;                  /* ps->start: 0 = copy L to R (aacsbr.c:1751-1756) */
    uint8_t num_env;
    uint8_t num_env_old;
    uint8_t iid_quant;
    uint8_t nr_iid_par;
    uint8_t nr_icc_par;
    uint8_t nr_ipdopd_par;
    uint8_t icc_mode;
    uint8_t enable_ipdopd;
    uint8_t is34bands;
    uint8_t is34bands_old;
    uint8_t reserved0;
    int8_t  border_position[6];
    uint8_t reserved1[2];
    int8_t  iid_par[5][34];
    int8_t  icc_par[5][34];
    int8_t  ipd_par[5][34];
    int8_t  opd_par[5][34];
} HeaacPsFrame;

/* ---- per-frame side info: channel coupling -------------------------------------------
 * A coupling channel element (HEAAC_TYPE_CCE in cfg.elem_type) is one more coded channel of the stream -- it takes a
 * coefficient / ICS / TNS slot like an SCE -- but has no output channel of its own: its decoded signal is ADDED to other
 * channels.  This struct mirrors what decode_cce leaves in ChannelCoupling (aac.h:165-177, aacdec.c:1503-1565) for an
 * INDEPENDENTLY SWITCHED element (coupling_point == AFTER_IMDCT), with the element / channel selection of
 * apply_channel_coupling (aacdec.c:1870-1898) already resolved by the front end into a list of axpy operations
 *     ret[target_ch[i]][n] += gain[i] * (ret[cce][n] - add_bias)        (apply_independent_coupling, aacdec.c:1849-1862)
 * applied in list order, CCEs in the order of their instance tags.  Dependent coupling (in the spectral domain, before or
 * after TNS) is not part of this ABI: a front end meeting it reports HEAAC_E_UNSUPPORTED. */
#define HEAAC_CCE_MAX_TARGETS 16
typedef struct HeaacCce {
    uint8_t coupling_point;         /* coup.coupling_point: must be 3 (AFTER_IMDCT); anything else = the element is ignored */
    uint8_t n_targets;              /* axpy operations of this frame                  */
    uint8_t reserved[2];
    uint8_t target_ch[HEAAC_CCE_MAX_TARGETS];   /* core channel index (cfg.elem_type order) */
    float   gain[HEAAC_CCE_MAX_TARGETS];        /* coup.gain[index][0]                      */
} HeaacCce;

/* ---- context configuration ---------------------------------------------------------
 * A context decodes `n_streams` independent streams that share one channel layout
 * (the BASELINE configs: stereo CPE, mono SCE + PS, 5.1 = SCE CPE CPE LFE, ...).
 * Channel c of a stream is the c-th core channel in element (bitstream) order; cfg.out_slot[c]
 * gives its position in the interleaved / planar output. */
typedef struct HeaacConfig {
    uint32_t abi_version;           /* HEAAC_ABI_VERSION                             */
    int32_t  device;                /* CUDA device ordinal                           */
    int32_t  n_streams;
    int32_t  n_elems;               /* elements per stream                           */
    uint8_t  elem_type[HEAAC_MAX_ELEMS]; /* HEAAC_TYPE_*                             */
    int32_t  sr_index;              /* MPEG-4 sampling frequency index of the CORE (aactab.c:1174-1207 tables) */
    int32_t  sbr;                   /* 1: every SCE/CPE element carries SBR (m4ac.sbr == 1) */
    int32_t  ps;                    /* 1: mono SCE + Parametric Stereo (m4ac.ps == 1) */
    /* Scaling convention chosen by the front end (aacdec.c:573-581):
     *   add_bias = 0,   sf_scale = -1/1024          -> samples in s16 units (SIMD path; default)
     *   add_bias = 385, sf_scale = -1/(1024*32768)  -> C "bias" path
     * coeffs[] must arrive already multiplied by sf_scale (as the reference's dequantiser does). */
    float    add_bias;
    float    sf_scale;
    int32_t  max_frames_per_call;   /* T: capacity of the device staging buffers     */
    int32_t  want_float_out;        /* also produce planar float PCM (debug / float parity) */
    int32_t  keep_intermediates;    /* keep per-stage arrays readable via heaac_b200_debug_read */
    /* Interleaved output position of core channel c.  The reference orders its output by
     * aac_channel_layout_map (aacdectab.h:74-82, output_configure aacdec.c:214-272), e.g. 5.1 =
     * {CPE0, SCE0, LFE0, CPE1} while the bitstream order is {SCE0, CPE0, CPE1, LFE0}.  All zero = identity. */
    uint8_t  out_slot[HEAAC_MAX_CHANNELS];
    int32_t  chunk_frames;          /* frames per staging chunk of the host-pointer path (1..32); 0 = automatic */
    int32_t  sbr_downsampled;       /* 1: downsampled SBR (m4ac.ext_sample_rate < 2 x core rate, aacsbr.c:1719): 32-band
                                     * synthesis QMF, 1024 output samples per frame at the core rate */
} HeaacConfig;

/* One call = `n_frames` consecutive frames of every stream.  Arrays are frame-major:
 *   coeffs [n_frames][n_streams][n_core_ch][1024]   float
 *   ics    [n_frames][n_streams][n_core_ch]
 *   tns    [n_frames][n_streams][n_core_ch]         (may be NULL if no channel has tns_present)
 *   sbr    [n_frames][n_streams][n_sbr_elems]       (NULL unless cfg.sbr)
 *   ps     [n_frames][n_streams]                    (NULL unless cfg.ps)
 *   pcm    [n_frames][n_streams][frame_out_len][n_out_ch]  int16, interleaved
 *   pcm_f32[n_frames][n_streams][n_out_ch][frame_out_len]  float, planar (optional)
 *   cce    [n_frames][n_streams][n_cce]             (NULL unless the layout has coupling channel elements)
 * frame_out_len = 1024 (no SBR) or 2048 (SBR).  n_out_ch = n_core_ch, or 2 for ps.
 * `on_device` != 0 means every pointer is a CUDA device pointer on cfg.device (no copies);
 * otherwise they are host pointers and the call copies H2D / D2H itself. */
typedef struct HeaacBatch {
    int32_t n_frames;
    int32_t on_device;
    const float        *coeffs;
    const HeaacIcs     *ics;
    const HeaacTns     *tns;
    const HeaacSbrElem *sbr;
    const HeaacPsFrame *ps;
    int16_t            *pcm;
    float              *pcm_f32;
    const struct HeaacCce *cce;     /* [n_frames][n_streams][n_cce elements] (NULL unless cfg.elem_type has HEAAC_TYPE_CCE) */
} HeaacBatch;

typedef struct HeaacContext HeaacContext;

/* Lifetime.  create: replaces aac_decode_init's back-end half (aacdec.c:537-600: ff_mdct_init x2,
 * window tables, ff_aac_sbr_init tables) + che_configure/ff_aac_sbr_ctx_init per stream
 * (aacdec.c:192-214, aacsbr.c:128-137).  State starts as the reference's post-init state
 * (kx' = 32, e_a[1] = -1, all histories zero). */
int  heaac_b200_create(const HeaacConfig *cfg, HeaacContext **out);
void heaac_b200_destroy(HeaacContext *ctx);
/* Re-initialise all per-stream state (as if freshly created). */
int  heaac_b200_reset_state(HeaacContext *ctx);

/* Layout queries derived from the config. */
int  heaac_b200_n_core_channels(const HeaacContext *ctx);
int  heaac_b200_n_out_channels(const HeaacContext *ctx);
int  heaac_b200_frame_out_len(const HeaacContext *ctx);
int  heaac_b200_n_sbr_elems(const HeaacContext *ctx);

/* The hot path: replaces spectral_to_sample + float_to_int16_interleave
 * (aacdec.c:2078, 2096-2097) for n_frames x n_streams frames.  Asynchronous on the
 * context's stream when on_device != 0; heaac_b200_sync waits.  Host-pointer calls
 * return after the PCM has landed in host memory. */
int  heaac_b200_synth(HeaacContext *ctx, const HeaacBatch *batch);
int  heaac_b200_sync(HeaacContext *ctx);
/* The same call split for pipelined callers: submit() only enqueues (host-pointer batches are streamed
 * through the staging sets chunk by chunk); inputs_consumed() returns once every host INPUT buffer submitted so
 * far has been read and may be overwritten; wait() returns once all PCM submitted so far is in host memory.
 * heaac_b200_synth == submit + wait for host-pointer batches. */
int  heaac_b200_synth_submit(HeaacContext *ctx, const HeaacBatch *batch);
int  heaac_b200_inputs_consumed(HeaacContext *ctx);
int  heaac_b200_wait(HeaacContext *ctx);
/* A ticket marks "all host-pointer work submitted so far": ticket_wait(t) returns once everything submitted before
 * ticket t was taken has its PCM in host memory, without waiting for later submissions (8 may be outstanding). */
long long heaac_b200_ticket(HeaacContext *ctx);
int  heaac_b200_ticket_wait(HeaacContext *ctx, long long ticket);
/* frames per staging chunk of the host-pointer path (derived from max_frames_per_call) */
int  heaac_b200_chunk_frames(const HeaacContext *ctx);
/* cfg.out_slot[] for the reference's output channel order (aac_channel_layout_map, aacdectab.h:74-82;
 * output_configure, aacdec.c:214-272); identity outside the default channel configurations.  Returns the
 * number of core channels. */
int  heaac_b200_reference_out_slots(const uint8_t *elem_type, int n_elems, uint8_t out_slot[HEAAC_MAX_CHANNELS]);

/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
long long heaac_b200_launch_count(const HeaacContext *ctx);
/* The cudaStream_t the context launches on, as an integer handle (for event timing). */
uintptr_t heaac_b200_stream(const HeaacContext *ctx);

/* Device memory helpers so a C/ctypes caller can stage device-resident batches without
 * linking the CUDA runtime itself. */
int  heaac_b200_dev_alloc(HeaacContext *ctx, size_t bytes, void **dptr);
int  heaac_b200_dev_free(HeaacContext *ctx, void *dptr);
int  heaac_b200_dev_upload(HeaacContext *ctx, void *dptr, const void *hptr, size_t bytes);
int  heaac_b200_dev_download(HeaacContext *ctx, void *hptr, const void *dptr, size_t bytes);
/* page-locked host memory (what the H2D / D2H copies of the host-pointer path overlap best with) */
int  heaac_b200_host_alloc(HeaacContext *ctx, size_t bytes, void **hptr);
int  heaac_b200_host_free(HeaacContext *ctx, void *hptr);

/* Event timing on the context's stream (CUDA events; milliseconds). */
int  heaac_b200_timer_start(HeaacContext *ctx);
int  heaac_b200_timer_stop(HeaacContext *ctx, float *ms);

/* Per-kernel event timing on the context's stream.  enable(capacity > 0) records up to `capacity`
 * launches (0 disables); read() synchronises, returns total ms and launch counts per kernel kind
 * (0 core, 1 sbr, 2 ps, 3 synth) and clears the record. */
int  heaac_b200_profile_enable(HeaacContext *ctx, int capacity);
int  heaac_b200_profile_read(HeaacContext *ctx, float ms[4], long long counts[4]);

/* Stage-level inspection (requires cfg.keep_intermediates).  Copies the named array of
 * the LAST frame processed to host.  Names and layouts: see DESIGN.md "intermediates". */
int  heaac_b200_debug_read(HeaacContext *ctx, const char *name, void *dst, size_t bytes);

/* ---- host front-end helper: SBR frequency tables -------------------------------------------
 * What the reference's sbr_reset() derives when a header changes (aacsbr.c:1023-1034 ->
 * sbr_make_f_master 303-488, sbr_make_f_derived 541-593, sbr_hf_calc_npatches 490-539,
 * sbr_make_f_tablelim 159-205).  Pure host code, no GPU needed.  Fills k0, kx[1], m[1], n[], n_q,
 * n_lim, f_table*, patch_* of *out; returns HEAAC_E_INVAL where the reference rejects the header. */
int  heaac_b200_sbr_tables(int core_sample_rate, int bs_start_freq, int bs_stop_freq, int bs_xover_band,
                           int bs_freq_scale, int bs_alter_scale, int bs_noise_bands, int bs_limiter_bands,
                           HeaacSbrElem *out);

/* ---- host front end (own C): packets -> per-frame structures -------------------------------
 * One HeaacFrontend per stream plays the role of the reference's parsers up to the spectral_to_sample()
 * call site (aacdec.c:1973-2076 element loop, decode_ics / decode_cpe, ff_decode_sbr_extension,
 * ff_ps_read_data): heaac_fe_parse() takes one packet (bare raw_data_block, or an ADTS frame) and fills the
 * frame's coeffs / HeaacIcs / HeaacTns / HeaacSbrElem / HeaacPsFrame, ready for heaac_b200_synth().
 * Pure host code (no GPU); different front ends may be driven from different threads.
 * extradata = AudioSpecificConfig (may be NULL for ADTS); bias_mode selects the scaling convention
 * (0: s16 units, sf_scale = -1/1024; 1: the C "bias" path, aacdec.c:573-576). */
typedef struct HeaacFrontend HeaacFrontend;
int  heaac_fe_open(HeaacFrontend **fe, const uint8_t *extradata, int extradata_size, int sample_rate, int channels,
                   int bias_mode);
void heaac_fe_close(HeaacFrontend *fe);
/* returns bytes consumed (>= 0) or a negative HEAAC_E_* code; sbr / ps / tns may be NULL.  Arrays are indexed by
 * core channel (coeffs [n_core_ch][1024], ics, tns) and by element (sbr); a stream configured in-band (ADTS)
 * can change layout inside the call, so size them for 16 channels / HEAAC_MAX_ELEMS elements in that case. */
int  heaac_fe_parse(HeaacFrontend *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                    HeaacSbrElem *sbr, HeaacPsFrame *ps);
/* heaac_fe_parse computed the way the device front end does it (a stateless pass over the core elements, then a
 * sequential pass over the SBR / PS payloads); same results, kept callable for tests.  tns must not be NULL. */
int  heaac_fe_parse_two_phase(HeaacFrontend *fe, const uint8_t *pkt, int size, float *coeffs, HeaacIcs *ics, HeaacTns *tns,
                              HeaacSbrElem *sbr, HeaacPsFrame *ps);
/* The device front end's treatment of a chunk of up to 32 consecutive frames of one stream (entropy decoding of the core elements
 * frame by frame without the stream's history, the SBR payloads of all frames read side by side, then the frames finished in
 * order), on the host: same results as n_frames calls of heaac_fe_parse, kept callable for tests.  Frame-major arrays for the
 * stream's layout (which must be known: extradata); consumed[f] < 0 marks a frame that failed (its outputs are zeroed). */
int  heaac_fe_parse_chunk(HeaacFrontend *fe, int n_frames, const uint8_t *const *pkt, const int *size, float *coeffs,
                          HeaacIcs *ics, HeaacTns *tns, HeaacSbrElem *sbr, HeaacPsFrame *ps, int *consumed);
/* how the calling thread's last heaac_fe_parse_chunk went: out[0] = passes of the SBR payloads' side-by-side step, out[1] = frames
 * that went through the one-pass parser, out[2] = of those: because their header changes how SBR data is read */
void heaac_fe_chunk_stats(int out[4]);
/* 1 if the stream signals downsampled SBR (extension rate below twice the core rate): HeaacConfig.sbr_downsampled */
int  heaac_fe_sbr_downsampled(const HeaacFrontend *fe);
/* copy sample rate and channel layout (not the stream state) from a front end that has seen a frame */
int  heaac_fe_copy_layout(HeaacFrontend *dst, const HeaacFrontend *src);
/* batch use: copy_layout and FIX the layout -- afterwards a packet whose ADTS header announces another channel
 * configuration or sampling-rate index fails with HEAAC_E_INVAL before anything is written (the caller's per-stream
 * output slots are sized for the fixed layout; the reference would reconfigure, aacdec.c:1994-2006) */
int  heaac_fe_fix_layout(HeaacFrontend *dst, const HeaacFrontend *src);
/* layout of the stream: returns the number of elements and fills the optional outputs; sbr / ps reflect
 * implicit signalling once the first frame has been parsed (aacdec.c:1659-1676) */
int  heaac_fe_info(const HeaacFrontend *fe, int *n_core_ch, int *n_sbr_elems, int *sbr, int *ps, int *sr_index,
                   uint8_t elem_type[HEAAC_MAX_ELEMS]);
/* Layouts given as a program config element (decode_pce, aacdec.c:280-340: AudioSpecificConfig with channel_configuration 0,
 * or in band in the first raw_data_block) may hold coupling channel elements (decode_cce, aacdec.c:1503-1568).  A CCE takes a
 * core channel slot like an SCE (HEAAC_TYPE_CCE in elem_type); heaac_fe_cce() returns the layout's number of CCEs and copies
 * the coupling side info of the frame heaac_fe_parse has just returned -- one HeaacCce per CCE in layout order, with the
 * target selection of apply_channel_coupling (aacdec.c:1870-1898) resolved -- which is what HeaacBatch.cce takes.  Only
 * independently switched coupling is inside the ABI; a frame with another kind fails with HEAAC_E_UNSUPPORTED.
 * heaac_fe_layout_from_pce() = 1: elements and output channels are in the reference's (instance tag, type) order
 * (output_configure, aacdec.c:246-262) and HeaacConfig.out_slot is the identity. */
int  heaac_fe_cce(const HeaacFrontend *fe, struct HeaacCce *out);
int  heaac_fe_layout_from_pce(const HeaacFrontend *fe);

/* ---- containers (host C): ADTS framing and the MP4 `esds` descriptor ------------------------------------
 * adts_write_header: 7-byte header (no CRC) for one raw_data_block, as adts_write_frame_header (adtsenc.c:84-147);
 * adts_parse_header: ff_aac_parse_header (aac_parser.c:29-67); returns the header size (7 / 9) or < 0;
 * adts_split: offsets and sizes of back-to-back ADTS frames in a buffer (a .aac file), returns the count;
 * esds_find_asc: ES_Descriptor -> DecoderConfigDescriptor -> DecoderSpecificInfo (isom.c ff_mp4_read_descr*): the
 * AudioSpecificConfig to pass as extradata. */
int  heaac_adts_write_header(uint8_t out[7], int sr_index, int chan_config, int payload_bytes);
int  heaac_adts_parse_header(const uint8_t *buf, int size, int *sr_index, int *chan_config, int *frame_bytes, int *n_blocks);
int  heaac_adts_split(const uint8_t *buf, int size, int *offset, int *frame_size, int max_frames);
int  heaac_esds_find_asc(const uint8_t *esds, int size, const uint8_t **asc, int *asc_size);

/* ---- batched decoder: packets -> PCM ------------------------------------------------------------
 * S front ends + one back end behind the reference's AVCodec entry points (aacdec.c:2128-2142):
 *   heaac_dec_open   ~ init   (aac_decode_init,  aacdec.c:537-598)
 *   heaac_dec_decode ~ decode (aac_decode_frame, aacdec.c:1962-2106) for n_frames x n_streams packets
 *   heaac_dec_close  ~ close  (aac_decode_close, aacdec.c:2108-2124)
 * All streams of a decoder share one configuration (extradata, or ADTS headers that agree).  Entropy decoding
 * runs on `n_threads` host threads and overlaps the GPU work of the previous chunk of frames.
 * pkt / pkt_size / consumed are frame-major [n_frames][n_streams]; consumed[i] = bytes used or a negative
 * HEAAC_E_* for a packet that failed to parse (that frame is then decoded as silence; the call goes on).
 * pcm: interleaved s16 [n_frames][n_streams][frame_out_len][n_out_ch] (caller-owned, page-locked preferred);
 * pcm_f32 (optional, needs want_float_out): planar float [n_frames][n_streams][n_out_ch][frame_out_len]. */
typedef struct HeaacDecoder HeaacDecoder;
typedef struct HeaacDecoderConfig {
    uint32_t abi_version;           /* HEAAC_ABI_VERSION */
    int32_t  device;
    int32_t  n_streams;
    int32_t  sample_rate;           /* AVCodecContext.sample_rate / channels: used only without extradata */
    int32_t  channels;
    int32_t  max_frames_per_call;
    int32_t  n_threads;             /* host parser threads; 0 = one per hardware thread */
    int32_t  want_float_out;
    const uint8_t *extradata;       /* AudioSpecificConfig, or NULL for ADTS input */
    int32_t  extradata_size;
    int32_t  gpu_parse;             /* 1: entropy decoding on the GPU too (heaac_b200_parser_*): only packets cross PCIe */
    int32_t  chunk_frames;          /* frames per pipeline chunk (parse | copy | synthesise overlap chunk-wise); 0 = automatic */
} HeaacDecoderConfig;
int  heaac_dec_open(const HeaacDecoderConfig *cfg, HeaacDecoder **dec);
void heaac_dec_close(HeaacDecoder *dec);
int  heaac_dec_decode(HeaacDecoder *dec, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                      int16_t *pcm, float *pcm_f32, int *consumed);
/* The same call split for streaming callers: submit() returns as soon as the packets have been taken over (parsed,
 * or copied for the GPU parser) with a ticket >= 0; pcm / pcm_f32 / consumed are complete once wait(ticket)
 * returns.  Up to 8 calls may be outstanding, each with its own output buffers.  decode == submit + wait. */
long long heaac_dec_submit(HeaacDecoder *dec, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                           int16_t *pcm, float *pcm_f32, int *consumed);
int  heaac_dec_wait(HeaacDecoder *dec, long long ticket);
/* returns 1 once the first call has configured the decoder (0 before), and fills the optional outputs */
int  heaac_dec_info(const HeaacDecoder *dec, int *n_out_channels, int *frame_out_len, int *sbr, int *ps,
                    long long *parse_errors, long long *gpu_launches);
/* gpu_parse decoders: average k_parse duration in ms per launch (0 otherwise) */
double heaac_dec_parse_kernel_ms(const HeaacDecoder *dec);

/* ---- device front end: the parser of heaac_fe_* compiled for the GPU, one thread per stream ---------
 * parser_create takes a host front end that was opened for the streams' configuration and not used yet (its
 * state becomes the initial state of every stream); parser_decode is heaac_dec_decode's GPU-parse path:
 * packets are gathered into page-locked memory, copied, parsed by k_parse straight into the back end's staging
 * sets, synthesised, and the PCM copied back; chunks of frames are pipelined across three streams.
 * max_packet_bytes 0 = the AAC limit of 768 bytes per channel. */
typedef struct HeaacGpuParser HeaacGpuParser;
int  heaac_b200_parser_create(HeaacContext *ctx, const HeaacFrontend *fe, size_t max_packet_bytes, HeaacGpuParser **out);
void heaac_b200_parser_destroy(HeaacGpuParser *p);
int  heaac_b200_parser_reset(HeaacGpuParser *p);
int  heaac_b200_parser_decode(HeaacGpuParser *p, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                              int16_t *pcm, float *pcm_f32, int *consumed);
/* asynchronous form: submit returns a ticket (>= 0) once the packets are copied; outputs are complete after wait */
long long heaac_b200_parser_submit(HeaacGpuParser *p, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                                   int16_t *pcm, float *pcm_f32, int *consumed);
int  heaac_b200_parser_wait(HeaacGpuParser *p, long long ticket);
long long heaac_b200_parser_launch_count(const HeaacGpuParser *p);
/* average k_parse duration in ms (CUDA events on its stream) and how many launches that average covers */
int  heaac_b200_parser_stats(const HeaacGpuParser *p, double *ms_per_launch, long long *timed_launches);

/* test hook: heaac_b200_sbr_tables evaluated by the DEVICE build of the same source for n headers at once;
 * params[i] = { start_freq, stop_freq, xover_band, freq_scale, alter_scale, noise_bands, limiter_bands, 0 }
 * (host pointers).  tests hold it to the host result over the whole header space. */
int  heaac_b200_sbr_tables_device(HeaacContext *ctx, int core_sample_rate, int n, const uint8_t *params,
                                  HeaacSbrElem *out, int *rc);

/* ---- batch-of-one shims with the reference's own transform signatures -----------------
 * ff_mdct_init / ff_imdct_half / ff_mdct_end (fft.h:44-48,166-186; mdct.c:61,124,226). */
typedef struct HeaacMdct HeaacMdct;
int  heaac_b200_mdct_init(HeaacMdct **s, int nbits, int inverse, double scale);
void heaac_b200_imdct_half(HeaacMdct *s, float *output, const float *input);  /* host pointers, N/2 in, N/2 out */
void heaac_b200_mdct_end(HeaacMdct *s);

/* ff_sbr_apply / ff_ps_apply (aacsbr.h:45-46, aacps.h:80) with plain arrays; the state the reference keeps inside
 * SpectralBandReplication / PSContext between calls lives in the shim (one stream).
 *   sbr_init  ~ ff_aac_sbr_ctx_init (+ ff_ps_ctx_init when ps): id_aac = HEAAC_TYPE_SCE / _CPE, core sampling index,
 *               scaling convention as in HeaacConfig
 *   sbr_apply ~ ff_sbr_apply(ac, sbr, id_aac, L, R): L (and R for a CPE) = the element's ret[]: 1024 core samples in,
 *               2048 output samples out; with PS, R is output only.  elem / psf = this frame's side info.
 *   ps_apply  ~ ff_ps_apply(avctx, ps, L, R, top): L[2][38][64] in place for slots 0..31, R[2][38][64] out. */
typedef struct HeaacSbrShim HeaacSbrShim;
int  heaac_b200_sbr_init(HeaacSbrShim **s, int id_aac, int core_sr_index, int ps, float sf_scale, float add_bias);
int  heaac_b200_sbr_apply(HeaacSbrShim *s, const HeaacSbrElem *elem, const HeaacPsFrame *psf, float *L, float *R);
int  heaac_b200_ps_apply(HeaacSbrShim *s, const HeaacPsFrame *psf, float *L, float *R, int top);
void heaac_b200_sbr_close(HeaacSbrShim *s);

/* sizeof() of an ABI or device-state struct by name ("HeaacSbrElem", ...), -1 if unknown; lets
 * language bindings assert that their struct mirrors match this build. */
int  heaac_b200_sizeof(const char *name);

const char *heaac_b200_version(void);
const char *heaac_b200_strerror(int err);

#ifdef __cplusplus
}
#endif
#endif /* HEAAC_B200_H */
