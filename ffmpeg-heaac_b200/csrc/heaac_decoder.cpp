/*
 * heaac_decoder.cpp -- batched packets -> PCM decoder: S host front ends + one GPU back end.
 *
 * Plays the role of the reference's AVCodec entry points for a batch of independent streams:
 *   heaac_dec_open    ~ aac_decode_init   (aacdec.c:537-598)
 *   heaac_dec_decode  ~ aac_decode_frame  (aacdec.c:1962-2106), for n_frames x n_streams packets per call
 *   heaac_dec_close   ~ aac_decode_close  (aacdec.c:2108-2124)
 * The element loop / entropy decoding runs on host threads (heaac_frontend.c, one HeaacFrontend per stream, streams
 * split across a persistent worker pool); everything from spectral_to_sample() down runs on the GPU
 * (heaac_b200_synth_submit).  Parsing of chunk k+1 overlaps the H2D / kernels / D2H of chunk k: two page-locked
 * parameter sets, heaac_b200_inputs_consumed() as the hand-back point.
 *
 * No CUDA calls in this file: the device is reached through the C-ABI only.
 */
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/heaac_b200.h"

namespace {

struct ParamSet {                    /* one chunk of parsed frames, frame-major [t][s]..., page-locked */
    float *coeffs = nullptr;
    HeaacIcs *ics = nullptr;
    HeaacTns *tns = nullptr;
    HeaacSbrElem *sbr = nullptr;
    HeaacPsFrame *ps = nullptr;
    HeaacCce *cce = nullptr;         /* layouts with coupling channel elements */
};

struct Job {                         /* what the pool is working on */
    const uint8_t *const *pkt = nullptr;
    const int *size = nullptr;
    int *consumed = nullptr;
    int n_frames = 0;                /* frames in this chunk */
    ParamSet *set = nullptr;
};

}  // namespace

struct HeaacDecoder {
    HeaacDecoderConfig cfg;
    std::vector<uint8_t> extradata;
    std::vector<HeaacFrontend *> fe;
    HeaacContext *be = nullptr;
    HeaacGpuParser *gpu = nullptr;                    /* cfg.gpu_parse: the device front end replaces fe[] + workers */
    std::vector<int> scratch_consumed[8];
    int *pending_consumed[8] = { nullptr };
    size_t pending_n[8] = { 0 };
    long long pending_ticket[8] = { 0 };
    long long n_submitted = 0, n_chunks = 0;
    int ncc = 0, nse = 0, sbr = 0, ps = 0, sr_index = 0, n_elems = 0, n_cce = 0, chunk = 1;
    uint8_t elem_type[HEAAC_MAX_ELEMS] = { 0 };
    ParamSet set[2];
    std::atomic<long long> parse_errors{0};

    /* worker pool */
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    Job job;
    unsigned long long generation = 0;
    int pending = 0;
    bool quit = false;
    bool configured = false;                          /* set only when configure() succeeded as a whole */
};

namespace {

/* a frame that failed to parse is replaced by digital silence so that the lock-step batch can go on; the error is
 * reported through consumed[] (the reference returns -1 for such a packet and emits nothing, aacdec.c:2064-2070) */
void silent_frame(HeaacDecoder *d, ParamSet *p, size_t idx)
{
    memset(p->coeffs + idx * d->ncc * 1024, 0, sizeof(float) * d->ncc * 1024);
    memset(p->ics + idx * d->ncc, 0, sizeof(HeaacIcs) * d->ncc);
    memset(p->tns + idx * d->ncc, 0, sizeof(HeaacTns) * d->ncc);
    if (d->sbr) memset(p->sbr + idx * d->nse, 0, sizeof(HeaacSbrElem) * d->nse);
    if (d->ps)  memset(p->ps + idx, 0, sizeof(HeaacPsFrame));
    if (d->n_cce) memset(p->cce + idx * d->n_cce, 0, sizeof(HeaacCce) * d->n_cce);
}

void parse_range(HeaacDecoder *d, const Job &j, int s0, int s1)
{
    const size_t S = d->cfg.n_streams;
    for (int s = s0; s < s1; s++) {
        for (int t = 0; t < j.n_frames; t++) {
            const size_t idx = (size_t)t * S + s;
            ParamSet *p = j.set;
            int r = heaac_fe_parse(d->fe[s], j.pkt[idx], j.size[idx], p->coeffs + idx * d->ncc * 1024, p->ics + idx * d->ncc,
                                   p->tns + idx * d->ncc, d->sbr ? p->sbr + idx * d->nse : nullptr, d->ps ? p->ps + idx : nullptr);
            if (r < 0) { silent_frame(d, p, idx); d->parse_errors++; }
            else if (d->n_cce) heaac_fe_cce(d->fe[s], p->cce + idx * d->n_cce);
            if (j.consumed) j.consumed[idx] = r;
        }
    }
}

void worker_main(HeaacDecoder *d, int me, int n)
{
    unsigned long long seen = 0;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_go.wait(lk, [&] { return d->quit || d->generation != seen; });
            if (d->quit) return;
            seen = d->generation;
            j = d->job;
        }
        const int S = d->cfg.n_streams;
        const int s0 = (int)((long long)S * me / n), s1 = (int)((long long)S * (me + 1) / n);
        parse_range(d, j, s0, s1);
        {
            std::lock_guard<std::mutex> lk(d->mu);
            if (--d->pending == 0) d->cv_done.notify_all();
        }
    }
}

void run_job(HeaacDecoder *d, const Job &j)
{
    if (d->workers.empty()) { parse_range(d, j, 0, d->cfg.n_streams); return; }
    std::unique_lock<std::mutex> lk(d->mu);
    d->job = j;
    d->pending = (int)d->workers.size();
    d->generation++;
    d->cv_go.notify_all();
    d->cv_done.wait(lk, [&] { return d->pending == 0; });
}

void free_sets(HeaacDecoder *d)
{
    for (auto &p : d->set) {
        if (!d->be) break;
        if (p.coeffs) heaac_b200_host_free(d->be, p.coeffs);
        if (p.ics) heaac_b200_host_free(d->be, p.ics);
        if (p.tns) heaac_b200_host_free(d->be, p.tns);
        if (p.sbr) heaac_b200_host_free(d->be, p.sbr);
        if (p.cce) heaac_b200_host_free(d->be, p.cce);
        if (p.ps) heaac_b200_host_free(d->be, p.ps);
        p = ParamSet();
    }
}

/* Stream layout, SBR and PS presence are known for certain only after one frame has been parsed (ADTS
 * configuration, implicit SBR / PS signalling: aacdec.c:1659-1676, 1994-2006): parse the first packet of stream 0
 * with a scratch front end, then build the back end to match. */
int configure(HeaacDecoder *d, const uint8_t *pkt, int size)
{
    HeaacFrontend *probe = nullptr;
    const uint8_t *ed = d->extradata.empty() ? nullptr : d->extradata.data();
    int r = heaac_fe_open(&probe, ed, (int)d->extradata.size(), d->cfg.sample_rate, d->cfg.channels, 0);
    if (r) return r;
    {
        std::vector<float> co(16 * 1024);
        std::vector<HeaacIcs> ic(16);
        std::vector<HeaacTns> tn(16);
        std::vector<HeaacSbrElem> sb(HEAAC_MAX_ELEMS);
        HeaacPsFrame ps;
        r = heaac_fe_parse(probe, pkt, size, co.data(), ic.data(), tn.data(), sb.data(), &ps);
    }
    if (r >= 0) {
        d->n_elems = heaac_fe_info(probe, &d->ncc, &d->nse, &d->sbr, &d->ps, &d->sr_index, d->elem_type);
        d->n_cce = heaac_fe_cce(probe, nullptr);
        r = d->n_elems > 0 ? 0 : HEAAC_E_INVAL;
    }
    if (r < 0) { heaac_fe_close(probe); return r; }
    struct ProbeGuard { HeaacFrontend *p; ~ProbeGuard() { heaac_fe_close(p); } } guard{probe};

    HeaacConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = HEAAC_ABI_VERSION;
    cfg.device = d->cfg.device;
    cfg.n_streams = d->cfg.n_streams;
    cfg.n_elems = d->n_elems;
    memcpy(cfg.elem_type, d->elem_type, sizeof(cfg.elem_type));
    cfg.sr_index = d->sr_index;
    cfg.sbr = d->sbr;
    cfg.ps = d->ps;
    cfg.sbr_downsampled = heaac_fe_sbr_downsampled(probe);
    cfg.add_bias = 0.0f;
    cfg.sf_scale = -1.0f / 1024.0f;                   /* s16-unit convention, aacdec.c:577-581 */
    cfg.max_frames_per_call = d->cfg.max_frames_per_call;
    cfg.want_float_out = d->cfg.want_float_out;
    /* the device front end takes channel_configuration layouts; streams with a program config element are parsed on the host */
    if (heaac_fe_layout_from_pce(probe)) d->cfg.gpu_parse = 0;
    /* the device parser spreads a chunk's frames over the lanes of a warp: longer chunks, more lanes at work */
    if (d->cfg.gpu_parse && d->cfg.max_frames_per_call >= 32) cfg.chunk_frames = d->cfg.max_frames_per_call >= 64 ? 32 : 16;
    if (d->cfg.chunk_frames > 0) cfg.chunk_frames = d->cfg.chunk_frames;
    /* channel_configuration layouts: the reference's speaker order; program config element layouts: the elements already are
     * in the reference's order (out_slot all zero = output channels in element order) */
    if (!heaac_fe_layout_from_pce(probe)) heaac_b200_reference_out_slots(cfg.elem_type, cfg.n_elems, cfg.out_slot);
    if ((r = heaac_b200_create(&cfg, &d->be))) return r;

    d->chunk = heaac_b200_chunk_frames(d->be);
    if (d->cfg.gpu_parse) {
        HeaacFrontend *f = nullptr;
        if ((r = heaac_fe_open(&f, ed, (int)d->extradata.size(), d->cfg.sample_rate, d->cfg.channels, 0))) return r;
        /* every stream starts from the layout packet 0 announced, and is held to it: the staging slots are sized for it */
        if ((r = heaac_fe_fix_layout(f, probe))) { heaac_fe_close(f); return r; }
        r = heaac_b200_parser_create(d->be, f, 0, &d->gpu);
        heaac_fe_close(f);
        return r;
    }
    const size_t n = (size_t)d->chunk * d->cfg.n_streams;
    for (auto &p : d->set) {
        if (heaac_b200_host_alloc(d->be, n * d->ncc * 1024 * sizeof(float), (void **)&p.coeffs) ||
            heaac_b200_host_alloc(d->be, n * d->ncc * sizeof(HeaacIcs), (void **)&p.ics) ||
            heaac_b200_host_alloc(d->be, n * d->ncc * sizeof(HeaacTns), (void **)&p.tns) ||
            (d->sbr && heaac_b200_host_alloc(d->be, n * d->nse * sizeof(HeaacSbrElem), (void **)&p.sbr)) ||
            (d->ps && heaac_b200_host_alloc(d->be, n * sizeof(HeaacPsFrame), (void **)&p.ps)) ||
            (d->n_cce && heaac_b200_host_alloc(d->be, n * d->n_cce * sizeof(HeaacCce), (void **)&p.cce)))
            return HEAAC_E_NOMEM;
    }
    for (int s = 0; s < d->cfg.n_streams; s++) {
        HeaacFrontend *f = nullptr;
        if ((r = heaac_fe_open(&f, ed, (int)d->extradata.size(), d->cfg.sample_rate, d->cfg.channels, 0))) return r;
        d->fe.push_back(f);
        if ((r = heaac_fe_fix_layout(f, probe))) return r;     /* a packet announcing another layout is rejected, not decoded */
    }
    int nt = d->cfg.n_threads > 0 ? d->cfg.n_threads : (int)std::thread::hardware_concurrency();
    if (nt > d->cfg.n_streams) nt = d->cfg.n_streams;
    if (nt > 1)
        for (int i = 0; i < nt; i++) d->workers.emplace_back(worker_main, d, i, nt);
    return 0;
}

/* undo a configure() that failed half way: the next call starts from scratch instead of running on a half-built decoder */
void unconfigure(HeaacDecoder *d)
{
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->quit = true;
        d->cv_go.notify_all();
    }
    for (auto &t : d->workers) t.join();
    d->workers.clear();
    d->quit = false;
    if (d->gpu) { heaac_b200_parser_destroy(d->gpu); d->gpu = nullptr; }
    free_sets(d);
    for (auto f : d->fe) heaac_fe_close(f);
    d->fe.clear();
    if (d->be) { heaac_b200_destroy(d->be); d->be = nullptr; }
    d->configured = false;
}

}  // namespace

extern "C" int heaac_dec_open(const HeaacDecoderConfig *cfg, HeaacDecoder **out)
{
    if (!cfg || !out || cfg->abi_version != HEAAC_ABI_VERSION || cfg->n_streams <= 0 || cfg->max_frames_per_call <= 0)
        return HEAAC_E_INVAL;
    if (cfg->extradata_size < 0 || (cfg->extradata_size > 0 && !cfg->extradata)) return HEAAC_E_INVAL;
    HeaacDecoder *d = new (std::nothrow) HeaacDecoder();
    if (!d) return HEAAC_E_NOMEM;
    d->cfg = *cfg;
    if (cfg->extradata_size) d->extradata.assign(cfg->extradata, cfg->extradata + cfg->extradata_size);
    d->cfg.extradata = nullptr;
    if (!d->extradata.empty()) {                      /* reject an unusable configuration now, like aac_decode_init */
        HeaacFrontend *probe = nullptr;
        int r = heaac_fe_open(&probe, d->extradata.data(), (int)d->extradata.size(), cfg->sample_rate, cfg->channels, 0);
        if (r) { delete d; return r; }
        heaac_fe_close(probe);
    }
    *out = d;
    return 0;
}

extern "C" void heaac_dec_close(HeaacDecoder *d)
{
    if (!d) return;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->quit = true;
        d->cv_go.notify_all();
    }
    for (auto &t : d->workers) t.join();
    if (d->be) heaac_b200_wait(d->be);
    if (d->gpu) heaac_b200_parser_destroy(d->gpu);
    free_sets(d);
    for (auto f : d->fe) heaac_fe_close(f);
    if (d->be) heaac_b200_destroy(d->be);
    delete d;
}

/* Asynchronous form: the packets are consumed (parsed, or copied for the GPU parser) before the call returns;
 * pcm / pcm_f32 / consumed are complete once heaac_dec_wait(ticket) returns.  Up to 8 calls may be outstanding. */
extern "C" long long heaac_dec_submit(HeaacDecoder *d, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                                      int16_t *pcm, float *pcm_f32, int *consumed)
{
    if (!d || n_frames <= 0 || !pkt || !pkt_size || (!pcm && !pcm_f32)) return HEAAC_E_INVAL;
    if (n_frames > d->cfg.max_frames_per_call) return HEAAC_E_INVAL;
    int r;
    if (!d->configured) {
        if ((r = configure(d, pkt[0], pkt_size[0]))) { unconfigure(d); return r; }
        d->configured = true;
    }
    if (pcm_f32 && !d->cfg.want_float_out) return HEAAC_E_INVAL;
    const size_t S = d->cfg.n_streams;
    if (d->gpu) {
        int *cons = consumed;
        const int slot = (int)(d->n_submitted & 7);
        /* 8 calls may be outstanding; the 9th first completes the oldest one that was never waited for */
        if (d->pending_consumed[slot] && (r = heaac_dec_wait(d, d->pending_ticket[slot]))) return r;
        if (!cons) { d->scratch_consumed[slot].resize((size_t)n_frames * S); cons = d->scratch_consumed[slot].data(); }
        long long t = heaac_b200_parser_submit(d->gpu, n_frames, pkt, pkt_size, pcm, pcm_f32, cons);
        if (t < 0) return t;
        d->pending_consumed[slot] = cons; d->pending_n[slot] = (size_t)n_frames * S; d->pending_ticket[slot] = t;
        d->n_submitted++;
        return t;
    }
    const size_t out_frame = S * heaac_b200_frame_out_len(d->be) * heaac_b200_n_out_channels(d->be);
    for (int f0 = 0; f0 < n_frames; f0 += d->chunk, d->n_chunks++) {
        const int n = n_frames - f0 < d->chunk ? n_frames - f0 : d->chunk;
        ParamSet *p = &d->set[d->n_chunks & 1];
        /* this set was handed to the GPU two chunks ago: its H2D copies must be over before it is overwritten */
        if (d->n_chunks >= 2 && (r = heaac_b200_inputs_consumed(d->be))) return r;
        Job j;
        j.pkt = pkt + (size_t)f0 * S; j.size = pkt_size + (size_t)f0 * S;
        j.consumed = consumed ? consumed + (size_t)f0 * S : nullptr;
        j.n_frames = n; j.set = p;
        run_job(d, j);
        HeaacBatch b;
        memset(&b, 0, sizeof(b));
        b.n_frames = n;
        b.coeffs = p->coeffs; b.ics = p->ics; b.tns = p->tns; b.sbr = p->sbr; b.ps = p->ps; b.cce = p->cce;
        b.pcm = pcm ? pcm + (size_t)f0 * out_frame : nullptr;
        b.pcm_f32 = pcm_f32 ? pcm_f32 + (size_t)f0 * out_frame : nullptr;
        if ((r = heaac_b200_synth_submit(d->be, &b))) return r;
    }
    d->n_submitted++;
    return heaac_b200_ticket(d->be);
}

extern "C" int heaac_dec_wait(HeaacDecoder *d, long long ticket)
{
    if (!d || !d->configured) return HEAAC_E_INVAL;
    if (!d->gpu) return heaac_b200_ticket_wait(d->be, ticket);
    int r = heaac_b200_parser_wait(d->gpu, ticket);
    if (r) return r;
    for (int slot = 0; slot < 8; slot++)
        if (d->pending_consumed[slot] && d->pending_ticket[slot] == ticket) {
            long long bad = 0;
            for (size_t i = 0; i < d->pending_n[slot]; i++) bad += d->pending_consumed[slot][i] < 0;
            d->parse_errors += bad;
            d->pending_consumed[slot] = nullptr;
        }
    return 0;
}

extern "C" int heaac_dec_decode(HeaacDecoder *d, int n_frames, const uint8_t *const *pkt, const int *pkt_size,
                                int16_t *pcm, float *pcm_f32, int *consumed)
{
    long long t = heaac_dec_submit(d, n_frames, pkt, pkt_size, pcm, pcm_f32, consumed);
    return t < 0 ? (int)t : heaac_dec_wait(d, t);
}

extern "C" int heaac_dec_info(const HeaacDecoder *d, int *n_out_channels, int *frame_out_len, int *sbr, int *ps,
                              long long *parse_errors, long long *gpu_launches)
{
    if (!d) return HEAAC_E_INVAL;
    if (parse_errors) *parse_errors = d->parse_errors.load();
    if (!d->configured) return 0;                      /* not configured yet: nothing decoded so far */
    if (n_out_channels) *n_out_channels = heaac_b200_n_out_channels(d->be);
    if (frame_out_len) *frame_out_len = heaac_b200_frame_out_len(d->be);
    if (sbr) *sbr = d->sbr;
    if (ps) *ps = d->ps;
    if (gpu_launches) *gpu_launches = heaac_b200_launch_count(d->be) + heaac_b200_parser_launch_count(d->gpu);
    return 1;
}

extern "C" double heaac_dec_parse_kernel_ms(const HeaacDecoder *d)
{
    double ms = 0.0;
    if (d && d->gpu) heaac_b200_parser_stats(d->gpu, &ms, nullptr);
    return ms;
}
