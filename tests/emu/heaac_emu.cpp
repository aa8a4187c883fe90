/*
 * tests/emu/heaac_emu.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the product's CTA-level kernel source (ffmpeg-heaac_b200/csrc/heaac_kernels.cuh)
 * as plain C++ with HEAAC_HOST_EMU (one "thread" per CTA, barriers are no-ops) so that the
 * kernel LOGIC (indexing, state hand-over, parameter handling) can be checked against the
 * oracle on machines without a GPU.  It is not a fallback: the product library
 * (libheaac_b200.so) never contains, loads or calls this, and data races / barrier bugs can
 * only be found on the GPU (tests/ -m gpu).  Exports heaac_emu_* mirroring heaac_b200_*.
 */
#define HEAAC_HOST_EMU 1
#include <stdlib.h>
#include <string.h>
#include <new>
#include "../../ffmpeg-heaac_b200/csrc/heaac_chain.cuh"

extern "C" void heaac_build_tables(HeaacDevTables *t);

struct EmuCtx {
    HeaacConfig cfg; HeaacGeom g;
    HeaacDevTables *T;
    HeaacChanState *chan; HeaacSynthState *synth; HeaacPsState *ps;
    float *W, *X;
    HeaacCoreSmem *sm_core; HeaacPsSmem *sm_ps; HeaacSynthSmem *sm_synth;
    HeaacChainSmem *sm_chain;                        /* the fused chain's "shared memory" */
    float *couple;                                   /* [S][ncc][1024]: k_core -> k_couple */
};

static void build_geom(const HeaacConfig *cfg, HeaacGeom *g)
{
    memset(g, 0, sizeof(*g));
    g->n_streams = cfg->n_streams; g->n_elems = cfg->n_elems;
    int ch = 0, se = 0;
    for (int e = 0; e < cfg->n_elems; e++) {
        g->elem_type[e] = cfg->elem_type[e];
        int nch = cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
        for (int c = 0; c < nch; c++, ch++) {
            g->ch_elem[ch] = e; g->ch_in_elem[ch] = c;
            g->ch_sbr_elem[ch] = cfg->elem_type[e] == HEAAC_TYPE_LFE ? 0xff : se;
            g->elem_is_cpe[ch] = nch == 2;
        }
        if (cfg->elem_type[e] != HEAAC_TYPE_LFE) se++;
    }
    for (int e = 0, c0 = 0; e < cfg->n_elems; c0 += cfg->elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1, e++)
        if (cfg->elem_type[e] == HEAAC_TYPE_CCE) g->cce_ch[g->n_cce++] = (uint8_t)c0;
    g->ncc = ch; g->nse = se; g->nout = cfg->ps ? 2 : ch - g->n_cce; g->ds = cfg->sbr && cfg->sbr_downsampled; g->out_len = cfg->sbr && !g->ds ? 2048 : 1024;
    g->sr_index = cfg->sr_index; g->sbr = cfg->sbr != 0; g->ps = cfg->ps != 0;
    g->add_bias = cfg->add_bias; g->sf_scale = cfg->sf_scale;
    int any = 0;
    for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) any |= cfg->out_slot[i];
    for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) g->out_slot[i] = (any && !cfg->ps) ? cfg->out_slot[i] : (uint8_t)i;
    if (g->n_cce) {
        int o = 0;
        for (int i = 0; i < ch; i++) {
            const bool is_cce = cfg->elem_type[g->ch_elem[i]] == HEAAC_TYPE_CCE;
            g->out_slot[i] = is_cce ? 0xff : (any ? cfg->out_slot[i] : (uint8_t)o);
            o += !is_cce;
        }
    }
}

extern "C" EmuCtx *heaac_emu_create(const HeaacConfig *cfg)
{
    EmuCtx *c = (EmuCtx *)calloc(1, sizeof(EmuCtx));
    c->cfg = *cfg; build_geom(cfg, &c->g);
    c->T = (HeaacDevTables *)malloc(sizeof(HeaacDevTables));
    heaac_build_tables(c->T);
    size_t S = c->g.n_streams;
    c->chan = (HeaacChanState *)calloc(S * c->g.ncc, sizeof(HeaacChanState));
    c->synth = (HeaacSynthState *)calloc(S * c->g.nout, sizeof(HeaacSynthState));
    c->ps = (HeaacPsState *)calloc(S, sizeof(HeaacPsState));
    c->W = (float *)calloc(S * c->g.ncc * 2048, sizeof(float));
    c->X = (float *)calloc(S * c->g.nout * HEAAC_X_FLOATS, sizeof(float));
    c->sm_core = (HeaacCoreSmem *)calloc(1, sizeof(HeaacCoreSmem));
    c->sm_ps = (HeaacPsSmem *)calloc(1, sizeof(HeaacPsSmem));
    c->sm_synth = (HeaacSynthSmem *)calloc(1, sizeof(HeaacSynthSmem));
    c->sm_chain = (HeaacChainSmem *)calloc(1, sizeof(HeaacChainSmem));
    c->couple = (float *)calloc(S * c->g.ncc * 1024, sizeof(float));
    return c;
}

extern "C" void heaac_emu_destroy(EmuCtx *c)
{
    if (!c) return;
    free(c->T); free(c->chan); free(c->synth); free(c->ps); free(c->W); free(c->X);
    free(c->sm_core); free(c->sm_ps); free(c->sm_synth); free(c->sm_chain); free(c->couple); free(c);
}

/* one coded channel through the fused chain (what one CTA of k_chain does), frames [f0, f0 + n) of the batch */
static void emu_chain(EmuCtx *c, const HeaacBatch *b, size_t gc, int f0, int n, int emit)
{
    const HeaacGeom g = c->g;
    const size_t S = g.n_streams;
    const size_t s = gc / g.ncc; const int ch = (int)(gc - s * g.ncc);
    const int se = g.ch_sbr_elem[ch];
    HeaacChainArgs a;
    memset(&a, 0, sizeof(a));
    a.n_frames = n; a.emit = emit;
    a.frame_ch = S * g.ncc; a.frame_elems = S * g.nse; a.pcm_frame = S * 2048 * g.nout;
    a.coeffs = b->coeffs + ((size_t)f0 * a.frame_ch + gc) * 1024;
    a.ics = b->ics + (size_t)f0 * a.frame_ch + gc;
    a.tns = b->tns ? b->tns + (size_t)f0 * a.frame_ch + gc : nullptr;
    a.sbr = se == 0xff ? nullptr : b->sbr + (size_t)f0 * a.frame_elems + s * g.nse + se;
    a.st = c->chan + gc;
    a.sy = c->synth + s * g.nout + ch;
    a.X = c->X + (s * g.nout + ch) * HEAAC_X_FLOATS;
    a.pcm = b->pcm ? b->pcm + (size_t)f0 * a.pcm_frame + s * 2048 * g.nout : nullptr;
    a.pcm_f32 = b->pcm_f32 ? b->pcm_f32 + (size_t)f0 * a.pcm_frame + s * 2048 * g.nout : nullptr;
    a.ch_in_elem = g.ch_in_elem[ch]; a.is_cpe = g.elem_is_cpe[ch]; a.nout = g.nout; a.out_slot = g.out_slot[ch]; a.sr_index = g.sr_index;
    a.qmf_scale = -1024.0f * g.sf_scale; a.add_bias = g.add_bias;
    memset(c->sm_chain, 0xA5, sizeof(HeaacChainSmem));            /* poison "shared memory" */
    heaac_chain_channel<1>(c->sm_chain, a, c->T, 0);
}

extern "C" int heaac_emu_synth(EmuCtx *c, const HeaacBatch *b)
{
    const HeaacGeom g = c->g;
    const size_t S = g.n_streams;
    const float qmf_scale = -1024.0f * g.sf_scale;
    if (g.sbr && !g.ps && !g.ds) {                                /* HE-AAC v1: k_chain, the whole call per channel */
        for (size_t gc = 0; gc < S * g.ncc; gc++)
            emu_chain(c, b, gc, 0, b->n_frames, 0);
        return 0;
    }
    for (int f = 0; f < b->n_frames; f++) {
        const float *co = b->coeffs + (size_t)f * S * g.ncc * 1024;
        const HeaacIcs *ic = b->ics + (size_t)f * S * g.ncc;
        const HeaacTns *tn = b->tns ? b->tns + (size_t)f * S * g.ncc : nullptr;
        const HeaacSbrElem *sb = b->sbr ? b->sbr + (size_t)f * S * g.nse : nullptr;
        const HeaacPsFrame *ps = b->ps ? b->ps + (size_t)f * S : nullptr;
        size_t pcm_frame = S * g.out_len * g.nout;
        int16_t *pc = b->pcm ? b->pcm + (size_t)f * pcm_frame : nullptr;
        float *pf = b->pcm_f32 ? b->pcm_f32 + (size_t)f * pcm_frame : nullptr;
        for (size_t gc = 0; gc < S * g.ncc; gc++) {               /* k_core, or k_chain up to X */
            size_t s = gc / g.ncc; int ch = (int)(gc - s * g.ncc);
            if (g.sbr) {
                emu_chain(c, b, gc, f, 1, 1);
                continue;
            }
            static float out[1024];
            HeaacChanState *st = c->chan + gc;
            memset(c->sm_core, 0xA5, sizeof(HeaacCoreSmem));      /* poison "shared memory" */
            static float cbuf[1024];
            memcpy(cbuf, co + gc * 1024, sizeof(cbuf));           /* the kernel stages the coefficients with cp.async */
            heaac_core_channel<1>(c->sm_core, cbuf, ic + gc, tn ? tn + gc : nullptr, nullptr, st, out, g.add_bias,
                               g.sr_index, c->T, 0);
            if (g.n_cce) {                                        /* planar, one slot per coded channel: k_couple finishes */
                memcpy(c->couple + gc * 1024, out, sizeof(out));
                continue;
            }
            for (int i = 0; i < 1024; i++) {
                if (pc) pc[(s * 1024 + i) * g.nout + g.out_slot[ch]] = heaac_to_s16(out[i], g.add_bias != 0.0f);
                if (pf) pf[(s * g.nout + g.out_slot[ch]) * 1024 + i] = out[i];
            }
        }
        if (g.n_cce)
            for (size_t s = 0; s < S; s++)                        /* k_couple */
                heaac_couple_stream<1>(&g, c->couple + s * g.ncc * 1024, b->cce + ((size_t)f * S + s) * g.n_cce,
                                       pc ? pc + s * 1024 * g.nout : nullptr, pf ? pf + s * g.nout * 1024 : nullptr, 0);
        if (!g.sbr) continue;
        if (g.ps)
            for (size_t s = 0; s < S; s++) {                      /* k_ps */
                float *XL = c->X + s * 2 * (HEAAC_X_FLOATS), *XR = XL + HEAAC_X_FLOATS;
                const HeaacSbrElem *e = sb + s * g.nse;
                memset(c->sm_ps, 0xA5, sizeof(HeaacPsSmem));
                heaac_ps_stream<1>(c->sm_ps, ps + s, c->ps + s, XL, XR, (int)e->kx[1] + (int)e->m[1], c->T, 0);
            }
        for (size_t s = 0; s < S; s++) {                          /* k_synth: one CTA per (stream, SBR element) */
            int o = 0;
            for (int e = 0; e < g.n_elems; e++) {
                int nch = g.elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
                {
                    int no = g.ps ? 2 : nch;
                    size_t go = s * g.nout + o;
                    memset(c->sm_synth, 0xA5, sizeof(HeaacSynthSmem));
                    if (g.ds)
                        heaac_synth_stream_ds<1>(c->sm_synth, c->X + go * HEAAC_X_FLOATS, c->synth + go, no, g.nout,
                                                 g.out_slot[o], g.out_slot[o + (no > 1)],
                                                 pc ? pc + s * 1024 * g.nout : nullptr, pf ? pf + s * g.nout * 1024 : nullptr,
                                                 qmf_scale, g.add_bias, c->T, 0);
                    else
                        heaac_synth_stream<1>(c->sm_synth, c->X + go * HEAAC_X_FLOATS, c->synth + go, no, g.nout,
                                              g.out_slot[o], g.out_slot[o + (no > 1)],
                                              pc ? pc + s * 2048 * g.nout : nullptr, pf ? pf + s * g.nout * 2048 : nullptr,
                                              qmf_scale, g.add_bias, c->T, 0);
                }
                o += nch;
            }
        }
    }
    return 0;
}

extern "C" int heaac_emu_tables(HeaacDevTables *out) { heaac_build_tables(out); return (int)sizeof(HeaacDevTables); }

extern "C" int heaac_emu_peek(EmuCtx *c, const char *name, void *dst, int cap)
{
    const void *p = nullptr; size_t n = 0;
    size_t S = c->g.n_streams;
    if (!strcmp(name, "X")) { p = c->X; n = S * c->g.nout * HEAAC_X_FLOATS * sizeof(float); }
    else if (!strcmp(name, "W")) { p = c->W; n = S * c->g.ncc * 2048 * sizeof(float); }
    else if (!strcmp(name, "chan_state")) { p = c->chan; n = S * c->g.ncc * sizeof(HeaacChanState); }
    else if (!strcmp(name, "ps_smem")) { p = c->sm_ps; n = sizeof(HeaacPsSmem); }
    else if (!strcmp(name, "ps_state")) { p = c->ps; n = S * sizeof(HeaacPsState); }
    if (!p) return -1;
    if (n > (size_t)cap) n = cap;
    memcpy(dst, p, n);
    return (int)n;
}
