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
    HeaacCoreSmem *sm_core; HeaacHfSmem *sm_sbr; HeaacPsSmem *sm_ps; HeaacSynthSmem *sm_synth;
    HeaacHfScratch *sm_hf; float *u, *vh;            /* the fused chain's "shared memory" */
    HeaacSbrElem elem;
    int old_path;                                    /* 1: round-1 kernels (k_sbr + k_synth) instead of the chain */
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
    g->ncc = ch; g->nse = se; g->nout = cfg->ps ? 2 : ch; g->ds = cfg->sbr && cfg->sbr_downsampled; g->out_len = cfg->sbr && !g->ds ? 2048 : 1024;
    g->sr_index = cfg->sr_index; g->sbr = cfg->sbr != 0; g->ps = cfg->ps != 0;
    g->add_bias = cfg->add_bias; g->sf_scale = cfg->sf_scale;
    int any = 0;
    for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) any |= cfg->out_slot[i];
    for (int i = 0; i < HEAAC_MAX_CHANNELS; i++) g->out_slot[i] = (any && !cfg->ps) ? cfg->out_slot[i] : (uint8_t)i;
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
    c->sm_sbr = (HeaacHfSmem *)calloc(1, sizeof(HeaacHfSmem));
    c->sm_ps = (HeaacPsSmem *)calloc(1, sizeof(HeaacPsSmem));
    c->sm_synth = (HeaacSynthSmem *)calloc(1, sizeof(HeaacSynthSmem));
    c->sm_hf = (HeaacHfScratch *)calloc(1, sizeof(HeaacHfScratch));
    c->u = (float *)calloc(2 * HEAAC_U_PLANE, sizeof(float));
    c->vh = (float *)calloc(1152, sizeof(float));
    c->old_path = getenv("HEAAC_EMU_OLD") != nullptr;
    return c;
}

extern "C" void heaac_emu_destroy(EmuCtx *c)
{
    if (!c) return;
    free(c->T); free(c->chan); free(c->synth); free(c->ps); free(c->W); free(c->X);
    free(c->sm_core); free(c->sm_sbr); free(c->sm_ps); free(c->sm_synth); free(c->sm_hf); free(c->u); free(c->vh); free(c);
}

extern "C" int heaac_emu_synth(EmuCtx *c, const HeaacBatch *b)
{
    const HeaacGeom g = c->g;
    const size_t S = g.n_streams;
    const float qmf_scale = -1024.0f * g.sf_scale;
    for (int f = 0; f < b->n_frames; f++) {
        const float *co = b->coeffs + (size_t)f * S * g.ncc * 1024;
        const HeaacIcs *ic = b->ics + (size_t)f * S * g.ncc;
        const HeaacTns *tn = b->tns ? b->tns + (size_t)f * S * g.ncc : nullptr;
        const HeaacSbrElem *sb = b->sbr ? b->sbr + (size_t)f * S * g.nse : nullptr;
        const HeaacPsFrame *ps = b->ps ? b->ps + (size_t)f * S : nullptr;
        size_t pcm_frame = S * g.out_len * g.nout;
        int16_t *pc = b->pcm ? b->pcm + (size_t)f * pcm_frame : nullptr;
        float *pf = b->pcm_f32 ? b->pcm_f32 + (size_t)f * pcm_frame : nullptr;
        for (size_t gc = 0; gc < S * g.ncc; gc++) {               /* k_core (+ QMF analysis) */
            size_t s = gc / g.ncc; int ch = (int)(gc - s * g.ncc);
            static float x[1312 + 64 * 33];
            float *out = x + 288;
            float bias = g.sbr ? 0.0f : g.add_bias;
            HeaacChanState *st = c->chan + gc;
            bool ana = g.sbr && g.ch_sbr_elem[ch] != 0xff;
            memset(c->sm_core, 0xA5, sizeof(HeaacCoreSmem));      /* poison "shared memory" */
            memset(x, 0xA5, sizeof(x));
            if (ana) memcpy(x, st->ana_hist, 288 * sizeof(float));
            static float cbuf[1024];
            memcpy(cbuf, co + gc * 1024, sizeof(cbuf));           /* the kernel stages the coefficients with cp.async */
            heaac_core_channel<1>(c->sm_core, cbuf, ic + gc, tn ? tn + gc : nullptr, nullptr, st, out, bias,
                               g.sr_index, c->T, 0);
            if (ana && !c->old_path) {
                /* the fused chain of k_chain: analysis -> HF stage -> synthesis, all on U */
                const float ana_scale = 1 / qmf_scale;
                if (ana_scale != 1.0f) for (int i = 0; i < 1024; i++) out[i] = out[i] * ana_scale;
                memcpy(st->ana_hist, x + 1024, 288 * sizeof(float));
                for (int i = 0; i < 2 * HEAAC_U_PLANE; i++) { uint32_t poison = 0xA5A5A5A5u; memcpy(c->u + i, &poison, 4); }
                memset(c->sm_hf, 0xA5, sizeof(HeaacHfScratch));
                heaac_qmf_analysis_u<1>(x, x + 1312, c->u, c->T, 0);
                const int se = g.ch_sbr_elem[ch];
                c->elem = sb[s * g.nse + se];
                heaac_hf_stage<1>(c->u, c->sm_hf, &c->elem, g.ch_in_elem[ch], g.elem_is_cpe[ch], st, c->T, 0);
                if (g.ps || g.ds) {
                    heaac_x_emit<1>(c->u, c->X + (s * g.nout + ch) * HEAAC_X_FLOATS, 0);
                } else {
                    HeaacSynthState *sy = c->synth + gc;
                    memcpy(c->vh, sy->v, sizeof(sy->v));
                    heaac_synth_inplace<1>(c->u, c->vh, sy, 1, g.nout, g.out_slot[ch], pc ? pc + s * 2048 * g.nout : nullptr,
                                           pf ? pf + s * g.nout * 2048 : nullptr, qmf_scale, g.add_bias, c->T, 0);
                }
            } else if (ana) {
                const float ana_scale = 1 / qmf_scale;
                if (ana_scale != 1.0f) for (int i = 0; i < 1024; i++) out[i] = out[i] * ana_scale;
                memcpy(st->ana_hist, x + 1024, 288 * sizeof(float));
                heaac_qmf_analysis<1>(x, x + 1312, c->W + gc * 2048, c->T, 0);
            } else if (!g.sbr) for (int i = 0; i < 1024; i++) {
                if (pc) pc[(s * 1024 + i) * g.nout + g.out_slot[ch]] = heaac_to_s16(out[i], g.add_bias != 0.0f);
                if (pf) pf[(s * g.nout + g.out_slot[ch]) * 1024 + i] = out[i];
            }
        }
        if (!g.sbr) continue;
        for (size_t gc = 0; c->old_path && gc < S * g.ncc; gc++) {   /* k_sbr (round-1 kernels) */
            size_t s = gc / g.ncc; int ch = (int)(gc - s * g.ncc);
            int se = g.ch_sbr_elem[ch];
            if (se == 0xff) continue;
            memset(c->sm_sbr, 0xA5, sizeof(HeaacHfSmem));
            heaac_hf_channel<1>(c->sm_sbr, c->W + gc * 2048, sb + s * g.nse + se, g.ch_in_elem[ch], g.elem_is_cpe[ch],
                             c->chan + gc, c->X + (s * g.nout + ch) * HEAAC_X_FLOATS, c->T, 0);
        }
        if (g.ps)
            for (size_t s = 0; s < S; s++) {                      /* k_ps */
                float *XL = c->X + s * 2 * (HEAAC_X_FLOATS), *XR = XL + HEAAC_X_FLOATS;
                const HeaacSbrElem *e = sb + s * g.nse;
                memset(c->sm_ps, 0xA5, sizeof(HeaacPsSmem));
                heaac_ps_stream<1>(c->sm_ps, ps + s, c->ps + s, XL, XR, (int)e->kx[1] + (int)e->m[1], c->T, 0);
            }
        if (!c->old_path && !g.ps && !g.ds) continue;             /* the chain has already produced the PCM */
        for (size_t s = 0; s < S; s++) {                          /* k_synth: one CTA per (stream, SBR element) */
            int o = 0;
            for (int e = 0; e < g.n_elems; e++) {
                int nch = g.elem_type[e] == HEAAC_TYPE_CPE ? 2 : 1;
                if (g.elem_type[e] != HEAAC_TYPE_LFE) {
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
    else if (!strcmp(name, "sbr_smem")) { p = c->sm_sbr; n = sizeof(HeaacHfSmem); }
    else if (!strcmp(name, "ps_smem")) { p = c->sm_ps; n = sizeof(HeaacPsSmem); }
    else if (!strcmp(name, "ps_state")) { p = c->ps; n = S * sizeof(HeaacPsState); }
    if (!p) return -1;
    if (n > (size_t)cap) n = cap;
    memcpy(dst, p, n);
    return (int)n;
}
