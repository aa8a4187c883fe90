/*
 * heaac_chain.cuh -- the SBR channel pipeline on ONE shared-memory array (sm_100a).
 *
 * Everything ff_sbr_apply does for one channel-frame between the core decoder's time samples and the PCM
 * (aacsbr.c:1716-1771) works on a single QMF-domain array U[plane re/im][band 0..63][column] that never leaves
 * shared memory.  Columns: 0..8 hold the synthesis filterbank history (the 9 v blocks before the frame), column
 * HEAAC_UC0 + t holds time index t = 0..39 of the reference's X_low / X_high (t = QMF slot + 8 for analysis output,
 * t = i + 2 for output slot i: ENVELOPE_ADJUSTMENT_OFFSET, aacsbr.c:39).
 *
 *   sbr_qmf_analysis 1136-1169   window + fold into rows 32..63 (free at that time), transforms write W[k][slot] to
 *                                U[k][t = 8 + slot], k < 32                                   (heaac_ana_*)
 *   sbr_lf_gen 1337-1357         U[k][t < 8] <- the previous frame's W tail: rows k < 32 ARE X_low[k][t]
 *   inverse filter 1232-1313, chirp 1316-1334, dequant 1089-1128, mapping 1451-1496          (heaac_hf_*)
 *   sbr_hf_gen 1360-1409         X_high[k][t] goes to U[k][t], rows kx .. kx+M-1 (patches copy from rows below k0 <= kx)
 *   env_estimate 1499-1546, gain_calc 1552-1605
 *   sbr_hf_assemble 1608-1714    Y[i][k] = X_high[k][i + 2] g + noise | sinusoid, IN PLACE: U[k][i + 2]
 *   sbr_x_gen 1412-1446          X[i][k] = U[k][i + 2]: nothing to move; rows >= `top` are never written -- the
 *                                consumers (synthesis, X emit) read them as zero
 *   sbr_qmf_synthesis 1175-1230  the two N = 128 half-IMDCTs of slot i read column i + 2 and write b0 / b1 back
 *                                into it, the combine step turns the column into the slot's 128-sample v block in
 *                                place, and the 10-tap polyphase sum walks along two rows per output sample with a
 *                                register window that runs on into the history columns     (heaac_synth_inplace)
 *
 * Row stride 49 (odd): phases with lanes = slots and phases with lanes = bands are both free of bank conflicts.
 *
 * The N = 128 transforms of both filterbanks are spread over four threads each (heaac_regfft.cuh), so that the whole CTA
 * works on them.  Per-band lookup tables derived from the header survive from frame to frame and are rebuilt only when
 * the header part of the side info changes; everything else that depends only on the side info (dequantisation, mapping,
 * the q_m / s_m halves of the gain formulas) is computed per (envelope, band) item straight from the side info, without
 * intermediate arrays or barriers.
 *
 * Arithmetic: the expressions are those of the oracle, operation for operation (order-sensitive parts through the
 * non-contracting heaac_f* primitives) -- only the work decomposition and the data placement differ.
 */
#ifndef HEAAC_CHAIN_CUH
#define HEAAC_CHAIN_CUH

#include "heaac_kernels.cuh"

#define HEAAC_US 49                      /* row stride of U */
#define HEAAC_UC0 9                      /* column of t = 0 */
#define HEAAC_U_PLANE (64 * HEAAC_US)
#define HEAAC_U_FLOATS (2 * HEAAC_U_PLANE)
#define HEAAC_MAX_LIM 32                 /* f_tablelim has 29 entries (sbr.h:139) */
#define HEAAC_GS 49                      /* row stride of the per-envelope gain arrays (48 used): rows of different envelopes
                                          * start on different banks */
#define HEAAC_HDR_WORDS 35               /* HeaacSbrElem up to ch[]: 140 bytes */

/* side info reaches the kernels unchecked on device-resident batches: every index derived from it is cut to the
 * range of the arrays it addresses (garbage in, garbage out -- never out of bounds) */
struct HeaacHfDims {
    int start, reset, kx, kx0, m_max, m0, k0, nenv, nnoise, n_q, n_lim, n_lo, n_hi, npatch;
    int i_lo, i_hi, i_temp, h_SL, e_a0, e_a1, t_old2, coupled, top;
};

/* per sub-band lookup tables derived from the header (frequency tables, patches): rebuilt only when the header changes */
struct HeaacHfTables {
    uint8_t band_idx[2][48];        /* envelope band (low / high resolution table) containing kx + m */
    uint8_t noise_idx[48];          /* noise-floor band containing kx + m */
    uint8_t lim_idx[48];            /* limiter band containing kx + m */
    uint8_t patch_src[48];          /* low-band source of kx + m in the HF generator, 255 = beyond the last patch */
    uint8_t is_mid[48];             /* kx + m is the middle sub-band of its high-resolution band (where a sinusoid sits) */
};

/* small per-channel state that lives in shared memory for the length of a call (HeaacChanState holds it between calls) */
struct HeaacHfKeep {
    uint32_t f_indexnoise, f_indexsine;
    float bw_array[8];
    uint8_t s_index0[48];
};

struct HeaacHfScratch {
    float phi[32][5];               /* per low band: rs0, rs1, is1, rs2, is2 (sums over t = 1..37) */
    float alpha[32][4];             /* a0.re a0.im a1.re a1.im */
    float bw[8];
    float e_orig[5][HEAAC_GS], q_map[5][HEAAC_GS], e_curr[5][HEAAC_GS], s_m[5][HEAAC_GS];
    float gmax[5][HEAAC_MAX_LIM], gboost[5][HEAAC_MAX_LIM], sum0[5][HEAAC_MAX_LIM];
    float gall[10][HEAAC_GS], qall[10][HEAAC_GS];   /* rows 0..4: gain / q_m of this frame's envelopes; 5..8: the four g_temp / q_temp
                                                     * history rows the smoother can reach (aacsbr.c:1632-1647); 9: zeros */
    float gsm[5][HEAAC_GS], qsm[5][HEAAC_GS];       /* the 5-tap smoother's output where all taps lie in one envelope */
    uint8_t s_idx[5][48];           /* s_indexmapped[1..5] */
    uint8_t s_map[5][48];
    int8_t  env_of_slot[40];        /* envelope covering output slot i, -1 = none */
    uint8_t slot_mode[40];          /* bit 0: envelope adjacent to a transient border (l_A / l_APrev); bit 1: the 5-tap
                                     * smoother applies and reaches into another envelope or the history */
};

HEAAC_DEV void heaac_hf_dims(HeaacHfDims *d, const HeaacSbrElem *p, int ch, int is_cpe)
{
    const HeaacSbrChan *pc = &p->ch[ch];
    d->start = p->start != 0; d->reset = p->reset != 0;
    d->kx = HEAAC_MIN((int)p->kx[1], 32); d->kx0 = HEAAC_MIN((int)p->kx[0], 32);
    d->m_max = HEAAC_MIN((int)p->m[1], HEAAC_MIN(48, 64 - d->kx));
    d->m0 = HEAAC_MIN((int)p->m[0], 64 - d->kx0);
    d->k0 = HEAAC_MIN((int)p->k0, d->kx);
    d->nenv = HEAAC_MIN((int)pc->bs_num_env, 5);
    d->nnoise = HEAAC_MIN((int)pc->bs_num_noise, 2);
    d->n_q = HEAAC_MIN((int)p->n_q, 5); d->n_lim = HEAAC_MIN((int)p->n_lim, 28);
    d->n_lo = HEAAC_MIN((int)p->n[0], 24); d->n_hi = HEAAC_MIN((int)p->n[1], 48);
    d->npatch = HEAAC_MIN((int)p->num_patches, 6);
    if (d->nenv < 1 || d->n_q < 1 || d->n_lim < 1) d->start = 0;
    d->i_lo = d->start ? HEAAC_MIN(2 * (int)pc->t_env[0], 38) : 0;
    d->i_hi = d->start ? HEAAC_MIN(HEAAC_MAX(2 * (int)pc->t_env[d->nenv], d->i_lo), 38) : 0;
    d->t_old2 = 2 * (int)pc->t_env_num_env_old;
    d->i_temp = HEAAC_MIN(HEAAC_MAX(d->t_old2 - 32, 0), 32);
    d->h_SL = 4 * !p->bs_smoothing_mode;
    d->e_a0 = pc->e_a[0]; d->e_a1 = pc->e_a[1];
    d->coupled = is_cpe && p->bs_coupling;
    /* rows >= top hold nothing this frame: high band of this frame, and of the previous one where its last slots reach
     * into this frame (sbr_x_gen, aacsbr.c:1420-1431) */
    d->top = d->start ? d->kx + d->m_max : d->kx;
    if (d->i_temp > 0) d->top = HEAAC_MAX(d->top, d->kx0 + d->m0);
}

/* by scatter: task = one band of one frequency table */
template <int NT>
HEAAC_DEV void heaac_hf_build_tables(HeaacHfTables *tb, const HeaacSbrElem *p, const HeaacHfDims *d, int tid)
{
    constexpr int nt = NT;
    const int kx = d->kx;
    for (int w = tid; w < 6 * 48; w += nt) (&tb->band_idx[0][0])[w] = w >= 4 * 48 && w < 5 * 48 ? 255 : 0;
    HEAAC_SYNC();
    HEAAC_NOUNROLL
    for (int w = tid; w < 152; w += nt) {
        if (w < 49) {                                      /* f_tablehigh band w */
            if (w < d->n_hi) {
                const int lo = p->f_tablehigh[w], hi = p->f_tablehigh[w + 1], mid = (lo + hi) >> 1;
                for (int b = lo; b < hi; b++)
                    if ((unsigned)(b - kx) < 48u) tb->band_idx[1][b - kx] = (uint8_t)w;
                if ((unsigned)(mid - kx) < 48u) tb->is_mid[mid - kx] = 1;
            }
        } else if (w < 64) {
        } else if (w < 64 + 25) {                          /* f_tablelow */
            const int j = w - 64;
            if (j < d->n_lo)
                for (int b = p->f_tablelow[j]; b < (int)p->f_tablelow[j + 1]; b++)
                    if ((unsigned)(b - kx) < 48u) tb->band_idx[0][b - kx] = (uint8_t)j;
        } else if (w < 96) {
        } else if (w < 96 + 5) {                           /* noise-floor bands: the last one runs to the top of the range */
            const int j = w - 96;
            if (j < d->n_q) {
                const int lo = j ? (int)p->f_tablenoise[j] : kx, hi = j + 1 < d->n_q ? (int)p->f_tablenoise[j + 1] : kx + 48;
                for (int b = lo; b < hi; b++)
                    if ((unsigned)(b - kx) < 48u) tb->noise_idx[b - kx] = (uint8_t)j;
            }
        } else if (w < 104) {
        } else if (w < 104 + 28) {                         /* limiter bands */
            const int j = w - 104;
            if (j < d->n_lim) {
                const int lo = j ? (int)p->f_tablelim[j] : kx, hi = j + 1 < d->n_lim ? (int)p->f_tablelim[j + 1] : kx + 48;
                for (int b = lo; b < hi; b++)
                    if ((unsigned)(b - kx) < 48u) tb->lim_idx[b - kx] = (uint8_t)j;
            }
        } else if (w < 136) {
        } else if (w < 136 + 6) {                          /* patches (aacsbr.c:1371-1395); what no patch covers stays 255 */
            const int j = w - 136;
            if (j < d->npatch) {
                int acc = kx;
                for (int q = 0; q < j; q++) acc += p->patch_num_subbands[q];
                const int hi = acc + (int)p->patch_num_subbands[j];
                const int src0 = (int)p->patch_start_subband[j];
                for (int b = acc; b < hi; b++) {
                    const int q = src0 + (b - acc);
                    if ((unsigned)(b - kx) < 48u) tb->patch_src[b - kx] = (uint8_t)(q < 32 ? q : 255);
                }
            }
        }
    }
    HEAAC_SYNC();
}

/* env_facs[e + 1][b] / noise_facs[kq + 1][j] of sbr_dequant (aacsbr.c:1089-1128), one value at a time */
HEAAC_DEV float heaac_env_dequant(const HeaacSbrElem *p, int ch, int coupled, int e, int b)
{
    if (coupled) {
        const HeaacSbrChan *p0 = &p->ch[0], *p1 = &p->ch[1];
        const float alpha      = p0->bs_amp_res ?  1.0f :  0.5f;
        const float pan_offset = p0->bs_amp_res ? 12.0f : 24.0f;
        const float t1 = heaac_exp2_half((float)p0->env_q[e][b] * alpha + 7.0f);
        const float t2 = heaac_exp2_half((pan_offset - (float)p1->env_q[e][b]) * alpha);
        const float fac = heaac_fdiv(t1, 1.0f + t2);
        return ch ? fac * t2 : fac;
    }
    const HeaacSbrChan *pc = &p->ch[ch];
    const float alpha = pc->bs_amp_res ? 1.0f : 0.5f;
    return heaac_exp2_half(alpha * (float)pc->env_q[e][b] + 6.0f);
}
HEAAC_DEV float heaac_noise_dequant(const HeaacSbrElem *p, int ch, int coupled, int e, int j)
{
    if (coupled) {
        const HeaacSbrChan *p0 = &p->ch[0], *p1 = &p->ch[1];
        const float t1 = heaac_exp2_half(6.0f - (float)p0->noise_q[e][j] + 1);
        const float t2 = heaac_exp2_half(12 - (float)p1->noise_q[e][j]);
        const float fac = heaac_fdiv(t1, 1.0f + t2);
        return ch ? fac * t2 : fac;
    }
    return heaac_exp2_half(6.0f - (float)p->ch[ch].noise_q[e][j]);
}

/* s_indexmapped[e + 1][m] of sbr_mapping (aacsbr.c:1472-1480); s0 = s_indexmapped[0] (the previous frame's last row) */
HEAAC_DEV int heaac_sine_at(const HeaacSbrChan *pc, const HeaacHfTables *tb, const uint8_t *s0, int e_a1, int e, int m)
{
    if (!pc->bs_add_harmonic_flag || !tb->is_mid[m]) return 0;
    return (uint8_t)(pc->bs_add_harmonic[tb->band_idx[1][m]] * (e >= e_a1 || s0[m] == 1));
}

/* ------------------------------------------------------------------------------------
 * QMF analysis, step 1 (aacsbr.c:1149-1160): window, fold, shuffle.  x: 1312 floats (288 history + 1024 new, already
 * scaled).  Transform input j of slot i goes to row 32 + (j & 31) of plane j >> 5, column HEAAC_UC0 + i: the high-band
 * rows of U are free until the HF generator runs.  A thread keeps one sub-band sample index k (its five window
 * coefficients in registers) and walks over slots.
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_ana_window(const float *x, float *u, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    constexpr int NG = nt >= 512 ? 8 : nt >= 256 ? 4 : nt >= 128 ? 2 : 1;      /* slot groups */
    constexpr int SL = 32 / NG;
    HEAAC_NOUNROLL
    for (int w = tid; w < 64 * NG; w += nt) {
        const int k = w & 63, i0 = (w >> 6) * SL;
        const float c0 = T->qmf_ds[k], c1 = T->qmf_ds[k + 64], c2 = T->qmf_ds[k + 128], c3 = T->qmf_ds[k + 192], c4 = T->qmf_ds[k + 256];
        int pos; bool neg = false;                         /* t[0] = u[0]; t[2k-1] = u[k], t[2k] = -u[64-k]; t[63] = u[32] */
        if (k == 0) pos = 0;
        else if (k < 32) pos = 2 * k - 1;
        else if (k == 32) pos = 63;
        else { pos = 2 * (64 - k); neg = true; }
        float *dst = u + (pos >> 5) * HEAAC_U_PLANE + (32 + (pos & 31)) * HEAAC_US + HEAAC_UC0 + i0;
        const float *xs = x + 32 * i0 + 319 - k;
        HEAAC_UNROLL4
        for (int i = 0; i < SL; i++, xs += 32) {
            float f = c0 * xs[0];
            f = f + c1 * xs[-64];
            f = f + c2 * xs[-128];
            f = f + c3 * xs[-192];
            f = f + c4 * xs[-256];
            dst[i] = neg ? -f : f;
        }
    }
}

/* step 2 (aacsbr.c:1161-1167): one half-IMDCT (N = 128, scale -2) per slot, four threads per slot (heaac_regfft.cuh);
 * qx: exchange buffer, 4 * 2 * HEAAC_QX_FLOATS floats.  Threads 0 .. min(nt, 128) - 1 take part. */
template <int NT>
HEAAC_DEV void heaac_ana_transform(float *u, float *qx, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT < 128 ? NT : 128;
    if (tid < nt)
        HEAAC_NOUNROLL
        for (int w = tid; w < 128; w += nt) {
            const int q = w & 7, t = (w >> 3) & 3, grp = w >> 5, i = grp * 8 + q;
            const float *src = u + 32 * HEAAC_US + HEAAC_UC0 + i;
            float *xre = qx + grp * 2 * HEAAC_QX_FLOATS + q;
            heaac_imdct128_quad_a([&](int hi, int lo, int) { return src[(hi >> 5) * HEAAC_U_PLANE + ((hi & 31) + lo) * HEAAC_US]; },
                                  T->tcos_ana, T->tsin_ana, t, xre, xre + HEAAC_QX_FLOATS);
        }
    HEAAC_SYNCWARP();
    if (tid < nt)
        HEAAC_NOUNROLL
        for (int w = tid; w < 128; w += nt) {
            const int q = w & 7, t = (w >> 3) & 3, grp = w >> 5, i = grp * 8 + q;
            float *ure = u + HEAAC_UC0 + 8 + i, *uim = u + HEAAC_U_PLANE + HEAAC_UC0 + 8 + i;
            const float *xre = qx + grp * 2 * HEAAC_QX_FLOATS + q;
            heaac_imdct128_quad_b([&](int hi, int lo, float v) {       /* W[i][k] = (-y[63-k], y[k]), aacsbr.c:1163-1166 */
                                      if (hi < 32) uim[(hi + lo) * HEAAC_US] = v;
                                      else         ure[(63 - hi - lo) * HEAAC_US] = -v;
                                  }, T->tcos_ana, T->tsin_ana, T->fft_wre, T->fft_wim, t, xre, xre + HEAAC_QX_FLOATS);
        }
}

/* ------------------------------------------------------------------------------------
 * Everything of the HF stage that depends only on the side info and the state; runs beside the analysis transforms
 * on threads t = 0 .. n-1 (any n; no barriers inside, every item is computed from the side info alone).
 * ---------------------------------------------------------------------------------- */
HEAAC_DEV void heaac_hf_params(float *u, HeaacHfScratch *sm, const HeaacHfTables *tb, const HeaacHfDims *d,
                               const HeaacSbrElem *p, int ch, HeaacChanState *st, HeaacHfKeep *keep, int t, int n)
{
    float (*ure)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u);
    float (*uim)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u + HEAAC_U_PLANE);
    const HeaacSbrChan *pc = &p->ch[ch];
    const int m_max = d->m_max, nenv = d->nenv;
    HEAAC_NOUNROLL
    for (int w = t; w < 2 * 32 * 8; w += n) {              /* next frame's W tail: t = 32..39, all 32 bands (before anything overwrites them) */
        const int pl = w >> 8, k = (w >> 3) & 31, i = w & 7;
        st->w_tail[w] = pl ? uim[k][HEAAC_UC0 + 32 + i] : ure[k][HEAAC_UC0 + 32 + i];
    }
    if (!d->start) return;
    const int smooth = d->h_SL != 0;
    HEAAC_NOUNROLL
    for (int i = t; i < 40; i += n) {                      /* time grid: envelope of each output slot */
        int e = -1;
        for (int q2 = 0; q2 < nenv; q2++)
            if (i >= 2 * pc->t_env[q2] && i < 2 * pc->t_env[q2 + 1]) e = q2;
        if (i >= d->i_hi || i < d->i_lo) e = -1;
        const int border = e >= 0 && (e == d->e_a0 || e == d->e_a1);
        /* the smoother's oldest tap is slot i - 4: inside the same envelope => the pre-computed value applies */
        int mixed = 0;
        if (e >= 0 && smooth && !border) mixed = i - 4 < HEAAC_MAX(2 * (int)pc->t_env[e], d->i_lo);
        sm->env_of_slot[i] = (int8_t)e;
        sm->slot_mode[i] = (uint8_t)(border | (mixed << 1));
    }
    /* the g_temp / q_temp rows the smoother reads below the first envelope: rows 2 t_old .. 2 t_old + 3 of the
     * previous frame (aacsbr.c:1637-1640), kept in the state as rows 32..41 */
    HEAAC_NOUNROLL
    for (int w = t; w < 5 * 48; w += n) {
        const int r = w / 48, m = w - r * 48;
        const int hr = d->t_old2 - 32 + r;
        const bool have = r < 4 && smooth && !d->reset && hr >= 0 && hr < 10;
        sm->gall[5 + r][m] = have ? st->g_hist[hr * 48 + m] : 0.0f;
        sm->qall[5 + r][m] = have ? st->q_hist[hr * 48 + m] : 0.0f;
    }
    HEAAC_NOUNROLL
    for (int i = t; i < d->n_q; i += n) {                  /* sbr_chirp */
        float old = keep->bw_array[i], nb;
        const int m0_ = pc->bs_invf_mode[0][i], m1_ = pc->bs_invf_mode[1][i];
        if (m0_ + m1_ == 1) nb = 0.6f;
        else nb = m0_ == 0 ? 0.0f : m0_ == 1 ? 0.75f : m0_ == 2 ? 0.9f : 0.98f;
        if (nb < old) nb = heaac_fadd(heaac_fmul(0.75f, nb), heaac_fmul(0.25f, old));
        else          nb = heaac_fadd(heaac_fmul(0.90625f, nb), heaac_fmul(0.09375f, old));
        nb = nb < 0.015625f ? 0.0f : nb;
        sm->bw[i] = nb;
        keep->bw_array[i] = nb;
    }
    /* sbr_dequant + sbr_mapping + the signal-independent halves of sbr_gain_calc (aacsbr.c:1566-1571), per (envelope, band) */
    HEAAC_NOUNROLL
    for (int w = t; w < nenv * 48; w += n) {
        const int e = w / 48, m = w - e * 48;
        if (m >= m_max) continue;
        const int fr = pc->bs_freq_res[e + 1] & 1;
        const int b = tb->band_idx[fr][m];
        const int kq = (d->nnoise > 1) && (pc->t_env[e] >= pc->t_q[1]);
        const float eo = heaac_env_dequant(p, ch, d->coupled, e, b);
        const float qm = heaac_noise_dequant(p, ch, d->coupled, kq, tb->noise_idx[m]);
        const int sine = heaac_sine_at(pc, tb, keep->s_index0, d->e_a1, e, m);
        const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
        int present = 0;
        if (pc->bs_add_harmonic_flag)
            for (int mm = tab[b]; mm < (int)tab[b + 1]; mm++)
                if ((unsigned)(mm - d->kx) < 48u && heaac_sine_at(pc, tb, keep->s_index0, d->e_a1, e, mm - d->kx)) { present = 1; break; }
        const float temp = heaac_fdiv(eo, heaac_fadd(1.0f, qm));
        sm->e_orig[e][m] = eo;
        sm->q_map[e][m] = qm;
        sm->qall[e][m] = heaac_fsqrt(heaac_fmul(temp, qm));
        sm->s_m[e][m] = heaac_fsqrt(heaac_fmul(temp, (float)sine));
        sm->s_idx[e][m] = (uint8_t)sine;
        sm->s_map[e][m] = (uint8_t)present;
    }
}

/* g_filt / q_filt of one (slot, band) when the smoother's taps straddle envelopes (aacsbr.c:1653-1658, 1685-1690);
 * row[j] = row of gall / qall feeding tap j */
HEAAC_DEV void heaac_smooth5(const HeaacHfScratch *sm, const int row[5], int m, float *g_o, float *q_o)
{
    float g_filt = 0.0f, q_filt = 0.0f;
    HEAAC_UNROLL
    for (int j = 0; j <= 4; j++) {
        g_filt = heaac_fadd(g_filt, heaac_fmul(sm->gall[row[j]][m], heaac_h_smooth[j]));
        q_filt = heaac_fadd(q_filt, heaac_fmul(sm->qall[row[j]][m], heaac_h_smooth[j]));
    }
    *g_o = g_filt; *q_o = q_filt;
}

/* limiter of sbr_gain_calc (aacsbr.c:1582-1588) for one (envelope, band): gain and q_m after the gain_max clamp */
HEAAC_DEV void heaac_limit(const HeaacHfScratch *sm, const HeaacHfTables *tb, int e, int m, float *g_o, float *q_o)
{
    const float gain_max = sm->gmax[e][tb->lim_idx[m]];
    const float g = sm->gall[e][m], q = sm->qall[e][m];
    const float q_m_max = heaac_fdiv(heaac_fmul(q, gain_max), g);
    *q_o = HEAAC_MIN(q, q_m_max);
    *g_o = HEAAC_MIN(g, gain_max);
}

/* ------------------------------------------------------------------------------------
 * sbr_hf_inverse_filter (aacsbr.c:1261-1313) with its autocorrelation (1232-1255): one thread per low band, the five sums of
 * a band share the loads of x[t], x[t+1], x[t+2] and each keeps the reference's order over t = 1..37.  Runs on one warp
 * beside heaac_hf_params.  Threads t = 0 .. n-1.
 * ---------------------------------------------------------------------------------- */
HEAAC_DEV void heaac_hf_lpc(const float *u, HeaacHfScratch *sm, const HeaacHfDims *d, int t, int n)
{
    if (!d->start) return;
    constexpr int C0 = HEAAC_UC0;
    HEAAC_NOUNROLL
    for (int k = t; k < d->k0; k += n) {
        const float *xr = u + k * HEAAC_US + C0, *xi = xr + HEAAC_U_PLANE;
        float rs0 = 0.0f, rs1 = 0.0f, rs2 = 0.0f, is1 = 0.0f, is2 = 0.0f;
        {
            float r1 = xr[1], i1 = xi[1], r2 = xr[2], i2 = xi[2];
            HEAAC_UNROLL4
            for (int i = 1; i < 38; i++) {
                const float r3 = xr[i + 2], i3 = xi[i + 2];
                rs0 = heaac_fadd(rs0, heaac_fadd(heaac_fmul(r1, r1), heaac_fmul(i1, i1)));
                rs1 = heaac_fadd(rs1, heaac_fadd(heaac_fmul(r1, r2), heaac_fmul(i1, i2)));
                is1 = heaac_fadd(is1, heaac_fsub(heaac_fmul(r1, i2), heaac_fmul(i1, r2)));
                rs2 = heaac_fadd(rs2, heaac_fadd(heaac_fmul(r1, r3), heaac_fmul(i1, i3)));
                is2 = heaac_fadd(is2, heaac_fsub(heaac_fmul(r1, i3), heaac_fmul(i1, r3)));
                r1 = r2; i1 = i2; r2 = r3; i2 = i3;
            }
        }
            float phi210 = heaac_fadd(heaac_fadd(rs0, heaac_fmul(xr[0], xr[0])), heaac_fmul(xi[0], xi[0]));
            float phi100 = heaac_fadd(heaac_fadd(rs0, heaac_fmul(xr[38], xr[38])), heaac_fmul(xi[38], xi[38]));
            float phi110 = heaac_fadd(heaac_fadd(rs1, heaac_fmul(xr[0], xr[1])), heaac_fmul(xi[0], xi[1]));
            float phi111 = heaac_fsub(heaac_fadd(is1, heaac_fmul(xr[0], xi[1])), heaac_fmul(xi[0], xr[1]));
            float phi000 = heaac_fadd(heaac_fadd(rs1, heaac_fmul(xr[38], xr[39])), heaac_fmul(xi[38], xi[39]));
            float phi001 = heaac_fsub(heaac_fadd(is1, heaac_fmul(xr[38], xi[39])), heaac_fmul(xi[38], xr[39]));
            float phi010 = heaac_fadd(heaac_fadd(rs2, heaac_fmul(xr[0], xr[2])), heaac_fmul(xi[0], xi[2]));
            float phi011 = heaac_fsub(heaac_fadd(is2, heaac_fmul(xr[0], xi[2])), heaac_fmul(xi[0], xr[2]));
            float dk = heaac_fsub(heaac_fmul(phi210, phi100),
                                  heaac_fdiv(heaac_fadd(heaac_fmul(phi110, phi110), heaac_fmul(phi111, phi111)), 1.000001f));
            float a1r, a1i, a0r, a0i;
            if (!dk) {
                a1r = a1i = 0.0f;
            } else {
                float tr = heaac_fsub(heaac_fsub(heaac_fmul(phi000, phi110), heaac_fmul(phi001, phi111)), heaac_fmul(phi010, phi100));
                float ti = heaac_fsub(heaac_fadd(heaac_fmul(phi000, phi111), heaac_fmul(phi001, phi110)), heaac_fmul(phi011, phi100));
                a1r = heaac_fdiv(tr, dk);
                a1i = heaac_fdiv(ti, dk);
            }
            if (!phi100) {
                a0r = a0i = 0.0f;
            } else {
                float tr = heaac_fadd(heaac_fadd(phi000, heaac_fmul(a1r, phi110)), heaac_fmul(a1i, phi111));
                float ti = heaac_fsub(heaac_fadd(phi001, heaac_fmul(a1i, phi110)), heaac_fmul(a1r, phi111));
                a0r = heaac_fdiv(-tr, phi100);
                a0i = heaac_fdiv(-ti, phi100);
            }
            if (heaac_fadd(heaac_fmul(a1r, a1r), heaac_fmul(a1i, a1i)) >= 16.0f ||
                heaac_fadd(heaac_fmul(a0r, a0r), heaac_fmul(a0i, a0i)) >= 16.0f)
                a1r = a1i = a0r = a0i = 0.0f;
            sm->alpha[k][0] = a0r; sm->alpha[k][1] = a0i; sm->alpha[k][2] = a1r; sm->alpha[k][3] = a1i;
    }
}

/* u    : U of this channel; rows < 32: X_low complete (analysis output of this frame in t = 8..39, the previous frame's
 *        tail in t < 8), heaac_hf_params done
 * After the call U[k][i + 2], i = 0..37, k < d->top is X[i][k] of sbr_x_gen (both planes). */
template <int NT>
HEAAC_DEV void heaac_hf_stage(float *u, HeaacHfScratch *sm, const HeaacHfTables *tb, const HeaacHfDims *d,
                              const HeaacSbrElem *p, int ch, HeaacChanState *st, const HeaacHfKeep *keep, int zero_to,
                              const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    float (*ure)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u);
    float (*uim)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u + HEAAC_U_PLANE);
    const HeaacSbrChan *pc = &p->ch[ch];
    const int kx = d->kx, m_max = d->m_max, nenv = d->nenv, i_lo = d->i_lo, i_hi = d->i_hi;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    const int smooth = d->start && d->h_SL;
    constexpr int C0 = HEAAC_UC0;

    /* (every barrier below is reached by every thread whatever the side info says: two channels with different side info may share
     * a CTA, heaac_chain_channel) */
    const bool on = d->start;
    if (on) {

        /* ================= phase 4: sbr_hf_gen, lanes = target bands, each warp item one stretch of time slots with the
         * two previous low-band samples carried in registers ================= */
        {
            constexpr int NCH = nt >= 256 ? 8 : nt >= 128 ? 4 : nt >= 64 ? 2 : 1;         /* stretches of time */
            const int t_first = i_lo + 2, t_end = i_hi + 2;
            const int len = (t_end - t_first + NCH - 1) / NCH;
            const int groups = (m_max + 31) >> 5;
            HEAAC_NOUNROLL
            for (int w = tid; w < NCH * groups * 32; w += nt) {
                const int m = (w & 31) + 32 * ((w >> 5) / NCH), c = (w >> 5) % NCH;
                if (m >= m_max) continue;
                const int ta = t_first + c * len, tb_ = HEAAC_MIN(ta + len, t_end);
                const int q = tb->patch_src[m];
                float *hr = &ure[kx + m][C0], *hi = &uim[kx + m][C0];
                if (q == 255) {                            /* beyond the last patch: X_high = 0 (memset, aacsbr.c:1403-1406) */
                    for (int t = ta; t < tb_; t++) { hr[t] = 0.0f; hi[t] = 0.0f; }
                    continue;
                }
                const float bw = sm->bw[tb->noise_idx[m]];
                const float a0 = sm->alpha[q][2] * bw * bw, a1 = sm->alpha[q][3] * bw * bw;
                const float a2 = sm->alpha[q][0] * bw,      a3 = sm->alpha[q][1] * bw;
                const float *xr = &ure[q][C0], *xi = &uim[q][C0];
                if (ta >= tb_) continue;
                float r2 = xr[ta - 2], i2 = xi[ta - 2], r1 = xr[ta - 1], i1 = xi[ta - 1];
                for (int t = ta; t < tb_; t++) {
                    const float r0 = xr[t], i0 = xi[t];
                    hr[t] = r2 * a0 - i2 * a1 + r1 * a2 - i1 * a3 + r0;
                    hi[t] = i2 * a0 + r2 * a1 + i1 * a2 + r1 * a3 + i0;
                    r2 = r1; i2 = i1; r1 = r0; i1 = i0;
                }
            }
        }
    }
    HEAAC_SYNC();
    if (on) {

        /* ================= phase 5: sbr_env_estimate, one item per (envelope, band), serial over the envelope's slots ================= */
        if (p->bs_interpol_freq) {
            HEAAC_NOUNROLL
            for (int w = tid; w < nenv * 48; w += nt) {
                const int e = w / 48, m = w - e * 48;
                if (m >= m_max) continue;
                const int ilb = HEAAC_MIN(pc->t_env[e] * 2 + 2, 40), iub = HEAAC_MIN(pc->t_env[e + 1] * 2 + 2, 40);
                const float recip = heaac_fdiv(0.5f, (float)(pc->t_env[e + 1] - pc->t_env[e]));
                const float *xr = &ure[kx + m][C0], *xi = &uim[kx + m][C0];
                float sum = 0.0f;
                for (int i = ilb; i < iub; i++)
                    sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(xr[i], xr[i]), heaac_fmul(xi[i], xi[i])));
                sm->e_curr[e][m] = heaac_fmul(sum, recip);
            }
        } else {
            HEAAC_NOUNROLL
            for (int w = tid; w < nenv * 64; w += nt) {    /* one item per envelope band: the whole group in the reference's order */
                const int e = w >> 6, b = w & 63;
                const int fr = pc->bs_freq_res[e + 1] & 1;
                const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
                const int nb = fr ? d->n_hi : d->n_lo;
                if (b >= nb) continue;
                const int ilb = HEAAC_MIN(pc->t_env[e] * 2 + 2, 40), iub = HEAAC_MIN(pc->t_env[e + 1] * 2 + 2, 40);
                const int env_size = 2 * (pc->t_env[e + 1] - pc->t_env[e]);
                const int lo = HEAAC_MAX((int)tab[b], kx), hi = HEAAC_MIN((int)tab[b + 1], kx + m_max);
                const int den = env_size * ((int)tab[b + 1] - (int)tab[b]);
                float sum = 0.0f;
                for (int k = lo; k < hi; k++)
                    for (int i = ilb; i < iub; i++)
                        sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(ure[k][C0 + i], ure[k][C0 + i]), heaac_fmul(uim[k][C0 + i], uim[k][C0 + i])));
                sum = heaac_fdiv(sum, (float)den);
                for (int k = lo; k < hi; k++) sm->e_curr[e][k - kx] = sum;
            }
        }
    }
    HEAAC_SYNC();
    if (on) {

        /* ================= phase 6: sbr_gain_calc, aacsbr.c:1552-1605 ================= */
        /* (a) the gain of every (envelope, band) and the limit of every (envelope, limiter band) */
        HEAAC_NOUNROLL
        for (int w = tid; w < nenv * 48; w += nt) {
            const int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            const int delta = !((e == d->e_a1) || (e == d->e_a0));
            const float eo = sm->e_orig[e][m], qm = sm->q_map[e][m], ec = sm->e_curr[e][m];
            if (!sm->s_map[e][m])
                sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(eo, heaac_fmul(heaac_fadd(1.0f, ec),
                                             heaac_fadd(1.0f, heaac_fmul(qm, (float)delta)))));
            else
                sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(heaac_fmul(eo, qm),
                                             heaac_fmul(heaac_fadd(1.0f, ec), heaac_fadd(1.0f, qm))));
        }
        HEAAC_NOUNROLL
        for (int w = nt - 1 - tid; w < nenv * HEAAC_MAX_LIM; w += nt) {   /* (from the far end of the CTA: those threads have no gain item) */
            const int e = w / HEAAC_MAX_LIM, k = w - e * HEAAC_MAX_LIM;
            if (k >= d->n_lim) continue;
            const int lo = HEAAC_MAX((int)p->f_tablelim[k] - kx, 0), hi = HEAAC_MIN((int)p->f_tablelim[k + 1] - kx, m_max);
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int m = lo; m < hi; m++) { sum0 = heaac_fadd(sum0, sm->e_orig[e][m]); sum1 = heaac_fadd(sum1, sm->e_curr[e][m]); }
            const float gain_max = heaac_fmul(heaac_limgain[p->bs_limiter_gains & 3],
                                              heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sum0),
                                                                     heaac_fadd(1.1920928955078125e-07f, sum1))));
            sm->gmax[e][k] = HEAAC_MIN(100000.0f, gain_max);
            sm->sum0[e][k] = sum0;
        }
    }
    HEAAC_SYNC();
    if (on) {
        /* (b) the boost of every (envelope, limiter band), from the limited gains (worked out on the fly) */
        HEAAC_NOUNROLL
        for (int w = tid; w < nenv * HEAAC_MAX_LIM; w += nt) {
            const int e = w / HEAAC_MAX_LIM, k = w - e * HEAAC_MAX_LIM;
            if (k >= d->n_lim) continue;
            const int delta = !((e == d->e_a1) || (e == d->e_a0));
            const int lo = HEAAC_MAX((int)p->f_tablelim[k] - kx, 0), hi = HEAAC_MIN((int)p->f_tablelim[k + 1] - kx, m_max);
            float sum1 = 0.0f;
            for (int m = lo; m < hi; m++) {
                float g, q;
                heaac_limit(sm, tb, e, m, &g, &q);
                const float s = sm->s_m[e][m];
                const float t3 = heaac_fmul(heaac_fmul((float)(delta && !s), q), q);
                const float acc = heaac_fadd(heaac_fadd(heaac_fmul(heaac_fmul(sm->e_curr[e][m], g), g), heaac_fmul(s, s)), t3);
                sum1 = heaac_fadd(sum1, acc);
            }
            const float gain_boost = heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sm->sum0[e][k]),
                                                            heaac_fadd(1.1920928955078125e-07f, sum1)));
            /* FFMIN(1.584893192 (double), gain_boost) then narrowed to float */
            sm->gboost[e][k] = (1.584893192 > (double)gain_boost) ? gain_boost : (float)1.584893192;
        }
    }
    HEAAC_SYNC();
    if (on) {
        /* (c) limit + boost of every (envelope, band) */
        HEAAC_NOUNROLL
        for (int w = tid; w < nenv * 48; w += nt) {
            const int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            float g, q;
            heaac_limit(sm, tb, e, m, &g, &q);
            const float gain_boost = sm->gboost[e][tb->lim_idx[m]];
            const float gf = heaac_fmul(g, gain_boost), qf = heaac_fmul(q, gain_boost);
            sm->gall[e][m] = gf;
            sm->qall[e][m] = qf;
            sm->s_m[e][m]  = heaac_fmul(sm->s_m[e][m], gain_boost);
            if (smooth) {                                  /* all five taps in this envelope: sum_j g h[j], in the smoother's own order */
                float g_filt = 0.0f, q_filt = 0.0f;
                HEAAC_UNROLL
                for (int j = 0; j <= 4; j++) {
                    g_filt = heaac_fadd(g_filt, heaac_fmul(gf, heaac_h_smooth[j]));
                    q_filt = heaac_fadd(q_filt, heaac_fmul(qf, heaac_h_smooth[j]));
                }
                sm->gsm[e][m] = g_filt; sm->qsm[e][m] = q_filt;
            }
        }
    }
    HEAAC_SYNC();
    if (on) {

        /* ================= phase 7: sbr_hf_assemble in place, U[kx + m][i + 2] <- Y[i][kx + m] ================= */
        const uint32_t noise0 = d->reset ? 0u : keep->f_indexnoise, sine0 = keep->f_indexsine;
        const int kx_sign = 1 - 2 * (kx & 1);
        {   /* (a) slots 0..31 whose gain is constant per (envelope, band): rows = bands, lanes = slots; everything that
             * depends on the slot only is worked out once per lane */
          for (int i = lane; i < 32; i += ls) {
            const int e = sm->env_of_slot[i];
            const int mode = e >= 0 ? sm->slot_mode[i] : 0;
            const bool active = e >= 0 && !(mode & 2);
            if (!active) continue;
            const int border = mode & 1;
            const float *grow = (smooth && !border) ? sm->gsm[e] : sm->gall[e];
            const float *qrow = (smooth && !border) ? sm->qsm[e] : sm->qall[e];
            const float *srow = sm->s_m[e];
            const int indexsine = (int)((sine0 + (uint32_t)(i - i_lo)) & 3);
            const float phi_re = indexsine == 0 ? 1.0f : indexsine == 2 ? -1.0f : 0.0f;
            const int phi_im = (indexsine == 1 ? 1 : indexsine == 3 ? -1 : 0) * kx_sign;
            const float phi_im_even = (float)phi_im, phi_im_odd = (float)-phi_im;
            const uint32_t nbase = noise0 + (uint32_t)(i - i_lo) * (uint32_t)m_max + 1u;
            float *yr = &ure[kx][C0 + i + 2], *yi = &uim[kx][C0 + i + 2];
            for (int m = warp; m < m_max; m += nwarps) {
                const float g_filt = grow[m], q_filt = qrow[m], s = srow[m];
                float yre = heaac_fmul(yr[m * HEAAC_US], g_filt);
                float yim = heaac_fmul(yi[m * HEAAC_US], g_filt);
                if (!border && !s) {
                    const uint32_t idx = (nbase + (uint32_t)m) & 0x1ff;
                    yre = heaac_fadd(yre, heaac_fmul(q_filt, T->noise[2 * idx]));
                    yim = heaac_fadd(yim, heaac_fmul(q_filt, T->noise[2 * idx + 1]));
                } else {
                    yre = heaac_fadd(yre, heaac_fmul(s, phi_re));
                    yim = heaac_fadd(yim, heaac_fmul(s, (m & 1) ? phi_im_odd : phi_im_even));
                }
                yr[m * HEAAC_US] = yre; yi[m * HEAAC_US] = yim;
            }
          }
        }
        {   /* (b) the other slots of the grid -- smoother taps reaching into another envelope or the history, and slots
             * 32..37 -- one slot per warp iteration, lanes = bands */
            for (int i = i_lo + warp; i < i_hi; i += nwarps) {
                const int e = sm->env_of_slot[i];
                const int mode = sm->slot_mode[i];
                if (e < 0 || (i < 32 && !(mode & 2))) continue;
                const int border = mode & 1, mixed = mode & 2;
                int row[5] = { e, e, e, e, e };
                if (mixed)
                    for (int j = 0; j <= 4; j++) {         /* g_temp row i + 4 - j: slot i - j of this frame, else history (aacsbr.c:1632-1647) */
                        const int src = i - j;
                        int r;
                        if (src >= i_lo) r = sm->env_of_slot[src] >= 0 ? sm->env_of_slot[src] : 9;
                        else if (d->reset) r = 0;
                        else r = 5 + (src + 4 - i_lo);
                        row[j] = r;
                    }
                const int indexsine = (int)((sine0 + (uint32_t)(i - i_lo)) & 3);
                const float phi_re = indexsine == 0 ? 1.0f : indexsine == 2 ? -1.0f : 0.0f;
                const int phi_im = indexsine == 1 ? 1 : indexsine == 3 ? -1 : 0;
                const uint32_t nbase = noise0 + (uint32_t)(i - i_lo) * (uint32_t)m_max + 1u;
                for (int m = lane; m < m_max; m += ls) {
                    float g_filt, q_filt;
                    if (mixed) heaac_smooth5(sm, row, m, &g_filt, &q_filt);
                    else if (smooth && !border) { g_filt = sm->gsm[e][m]; q_filt = sm->qsm[e][m]; }
                    else { g_filt = sm->gall[e][m]; q_filt = sm->qall[e][m]; }
                    float *yr = &ure[kx + m][C0 + i + 2], *yi = &uim[kx + m][C0 + i + 2];
                    const float s = sm->s_m[e][m];
                    float yre = heaac_fmul(*yr, g_filt);
                    float yim = heaac_fmul(*yi, g_filt);
                    if (!border && !s) {
                        const uint32_t idx = (nbase + (uint32_t)m) & 0x1ff;
                        yre = heaac_fadd(yre, heaac_fmul(q_filt, T->noise[2 * idx]));
                        yim = heaac_fadd(yim, heaac_fmul(q_filt, T->noise[2 * idx + 1]));
                    } else {
                        const int phi_sign = kx_sign * ((m & 1) ? -1 : 1);
                        yre = heaac_fadd(yre, heaac_fmul(s, phi_re));
                        yim = heaac_fadd(yim, heaac_fmul(s, (float)(phi_im * phi_sign)));
                    }
                    *yr = yre; *yi = yim;
                }
            }
        }
    }
    HEAAC_SYNC();

    /* ================= phase 8: sbr_x_gen -- cells no stage wrote, below `top` only ================= */
    {
        const int i_temp = d->i_temp, m0 = d->m0, kx0 = d->kx0, top = d->top;
        const int hf_lo = d->start ? kx : 64, hf_hi = d->start ? kx + m_max : 64;
        /* (a) slots below i_temp still belong to the previous frame's grid (aacsbr.c:1420-1431): its Y tail, zeros above its band range */
        if (i_temp > 0)
            HEAAC_NOUNROLL
            for (int w = tid; w < 64 * 8; w += nt) {
                const int k = w >> 3, j = w & 7;
                if (j >= i_temp || j >= 6 || k < kx0 || k >= top) continue;
                const bool hf_old = k < kx0 + m0;
                ure[k][C0 + j + 2] = hf_old ? st->y_tail[k * 8 + j] : 0.0f;
                uim[k][C0 + j + 2] = hf_old ? st->y_tail[512 + k * 8 + j] : 0.0f;
            }
        /* (b) slots >= i_temp: cells of rows kx .. top-1 that are neither low band nor covered by this frame's grid are silent.
         *     With a conformant grid (it starts where the previous one ended and covers the frame) there are none. */
        const bool covered = d->start && hf_hi == top && i_lo <= i_temp && i_hi >= zero_to;
        if (!covered)
            for (int k = kx + warp; k < top; k += nwarps) {
                const bool hf = k >= hf_lo && k < hf_hi;
                for (int i = i_temp + lane; i < 38; i += ls)
                    if (!hf || i < i_lo || i >= i_hi) { ure[k][C0 + i + 2] = 0.0f; uim[k][C0 + i + 2] = 0.0f; }
            }
        /* slots i_temp <= i < 8 of low-band rows kx' <= k < kx hold zeros from sbr_lf_gen (X_low[k][t < 8] = 0 for k >= kx') */
    }
    HEAAC_SYNC();
}

/* The state the next frame needs from this one (reads U slots 32..37 and the scratch only: may run beside the
 * synthesis transforms, which work on slots 0..31); threads t = 0 .. n-1 */
HEAAC_DEV void heaac_hf_save(const float *u, const HeaacHfScratch *sm, const HeaacHfDims *d, HeaacChanState *st, HeaacHfKeep *keep, int t, int n)
{
    if (!d->start) return;
    const float (*ure)[HEAAC_US] = reinterpret_cast<const float (*)[HEAAC_US]>(u);
    const float (*uim)[HEAAC_US] = reinterpret_cast<const float (*)[HEAAC_US]>(u + HEAAC_U_PLANE);
    const int kx = d->kx, m_max = d->m_max, i_lo = d->i_lo, i_hi = d->i_hi, nenv = d->nenv;
    constexpr int C0 = HEAAC_UC0;
    HEAAC_NOUNROLL
    for (int w = t; w < m_max * 8; w += n) {               /* the Y tail: slots 32..37 of the high band */
        const int k = kx + (w >> 3), j = w & 7;
        if (j < 6 && 32 + j >= i_lo && 32 + j < i_hi) {
            st->y_tail[k * 8 + j] = ure[k][C0 + 34 + j]; st->y_tail[512 + k * 8 + j] = uim[k][C0 + 34 + j];
        }
    }
    HEAAC_NOUNROLL
    for (int w = t; w < 10 * 48; w += n) {                 /* g_temp / q_temp rows 32..41 for the next frame, aacsbr.c:1642-1647 */
        const int r = 32 + w / 48, m = w % 48;
        const int src = r - d->h_SL;
        if (m < m_max && src >= i_lo && src < i_hi) {
            const int es = sm->env_of_slot[src];
            if (es >= 0) { st->g_hist[w] = sm->gall[es][m]; st->q_hist[w] = sm->qall[es][m]; }
        }
    }
    HEAAC_NOUNROLL
    for (int m = t; m < 48; m += n) keep->s_index0[m] = m < m_max ? sm->s_idx[nenv - 1][m] : 0;
    if (t == 0) {
        const uint32_t noise0 = d->reset ? 0u : keep->f_indexnoise;
        keep->f_indexnoise = (noise0 + (uint32_t)(i_hi - i_lo) * (uint32_t)m_max) & 0x1ff;
        keep->f_indexsine = (keep->f_indexsine + (uint32_t)(i_hi - i_lo)) & 3;
    }
}

/* ------------------------------------------------------------------------------------
 * U -> the inter-kernel X buffer (Parametric Stereo and downsampled synthesis read it): X[plane][band][HEAAC_XS],
 * slots 0..37; rows >= top are silent.
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_x_emit(const float *u, float *X_out, int top, int tid)
{
    constexpr int nt = NT;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    for (int r = warp; r < 128; r += nwarps) {             /* 2 planes x 64 bands: rows are contiguous in both arrays */
        const float *src = u + r * HEAAC_US + HEAAC_UC0 + 2;
        float *dst = X_out + r * HEAAC_XS;
        const bool live = (r & 63) < top;
        for (int i = lane; i < 38; i += ls) dst[i] = live ? src[i] : 0.0f;
    }
}

/* ------------------------------------------------------------------------------------
 * QMF synthesis (sbr_qmf_synthesis, aacsbr.c:1175-1230) + float->s16 of one channel, in place on U.
 *   history: columns 0..8 of U: column 8 - b is the v block b + 1 slots before this frame's first; on return they hold
 *   the 9 newest blocks of this frame.  v[n], n < 64, of a block lies in the imaginary plane's row n, v[64 + j] in the
 *   real plane's row j.
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_synth_transforms(float *u, float *qx, int top, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    /* the two half-IMDCTs (N = 128, scale 1/64) of every slot, four threads each, column in, column out; two rounds of 32
     * transforms (the exchange buffer holds 32).  Threads 0 .. min(nt, 128) - 1 take part. */
    constexpr int nt = NT < 128 ? NT : 128;
    for (int round = 0; round < 2; round++) {              /* round = plane */
        if (tid < nt)
            HEAAC_NOUNROLL
            for (int w = tid; w < 128; w += nt) {
                const int q = w & 7, t = (w >> 3) & 3, grp = w >> 5, i = grp * 8 + q;
                const float *col = u + round * HEAAC_U_PLANE + HEAAC_UC0 + 2 + i;
                float *xre = qx + grp * 2 * HEAAC_QX_FLOATS + q;
                const bool flip = round == 1;              /* imaginary plane: odd inputs change sign, aacsbr.c:1205-1207 */
                heaac_imdct128_quad_a([&](int hi, int lo, int odd) {
                                          const float xv = hi + lo < top ? col[(hi + lo) * HEAAC_US] : 0.0f;
                                          return odd && flip ? -xv : xv;
                                      }, T->tcos_syn, T->tsin_syn, t, xre, xre + HEAAC_QX_FLOATS);
            }
        HEAAC_SYNCWARP();
        if (tid < nt)
            HEAAC_NOUNROLL
            for (int w = tid; w < 128; w += nt) {
                const int q = w & 7, t = (w >> 3) & 3, grp = w >> 5, i = grp * 8 + q;
                float *col = u + round * HEAAC_U_PLANE + HEAAC_UC0 + 2 + i;
                const float *xre = qx + grp * 2 * HEAAC_QX_FLOATS + q;
                heaac_imdct128_quad_b([&](int hi, int lo, float v) { col[(hi + lo) * HEAAC_US] = v; },
                                      T->tcos_syn, T->tsin_syn, T->fft_wre, T->fft_wim, t, xre, xre + HEAAC_QX_FLOATS);
            }
        HEAAC_SYNCWARP();
    }
}

template <int NT>
HEAAC_DEV void heaac_synth_output(float *u, int nout, int out_slot, int16_t *pcm, float *pcm_f32, float scale, float bias,
                                  const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    float *ure = u, *uim = u + HEAAC_U_PLANE;
    constexpr int CS = HEAAC_UC0 + 2;                      /* column of slot 0 */
    /* ---- v[n] = -b0[63-n] + b1[n], v[127-n] = b0[63-n] + b1[n] (aacsbr.c:1210-1213): the column becomes the slot's v block ---- */
    for (int n = warp; n < 64; n += nwarps) {
        float *pr = ure + (63 - n) * HEAAC_US + CS, *pi = uim + n * HEAAC_US + CS;
        for (int i = lane; i < 32; i += ls) {
            const float p1 = pr[i], q1 = pi[i];
            pi[i] = -p1 + q1;
            pr[i] =  p1 + q1;
        }
    }
    HEAAC_SYNC();
    /* ---- out[n] = sum_j v[off_j + n] w[64 j + n] (aacsbr.c:1215-1227): tap j reads the block of slot i - j, row n of the
     *      imaginary (j even) or real (j odd) plane -- two rows per output sample, walked with a register window; the first
     *      group of slots runs on into the history columns (which lie 2 columns further left) ---- */
    constexpr int G = NT >= 256 ? 4 : NT >= 128 ? 2 : 1;   /* groups of consecutive slots */
    constexpr int L = 32 / G, CH = 8;                      /* a group is worked off in chunks of CH slots: the loop stays rolled */
    const int scale_and_bias = scale != 1.0f || bias != 0.0f;
    const int bias_mode = bias != 0.0f;
    HEAAC_NOUNROLL
    for (int w = tid; w < 64 * G; w += nt) {
        const int n = w & 63, i0 = (w >> 6) * L;
        float wj[10], im[CH + 9], re[CH + 9];
        HEAAC_UNROLL
        for (int j = 0; j < 10; j++) wj[j] = T->qmf_us[64 * j + n];
        const int gap = i0 == 0 ? 2 : 0;
        const float *rowi = uim + n * HEAAC_US + CS + i0, *rowr = ure + n * HEAAC_US + CS + i0;
        HEAAC_UNROLL
        for (int a = 0; a < 9; a++) {                      /* a <-> slot i0 - 9 + a */
            im[a] = rowi[a - 9 - gap];
            re[a] = rowr[a - 9 - gap];
        }
        int16_t *po = pcm ? pcm + ((size_t)i0 * 64 + n) * nout + out_slot : nullptr;
        float *pf = pcm_f32 ? pcm_f32 + (size_t)out_slot * 2048 + i0 * 64 + n : nullptr;
        HEAAC_NOUNROLL
        for (int c = 0; c < L; c += CH) {
            HEAAC_UNROLL
            for (int a = 0; a < CH; a++) { im[9 + a] = rowi[c + a]; re[9 + a] = rowr[c + a]; }
            HEAAC_UNROLL
            for (int ii = 0; ii < CH; ii++) {
                float acc = im[9 + ii] * wj[0] + 0.0f;
                HEAAC_UNROLL
                for (int j = 1; j < 10; j++)
                    acc = ((j & 1) ? re[9 + ii - j] : im[9 + ii - j]) * wj[j] + acc;
                if (scale_and_bias) acc = acc * scale + bias;
                if (po) po[(size_t)(c + ii) * 64 * nout] = heaac_to_s16(acc, bias_mode);
                if (pf) pf[(c + ii) * 64] = acc;
            }
            HEAAC_UNROLL
            for (int a = 0; a < 9; a++) { im[a] = im[CH + a]; re[a] = re[CH + a]; }
        }
    }
    HEAAC_SYNC();
    /* the newest 9 blocks (slots 23..31) become the history: columns CS + 23 .. CS + 31 -> 0 .. 8 */
    HEAAC_NOUNROLL
    for (int w = tid; w < 128 * 9; w += nt) {
        const int r = w / 9, b = w - r * 9;
        u[r * HEAAC_US + b] = u[r * HEAAC_US + CS + 23 + b];
    }
    HEAAC_SYNC();
}

/* history columns <-> HeaacSynthState (newest block first: v[b * 128 + idx], block b = column 8 - b) */
template <int NT>
HEAAC_DEV void heaac_synth_hist_load(float *u, const HeaacSynthState *st, int tid)
{
    constexpr int nt = NT;
    HEAAC_NOUNROLL
    for (int w = tid; w < 1152; w += nt) {
        const int b = w >> 7, idx = w & 127;
        u[(idx < 64 ? HEAAC_U_PLANE + idx * HEAAC_US : (idx - 64) * HEAAC_US) + 8 - b] = st->v[w];
    }
}
template <int NT>
HEAAC_DEV void heaac_synth_hist_store(const float *u, HeaacSynthState *st, int tid)
{
    constexpr int nt = NT;
    HEAAC_NOUNROLL
    for (int w = tid; w < 1152; w += nt) {
        const int b = w >> 7, idx = w & 127;
        st->v[w] = u[(idx < 64 ? HEAAC_U_PLANE + idx * HEAAC_US : (idx - 64) * HEAAC_US) + 8 - b];
    }
}

/* ------------------------------------------------------------------------------------
 * One channel, all frames of a call.
 * ---------------------------------------------------------------------------------- */
struct HeaacChainSmem {
    float u[HEAAC_U_FLOATS];
    HeaacSbrElem p2[2];                  /* SBR side info of the current frame and (being fetched) of the next one */
    HeaacHfTables tb;
    HeaacHfKeep keep;
    HeaacHfDims d;
    int hdr_changed;
    union alignas(16) {
        struct { HeaacCoreSmem core; alignas(16) float x[1312]; alignas(16) float coef[1024]; } c;   /* core transform + analysis window */
        HeaacHfScratch hf;                                                                           /* HF stage */
    } s;
    alignas(8) unsigned long long mbar[1];
#ifdef HEAAC_CHAIN_PAD_BYTES
    char occupancy_experiment_pad[HEAAC_CHAIN_PAD_BYTES];      /* tools/var_sweep.sh: fewer CTAs per SM */
#endif
};

struct HeaacChainArgs {
    int n_frames, emit;                  /* emit = 1 (Parametric Stereo, downsampled SBR): the chain stops at X, which goes to the
                                          * inter-kernel buffer; 2 (keep_intermediates): X is written out as well, and the chain goes on */
    size_t frame_ch;                     /* channels per frame (stride of the per-channel arrays between frames) */
    size_t frame_elems;                  /* SBR elements per frame */
    size_t pcm_frame;                    /* PCM samples per frame, all streams (2048 * nout * n_streams) */
    const float *coeffs;                 /* this channel, frame 0 (NULL: `time_in` holds the time samples, no core transform) */
    const float *time_in;
    const HeaacIcs *ics; const HeaacTns *tns; const float *tns_done;     /* this channel, frame 0 */
    const HeaacSbrElem *sbr;             /* this channel's element, frame 0 */
    HeaacChanState *st; HeaacSynthState *sy;
    float *X;                            /* this channel's inter-kernel X */
    size_t x_frame;                      /* floats between the X of consecutive frames (0: every frame goes to the same place) */
    int16_t *pcm; float *pcm_f32;        /* this stream, frame 0 */
    int ch_in_elem, is_cpe, nout, out_slot, sr_index;
    float qmf_scale, add_bias;
};

/* the asynchronous fetch of the next frame's coefficients (global -> shared bulk copy + mbarrier on the device) */
#ifdef HEAAC_HOST_EMU
static inline void heaac_fetch_setup(unsigned long long *) {}
static inline void heaac_fetch_issue(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
static inline void heaac_fetch_wait(unsigned long long *, unsigned) {}
#else
__device__ __forceinline__ unsigned heaac_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void heaac_fetch_setup(unsigned long long *b)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(heaac_smem_u32(b)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* one thread: dst, src and bytes multiples of 16; earlier generic-proxy accesses to dst are ordered before the copy */
__device__ __forceinline__ void heaac_fetch_issue(void *dst, const void *src, unsigned bytes, unsigned long long *b)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(heaac_smem_u32(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(heaac_smem_u32(dst)), "l"(src), "r"(bytes), "r"(heaac_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void heaac_fetch_wait(unsigned long long *b, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(heaac_smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
#endif

template <int NT>
HEAAC_DEV void heaac_chain_channel(HeaacChainSmem *sm, const HeaacChainArgs &a, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    constexpr bool SPLIT = nt > 32;                  /* one warp runs the inverse filter while the others do the side-info work */
    float *x = sm->s.c.x, *out = x + 288;
    const float ana_scale = 1 / a.qmf_scale;                                    /* aacsbr.c:1730 */
    HeaacChanState *st = a.st;
    constexpr int P_WORDS = (int)(sizeof(HeaacSbrElem) / 4);

    if (tid == 0) { heaac_fetch_setup(&sm->mbar[0]); sm->hdr_changed = 1; }
    HEAAC_SYNC();
    if (tid == 0 && a.coeffs) heaac_fetch_issue(sm->s.c.coef, a.coeffs, 4096, &sm->mbar[0]);
    if (a.emit != 1) heaac_synth_hist_load<NT>(sm->u, a.sy, tid);
    for (int i = tid; i < 8; i += nt) sm->keep.bw_array[i] = st->bw_array[i];
    for (int i = tid; i < 48; i += nt) sm->keep.s_index0[i] = st->s_index0[i];
    if (tid == 0) { sm->keep.f_indexnoise = st->f_indexnoise; sm->keep.f_indexsine = st->f_indexsine; }
    if (a.sbr) {
        for (int i = tid; i < P_WORDS; i += nt)      /* the first frame's side info */
            heaac_cp_async4(reinterpret_cast<float *>(&sm->p2[0]) + i, reinterpret_cast<const float *>(a.sbr) + i);
    } else {
        /* an element without SBR data in an SBR stream (LFE): the reference still runs ff_sbr_apply on it (aacdec.c:1923-1925) with
         * the SpectralBandReplication in its initial state -- start = 0, kx = kx' = 32, M = 0 (aacsbr.c:128-137): upsampling only */
        for (int i = tid; i < 2 * P_WORDS; i += nt) reinterpret_cast<uint32_t *>(&sm->p2[0])[i] = 0;
        HEAAC_SYNC();
        if (tid == 0) { sm->p2[0].kx[0] = sm->p2[0].kx[1] = 32; sm->p2[1].kx[0] = sm->p2[1].kx[1] = 32; }
    }
    heaac_cp_async_commit();
    for (int f = 0; f < a.n_frames; f++) {
        const size_t fo = (size_t)f * a.frame_ch;
        const HeaacSbrElem *p = &sm->p2[f & 1];
        /* fetched while the previous frame was synthesised: this frame's side info; now under way: the analysis history, the
         * overlap-add history (straight to where the windowing puts its output) and the low band's first 8 time slots
         * (sbr_lf_gen, aacsbr.c:1337-1357: the previous frame's W[k][24..31]) */
        heaac_cp_async_wait<0>();
        for (int i = tid; i < 72; i += nt) heaac_cp_async16(x + 4 * i, st->ana_hist + 4 * i);
        if (a.coeffs)
            for (int i = tid; i < 128; i += nt) heaac_cp_async16(out + 4 * i, st->saved + 4 * i);
        for (int w = tid; w < 2 * 32 * 8; w += nt)
            heaac_cp_async4(sm->u + (w >> 8) * HEAAC_U_PLANE + ((w >> 3) & 31) * HEAAC_US + HEAAC_UC0 + (w & 7), st->w_tail + w);
        heaac_cp_async_commit();
        HEAAC_SYNC();
        {   /* did the header part (frequency tables, patches) change? */
            if (f > 0) {
                const uint32_t *cur = reinterpret_cast<const uint32_t *>(p), *old = reinterpret_cast<const uint32_t *>(&sm->p2[(f & 1) ^ 1]);
                for (int i = tid; i < HEAAC_HDR_WORDS; i += nt)
                    if (cur[i] != old[i]) sm->hdr_changed = 1;
            }
            /* barrier; rebuild the tables when any channel of the CTA needs it (the barriers inside must be uniform) */
            const int rebuild = HEAAC_SYNC_OR(sm->hdr_changed || p->reset);
            if (tid == 0) heaac_hf_dims(&sm->d, p, a.ch_in_elem, a.is_cpe);
            if (rebuild) {
                HEAAC_SYNC();
                heaac_hf_build_tables<NT>(&sm->tb, p, &sm->d, tid);
            }
        }
        if (a.coeffs) {
            heaac_fetch_wait(&sm->mbar[0], (unsigned)f & 1u);    /* the coefficients have landed */
            heaac_core_channel<NT>(&sm->s.c.core, sm->s.c.coef, a.ics + fo, a.tns ? a.tns + fo : nullptr,
                                   a.tns_done ? a.tns_done + fo * 1024 : nullptr, st, out, 0.0f, a.sr_index, T, tid, out);
            if (ana_scale != 1.0f)
                for (int i = tid; i < 1024; i += nt) out[i] = out[i] * ana_scale;
        } else {
            for (int i = tid; i < 1024; i += nt) {
                const float v = a.time_in[fo * 1024 + i];
                out[i] = ana_scale != 1.0f ? v * ana_scale : v;
            }
            heaac_cp_async_wait<0>();
        }
        HEAAC_SYNC();
        if (tid == 0) sm->hdr_changed = 0;
        for (int i = tid; i < 288; i += nt) st->ana_hist[i] = x[1024 + i];
        for (int w = tid; w < 32 * 8; w += nt) {             /* X_low[k][t < 8] = 0 for k >= kx' */
            const int k = w >> 3, i = w & 7;
            if (k >= sm->d.kx0) { sm->u[k * HEAAC_US + HEAAC_UC0 + i] = 0.0f; sm->u[HEAAC_U_PLANE + k * HEAAC_US + HEAAC_UC0 + i] = 0.0f; }
        }
        heaac_ana_window<NT>(x, sm->u, T, tid);
        HEAAC_SYNC();
        /* (the exchange buffer of the transforms lies in the dead core scratch; the HF scratch takes its place afterwards) */
        heaac_ana_transform<NT>(sm->u, reinterpret_cast<float *>(&sm->s), T, tid);
        HEAAC_SYNC();
        if (!SPLIT || tid < 32) heaac_hf_lpc(sm->u, &sm->s.hf, &sm->d, tid, SPLIT ? 32 : nt);
        if (!SPLIT || tid >= 32)
            heaac_hf_params(sm->u, &sm->s.hf, &sm->tb, &sm->d, p, a.ch_in_elem, st, &sm->keep, SPLIT ? tid - 32 : tid, SPLIT ? nt - 32 : nt);
        HEAAC_SYNC();
        heaac_hf_stage<NT>(sm->u, &sm->s.hf, &sm->tb, &sm->d, p, a.ch_in_elem, st, &sm->keep, a.emit ? 38 : 32, T, tid);
        if (a.emit) heaac_x_emit<NT>(sm->u, a.X + (size_t)f * a.x_frame, sm->d.top, tid);
        heaac_hf_save(sm->u, &sm->s.hf, &sm->d, st, &sm->keep, tid, nt);
        HEAAC_SYNC();
        /* the scratch is free: fetch the next frame's coefficients and side info while the synthesis goes on (the exchange
         * buffer of the transforms lies below the coefficient buffer) */
        if (f + 1 < a.n_frames) {
            if (tid == 0 && a.coeffs) heaac_fetch_issue(sm->s.c.coef, a.coeffs + (fo + a.frame_ch) * 1024, 4096, &sm->mbar[0]);
        }
        if (f + 1 < a.n_frames && a.sbr) {
            const float *src = reinterpret_cast<const float *>(a.sbr + (size_t)(f + 1) * a.frame_elems);
            float *dst = reinterpret_cast<float *>(&sm->p2[(f & 1) ^ 1]);
            for (int i = tid; i < P_WORDS; i += nt) heaac_cp_async4(dst + i, src + i);
        }
        heaac_cp_async_commit();
        if (a.emit != 1) {
            heaac_synth_transforms<NT>(sm->u, reinterpret_cast<float *>(&sm->s), sm->d.top, T, tid);
            HEAAC_SYNC();
            heaac_synth_output<NT>(sm->u, a.nout, a.out_slot, a.pcm ? a.pcm + (size_t)f * a.pcm_frame : nullptr,
                                   a.pcm_f32 ? a.pcm_f32 + (size_t)f * a.pcm_frame : nullptr, a.qmf_scale, a.add_bias, T, tid);
        }
    }
    heaac_cp_async_wait<0>();
    if (a.emit != 1) heaac_synth_hist_store<NT>(sm->u, a.sy, tid);
    for (int i = tid; i < 8; i += nt) st->bw_array[i] = sm->keep.bw_array[i];
    for (int i = tid; i < 48; i += nt) st->s_index0[i] = sm->keep.s_index0[i];
    if (tid == 0) { st->f_indexnoise = sm->keep.f_indexnoise; st->f_indexsine = sm->keep.f_indexsine; }
}

#endif /* HEAAC_CHAIN_CUH */
