/*
 * heaac_chain.cuh -- the SBR channel pipeline on ONE shared-memory array (sm_100a).
 *
 * Everything ff_sbr_apply does for one channel-frame between the core decoder's time samples and the PCM
 * (aacsbr.c:1716-1771) works on a single QMF-domain array U[plane re/im][band 0..63][t 0..39] that never leaves
 * shared memory:
 *
 *   sbr_qmf_analysis 1136-1169   writes W[k][slot] to U[k][8 + slot], k < 32            (heaac_qmf_analysis_u)
 *   sbr_lf_gen 1337-1357         U[k][0..7] <- the previous frame's W tail: rows k < 32 ARE X_low[k][t]
 *   inverse filter 1232-1313, chirp 1316-1334, dequant 1089-1128, mapping 1451-1496        (heaac_hf_stage)
 *   sbr_hf_gen 1360-1409         X_high[k][t] goes to U[k][t], rows kx .. kx+M-1 (they do not overlap the rows it
 *                                reads: patches copy from below k0 <= kx)
 *   env_estimate 1499-1546, gain_calc 1552-1605
 *   sbr_hf_assemble 1608-1714    Y[i][k] = X_high[k][i + 2] g + noise | sinusoid, IN PLACE: U[k][i + 2]
 *   sbr_x_gen 1412-1446          X[i][k] = U[k][i + 2] for every band: nothing to move, only cells no stage wrote
 *                                (bands above kx + M, slots outside the envelope grid) are cleared / patched
 *   sbr_qmf_synthesis 1175-1230  the two N = 128 half-IMDCTs of slot i read column i + 2 and write b0 / b1 back
 *                                into it, the combine step turns the column into the slot's 128-sample v block in
 *                                place, and the 10-tap polyphase sum walks along two rows per output sample with a
 *                                register window (heaac_synth_inplace); 9 older v blocks come from the state.
 *
 * Time index t: X_low / X_high index of the reference (t = QMF slot + 8 for analysis output, t = i + 2 for output
 * slot i: ENVELOPE_ADJUSTMENT_OFFSET, aacsbr.c:39).  Row stride 41 (odd): phases with lanes = slots and phases
 * with lanes = bands are both free of bank conflicts.
 *
 * Arithmetic: the expressions are those of round 1's heaac_hf_channel / heaac_synth_stream (same operations in the
 * same order as the oracle; order-sensitive parts through the non-contracting heaac_f* primitives) -- only the
 * work decomposition and the data placement changed.
 */
#ifndef HEAAC_CHAIN_CUH
#define HEAAC_CHAIN_CUH

#include "heaac_kernels.cuh"

#define HEAAC_US 41                      /* row stride of U (40 used) */
#define HEAAC_U_PLANE (64 * HEAAC_US)
#define HEAAC_MAX_LIM 32                 /* f_tablelim has 29 entries (sbr.h:139) */

/* ------------------------------------------------------------------------------------
 * QMF analysis into U.  x: 1312 floats (288 history + 1024 new, already scaled); tb: 64 * 33 floats of scratch.
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_qmf_analysis_u(const float *x, float *tb, float *u, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    for (int w = tid; w < 32 * 64; w += nt) {              /* aacsbr.c:1149-1160 */
        const int i = w >> 6, k = w & 63;
        const float *xs = x + 32 * i + 319 - k;
        float f = T->qmf_ds[k] * xs[0];
        f = f + T->qmf_ds[k + 64]  * xs[-64];
        f = f + T->qmf_ds[k + 128] * xs[-128];
        f = f + T->qmf_ds[k + 192] * xs[-192];
        f = f + T->qmf_ds[k + 256] * xs[-256];
        int pos; float val = f;                            /* t[0] = u[0]; t[2k-1] = u[k], t[2k] = -u[64-k]; t[63] = u[32] */
        if (k == 0) pos = 0;
        else if (k < 32) pos = 2 * k - 1;
        else if (k == 32) pos = 63;
        else { pos = 2 * (64 - k); val = -f; }
        tb[pos * 33 + i] = val;
    }
    HEAAC_SYNC();
    for (int i = tid; i < 32; i += nt) {                   /* one half-IMDCT (N = 128, scale -2) per slot, lanes = slots */
        float *ure = u + 8 + i, *uim = u + HEAAC_U_PLANE + 8 + i;
        heaac_imdct128_reg<HEAAC_MDCT_ANA>(
            [&](int j) { return tb[j * 33 + i]; },
            [&](int j, float v) {                          /* W[i][k] = (-y[63-k], y[k]), aacsbr.c:1163-1166 */
                if (j < 32) uim[j * HEAAC_US] = v;
                else        ure[(63 - j) * HEAAC_US] = -v;
            });
    }
}

/* ------------------------------------------------------------------------------------ */
struct HeaacHfScratch {
    float phi[32][8];               /* per low band: rs0, rs1, is1, rs2, is2 (sums over t = 1..37) */
    float alpha[32][4];             /* a0.re a0.im a1.re a1.im */
    float env[5][48];               /* dequantised env_facs[1..5] */
    float noise[2][8];
    float e_orig[5][48], q_map[5][48], e_curr[5][48], s_m[5][48];
    float gmax[5][HEAAC_MAX_LIM], gboost[5][HEAAC_MAX_LIM];
    float bw[8];
    float gall[10][48], qall[10][48];   /* rows 0..4: gain / q_m of this frame's envelopes; 5..8: the four g_temp / q_temp
                                         * history rows the smoother can reach (aacsbr.c:1632-1647); 9: zeros */
    float gsm[5][48], qsm[5][48];   /* the 5-tap smoother's output where all taps lie in one envelope */
    uint8_t s_idx[6][48];           /* s_indexmapped[0..5] */
    uint8_t s_map[5][48];
    uint8_t band_idx[2][48];        /* envelope band (low / high resolution table) containing kx + m */
    uint8_t band_first[2][48];      /* 1 if kx + m is the first sub-band of its envelope band */
    uint8_t noise_idx[48];          /* noise-floor band containing kx + m */
    uint8_t lim_idx[48];            /* limiter band containing kx + m */
    uint8_t patch_src[48];          /* low-band source of kx + m in the HF generator, 255 = beyond the last patch */
    int8_t  env_of_slot[40];        /* envelope covering output slot i, -1 = none */
    uint8_t slot_mode[40];          /* bit 0: envelope adjacent to a transient border (l_A / l_APrev); bit 1: the 5-tap
                                     * smoother applies and reaches into another envelope or the history */
};

/* side info reaches the kernels unchecked on device-resident batches: every index derived from it is cut to the
 * range of the arrays it addresses (garbage in, garbage out -- never out of bounds) */
struct HeaacHfDims {
    int start, reset, kx, kx0, m_max, m0, k0, nenv, nnoise, n_q, n_lim, n_lo, n_hi, npatch;
    int i_lo, i_hi, i_temp, h_SL, e_a0, e_a1, t_old2;
};

HEAAC_DEV HeaacHfDims heaac_hf_dims(const HeaacSbrElem *p, const HeaacSbrChan *pc)
{
    HeaacHfDims d;
    d.start = p->start != 0; d.reset = p->reset != 0;
    d.kx = HEAAC_MIN((int)p->kx[1], 32); d.kx0 = HEAAC_MIN((int)p->kx[0], 32);
    d.m_max = HEAAC_MIN((int)p->m[1], HEAAC_MIN(48, 64 - d.kx));
    d.m0 = HEAAC_MIN((int)p->m[0], 64 - d.kx0);
    d.k0 = HEAAC_MIN((int)p->k0, d.kx);
    d.nenv = HEAAC_MIN((int)pc->bs_num_env, 5);
    d.nnoise = HEAAC_MIN((int)pc->bs_num_noise, 2);
    d.n_q = HEAAC_MIN((int)p->n_q, 5); d.n_lim = HEAAC_MIN((int)p->n_lim, 28);
    d.n_lo = HEAAC_MIN((int)p->n[0], 24); d.n_hi = HEAAC_MIN((int)p->n[1], 48);
    d.npatch = HEAAC_MIN((int)p->num_patches, 6);
    if (d.nenv < 1 || d.n_q < 1 || d.n_lim < 1) d.start = 0;
    d.i_lo = d.start ? HEAAC_MIN(2 * (int)pc->t_env[0], 38) : 0;
    d.i_hi = d.start ? HEAAC_MIN(HEAAC_MAX(2 * (int)pc->t_env[d.nenv], d.i_lo), 38) : 0;
    d.t_old2 = 2 * (int)pc->t_env_num_env_old;
    d.i_temp = HEAAC_MIN(HEAAC_MAX(d.t_old2 - 32, 0), 32);
    d.h_SL = 4 * !p->bs_smoothing_mode;
    d.e_a0 = pc->e_a[0]; d.e_a1 = pc->e_a[1];
    return d;
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

/* u    : U of this channel, analysis output of the current frame already in place (columns 8..39, rows < 32)
 * p    : the element's SBR side info, already in shared memory
 * After the call U[k][i + 2], i = 0..37, is X[i][k] of sbr_x_gen for all 64 bands (both planes). */
template <int NT>
HEAAC_DEV void heaac_hf_stage(float *u, HeaacHfScratch *sm, const HeaacSbrElem *p, int ch, int is_cpe,
                              HeaacChanState *st, const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    float (*ure)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u);
    float (*uim)[HEAAC_US] = reinterpret_cast<float (*)[HEAAC_US]>(u + HEAAC_U_PLANE);
    const HeaacSbrChan *pc = &p->ch[ch];
    const HeaacHfDims d = heaac_hf_dims(p, pc);
    const int kx = d.kx, kx0 = d.kx0, m_max = d.m_max, nenv = d.nenv, i_lo = d.i_lo, i_hi = d.i_hi;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    const int smooth = d.start && d.h_SL;

    /* ================= phase 1: everything that depends only on the side info and the state ================= */
    for (int w = tid; w < 32 * 8; w += nt) {               /* sbr_lf_gen: t < 8 = previous frame's W[k][24..31], k < kx' */
        const int k = w >> 3, i = w & 7, old = k < kx0;
        ure[k][i] = old ? st->w_tail[w] : 0.0f;
        uim[k][i] = old ? st->w_tail[256 + w] : 0.0f;
    }
    if (d.start) {
        /* per-sub-band lookup tables, by scatter: task = one band of one frequency table */
        for (int w = tid; w < 192; w += nt) {
            if (w < 49) {                                  /* f_tablehigh band w */
                if (w < d.n_hi)
                    for (int b = p->f_tablehigh[w]; b < (int)p->f_tablehigh[w + 1]; b++)
                        if ((unsigned)(b - kx) < 48u) { sm->band_idx[1][b - kx] = (uint8_t)w; sm->band_first[1][b - kx] = b == (int)p->f_tablehigh[w]; }
            } else if (w < 64) {
            } else if (w < 64 + 25) {                      /* f_tablelow */
                const int j = w - 64;
                if (j < d.n_lo)
                    for (int b = p->f_tablelow[j]; b < (int)p->f_tablelow[j + 1]; b++)
                        if ((unsigned)(b - kx) < 48u) { sm->band_idx[0][b - kx] = (uint8_t)j; sm->band_first[0][b - kx] = b == (int)p->f_tablelow[j]; }
            } else if (w < 96) {
            } else if (w < 96 + 5) {                       /* noise-floor bands: the last one runs to the top of the range */
                const int j = w - 96;
                if (j < d.n_q) {
                    const int lo = j ? (int)p->f_tablenoise[j] : kx, hi = j + 1 < d.n_q ? (int)p->f_tablenoise[j + 1] : kx + 48;
                    for (int b = lo; b < hi; b++)
                        if ((unsigned)(b - kx) < 48u) sm->noise_idx[b - kx] = (uint8_t)j;
                }
            } else if (w < 104) {
            } else if (w < 104 + 28) {                     /* limiter bands */
                const int j = w - 104;
                if (j < d.n_lim) {
                    const int lo = j ? (int)p->f_tablelim[j] : kx, hi = j + 1 < d.n_lim ? (int)p->f_tablelim[j + 1] : kx + 48;
                    for (int b = lo; b < hi; b++)
                        if ((unsigned)(b - kx) < 48u) sm->lim_idx[b - kx] = (uint8_t)j;
                }
            } else if (w < 136) {
            } else if (w < 136 + 7) {                      /* patches (aacsbr.c:1371-1395); task npatch clears what no patch covers */
                const int j = w - 136;
                if (j <= d.npatch) {
                    int acc = kx;
                    for (int q = 0; q < j; q++) acc += p->patch_num_subbands[q];
                    const int hi = j < d.npatch ? acc + (int)p->patch_num_subbands[j] : kx + 48;
                    const int src0 = j < d.npatch ? (int)p->patch_start_subband[j] : 255;
                    for (int b = acc; b < hi; b++) {
                        const int q = j < d.npatch ? src0 + (b - acc) : 255;
                        if ((unsigned)(b - kx) < 48u) sm->patch_src[b - kx] = (uint8_t)(q < 32 ? q : 255);
                    }
                }
            } else if (w < 152) {
            } else {                                       /* time grid: envelope of each output slot */
                const int i = w - 152;
                int e = -1;
                for (int q2 = 0; q2 < nenv; q2++)
                    if (i >= 2 * pc->t_env[q2] && i < 2 * pc->t_env[q2 + 1]) e = q2;
                if (i >= i_hi || i < i_lo) e = -1;
                const int border = e >= 0 && (e == d.e_a0 || e == d.e_a1);
                /* the smoother's oldest tap is slot i - 4: inside the same envelope => the pre-computed value applies */
                int mixed = 0;
                if (e >= 0 && smooth && !border) mixed = i - 4 < HEAAC_MAX(2 * (int)pc->t_env[e], i_lo);
                sm->env_of_slot[i] = (int8_t)e;
                sm->slot_mode[i] = (uint8_t)(border | (mixed << 1));
            }
        }
        /* the g_temp / q_temp rows the smoother reads below the first envelope: rows 2 t_old .. 2 t_old + 3 of the
         * previous frame (aacsbr.c:1637-1640), kept in the state as rows 32..41 */
        for (int w = tid; w < 5 * 48; w += nt) {
            const int r = w / 48, m = w - r * 48;
            const int hr = d.t_old2 - 32 + r;
            const bool have = r < 4 && smooth && !d.reset && hr >= 0 && hr < 10;
            sm->gall[5 + r][m] = have ? st->g_hist[hr * 48 + m] : 0.0f;
            sm->qall[5 + r][m] = have ? st->q_hist[hr * 48 + m] : 0.0f;
        }
        for (int w = tid; w < 6 * 48; w += nt) {
            const int e = w / 48, m = w - e * 48;
            sm->s_idx[e][m] = e == 0 ? st->s_index0[m] : 0;
        }
        for (int i = tid; i < d.n_q; i += nt) {            /* sbr_chirp */
            float old = st->bw_array[i], nb;
            const int m0_ = pc->bs_invf_mode[0][i], m1_ = pc->bs_invf_mode[1][i];
            if (m0_ + m1_ == 1) nb = 0.6f;
            else nb = m0_ == 0 ? 0.0f : m0_ == 1 ? 0.75f : m0_ == 2 ? 0.9f : 0.98f;
            if (nb < old) nb = heaac_fadd(heaac_fmul(0.75f, nb), heaac_fmul(0.25f, old));
            else          nb = heaac_fadd(heaac_fmul(0.90625f, nb), heaac_fmul(0.09375f, old));
            nb = nb < 0.015625f ? 0.0f : nb;
            sm->bw[i] = nb;
            st->bw_array[i] = nb;
        }
        /* sbr_dequant, aacsbr.c:1089-1128: one warp per envelope */
        if (is_cpe && p->bs_coupling) {
            const HeaacSbrChan *p0 = &p->ch[0], *p1 = &p->ch[1];
            const float alpha      = p0->bs_amp_res ?  1.0f :  0.5f;
            const float pan_offset = p0->bs_amp_res ? 12.0f : 24.0f;
            const int ne0 = HEAAC_MIN((int)p0->bs_num_env, 5), nn0 = HEAAC_MIN((int)p0->bs_num_noise, 2);
            for (int e = warp; e < ne0; e += nwarps) {
                const int nb = p->n[p0->bs_freq_res[e + 1] & 1];
                for (int k = lane; k < nb && k < 48; k += ls) {
                    float t1 = heaac_exp2_half((float)p0->env_q[e][k] * alpha + 7.0f);
                    float t2 = heaac_exp2_half((pan_offset - (float)p1->env_q[e][k]) * alpha);
                    float fac = heaac_fdiv(t1, 1.0f + t2);
                    sm->env[e][k] = ch ? fac * t2 : fac;
                }
            }
            for (int w = tid; w < nn0 * 8; w += nt) {
                const int e = w >> 3, k = w & 7;
                if (k >= d.n_q) continue;
                float t1 = heaac_exp2_half(6.0f - (float)p0->noise_q[e][k] + 1);
                float t2 = heaac_exp2_half(12 - (float)p1->noise_q[e][k]);
                float fac = heaac_fdiv(t1, 1.0f + t2);
                sm->noise[e][k] = ch ? fac * t2 : fac;
            }
        } else {
            const float alpha = pc->bs_amp_res ? 1.0f : 0.5f;
            for (int e = warp; e < nenv; e += nwarps) {
                const int nb = p->n[pc->bs_freq_res[e + 1] & 1];
                for (int k = lane; k < nb && k < 48; k += ls)
                    sm->env[e][k] = heaac_exp2_half(alpha * (float)pc->env_q[e][k] + 6.0f);
            }
            for (int w = tid; w < d.nnoise * 8; w += nt) {
                const int e = w >> 3, k = w & 7;
                if (k >= d.n_q) continue;
                sm->noise[e][k] = heaac_exp2_half(6.0f - (float)pc->noise_q[e][k]);
            }
        }
    }
    HEAAC_SYNC();

    /* ================= phase 2 ================= */
    for (int w = tid; w < 2 * 32 * 8; w += nt) {           /* next frame's W tail: t = 32..39, all 32 bands (before anything overwrites them) */
        const int pl = w >> 8, k = (w >> 3) & 31, i = w & 7;
        st->w_tail[w] = pl ? uim[k][32 + i] : ure[k][32 + i];
    }
    if (d.start) {
        /* autocorrelation: five independent 37-term chains per low band, aacsbr.c:1232-1255 */
        for (int w = tid; w < d.k0 * 5; w += nt) {
            const int k = w / 5, q = w - k * 5;
            const float *xr = ure[k], *xi = uim[k];
            const int lag = q == 0 ? 0 : (q <= 2 ? 1 : 2);
            float acc = 0.0f;
            if (q == 0 || q == 1 || q == 3) {
                for (int i = 1; i < 38; i++)
                    acc = heaac_fadd(acc, heaac_fadd(heaac_fmul(xr[i], xr[i + lag]), heaac_fmul(xi[i], xi[i + lag])));
            } else {
                for (int i = 1; i < 38; i++)
                    acc = heaac_fadd(acc, heaac_fsub(heaac_fmul(xr[i], xi[i + lag]), heaac_fmul(xi[i], xr[i + lag])));
            }
            sm->phi[k][q] = acc;
        }
        /* sbr_mapping: one warp per envelope */
        for (int e = warp; e < nenv; e += nwarps) {
            const int fr = pc->bs_freq_res[e + 1] & 1;
            const int kq = (d.nnoise > 1) && (pc->t_env[e] >= pc->t_q[1]);
            for (int m = lane; m < m_max; m += ls) {
                sm->e_orig[e][m] = sm->env[e][sm->band_idx[fr][m]];
                sm->q_map[e][m] = sm->noise[kq][sm->noise_idx[m]];
            }
            if (pc->bs_add_harmonic_flag)
                for (int i = lane; i < d.n_hi; i += ls) {
                    const int mid = (((int)p->f_tablehigh[i] + (int)p->f_tablehigh[i + 1]) >> 1) - kx;
                    if ((unsigned)mid < 48u)
                        sm->s_idx[e + 1][mid] = pc->bs_add_harmonic[i] * (e >= d.e_a1 || sm->s_idx[0][mid] == 1);
                }
        }
    }
    HEAAC_SYNC();

    if (d.start) {
        /* ================= phase 3: LPC per low band (aacsbr.c:1261-1313) | s_mapped ================= */
        for (int k = tid; k < d.k0; k += nt) {
            const float *xr = ure[k], *xi = uim[k];
            const float rs0 = sm->phi[k][0], rs1 = sm->phi[k][1], is1 = sm->phi[k][2], rs2 = sm->phi[k][3], is2 = sm->phi[k][4];
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
        for (int e = (warp + nwarps - 1) % nwarps; e < nenv; e += nwarps) {   /* (the LPC sits in warp 0: envelope 0 goes to warp 1) */
            const int fr = pc->bs_freq_res[e + 1] & 1;
            const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
            for (int m = lane; m < m_max; m += ls) {
                const int b = sm->band_idx[fr][m];
                int present = 0;
                for (int mm = tab[b]; mm < (int)tab[b + 1]; mm++)
                    if ((unsigned)(mm - kx) < 48u && sm->s_idx[e + 1][mm - kx]) { present = 1; break; }
                sm->s_map[e][m] = (uint8_t)present;
            }
        }
        HEAAC_SYNC();

        /* ================= phase 4: sbr_hf_gen, rows = target bands, lanes = t ================= */
        const int t_first = i_lo + 2, t_wide = HEAAC_MIN(i_hi + 2, i_lo + 34);    /* the first 32 slots of the grid */
        for (int m = warp; m < m_max; m += nwarps) {
            const int q = sm->patch_src[m];
            float *hr = ure[kx + m], *hi = uim[kx + m];
            if (q == 255) {                                /* beyond the last patch: X_high = 0 (memset, aacsbr.c:1403-1406) */
                for (int t = t_first + lane; t < t_wide; t += ls) { hr[t] = 0.0f; hi[t] = 0.0f; }
                continue;
            }
            const float bw = sm->bw[sm->noise_idx[m]];
            const float a0 = sm->alpha[q][2] * bw * bw, a1 = sm->alpha[q][3] * bw * bw;
            const float a2 = sm->alpha[q][0] * bw,      a3 = sm->alpha[q][1] * bw;
            const float *xr0 = ure[q], *xi0 = uim[q];
            for (int t = t_first + lane; t < t_wide; t += ls) {
                const float *xr = xr0 + t, *xi = xi0 + t;
                hr[t] = xr[-2] * a0 - xi[-2] * a1 + xr[-1] * a2 - xi[-1] * a3 + xr[0];
                hi[t] = xi[-2] * a0 + xr[-2] * a1 + xi[-1] * a2 + xr[-1] * a3 + xi[0];
            }
        }
        for (int w = tid; w < m_max * 8; w += nt) {        /* the (at most 6) remaining slots, all bands at once */
            const int m = w >> 3, t = i_lo + 34 + (w & 7);
            if (t >= i_hi + 2) continue;
            const int q = sm->patch_src[m];
            float *hr = ure[kx + m], *hi = uim[kx + m];
            if (q == 255) { hr[t] = 0.0f; hi[t] = 0.0f; continue; }
            const float bw = sm->bw[sm->noise_idx[m]];
            const float a0 = sm->alpha[q][2] * bw * bw, a1 = sm->alpha[q][3] * bw * bw;
            const float a2 = sm->alpha[q][0] * bw,      a3 = sm->alpha[q][1] * bw;
            const float *xr = ure[q] + t, *xi = uim[q] + t;
            hr[t] = xr[-2] * a0 - xi[-2] * a1 + xr[-1] * a2 - xi[-1] * a3 + xr[0];
            hi[t] = xi[-2] * a0 + xr[-2] * a1 + xi[-1] * a2 + xr[-1] * a3 + xi[0];
        }
        HEAAC_SYNC();

        /* ================= phase 5: sbr_env_estimate, lanes = bands, serial over the envelope's slots ================= */
        for (int e = warp; e < nenv; e += nwarps) {
            const int ilb = HEAAC_MIN(pc->t_env[e] * 2 + 2, 40), iub = HEAAC_MIN(pc->t_env[e + 1] * 2 + 2, 40);
            if (p->bs_interpol_freq) {
                const float recip = heaac_fdiv(0.5f, (float)(pc->t_env[e + 1] - pc->t_env[e]));
                for (int m = lane; m < m_max; m += ls) {
                    const float *xr = ure[kx + m], *xi = uim[kx + m];
                    float sum = 0.0f;
                    for (int i = ilb; i < iub; i++)
                        sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(xr[i], xr[i]), heaac_fmul(xi[i], xi[i])));
                    sm->e_curr[e][m] = heaac_fmul(sum, recip);
                }
            } else {
                const int fr = pc->bs_freq_res[e + 1] & 1;
                const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
                const int nb = fr ? d.n_hi : d.n_lo;
                const int env_size = 2 * (pc->t_env[e + 1] - pc->t_env[e]);
                for (int b = lane; b < nb; b += ls) {      /* one lane per envelope band: the whole group in the reference's order */
                    const int lo = HEAAC_MAX((int)tab[b], kx), hi = HEAAC_MIN((int)tab[b + 1], kx + m_max);
                    const int den = env_size * ((int)tab[b + 1] - (int)tab[b]);
                    float sum = 0.0f;
                    for (int k = lo; k < hi; k++)
                        for (int i = ilb; i < iub; i++)
                            sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(ure[k][i], ure[k][i]), heaac_fmul(uim[k][i], uim[k][i])));
                    sum = heaac_fdiv(sum, (float)den);
                    for (int k = lo; k < hi; k++) sm->e_curr[e][k - kx] = sum;
                }
            }
        }
        HEAAC_SYNC();

        /* ================= phase 6: sbr_gain_calc, aacsbr.c:1552-1605: element-wise passes + per-limiter-band sums ================= */
        for (int e = warp; e < nenv; e += nwarps) {
            const int delta = !((e == d.e_a1) || (e == d.e_a0));
            for (int m = lane; m < m_max; m += ls) {
                const float eo = sm->e_orig[e][m], qm = sm->q_map[e][m], ec = sm->e_curr[e][m];
                const float temp = heaac_fdiv(eo, heaac_fadd(1.0f, qm));
                sm->qall[e][m] = heaac_fsqrt(heaac_fmul(temp, qm));
                sm->s_m[e][m] = heaac_fsqrt(heaac_fmul(temp, (float)sm->s_idx[e + 1][m]));
                if (!sm->s_map[e][m])
                    sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(eo, heaac_fmul(heaac_fadd(1.0f, ec),
                                                 heaac_fadd(1.0f, heaac_fmul(qm, (float)delta)))));
                else
                    sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(heaac_fmul(eo, qm),
                                                 heaac_fmul(heaac_fadd(1.0f, ec), heaac_fadd(1.0f, qm))));
            }
        }
        for (int w = tid; w < nenv * d.n_lim; w += nt) {
            const int e = w / d.n_lim, k = w - e * d.n_lim;
            const int lo = HEAAC_MAX((int)p->f_tablelim[k] - kx, 0), hi = HEAAC_MIN((int)p->f_tablelim[k + 1] - kx, m_max);
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int m = lo; m < hi; m++) { sum0 = heaac_fadd(sum0, sm->e_orig[e][m]); sum1 = heaac_fadd(sum1, sm->e_curr[e][m]); }
            const float gain_max = heaac_fmul(heaac_limgain[p->bs_limiter_gains & 3],
                                              heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sum0),
                                                                     heaac_fadd(1.1920928955078125e-07f, sum1))));
            sm->gmax[e][k] = HEAAC_MIN(100000.0f, gain_max);
        }
        HEAAC_SYNC();
        for (int e = warp; e < nenv; e += nwarps)
            for (int m = lane; m < m_max; m += ls) {
                const float gain_max = sm->gmax[e][sm->lim_idx[m]];
                const float q_m_max = heaac_fdiv(heaac_fmul(sm->qall[e][m], gain_max), sm->gall[e][m]);
                sm->qall[e][m] = HEAAC_MIN(sm->qall[e][m], q_m_max);
                sm->gall[e][m] = HEAAC_MIN(sm->gall[e][m], gain_max);
            }
        HEAAC_SYNC();
        for (int w = tid; w < nenv * d.n_lim; w += nt) {
            const int e = w / d.n_lim, k = w - e * d.n_lim;
            const int delta = !((e == d.e_a1) || (e == d.e_a0));
            const int lo = HEAAC_MAX((int)p->f_tablelim[k] - kx, 0), hi = HEAAC_MIN((int)p->f_tablelim[k + 1] - kx, m_max);
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int m = lo; m < hi; m++) {
                const float g = sm->gall[e][m], s = sm->s_m[e][m], q = sm->qall[e][m];
                sum0 = heaac_fadd(sum0, sm->e_orig[e][m]);
                const float t3 = heaac_fmul(heaac_fmul((float)(delta && !s), q), q);
                const float acc = heaac_fadd(heaac_fadd(heaac_fmul(heaac_fmul(sm->e_curr[e][m], g), g), heaac_fmul(s, s)), t3);
                sum1 = heaac_fadd(sum1, acc);
            }
            const float gain_boost = heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sum0),
                                                            heaac_fadd(1.1920928955078125e-07f, sum1)));
            /* FFMIN(1.584893192 (double), gain_boost) then narrowed to float */
            sm->gboost[e][k] = (1.584893192 > (double)gain_boost) ? gain_boost : (float)1.584893192;
        }
        HEAAC_SYNC();
        for (int e = warp; e < nenv; e += nwarps)
            for (int m = lane; m < m_max; m += ls) {
                const float gain_boost = sm->gboost[e][sm->lim_idx[m]];
                const float gf = heaac_fmul(sm->gall[e][m], gain_boost), qf = heaac_fmul(sm->qall[e][m], gain_boost);
                sm->gall[e][m] = gf;
                sm->qall[e][m] = qf;
                sm->s_m[e][m]  = heaac_fmul(sm->s_m[e][m], gain_boost);
                if (smooth) {                              /* all five taps in this envelope: sum_j g h[j], in the smoother's own order */
                    float g_filt = 0.0f, q_filt = 0.0f;
                    HEAAC_UNROLL
                    for (int j = 0; j <= 4; j++) {
                        g_filt = heaac_fadd(g_filt, heaac_fmul(gf, heaac_h_smooth[j]));
                        q_filt = heaac_fadd(q_filt, heaac_fmul(qf, heaac_h_smooth[j]));
                    }
                    sm->gsm[e][m] = g_filt; sm->qsm[e][m] = q_filt;
                }
            }
        HEAAC_SYNC();

        /* ================= phase 7: sbr_hf_assemble in place, U[kx + m][i + 2] <- Y[i][kx + m] ================= */
        const uint32_t noise0 = d.reset ? 0u : st->f_indexnoise, sine0 = st->f_indexsine;
        const int kx_sign = 1 - 2 * (kx & 1);
        {   /* (a) slots 0..31 whose gain is constant per (envelope, band): rows = bands, lanes = slots; everything that
             * depends on the slot only is worked out once per lane */
          for (int i = lane; i < 32; i += ls) {
            const int e = sm->env_of_slot[i];
            const int mode = e >= 0 ? sm->slot_mode[i] : 0;
            const bool active = e >= 0 && !(mode & 2) && i < 32;
            const int border = mode & 1;
            const int es = e >= 0 ? e : 0;
            const float *grow = (smooth && !border) ? sm->gsm[es] : sm->gall[es];
            const float *qrow = (smooth && !border) ? sm->qsm[es] : sm->qall[es];
            const float *srow = sm->s_m[es];
            const int indexsine = (int)((sine0 + (uint32_t)(i - i_lo)) & 3);
            const float phi_re = indexsine == 0 ? 1.0f : indexsine == 2 ? -1.0f : 0.0f;
            const int phi_im = indexsine == 1 ? 1 : indexsine == 3 ? -1 : 0;
            const uint32_t nbase = noise0 + (uint32_t)(i - i_lo) * (uint32_t)m_max + 1u;
            for (int m = warp; m < m_max; m += nwarps) {
                if (!active) continue;
                float *yr = &ure[kx + m][i + 2], *yi = &uim[kx + m][i + 2];
                const float g_filt = grow[m], q_filt = qrow[m], s = srow[m];
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
                        else if (d.reset) r = 0;
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
                    float *yr = &ure[kx + m][i + 2], *yi = &uim[kx + m][i + 2];
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
        HEAAC_SYNC();
    }

    /* ================= phase 8: sbr_x_gen -- cells no stage wrote -- and the state for the next frame ================= */
    {
        const int i_temp = d.i_temp, m0 = d.m0;
        const int hf_lo = d.start ? kx : 64, hf_hi = d.start ? kx + m_max : 64;
        /* (a) the Y tail: slots 32..37 of the high band.  One work item per (band, slot - 32): it reads the previous
         *     frame's value (needed where this frame's first slots still belong to the previous grid, aacsbr.c:1420-1431)
         *     before it stores this frame's */
        for (int w = tid; w < 64 * 8; w += nt) {
            const int k = w >> 3, j = w & 7;
            if (j >= 6) continue;
            if (j < i_temp && k >= kx0) {
                const bool hf_old = k < kx0 + m0;
                ure[k][j + 2] = hf_old ? st->y_tail[k * 8 + j] : 0.0f;
                uim[k][j + 2] = hf_old ? st->y_tail[512 + k * 8 + j] : 0.0f;
            }
            if (k >= hf_lo && k < hf_hi && 32 + j >= i_lo && 32 + j < i_hi) {
                st->y_tail[k * 8 + j] = ure[k][34 + j]; st->y_tail[512 + k * 8 + j] = uim[k][34 + j];
            }
        }
        /* (b) slots >= i_temp: bands that are neither low band nor covered high band are silent */
        for (int k = kx + warp; k < 64; k += nwarps) {
            const bool hf = k >= hf_lo && k < hf_hi;
            for (int i = i_temp + lane; i < 38; i += ls)
                if (!hf || i < i_lo || i >= i_hi) { ure[k][i + 2] = 0.0f; uim[k][i + 2] = 0.0f; }
        }
        /* slots i_temp <= i < 8 of low-band rows kx' <= k < kx hold zeros from phase 1 (X_low[k][t < 8] = 0 for k >= kx') */
        if (d.start) {
            /* g_temp / q_temp rows 32..41 for the next frame, aacsbr.c:1642-1647 */
            for (int w = tid; w < 10 * 48; w += nt) {
                const int r = 32 + w / 48, m = w % 48;
                const int src = r - d.h_SL;
                if (m < m_max && src >= i_lo && src < i_hi) {
                    const int es = sm->env_of_slot[src];
                    if (es >= 0) { st->g_hist[w] = sm->gall[es][m]; st->q_hist[w] = sm->qall[es][m]; }
                }
            }
            for (int m = tid; m < 48; m += nt) st->s_index0[m] = sm->s_idx[nenv][m];
            if (tid == 0) {
                const uint32_t noise0 = d.reset ? 0u : st->f_indexnoise;
                st->f_indexnoise = (noise0 + (uint32_t)(i_hi - i_lo) * (uint32_t)m_max) & 0x1ff;
                st->f_indexsine = (st->f_indexsine + (uint32_t)(i_hi - i_lo)) & 3;
            }
        }
    }
    HEAAC_SYNC();
}

/* ------------------------------------------------------------------------------------
 * U -> the inter-kernel X buffer (Parametric Stereo and downsampled synthesis read it): X[plane][band][HEAAC_XS],
 * slots 0..37.
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_x_emit(const float *u, float *X_out, int tid)
{
    constexpr int nt = NT;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    for (int r = warp; r < 128; r += nwarps) {             /* 2 planes x 64 bands: rows are contiguous in both arrays */
        const float *src = u + r * HEAAC_US + 2;
        float *dst = X_out + r * HEAAC_XS;
        for (int i = lane; i < 38; i += ls) dst[i] = src[i];
    }
}

/* ------------------------------------------------------------------------------------
 * QMF synthesis (sbr_qmf_synthesis, aacsbr.c:1175-1230) + float->s16 of one channel, in place on U.
 *   vh  : the 9 v blocks before this frame (1152 floats, newest first) in shared memory; replaced by this frame's
 *   st  : receives the new history when `store_state`
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_synth_inplace(float *u, float *vh, HeaacSynthState *st, int store_state, int nout, int out_slot,
                                   int16_t *pcm, float *pcm_f32, float scale, float bias,
                                   const HeaacDevTables *HEAAC_RESTRICT T, int tid)
{
    constexpr int nt = NT;
    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    float *ure = u, *uim = u + HEAAC_U_PLANE;
    /* ---- the two half-IMDCTs (N = 128, scale 1/64) of every slot: thread = (plane, slot), column in, column out ---- */
    for (int w = tid; w < 64; w += nt) {
        const int plane = w >> 5, i = w & 31;
        float *col = (plane ? uim : ure) + (i + 2);
        if (plane == 0)
            heaac_imdct128_reg<HEAAC_MDCT_SYN>(
                [&](int j) { return col[j * HEAAC_US]; },
                [&](int j, float b) { col[j * HEAAC_US] = b; });
        else
            heaac_imdct128_reg<HEAAC_MDCT_SYN>(
                [&](int j) { const float xv = col[j * HEAAC_US]; return (j & 1) ? -xv : xv; },     /* aacsbr.c:1205-1207 */
                [&](int j, float b) { col[j * HEAAC_US] = b; });
    }
    HEAAC_SYNC();
    /* ---- v[n] = -b0[63-n] + b1[n], v[127-n] = b0[63-n] + b1[n] (aacsbr.c:1210-1213): the column becomes the slot's v block,
     *      v[n] in the imaginary plane's row n, v[64 + j] in the real plane's row j ---- */
    for (int n = warp; n < 64; n += nwarps)
        for (int i = lane; i < 32; i += ls) {
            float *pr = ure + (63 - n) * HEAAC_US + (i + 2), *pi = uim + n * HEAAC_US + (i + 2);
            const float p1 = *pr, q1 = *pi;
            *pi = -p1 + q1;
            *pr =  p1 + q1;
        }
    HEAAC_SYNC();
    /* ---- out[n] = sum_j v[off_j + n] w[64 j + n] (aacsbr.c:1215-1227): tap j reads the block of slot i - j, row n of the
     *      imaginary (j even) or real (j odd) plane -- two rows per output sample, walked with a register window ---- */
    constexpr int G = NT >= 256 ? 4 : NT >= 128 ? 2 : 1;   /* groups of consecutive slots */
    constexpr int L = 32 / G;
    const int scale_and_bias = scale != 1.0f || bias != 0.0f;
    const int bias_mode = bias != 0.0f;
    for (int w = tid; w < 64 * G; w += nt) {
        const int n = w & 63, i0 = (w >> 6) * L;
        float wj[10], im[L + 9], re[L + 9];
        HEAAC_UNROLL
        for (int j = 0; j < 10; j++) wj[j] = T->qmf_us[64 * j + n];
        const float *rowi = uim + n * HEAAC_US + 2, *rowr = ure + n * HEAAC_US + 2;
        HEAAC_UNROLL
        for (int a = 0; a < L + 9; a++) {                  /* a <-> slot i0 - 9 + a */
            const int sl = i0 - 9 + a;
            im[a] = sl >= 0 ? rowi[sl] : vh[(-1 - sl) * 128 + n];
            re[a] = sl >= 0 ? rowr[sl] : vh[(-1 - sl) * 128 + 64 + n];
        }
        HEAAC_UNROLL
        for (int ii = 0; ii < L; ii++) {
            float acc = im[9 + ii] * wj[0] + 0.0f;
            HEAAC_UNROLL
            for (int j = 1; j < 10; j++)
                acc = ((j & 1) ? re[9 + ii - j] : im[9 + ii - j]) * wj[j] + acc;
            if (scale_and_bias) acc = acc * scale + bias;
            const int o = (i0 + ii) * 64 + n;
            if (pcm)     pcm[(size_t)o * nout + out_slot] = heaac_to_s16(acc, bias_mode);
            if (pcm_f32) pcm_f32[(size_t)out_slot * 2048 + o] = acc;
        }
    }
    HEAAC_SYNC();
    for (int w = tid; w < 1152; w += nt) {                 /* the newest 9 blocks (slots 31..23) are the next frame's history */
        const int b = w >> 7, idx = w & 127;
        const float v = idx < 64 ? uim[idx * HEAAC_US + (31 - b) + 2] : ure[(idx - 64) * HEAAC_US + (31 - b) + 2];
        vh[w] = v;
        if (store_state) st->v[w] = v;
    }
    HEAAC_SYNC();
}

#endif /* HEAAC_CHAIN_CUH */
