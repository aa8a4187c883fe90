/*
 * heaac_sbr.cuh -- SBR stages of the synthesis path as CTA-level routines (sm_100a).
 *
 *   heaac_qmf_analysis : sbr_qmf_analysis (aacsbr.c:1136-1169).  Runs inside the core kernel right after
 *                        the IMDCT, on the time samples still in shared memory.  Windowing with lanes =
 *                        sub-band sample k, then one register-resident half-IMDCT per time slot with
 *                        lanes = slots; W leaves the CTA band-major (slot fastest), coalesced.
 *   heaac_hf_channel   : sbr_lf_gen 1337-1357, sbr_hf_inverse_filter 1232-1313, sbr_chirp 1316-1334,
 *                        sbr_dequant 1089-1128, sbr_hf_gen 1360-1409, sbr_mapping 1451-1496,
 *                        sbr_env_estimate 1499-1546, sbr_gain_calc 1552-1605, sbr_hf_assemble 1608-1714 and
 *                        sbr_x_gen 1412-1446 for one channel-frame.  All QMF-domain arrays are planar
 *                        re/im with the time slot as the fastest index and an odd row stride (41), so that
 *                        both "lanes = slots" and "lanes = bands" phases are free of bank conflicts.
 *
 * Order-sensitive arithmetic (autocorrelation, LPC, gains) keeps the reference's evaluation order and uses
 * the non-contracting heaac_f* primitives.
 */
#ifndef HEAAC_SBR_CUH
#define HEAAC_SBR_CUH

#include "heaac_dev.cuh"
#include "heaac_regfft.cuh"

#define HEAAC_QS 41                 /* row stride of the [band][slot] arrays (40 slots used) */

/* (row, lane) loop nests: rows are dealt to warps, the fastest index to lanes, so that everything that
 * depends only on the row is computed once per warp iteration.  With nt = 1 (host emulation) they
 * degenerate to plain nested loops. */
#define HEAAC_WARP(tid, nt)    ((nt) >= 32 ? (tid) >> 5 : 0)
#define HEAAC_NWARPS(nt)       ((nt) >= 32 ? (nt) >> 5 : 1)
#define HEAAC_LANE(tid, nt)    ((nt) >= 32 ? (tid) & 31 : 0)
#define HEAAC_LSTRIDE(nt)      ((nt) >= 32 ? 32 : 1)

/* ------------------------------------------------------------------------------------
 * QMF analysis.  x: 1312 floats in shared memory (288 history + 1024 new, already scaled).
 * tb: 64 * 33 floats of scratch (folded + shuffled transform inputs, [input index][slot], stride 33).
 * W_out: this channel's W buffer in global memory, [2 planes][32 bands][32 slots].
 * ---------------------------------------------------------------------------------- */
template <int NT>
HEAAC_DEV void heaac_qmf_analysis(const float *x, float *tb, float *W_out,
                                  const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
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
        /* t[0] = u[0]; t[2k-1] = u[k], t[2k] = -u[64-k] (k = 1..31); t[63] = u[32] */
        int pos; float val = f;
        if (k == 0) pos = 0;
        else if (k < 32) pos = 2 * k - 1;
        else if (k == 32) pos = 63;
        else { pos = 2 * (64 - k); val = -f; }
        tb[pos * 33 + i] = val;
    }
    HEAAC_SYNC();
    for (int i = tid; i < 32; i += nt) {                   /* one half-IMDCT (N = 128, scale -2) per slot */
        float *wre = W_out + i, *wim = W_out + 32 * 32 + i;
        heaac_imdct128_reg<HEAAC_MDCT_ANA>(
            [&](int j) { return tb[j * 33 + i]; },
            [&](int j, float v) {                          /* W[i][k] = (-y[63-k], y[k]), aacsbr.c:1163-1166 */
                if (j < 32) wim[j * 32] = v;
                else        wre[(63 - j) * 32] = -v;
            });
    }
}

/* ------------------------------------------------------------------------------------ */
struct HeaacHfSmem {
    float xl_re[32][HEAAC_QS], xl_im[32][HEAAC_QS];      /* X_low[k][i] */
    float xh_re[48][HEAAC_QS], xh_im[48][HEAAC_QS];      /* X_high[kx + m][i] */
    float phi[32][8];               /* per band: sums rs0, rs1, is1, rs2, is2 over i = 1..37 */
    float alpha[32][4];             /* a0.re a0.im a1.re a1.im */
    float env[5][48];               /* dequantised env_facs[1..5] */
    float noise[2][8];
    float e_orig[5][48], q_map[5][48], e_curr[5][48], s_m[5][48];     /* gain / q_m live in gall / qall rows 0..4 */
    float gmax[5][8], gboost[5][8];
    float bw[8];
    float ynew_re[48][8], ynew_im[48][8];   /* Y[1][32..37][kx + m] of this frame, [m][slot - 32] */
    uint8_t s_idx[6][48];           /* s_indexmapped[0..5] */
    uint8_t s_map[5][48];
    uint8_t band_idx[2][48];        /* envelope band (low / high resolution table) containing kx + m */
    uint8_t band_first[2][48];      /* 1 if kx + m is the first sub-band of its envelope band */
    uint8_t noise_idx[48];          /* noise-floor band containing kx + m */
    uint8_t lim_idx[48];            /* limiter band containing kx + m */
    uint8_t patch_src[48];          /* low-band source of kx + m in the HF generator, 255 = beyond the last patch */
    uint8_t ynew_valid[48][8];
    int8_t  env_of_slot[48];
    uint8_t slot_border[48];        /* slot lies in an envelope adjacent to a transient border (l_A / l_APrev) */
    uint8_t smooth_src[40][5];      /* row of gall/qall feeding tap j of the 5-tap gain smoother at slot i */
    float gall[16][48], qall[16][48];   /* rows 0..4: this frame's envelopes, 5..14: g_temp/q_temp history rows 32..41, 15: zero */
    HeaacSbrElem p;
};

/* exact 2^(k/2): exp2f of a multiple of 0.5 equals ldexp(1 or sqrt2, n) (checked against glibc) */
HEAAC_DEV float heaac_exp2_half(float x)
{
    float fl = floorf(x);
    float m = (x != fl) ? 1.41421356237309504880f : 1.0f;
    return ldexpf(m, (int)fl);
}

template <int NT>
HEAAC_DEV void heaac_sbr_dequant(HeaacHfSmem *sm, int ch, int is_cpe,  int tid)
{
    constexpr int nt = NT;
    const HeaacSbrElem *p = &sm->p;
    const HeaacSbrChan *pc = &p->ch[ch];
    if (is_cpe && p->bs_coupling) {
        const HeaacSbrChan *p0 = &p->ch[0], *p1 = &p->ch[1];
        float alpha      = p0->bs_amp_res ?  1.0f :  0.5f;
        float pan_offset = p0->bs_amp_res ? 12.0f : 24.0f;
        for (int w = tid; w < p0->bs_num_env * 48; w += nt) {
            int e = w / 48, k = w - e * 48;
            if (k >= p->n[p0->bs_freq_res[e + 1]]) continue;
            float t1 = heaac_exp2_half((float)p0->env_q[e][k] * alpha + 7.0f);
            float t2 = heaac_exp2_half((pan_offset - (float)p1->env_q[e][k]) * alpha);
            float fac = heaac_fdiv(t1, 1.0f + t2);
            sm->env[e][k] = ch ? fac * t2 : fac;
        }
        for (int w = tid; w < p0->bs_num_noise * 8; w += nt) {
            int e = w >> 3, k = w & 7;
            if (k >= p->n_q) continue;
            float t1 = heaac_exp2_half(6.0f - (float)p0->noise_q[e][k] + 1);
            float t2 = heaac_exp2_half(12 - (float)p1->noise_q[e][k]);
            float fac = heaac_fdiv(t1, 1.0f + t2);
            sm->noise[e][k] = ch ? fac * t2 : fac;
        }
    } else {
        float alpha = pc->bs_amp_res ? 1.0f : 0.5f;
        for (int w = tid; w < pc->bs_num_env * 48; w += nt) {
            int e = w / 48, k = w - e * 48;
            if (k >= p->n[pc->bs_freq_res[e + 1]]) continue;
            sm->env[e][k] = heaac_exp2_half(alpha * (float)pc->env_q[e][k] + 6.0f);
        }
        for (int w = tid; w < pc->bs_num_noise * 8; w += nt) {
            int e = w >> 3, k = w & 7;
            if (k >= p->n_q) continue;
            sm->noise[e][k] = heaac_exp2_half(6.0f - (float)pc->noise_q[e][k]);
        }
    }
}

struct HeaacYCtx {
    int h_SL, i_lo, i_hi, t_old2, kx, m_max, reset, e_a0, e_a1;
    uint32_t noise0, sine0;
};

/* Y[1][i][kx + m] of sbr_hf_assemble (aacsbr.c:1649-1710) for one (slot, band).  The g_temp/q_temp rows the
 * 5-tap smoother reads (aacsbr.c:1632-1647, 1653-1658) are resolved per slot into sm->smooth_src. */
HEAAC_DEV void heaac_sbr_y(const HeaacHfSmem *sm, const HeaacYCtx *y,
                           const HeaacDevTables *HEAAC_RESTRICT T, int i, int m, float *yre_o, float *yim_o)
{
    const int e = sm->env_of_slot[i];
    const int at_border = sm->slot_border[i];
    float g_filt, q_filt;
    if (y->h_SL && !at_border) {
        g_filt = 0.0f; q_filt = 0.0f;
        for (int j = 0; j <= 4; j++) {
            const int r = sm->smooth_src[i][j];
            g_filt = heaac_fadd(g_filt, heaac_fmul(sm->gall[r][m], heaac_h_smooth[j]));
            q_filt = heaac_fadd(q_filt, heaac_fmul(sm->qall[r][m], heaac_h_smooth[j]));
        }
    } else {
        g_filt = sm->gall[e][m];
        q_filt = sm->qall[e][m];
    }
    float yre = heaac_fmul(sm->xh_re[m][i + 2], g_filt);
    float yim = heaac_fmul(sm->xh_im[m][i + 2], g_filt);
    const int indexsine = (int)((y->sine0 + (uint32_t)(i - y->i_lo)) & 3);
    const int phi_sign = (1 - 2 * (y->kx & 1)) * ((m & 1) ? -1 : 1);
    const int phi_re = indexsine == 0 ? 1 : indexsine == 2 ? -1 : 0;
    const int phi_im = indexsine == 1 ? 1 : indexsine == 3 ? -1 : 0;
    const float s = sm->s_m[e][m];
    if (!at_border && !s) {
        uint32_t idx = (y->noise0 + (uint32_t)(i - y->i_lo) * (uint32_t)y->m_max + (uint32_t)m + 1u) & 0x1ff;
        yre = heaac_fadd(yre, heaac_fmul(q_filt, T->noise[2 * idx]));
        yim = heaac_fadd(yim, heaac_fmul(q_filt, T->noise[2 * idx + 1]));
    } else {
        yre = heaac_fadd(yre, heaac_fmul(s, (float)phi_re));
        yim = heaac_fadd(yim, heaac_fmul(s, (float)(phi_im * phi_sign)));
    }
    *yre_o = yre; *yim_o = yim;
}

/* W_in  : this channel's W of the current frame, [2][32 bands][32 slots] (global, written by the core kernel)
 * X_out : this channel's X[2][64][HEAAC_XS] (band-major) buffer in global scratch. */
template <int NT>
HEAAC_DEV void heaac_hf_channel(HeaacHfSmem *sm, const float *HEAAC_RESTRICT W_in,
                                const HeaacSbrElem *HEAAC_RESTRICT elem, int ch, int is_cpe,
                                HeaacChanState *st, float *X_out,
                                const HeaacDevTables *HEAAC_RESTRICT T,  int tid)
{
    constexpr int nt = NT;
    /* this frame's W goes straight into X_low slots 8..39, asynchronously; bands >= kx are cleared once kx is known */
    for (int w = tid; w < 2 * 32 * 32; w += nt) {
        const int pl = w >> 10, k = (w >> 5) & 31, i = w & 31;
        heaac_cp_async4(pl ? &sm->xl_im[k][8 + i] : &sm->xl_re[k][8 + i], W_in + w);
    }
    heaac_cp_async_commit();
    {
        const uint32_t *src = (const uint32_t *)elem;
        uint32_t *dst = (uint32_t *)&sm->p;
        for (int i = tid; i < (int)(sizeof(HeaacSbrElem) / 4); i += nt) dst[i] = src[i];
    }
    HEAAC_SYNC();
    const HeaacSbrElem *p = &sm->p;
    const HeaacSbrChan *pc = &p->ch[ch];
    const int kx = p->kx[1], kx0 = p->kx[0], m_max = p->m[1], m0 = p->m[0];
    const int nenv = pc->bs_num_env;
    const int i_lo = p->start ? 2 * pc->t_env[0] : 0, i_hi = p->start ? 2 * pc->t_env[nenv] : 0;

    const int warp = HEAAC_WARP(tid, nt), nwarps = HEAAC_NWARPS(nt), lane = HEAAC_LANE(tid, nt), ls = HEAAC_LSTRIDE(nt);
    /* ---- sbr_lf_gen: X_low[k][i] = W[k][i - 8] (i >= 8, k < kx) | previous frame's W[k][24 + i] (i < 8, k < kx') ---- */
    for (int w = tid; w < 32 * 8; w += nt) {               /* slots 0..7 = previous frame's W[k][24..31] */
        const int k = w >> 3, i = w & 7, old = k < kx0;
        sm->xl_re[k][i] = old ? st->w_tail[w] : 0.0f;
        sm->xl_im[k][i] = old ? st->w_tail[256 + w] : 0.0f;
    }
    /* per-sub-band lookup tables */
    for (int m = tid; m < 48; m += nt) {
        const int band = kx + m;
        int q = 255;
        if (m < m_max) {
            int acc = kx;
            for (int j = 0; j < p->num_patches; j++) {
                int nsb = p->patch_num_subbands[j];
                if (band < acc + nsb) { q = p->patch_start_subband[j] + (band - acc); break; }
                acc += nsb;
            }
        }
        sm->patch_src[m] = (uint8_t)q;
        int g = 0;
        while (g + 1 < p->n_q && band >= p->f_tablenoise[g + 1]) g++;
        sm->noise_idx[m] = (uint8_t)g;
        int l = 0;
        while (l + 1 < p->n_lim && band >= p->f_tablelim[l + 1]) l++;
        sm->lim_idx[m] = (uint8_t)l;
        for (int fr = 0; fr < 2; fr++) {
            const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
            int b = 0;
            while (b + 1 < p->n[fr] && band >= tab[b + 1]) b++;
            sm->band_idx[fr][m] = (uint8_t)b;
            sm->band_first[fr][m] = band == tab[b];
        }
        int e = -1;
        for (int q2 = 0; q2 < nenv; q2++)
            if (m >= 2 * pc->t_env[q2] && m < 2 * pc->t_env[q2 + 1]) e = q2;
        sm->env_of_slot[m] = (int8_t)e;                    /* (index m doubles as the slot index here) */
        sm->slot_border[m] = (uint8_t)(e == pc->e_a[0] || e == pc->e_a[1]);
        if (m < 40) {
            /* g_temp row (i + 4 - j) of the smoother: slot i - j of this frame, else a history row, aacsbr.c:1632-1647 */
            const int t_old2_ = 2 * pc->t_env_num_env_old;
            for (int j = 0; j <= 4; j++) {
                int src = m - j, r;
                if (src >= i_lo) {
                    r = 15;
                    for (int q2 = 0; q2 < nenv; q2++)
                        if (src >= 2 * pc->t_env[q2] && src < 2 * pc->t_env[q2 + 1]) r = q2;
                } else if (p->reset) {
                    r = 0;
                } else {
                    int hr = t_old2_ + (src + 4 - i_lo) - 32;      /* history row relative to row 32 */
                    r = (hr >= 0 && hr < 10) ? 5 + hr : 15;
                }
                sm->smooth_src[m][j] = (uint8_t)r;
            }
        }
    }
    for (int w = tid; w < 11 * 48; w += nt) {              /* history rows of g_temp / q_temp, plus the zero row */
        const int r = w / 48, m = w - r * 48;
        sm->gall[5 + r][m] = r < 10 ? st->g_hist[w] : 0.0f;
        sm->qall[5 + r][m] = r < 10 ? st->q_hist[w] : 0.0f;
    }
    HEAAC_SYNC();
    heaac_cp_async_wait<0>();
    HEAAC_SYNC();
    for (int w = tid; w < 2 * 32 * 8; w += nt) {           /* next frame's W tail: slots 24..31, all 32 bands */
        const int pl = w >> 8, k = (w >> 3) & 31, i = w & 7;
        st->w_tail[pl * 256 + k * 8 + i] = pl ? sm->xl_im[k][32 + i] : sm->xl_re[k][32 + i];
    }
    /* (sbr_lf_gen zeroes bands >= kx; nothing below ever reads those rows: the LPC uses k < k0 <= kx, the
     *  patches copy from below k0, and X generation takes the low band only for k < kx) */

    if (p->start) {
        /* ---- autocorrelation: five independent 37-term chains per low band, aacsbr.c:1232-1255 ---- */
        for (int w = tid; w < p->k0 * 5; w += nt) {
            const int k = w / 5, q = w - k * 5;
            const float *xr = sm->xl_re[k], *xi = sm->xl_im[k];
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
        for (int i = tid; i < p->n_q; i += nt) {           /* sbr_chirp */
            float old = st->bw_array[i], nb;
            int m0_ = pc->bs_invf_mode[0][i], m1_ = pc->bs_invf_mode[1][i];
            if (m0_ + m1_ == 1) nb = 0.6f;
            else nb = m0_ == 0 ? 0.0f : m0_ == 1 ? 0.75f : m0_ == 2 ? 0.9f : 0.98f;
            if (nb < old) nb = heaac_fadd(heaac_fmul(0.75f, nb), heaac_fmul(0.25f, old));
            else          nb = heaac_fadd(heaac_fmul(0.90625f, nb), heaac_fmul(0.09375f, old));
            nb = nb < 0.015625f ? 0.0f : nb;
            sm->bw[i] = nb;
            st->bw_array[i] = nb;
        }
        heaac_sbr_dequant<NT>(sm, ch, is_cpe, tid);
        for (int w = tid; w < 6 * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            sm->s_idx[e][m] = e == 0 ? st->s_index0[m] : 0;
        }
        HEAAC_SYNC();
        /* ---- second-order covariance LPC per low band, aacsbr.c:1261-1313 ---- */
        for (int k = tid; k < p->k0; k += nt) {
            const float *xr = sm->xl_re[k], *xi = sm->xl_im[k];
            const float rs0 = sm->phi[k][0], rs1 = sm->phi[k][1], is1 = sm->phi[k][2], rs2 = sm->phi[k][3], is2 = sm->phi[k][4];
            /* phi[a][b][c] as in autocorrelate(), aacsbr.c:1242-1253 */
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
        /* ---- sbr_mapping (independent of the LPC) ---- */
        for (int w = tid; w < nenv * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            int fr = pc->bs_freq_res[e + 1];
            sm->e_orig[e][m] = sm->env[e][sm->band_idx[fr][m]];
            int kq = (pc->bs_num_noise > 1) && (pc->t_env[e] >= pc->t_q[1]);
            sm->q_map[e][m] = sm->noise[kq][sm->noise_idx[m]];
        }
        if (pc->bs_add_harmonic_flag) {
            for (int w = tid; w < nenv * 48; w += nt) {
                int e = w / 48, i = w - e * 48;
                if (i >= p->n[1]) continue;
                int mid = ((int)p->f_tablehigh[i] + (int)p->f_tablehigh[i + 1]) >> 1;
                sm->s_idx[e + 1][mid - kx] = pc->bs_add_harmonic[i] *
                    (e >= pc->e_a[1] || sm->s_idx[0][mid - kx] == 1);
            }
        }
        HEAAC_SYNC();
        /* ---- sbr_hf_gen (rows = target bands, lanes = slots) ---- */
        for (int m = warp; m < m_max; m += nwarps) {
            const int q = sm->patch_src[m];
            if (q == 255) {                                /* beyond the last patch: memset(X_high + k, 0) */
                for (int d = lane; d < 40; d += ls) { sm->xh_re[m][d] = 0.0f; sm->xh_im[m][d] = 0.0f; }
                continue;
            }
            const float bw = sm->bw[sm->noise_idx[m]];
            const float a0 = sm->alpha[q][2] * bw * bw, a1 = sm->alpha[q][3] * bw * bw;
            const float a2 = sm->alpha[q][0] * bw,      a3 = sm->alpha[q][1] * bw;
            for (int d = i_lo + 2 + lane; d < i_lo + 34 && d < i_hi + 2; d += ls) {   /* first 32 slots of the grid */
                const float *xr = sm->xl_re[q] + d, *xi = sm->xl_im[q] + d;
                sm->xh_re[m][d] = xr[-2] * a0 - xi[-2] * a1 + xr[-1] * a2 - xi[-1] * a3 + xr[0];
                sm->xh_im[m][d] = xi[-2] * a0 + xr[-2] * a1 + xi[-1] * a2 + xr[-1] * a3 + xi[0];
            }
        }
        for (int w = tid; w < m_max * 8; w += nt) {        /* the (at most 6) remaining slots, all bands at once */
            const int m = w >> 3, d = i_lo + 34 + (w & 7);
            const int q = sm->patch_src[m];
            if (q == 255 || d >= i_hi + 2) continue;
            const float bw = sm->bw[sm->noise_idx[m]];
            const float a0 = sm->alpha[q][2] * bw * bw, a1 = sm->alpha[q][3] * bw * bw;
            const float a2 = sm->alpha[q][0] * bw,      a3 = sm->alpha[q][1] * bw;
            const float *xr = sm->xl_re[q] + d, *xi = sm->xl_im[q] + d;
            sm->xh_re[m][d] = xr[-2] * a0 - xi[-2] * a1 + xr[-1] * a2 - xi[-1] * a3 + xr[0];
            sm->xh_im[m][d] = xi[-2] * a0 + xr[-2] * a1 + xi[-1] * a2 + xr[-1] * a3 + xi[0];
        }
        for (int w = tid; w < nenv * 48; w += nt) {        /* s_mapped */
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            int fr = pc->bs_freq_res[e + 1];
            const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
            int b = sm->band_idx[fr][m], present = 0;
            for (int mm = tab[b]; mm < tab[b + 1]; mm++)
                if (sm->s_idx[e + 1][mm - kx]) { present = 1; break; }
            sm->s_map[e][m] = (uint8_t)present;
        }
        HEAAC_SYNC();
        /* ---- sbr_env_estimate (lanes = bands, serial over the envelope's slots) ---- */
        for (int w = tid; w < nenv * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            int ilb = pc->t_env[e] * 2 + 2, iub = pc->t_env[e + 1] * 2 + 2;
            if (p->bs_interpol_freq) {
                float recip = heaac_fdiv(0.5f, (float)(pc->t_env[e + 1] - pc->t_env[e]));
                float sum = 0.0f;
                const float *xr = sm->xh_re[m], *xi = sm->xh_im[m];
                for (int i = ilb; i < iub; i++)
                    sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(xr[i], xr[i]), heaac_fmul(xi[i], xi[i])));
                sm->e_curr[e][m] = heaac_fmul(sum, recip);
            } else {
                int fr = pc->bs_freq_res[e + 1];
                if (!sm->band_first[fr][m]) continue;      /* the first sub-band of a group does the whole group */
                const uint8_t *tab = fr ? p->f_tablehigh : p->f_tablelow;
                int b = sm->band_idx[fr][m];
                int env_size = 2 * (pc->t_env[e + 1] - pc->t_env[e]);
                int den = env_size * (tab[b + 1] - tab[b]);
                float sum = 0.0f;
                for (int k = tab[b]; k < tab[b + 1]; k++)
                    for (int i = ilb; i < iub; i++)
                        sum = heaac_fadd(sum, heaac_fadd(heaac_fmul(sm->xh_re[k - kx][i], sm->xh_re[k - kx][i]),
                                                         heaac_fmul(sm->xh_im[k - kx][i], sm->xh_im[k - kx][i])));
                sum = heaac_fdiv(sum, (float)den);
                for (int k = tab[b]; k < tab[b + 1]; k++) sm->e_curr[e][k - kx] = sum;
            }
        }
        HEAAC_SYNC();
        /* ---- sbr_gain_calc, aacsbr.c:1552-1605, as element-wise passes + per-limiter-band reductions ---- */
        for (int w = tid; w < nenv * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            int delta = !((e == pc->e_a[1]) || (e == pc->e_a[0]));
            float eo = sm->e_orig[e][m], qm = sm->q_map[e][m], ec = sm->e_curr[e][m];
            float temp = heaac_fdiv(eo, heaac_fadd(1.0f, qm));
            sm->qall[e][m] = heaac_fsqrt(heaac_fmul(temp, qm));
            sm->s_m[e][m] = heaac_fsqrt(heaac_fmul(temp, (float)sm->s_idx[e + 1][m]));
            if (!sm->s_map[e][m])
                sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(eo, heaac_fmul(heaac_fadd(1.0f, ec),
                                             heaac_fadd(1.0f, heaac_fmul(qm, (float)delta)))));
            else
                sm->gall[e][m] = heaac_fsqrt(heaac_fdiv(heaac_fmul(eo, qm),
                                             heaac_fmul(heaac_fadd(1.0f, ec), heaac_fadd(1.0f, qm))));
        }
        for (int w = tid; w < nenv * p->n_lim; w += nt) {
            int e = w / p->n_lim, k = w - e * p->n_lim;
            int lo = p->f_tablelim[k] - kx, hi = p->f_tablelim[k + 1] - kx;
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int m = lo; m < hi; m++) { sum0 = heaac_fadd(sum0, sm->e_orig[e][m]); sum1 = heaac_fadd(sum1, sm->e_curr[e][m]); }
            float gain_max = heaac_fmul(heaac_limgain[p->bs_limiter_gains & 3],
                                        heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sum0),
                                                               heaac_fadd(1.1920928955078125e-07f, sum1))));
            sm->gmax[e][k] = HEAAC_MIN(100000.0f, gain_max);
        }
        HEAAC_SYNC();
        for (int w = tid; w < nenv * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            float gain_max = sm->gmax[e][sm->lim_idx[m]];
            float q_m_max = heaac_fdiv(heaac_fmul(sm->qall[e][m], gain_max), sm->gall[e][m]);
            sm->qall[e][m]  = HEAAC_MIN(sm->qall[e][m], q_m_max);
            sm->gall[e][m] = HEAAC_MIN(sm->gall[e][m], gain_max);
        }
        HEAAC_SYNC();
        for (int w = tid; w < nenv * p->n_lim; w += nt) {
            int e = w / p->n_lim, k = w - e * p->n_lim;
            int delta = !((e == pc->e_a[1]) || (e == pc->e_a[0]));
            int lo = p->f_tablelim[k] - kx, hi = p->f_tablelim[k + 1] - kx;
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int m = lo; m < hi; m++) {
                float g = sm->gall[e][m], s = sm->s_m[e][m], q = sm->qall[e][m];
                sum0 = heaac_fadd(sum0, sm->e_orig[e][m]);
                float t3 = heaac_fmul(heaac_fmul((float)(delta && !s), q), q);
                float acc = heaac_fadd(heaac_fadd(heaac_fmul(heaac_fmul(sm->e_curr[e][m], g), g), heaac_fmul(s, s)), t3);
                sum1 = heaac_fadd(sum1, acc);
            }
            float gain_boost = heaac_fsqrt(heaac_fdiv(heaac_fadd(1.1920928955078125e-07f, sum0),
                                                      heaac_fadd(1.1920928955078125e-07f, sum1)));
            /* FFMIN(1.584893192 (double), gain_boost) then narrowed to float */
            sm->gboost[e][k] = (1.584893192 > (double)gain_boost) ? gain_boost : (float)1.584893192;
        }
        HEAAC_SYNC();
        for (int w = tid; w < nenv * 48; w += nt) {
            int e = w / 48, m = w - e * 48;
            if (m >= m_max) continue;
            float gain_boost = sm->gboost[e][sm->lim_idx[m]];
            const float gf = heaac_fmul(sm->gall[e][m], gain_boost), qf = heaac_fmul(sm->qall[e][m], gain_boost);
            sm->gall[e][m] = gf;
            sm->qall[e][m] = qf;
            sm->s_m[e][m]  = heaac_fmul(sm->s_m[e][m], gain_boost);
        }
        HEAAC_SYNC();
    }

    /* ---- sbr_hf_assemble fused with sbr_x_gen: X[plane][band][slot], slot fastest ---- */
    {
        HeaacYCtx y;
        y.h_SL = 4 * !p->bs_smoothing_mode;
        y.i_lo = i_lo; y.i_hi = i_hi;
        y.t_old2 = 2 * pc->t_env_num_env_old;
        y.noise0 = p->reset ? 0u : st->f_indexnoise;
        y.sine0 = st->f_indexsine;
        y.kx = kx; y.m_max = m_max; y.reset = p->reset;
        y.e_a0 = pc->e_a[0]; y.e_a1 = pc->e_a[1];
        const int i_Temp = HEAAC_MAX(y.t_old2 - 32, 0);
        float *Xre = X_out, *Xim = X_out + HEAAC_X_PLANE;
        for (int k = warp; k < 64; k += nwarps) {          /* rows = QMF bands, lanes = slots 0..31 */
            const int low_cur = k < kx, low_old = k < kx0;
            const int hf_cur = p->start && k >= kx && k < kx + m_max, hf_old = k >= kx0 && k < kx0 + m0;
            const int m = k - kx;
            for (int i = lane; i < 32; i += ls) {
                float re = 0.0f, im = 0.0f;
                /* sbr_x_gen, aacsbr.c:1412-1446 */
                if (i < i_Temp) {
                    if (low_old) { re = sm->xl_re[k][i + 2]; im = sm->xl_im[k][i + 2]; }
                    else if (hf_old) { re = st->y_tail[k * 8 + i]; im = st->y_tail[512 + k * 8 + i]; }
                } else {
                    if (low_cur) { re = sm->xl_re[k][i + 2]; im = sm->xl_im[k][i + 2]; }
                    else if (hf_cur && i >= i_lo && i < i_hi) heaac_sbr_y(sm, &y, T, i, m, &re, &im);
                    /* (slots of Y[1] the envelope grid does not cover would hold stale data of an older
                     *  frame in the reference; a conformant grid always covers [i_Temp, 32)) */
                }
                Xre[k * HEAAC_XS + i] = re; Xim[k * HEAAC_XS + i] = im;
            }
        }
        for (int w = tid; w < 64 * 8; w += nt) {           /* slots 32..37: low band for PS look-ahead; Y tail for the next frame */
            const int k = w >> 3, j = w & 7, i = 32 + j;
            if (j >= 6) continue;
            float re = 0.0f, im = 0.0f, yre = 0.0f, yim = 0.0f;
            const int in_hf = p->start && k >= kx && k < kx + m_max && i < i_hi;
            if (in_hf) heaac_sbr_y(sm, &y, T, i, k - kx, &yre, &yim);
            if (k < kx) { re = sm->xl_re[k][i + 2]; im = sm->xl_im[k][i + 2]; }
            Xre[k * HEAAC_XS + i] = re; Xim[k * HEAAC_XS + i] = im;
            if (k >= kx && k < kx + 48) {
                sm->ynew_re[k - kx][j] = yre; sm->ynew_im[k - kx][j] = yim;
                sm->ynew_valid[k - kx][j] = (uint8_t)in_hf;
            }
        }
        HEAAC_SYNC();      /* every read of the old y_tail / g_hist / q_hist is done */
        for (int w = tid; w < 48 * 8; w += nt) {
            const int m = w >> 3, j = w & 7, k = kx + m;
            if (j < 6 && k < 64 && sm->ynew_valid[m][j]) {
                st->y_tail[k * 8 + j] = sm->ynew_re[m][j]; st->y_tail[512 + k * 8 + j] = sm->ynew_im[m][j];
            }
        }
        if (p->start) {
            /* g_temp / q_temp rows 32..41 for the next frame, aacsbr.c:1642-1647 */
            for (int w = tid; w < 10 * 48; w += nt) {
                int r = 32 + w / 48, m = w % 48;
                int src = r - y.h_SL;
                if (m < m_max && src >= i_lo && src < i_hi) {
                    int es = sm->env_of_slot[src];
                    st->g_hist[w] = sm->gall[es][m];
                    st->q_hist[w] = sm->qall[es][m];
                }
            }
            for (int m = tid; m < 48; m += nt) st->s_index0[m] = sm->s_idx[nenv][m];
            if (tid == 0) {
                st->f_indexnoise = (y.noise0 + (uint32_t)(i_hi - i_lo) * (uint32_t)m_max) & 0x1ff;
                st->f_indexsine = (y.sine0 + (uint32_t)(i_hi - i_lo)) & 3;
            }
        }
        HEAAC_SYNC();
    }
}

#endif /* HEAAC_SBR_CUH */
