/*
 * heaac_sbr_tables_core.h -- derivation of the SBR frequency band tables from header parameters
 * (ISO/IEC 14496-3 4.6.18.3), written once for two targets: the host front end (heaac_sbr_tables.c,
 * heaac_frontend.c) and the device front end (heaac_parse_gpu.cu, FE_DEVICE defined).
 *
 * Follows the behaviour of the reference's sbr_reset() chain so that a front end built on it
 * hands the synthesis kernels exactly the tables the reference would:
 *   sbr_make_f_master    libavcodec/aacsbr.c:303-488
 *   sbr_make_f_derived   libavcodec/aacsbr.c:541-593
 *   sbr_hf_calc_npatches libavcodec/aacsbr.c:490-539
 *   sbr_make_f_tablelim  libavcodec/aacsbr.c:159-205
 * Serial integer/float table math that runs only when a header changes.  Same libm calls as the reference
 * (powf, log2f, lrintf); the device build is checked against the host build over the whole header space
 * (tests/test_gpu_parser.py).
 */
#ifndef HEAAC_SBR_TABLES_CORE_H
#define HEAAC_SBR_TABLES_CORE_H
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/heaac_b200.h"

#ifndef FE_FN
#ifdef FE_DEVICE
#define FE_FN static __device__
#define FE_TAB static __device__ const
#else
#define FE_FN static
#define FE_TAB static const
#endif
#endif

/* Table 4.A.? "start frequency offsets" per sampling-rate class (aacsbrdata.h:260-267) */
FE_TAB int8_t fe_start_offset[6][16] = {
    { -8, -7, -6, -5, -4, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7 },   /* 16000 */
    { -5, -4, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13 },   /* 22050 */
    { -5, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16 },   /* 24000 */
    { -6, -4, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16 },   /* 32000 */
    { -4, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16, 20 },   /* 44100..64000 */
    { -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16, 20, 24 },   /* > 64000 */
};

/* ascending sort of int16 values (the reference calls the C library sort with an int16 comparator) */
FE_FN void fe_sort_i16(void *base, int n)
{
    int16_t *a = (int16_t *)base;
    for (int i = 1; i < n; i++) {
        int16_t v = a[i];
        int j = i - 1;
        for (; j >= 0 && a[j] > v; j--) a[j + 1] = a[j];
        a[j + 1] = v;
    }
}

FE_FN void make_bands(int16_t *bands, int start, int stop, int num_bands)     /* aacsbr.c:269-285 */
{
    float base = powf((float)stop / start, 1.0f / num_bands);
    float prod = start;
    int previous = start;
    for (int k = 0; k < num_bands - 1; k++) {
        prod *= base;
        int present = lrintf(prod);
        bands[k] = present - previous;
        previous = present;
    }
    bands[num_bands - 1] = stop - previous;
}

FE_FN int in_table(const int16_t *t, int last, int16_t v)
{
    for (int i = 0; i <= last; i++) if (t[i] == v) return 1;
    return 0;
}

typedef struct Work {
    int k0, k1, k2, n_master;
    uint16_t f_master[49];
} Work;

FE_FN int f_master(Work *w, int sbr_rate, int start_freq, int stop_freq, int freq_scale, int alter_scale, int xover)
{
    unsigned temp = sbr_rate < 32000 ? 3000 : sbr_rate < 64000 ? 4000 : 5000;
    unsigned start_min = ((temp << 7) + (sbr_rate >> 1)) / sbr_rate;
    unsigned stop_min  = ((temp << 8) + (sbr_rate >> 1)) / sbr_rate;
    int row;
    switch (sbr_rate) {
    case 16000: row = 0; break;
    case 22050: row = 1; break;
    case 24000: row = 2; break;
    case 32000: row = 3; break;
    case 44100: case 48000: case 64000: row = 4; break;
    case 88200: case 96000: case 128000: case 176400: case 192000: row = 5; break;
    default: return -1;
    }
    w->k0 = start_min + fe_start_offset[row][start_freq];
    if (stop_freq < 14) {
        int16_t dk[13];
        w->k2 = stop_min;
        make_bands(dk, stop_min, 64, 13);
        fe_sort_i16(dk, 13);
        for (int k = 0; k < stop_freq; k++) w->k2 += dk[k];
    } else if (stop_freq == 14) w->k2 = 2 * w->k0;
    else if (stop_freq == 15) w->k2 = 3 * w->k0;
    else return -1;
    if (w->k2 > 64) w->k2 = 64;
    unsigned max_bands = sbr_rate <= 32000 ? 48 : sbr_rate == 44100 ? 35 : 32;
    if (w->k2 - w->k0 > (int)max_bands) return -1;

    if (!freq_scale) {
        int dk = alter_scale + 1;
        w->n_master = ((w->k2 - w->k0 + (dk & 2)) >> dk) << 1;
        if (w->n_master <= 0 || xover >= w->n_master) return -1;
        for (int k = 1; k <= w->n_master; k++) w->f_master[k] = dk;
        int k2diff = w->k2 - w->k0 - w->n_master * dk;
        if (k2diff < 0) {
            w->f_master[1]--;
            w->f_master[2] -= (k2diff < -1);
        } else if (k2diff) {
            w->f_master[w->n_master]++;
        }
        w->f_master[0] = w->k0;
        for (int k = 1; k <= w->n_master; k++) w->f_master[k] += w->f_master[k - 1];
    } else {
        int half_bands = 7 - freq_scale;
        int two_regions, num_bands_0, vdk0_max;
        int16_t vk0[49];
        if (49 * w->k2 > 110 * w->k0) { two_regions = 1; w->k1 = 2 * w->k0; }
        else                          { two_regions = 0; w->k1 = w->k2; }
        num_bands_0 = lrintf(half_bands * log2f(w->k1 / (float)w->k0)) * 2;
        if (num_bands_0 <= 0) return -1;
        vk0[0] = 0;
        make_bands(vk0 + 1, w->k0, w->k1, num_bands_0);
        fe_sort_i16(vk0 + 1, num_bands_0);
        vdk0_max = vk0[num_bands_0];
        vk0[0] = w->k0;
        for (int k = 1; k <= num_bands_0; k++) {
            if (vk0[k] <= 0) return -1;
            vk0[k] += vk0[k - 1];
        }
        if (two_regions) {
            int16_t vk1[49];
            float invwarp = alter_scale ? 0.76923076923076923077f : 1.0f;
            int num_bands_1 = lrintf(half_bands * invwarp * log2f(w->k2 / (float)w->k1)) * 2;
            make_bands(vk1 + 1, w->k1, w->k2, num_bands_1);
            int vdk1_min = vk1[1];
            for (int i = 2; i <= num_bands_1; i++) if (vk1[i] < vdk1_min) vdk1_min = vk1[i];
            if (vdk1_min < vdk0_max) {
                fe_sort_i16(vk1 + 1, num_bands_1);
                int a = vdk0_max - vk1[1], b = (vk1[num_bands_1] - vk1[1]) >> 1;
                int change = a > b ? b : a;
                vk1[1] += change;
                vk1[num_bands_1] -= change;
            }
            fe_sort_i16(vk1 + 1, num_bands_1);
            vk1[0] = w->k1;
            for (int k = 1; k <= num_bands_1; k++) {
                if (vk1[k] <= 0) return -1;
                vk1[k] += vk1[k - 1];
            }
            w->n_master = num_bands_0 + num_bands_1;
            if (w->n_master <= 0 || xover >= w->n_master || w->n_master > 48) return -1;
            for (int k = 0; k <= num_bands_0; k++) w->f_master[k] = vk0[k];
            for (int k = 1; k <= num_bands_1; k++) w->f_master[num_bands_0 + k] = vk1[k];
        } else {
            w->n_master = num_bands_0;
            if (w->n_master <= 0 || xover >= w->n_master || w->n_master > 48) return -1;
            for (int k = 0; k <= num_bands_0; k++) w->f_master[k] = vk0[k];
        }
    }
    return 0;
}

FE_FN void f_tablelim(HeaacSbrElem *o, int limiter_bands)                        /* aacsbr.c:159-205 */
{
    uint16_t lim[29];
    memset(lim, 0, sizeof(lim));
    if (limiter_bands > 0) {
        const float lim_bands_per_octave_warped = limiter_bands == 1 ? 1.32715174233856803909f :
                                                  limiter_bands == 2 ? 1.18509277094158210129f : 1.11987160404675912501f;
        int16_t patch_borders[7];
        uint16_t *in = lim + 1, *out = lim;
        patch_borders[0] = o->kx[1];
        for (int k = 1; k <= o->num_patches; k++)
            patch_borders[k] = patch_borders[k - 1] + o->patch_num_subbands[k - 1];
        for (int k = 0; k <= o->n[0]; k++) lim[k] = o->f_tablelow[k];
        for (int k = 1; k < o->num_patches; k++) lim[o->n[0] + k] = patch_borders[k];
        fe_sort_i16(lim, o->num_patches + o->n[0]);
        int n_lim = o->n[0] + o->num_patches - 1;
        while (out < lim + n_lim) {
            if (*in >= *out * lim_bands_per_octave_warped) {
                *++out = *in++;
            } else if (*in == *out || !in_table(patch_borders, o->num_patches, *in)) {
                in++; n_lim--;
            } else if (!in_table(patch_borders, o->num_patches, *out)) {
                *out = *in++; n_lim--;
            } else {
                *++out = *in++;
            }
        }
        o->n_lim = n_lim;
    } else {
        lim[0] = o->f_tablelow[0];
        lim[1] = o->f_tablelow[o->n[0]];
        o->n_lim = 1;
    }
    for (int k = 0; k < 29; k++) o->f_tablelim[k] = (uint8_t)lim[k];
}

/*
 * Fill the table part of *out (k0, kx[1], m[1], n[], n_q, n_lim, f_table*, patches) for the given
 * SBR header.  core_sample_rate is the AAC-LC rate; SBR runs at twice that (aacsbr.c:1054-1055).
 * kx[0] / m[0] (previous frame's values) and all per-frame fields are left untouched.
 * Returns 0, or HEAAC_E_INVAL when the reference would reject the header ("pure upsampling").
 */
FE_FN int fe_sbr_tables(int core_sample_rate, int bs_start_freq, int bs_stop_freq, int bs_xover_band,
                          int bs_freq_scale, int bs_alter_scale, int bs_noise_bands, int bs_limiter_bands,
                          HeaacSbrElem *out)
{
    if (!out || bs_start_freq < 0 || bs_start_freq > 15 || bs_stop_freq < 0 || bs_stop_freq > 15 ||
        bs_xover_band < 0 || bs_xover_band > 7 || bs_freq_scale < 0 || bs_freq_scale > 3 ||
        bs_noise_bands < 0 || bs_noise_bands > 3 || bs_limiter_bands < 0 || bs_limiter_bands > 3)
        return HEAAC_E_INVAL;
    const int sbr_rate = 2 * core_sample_rate;
    Work w;
    memset(&w, 0, sizeof(w));
    if (f_master(&w, sbr_rate, bs_start_freq, bs_stop_freq, bs_freq_scale, bs_alter_scale, bs_xover_band) < 0)
        return HEAAC_E_INVAL;

    /* derived tables, aacsbr.c:541-593 */
    int n_high = w.n_master - bs_xover_band, n_low = (n_high + 1) >> 1;
    uint16_t f_high[49], f_low[25], f_noise[6];
    memset(f_high, 0, sizeof(f_high)); memset(f_low, 0, sizeof(f_low)); memset(f_noise, 0, sizeof(f_noise));
    for (int k = 0; k <= n_high; k++) f_high[k] = w.f_master[bs_xover_band + k];
    int m = f_high[n_high] - f_high[0], kx = f_high[0];
    if (kx + m > 64 || kx > 32) return HEAAC_E_INVAL;
    f_low[0] = f_high[0];
    int odd = n_high & 1;
    for (int k = 1; k <= n_low; k++) f_low[k] = f_high[2 * k - odd];
    long nq = lrintf(bs_noise_bands * log2f(w.k2 / (float)kx));
    int n_q = nq > 1 ? (int)nq : 1;
    if (n_q > 5) return HEAAC_E_INVAL;
    f_noise[0] = f_low[0];
    int temp = 0;
    for (int k = 1; k <= n_q; k++) {
        temp += (n_low - temp) / (n_q + 1 - k);
        f_noise[k] = f_low[temp];
    }
    out->k0 = w.k0; out->kx[1] = kx; out->m[1] = m;
    out->n[0] = n_low; out->n[1] = n_high; out->n_q = n_q;
    for (int k = 0; k < 25; k++) out->f_tablelow[k] = (uint8_t)f_low[k];
    for (int k = 0; k < 49; k++) out->f_tablehigh[k] = (uint8_t)f_high[k];
    for (int k = 0; k < 6; k++)  out->f_tablenoise[k] = (uint8_t)f_noise[k];

    /* patch construction, aacsbr.c:490-539 */
    {
        int i, k, sb = 0, msb = w.k0, usb = kx;
        int goal_sb = ((1000 << 11) + (sbr_rate >> 1)) / sbr_rate;
        int num_patches = 0;
        uint8_t pns[6] = { 0 }, pss[6] = { 0 };
        if (goal_sb < kx + m) {
            for (k = 0; w.f_master[k] < goal_sb; k++) ;
        } else
            k = w.n_master;
        do {
            int odd_ = 0;
            for (i = k; i == k || sb > (w.k0 - 1 + msb - odd_); i--) {
                sb = w.f_master[i];
                odd_ = (sb + w.k0) & 1;
            }
            if (num_patches > 5) return HEAAC_E_INVAL;
            pns[num_patches] = sb - usb > 0 ? sb - usb : 0;
            pss[num_patches] = w.k0 - odd_ - pns[num_patches];
            if (pns[num_patches] > 0) {
                usb = sb; msb = sb; num_patches++;
            } else
                msb = kx;
            if (w.f_master[k] - sb < 3) k = w.n_master;
        } while (sb != kx + m);
        if (pns[num_patches - 1] < 3 && num_patches > 1) num_patches--;
        out->num_patches = num_patches;
        memcpy(out->patch_num_subbands, pns, 6);
        memcpy(out->patch_start_subband, pss, 6);
    }
    f_tablelim(out, bs_limiter_bands);
    return 0;
}

#endif
