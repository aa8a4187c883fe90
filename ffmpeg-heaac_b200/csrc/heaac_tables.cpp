/*
 * heaac_tables.cpp -- host-side construction of HeaacDevTables (copied to HBM once per context).
 *
 * Tables that have a closed form are computed here with the same formulas and precision as the
 * reference's init code so that the device sees bit-identical constants:
 *   sine windows     libavcodec/mdct_tablegen.h:49-53   sinf((i + 0.5) * pi / 2n)
 *   KBD windows      libavcodec/mdct.c:35-55            alpha = 4 (1024), 6 (128); aacdec.c:593-594
 *   MDCT twiddles    libavcodec/mdct.c:94-100           -cos/-sin(2 pi (i + 1/8 + [scale<0] N/4)/N) sqrt|scale|
 *   PS tables        libavcodec/aacps_tablegen.h:87-209
 * Normative data tables (QMF prototype, noise table, scalefactor-band offsets) come from the
 * generated heaac_tables_data.h.  tests/test_tables.py compares every entry with the reference.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "heaac_dev.cuh"
#include "heaac_tables_data.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static void sine_window(float *w, int n)
{
    for (int i = 0; i < n; i++)
        w[i] = sinf((i + 0.5) * (M_PI / (2.0 * n)));
}

static void kbd_window(float *w, float alpha, int n)
{
    double *cum = (double *)malloc(sizeof(double) * n);
    double a2 = (alpha * M_PI / n) * (alpha * M_PI / n), sum = 0.0;
    for (int i = 0; i < n; i++) {
        double t = i * (n - i) * a2, bessel = 1.0;
        for (int j = 50; j > 0; j--)
            bessel = bessel * t / (j * j) + 1;
        sum += bessel;
        cum[i] = sum;
    }
    sum++;
    for (int i = 0; i < n; i++)
        w[i] = sqrt(cum[i] / sum);
    free(cum);
}

static void mdct_twiddles(float *tcos, float *tsin, int n, double scale)
{
    int n4 = n >> 2;
    double theta = 1.0 / 8.0 + (scale < 0 ? n4 : 0);
    double sc = sqrt(fabs(scale));
    for (int i = 0; i < n4; i++) {
        double alpha = 2 * M_PI * (i + theta) / n;
        tcos[i] = -cos(alpha) * sc;
        tsin[i] = -sin(alpha) * sc;
    }
}

static void ps_proto_filters(float (*filter)[7][2], const float *proto, int bands)
{
    for (int q = 0; q < bands; q++)
        for (int n = 0; n < 7; n++) {
            double theta = 2 * M_PI * (q + 0.5) * (n - 6) / bands;
            filter[q][n][0] = proto[n] *  cos(theta);
            filter[q][n][1] = proto[n] * -sin(theta);
        }
}

static void ps_tables(HeaacDevTables *t)
{
    static const float pd_sin[8] = { 0, (float)M_SQRT1_2, 1, (float)M_SQRT1_2, 0, -(float)M_SQRT1_2, -1, -(float)M_SQRT1_2 };
    static const float pd_cos[8] = { 1, (float)M_SQRT1_2, 0, -(float)M_SQRT1_2, -1, -(float)M_SQRT1_2, 0, (float)M_SQRT1_2 };
    /* ISO/IEC 14496-3 Tables 8.25/8.26 (IID dequantisation, default then fine) and 8.27 (ICC) */
    static const float iid_lin[46] = {
        0.05623413251903, 0.12589254117942, 0.19952623149689, 0.31622776601684, 0.44668359215096,
        0.63095734448019, 0.79432823472428, 1, 1.25892541179417, 1.58489319246111,
        2.23872113856834, 3.16227766016838, 5.01187233627272, 7.94328234724282, 17.7827941003892,
        0.00316227766017, 0.00562341325190, 0.01, 0.01778279410039, 0.03162277660168,
        0.05623413251903, 0.07943282347243, 0.11220184543020, 0.15848931924611, 0.22387211385683,
        0.31622776601684, 0.39810717055350, 0.50118723362727, 0.63095734448019, 0.79432823472428,
        1, 1.25892541179417, 1.58489319246111, 1.99526231496888, 2.51188643150958,
        3.16227766016838, 4.46683592150963, 6.30957344480193, 8.91250938133745, 12.5892541179417,
        17.7827941003892, 31.6227766016838, 56.2341325190349, 100, 177.827941003892, 316.227766016837 };
    static const float icc_q[8] = { 1, 0.937, 0.84118, 0.60092, 0.36764, 0, -0.589, -1 };
    static const float icc_acos[8] = { 0, 0.35685527, 0.57133466, 0.92614472, 1.1943263, (float)(M_PI / 2), 2.2006171, (float)M_PI };
    static const int8_t centre20[10] = { -3, -1, 1, 3, 5, 7, 10, 14, 18, 22 };
    static const int8_t centre34[32] = { 2, 6, 10, 14, 18, 22, 26, 30, 34, -10, -6, -2, 51, 57, 15, 21,
                                         27, 33, 39, 45, 54, 66, 78, 42, 102, 66, 78, 90, 102, 114, 126, 90 };
    static const float link_frac[3] = { 0.43f, 0.75f, 0.347f };
    const float gain_frac = 0.39f;
    static const float proto_q8[7]   = { 0.00746082949812f, 0.02270420949825f, 0.04546865930473f, 0.07266113929591f,
                                         0.09885108575264f, 0.11793710567217f, 0.125f };
    static const float proto_q12[7]  = { 0.04081179924692f, 0.03812810994926f, 0.05144908135699f, 0.06399831151592f,
                                         0.07428313801106f, 0.08100347892914f, 0.08333333333333f };
    static const float proto1_q8[7]  = { 0.01565675600122f, 0.03752716391991f, 0.05417891378782f, 0.08417044116767f,
                                         0.10307344158036f, 0.12222452249753f, 0.125f };
    static const float proto2_q4[7]  = { -0.05908211155639f, -0.04871498374946f, 0.0f, 0.07778723915851f,
                                         0.16486303567403f, 0.23279856662996f, 0.25f };
    static const float proto1_q2[7]  = { 0.0f, 0.01899487526049f, 0.0f, -0.07293139167538f, 0.0f, 0.30596630545168f, 0.5f };
    static const int8_t map20[71] = {
         1,  0,  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 14, 15,
        15, 15, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 18, 18,
        18, 18, 18, 18, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19,
        19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19 };
    static const int8_t map34[91] = {
         0,  1,  2,  3,  4,  5,  6,  6,  7,  2,  1,  0, 10, 10,  4,  5,  6,  7,  8,
         9, 10, 11, 12,  9, 14, 11, 12, 13, 14, 15, 16, 13, 16, 17, 18, 19, 20, 21,
        22, 22, 23, 23, 24, 24, 25, 25, 26, 26, 27, 27, 27, 28, 28, 28, 29, 29, 29,
        30, 30, 30, 31, 31, 31, 31, 32, 32, 32, 32, 33, 33, 33, 33, 33, 33, 33, 33,
        33, 33, 33, 33, 33, 33, 33, 33, 33, 33, 33, 33, 33, 33, 33 };

    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++)
            for (int c = 0; c < 8; c++) {
                float re = 0.25f * pd_cos[a] + 0.5f * pd_cos[b] + pd_cos[c];
                float im = 0.25f * pd_sin[a] + 0.5f * pd_sin[b] + pd_sin[c];
                float mag = 1 / sqrt((double)(im * im + re * re));   /* C semantics: double sqrt of a float expression */
                t->pd_re_smooth[a * 64 + b * 8 + c] = re * mag;
                t->pd_im_smooth[a * 64 + b * 8 + c] = im * mag;
            }
    for (int iid = 0; iid < 46; iid++) {
        float c = iid_lin[iid];
        float c1 = (float)M_SQRT2 / sqrtf(1.0f + c * c);
        float c2 = c * c1;
        for (int icc = 0; icc < 8; icc++) {
            /* mixing procedure Ra */
            float alpha = 0.5f * icc_acos[icc];
            float beta  = alpha * (c1 - c2) * (float)M_SQRT1_2;
            t->HA[iid][icc][0] = c2 * cosf(beta + alpha);
            t->HA[iid][icc][1] = c1 * cosf(beta - alpha);
            t->HA[iid][icc][2] = c2 * sinf(beta + alpha);
            t->HA[iid][icc][3] = c1 * sinf(beta - alpha);
            /* mixing procedure Rb */
            float rho = icc_q[icc] > 0.05f ? icc_q[icc] : 0.05f;
            float al = 0.5f * atan2f(2.0f * c * rho, c * c - 1.0f);
            float mu = c + 1.0f / c;
            mu = sqrtf(1 + (4 * rho * rho - 4) / (mu * mu));
            float gamma = atanf(sqrtf((1.0f - mu) / (1.0f + mu)));
            if (al < 0) al += M_PI / 2;
            float ac = cosf(al), as = sinf(al), gc = cosf(gamma), gs = sinf(gamma);
            t->HB[iid][icc][0] =  M_SQRT2 * ac * gc;
            t->HB[iid][icc][1] =  M_SQRT2 * as * gc;
            t->HB[iid][icc][2] = -M_SQRT2 * as * gs;
            t->HB[iid][icc][3] =  M_SQRT2 * ac * gs;
        }
    }
    for (int is34 = 0; is34 < 2; is34++) {
        int nb = is34 ? 50 : 30;
        for (int k = 0; k < nb; k++) {
            double fc;
            if (!is34) fc = k < 10 ? centre20[k] * 0.125 : k - 6.5f;
            else       fc = k < 32 ? centre34[k] / 24.  : k - 26.5f;
            for (int m = 0; m < 3; m++) {
                double theta = -M_PI * link_frac[m] * fc;
                t->Q_fract_allpass[is34][k][m][0] = cos(theta);
                t->Q_fract_allpass[is34][k][m][1] = sin(theta);
            }
            double theta = -M_PI * gain_frac * fc;
            t->phi_fract[is34][k][0] = cos(theta);
            t->phi_fract[is34][k][1] = sin(theta);
        }
    }
    ps_proto_filters(t->f20_0_8,  proto_q8,   8);
    ps_proto_filters(t->f34_0_12, proto_q12, 12);
    ps_proto_filters(t->f34_1_8,  proto1_q8,  8);
    ps_proto_filters(t->f34_2_4,  proto2_q4,  4);
    memcpy(t->g1_Q2, proto1_q2, sizeof(proto1_q2));
    memcpy(t->k_to_i_20, map20, sizeof(map20));
    memcpy(t->k_to_i_34, map34, sizeof(map34));
    for (int mode = 0; mode < 2; mode++) {
        const int8_t *map = mode ? t->k_to_i_34 : t->k_to_i_20;
        const int nb = mode ? 91 : 71, np = mode ? 34 : 20;
        int pos = 0;
        for (int i = 0; i < 36; i++) {
            t->ps_inv_off[mode][i] = (uint8_t)pos;
            if (i < np) for (int k = 0; k < nb; k++) if (map[k] == i) t->ps_inv_k[mode][pos++] = (uint8_t)k;
        }
    }
}

extern "C" void heaac_build_tables(HeaacDevTables *t)
{
    memset(t, 0, sizeof(*t));
    sine_window(t->sine_1024, 1024);
    sine_window(t->sine_128, 128);
    kbd_window(t->kbd_1024, 4.0f, 1024);
    kbd_window(t->kbd_128, 6.0f, 128);
    mdct_twiddles(t->tcos_2048, t->tsin_2048, 2048, 1.0);
    mdct_twiddles(t->tcos_256,  t->tsin_256,  256,  1.0);
    mdct_twiddles(t->tcos_syn,  t->tsin_syn,  128,  1.0 / 64);
    mdct_twiddles(t->tcos_ana,  t->tsin_ana,  128, -2.0);
    for (int j = 0; j < 256; j++) {
        t->fft_wre[j] = cos(2 * M_PI * j / 512);
        t->fft_wim[j] = sin(2 * M_PI * j / 512);
    }
    for (int i = 0; i < 512; i++) {
        int r = 0;
        for (int b = 0; b < 9; b++) r |= ((i >> b) & 1) << (8 - b);
        t->bitrev512[i] = (uint16_t)r;
    }
    for (int i = 0; i < 32; i++) {
        int r = 0;
        for (int b = 0; b < 5; b++) r |= ((i >> b) & 1) << (4 - b);
        t->bitrev32[i] = (uint8_t)r;
    }
    memcpy(t->qmf_us, heaac_qmf_window_us, sizeof(t->qmf_us));
    for (int n = 0; n < 320; n++) t->qmf_ds[n] = heaac_qmf_window_us[2 * n];
    memcpy(t->noise, heaac_sbr_noise_table, sizeof(t->noise));
    memcpy(t->swb_offset_1024, heaac_swb_offset_1024, sizeof(t->swb_offset_1024));
    memcpy(t->swb_offset_128, heaac_swb_offset_128, sizeof(t->swb_offset_128));
    memcpy(t->num_swb_1024, heaac_num_swb_1024, 13);
    memcpy(t->num_swb_128, heaac_num_swb_128, 13);
    memcpy(t->tns_max_bands_1024, heaac_tns_max_bands_1024, 13);
    memcpy(t->tns_max_bands_128, heaac_tns_max_bands_128, 13);
    ps_tables(t);
}
