/*
 * ref_tables.c -- TEST INFRASTRUCTURE ONLY.
 * Exposes the reference's static data tables by name so that (a) tools/gen_tables.py can
 * emit the normative ISO/IEC 14496-3 tables (QMF prototype, SBR noise table, Huffman
 * codebooks, scalefactor-band offsets) into the product's generated data header, and
 * (b) tests can check every table the product COMPUTES (KBD/sine windows, MDCT twiddles,
 * PS filters and mixing matrices) against the reference's initialised copy.
 * Compiled only into oracle/_ref/libheaac_ref.so; includes the reference headers in place.
 */
#include <string.h>
#include <stdint.h>
#include <stdlib.h>
#include "libavcodec/avcodec.h"
#include "libavcodec/get_bits.h"
#include "libavcodec/aac.h"
#include "libavcodec/aactab.h"
#include "libavcodec/aacdectab.h"
#include "libavcodec/aacsbrdata.h"
#include "libavcodec/aacps_tablegen.h"
#include "libavcodec/aacpsdata.c"
extern float ff_aac_pow2sf_tab[428]; void ff_aac_tableinit(void);
#include "libavcodec/cbrt_tablegen.h"

static int inited;
static float qmf_us[640], qmf_ds[320];
static uint16_t swb1024[13][64], swb128[13][32];
static float tns_map[4][16];

static void init_once(void)
{
    if (inited) return;
    inited = 1;
    /* aacsbr.c:117-123 */
    memcpy(qmf_us, sbr_qmf_window_us, sizeof(qmf_us));
    for (int n = 1; n < 320; n++) qmf_us[320 + n] = qmf_us[320 - n];
    qmf_us[384] = -qmf_us[384];
    qmf_us[512] = -qmf_us[512];
    for (int n = 0; n < 320; n++) qmf_ds[n] = qmf_us[2 * n];
    ps_tableinit();
    ff_aac_tableinit();
    cbrt_tableinit();
    for (int i = 0; i < 13; i++) {
        for (int k = 0; k <= ff_aac_num_swb_1024[i]; k++) swb1024[i][k] = ff_swb_offset_1024[i][k];
        for (int k = 0; k <= ff_aac_num_swb_128[i];  k++) swb128[i][k]  = ff_swb_offset_128[i][k];
    }
    static const int tsz[4] = { 8, 16, 4, 8 };   /* tns_tmp2_map[2*compress + res] */
    for (int i = 0; i < 4; i++) memcpy(tns_map[i], tns_tmp2_map[i], tsz[i] * sizeof(float));
}

#define T(nm, obj) if (!strcmp(name, nm)) { *p = (const void *)(obj); return (int)sizeof(obj); }
int ref_table(const char *name, const void **p)
{
    init_once();
    T("sbr_qmf_window_us", qmf_us)
    T("sbr_qmf_window_ds", qmf_ds)
    T("sbr_noise_table", sbr_noise_table)
    T("sbr_offset", sbr_offset)
    T("swb_offset_1024", swb1024)
    T("swb_offset_128", swb128)
    T("num_swb_1024", *(const uint8_t (*)[13])ff_aac_num_swb_1024)
    T("num_swb_128", *(const uint8_t (*)[13])ff_aac_num_swb_128)
    T("tns_max_bands_1024", ff_tns_max_bands_1024)
    T("tns_max_bands_128", ff_tns_max_bands_128)
    T("tns_tmp2_map", tns_map)
    T("scalefactor_code", ff_aac_scalefactor_code)
    T("scalefactor_bits", ff_aac_scalefactor_bits)
    T("spectral_sizes", ff_aac_spectral_sizes)
    T("pow2sf_tab", ff_aac_pow2sf_tab)
    T("cbrt_tab", cbrt_tab)
    /* SBR huffman */
    T("t_huffman_env_1_5dB_bits", t_huffman_env_1_5dB_bits) T("t_huffman_env_1_5dB_codes", t_huffman_env_1_5dB_codes)
    T("f_huffman_env_1_5dB_bits", f_huffman_env_1_5dB_bits) T("f_huffman_env_1_5dB_codes", f_huffman_env_1_5dB_codes)
    T("t_huffman_env_bal_1_5dB_bits", t_huffman_env_bal_1_5dB_bits) T("t_huffman_env_bal_1_5dB_codes", t_huffman_env_bal_1_5dB_codes)
    T("f_huffman_env_bal_1_5dB_bits", f_huffman_env_bal_1_5dB_bits) T("f_huffman_env_bal_1_5dB_codes", f_huffman_env_bal_1_5dB_codes)
    T("t_huffman_env_3_0dB_bits", t_huffman_env_3_0dB_bits) T("t_huffman_env_3_0dB_codes", t_huffman_env_3_0dB_codes)
    T("f_huffman_env_3_0dB_bits", f_huffman_env_3_0dB_bits) T("f_huffman_env_3_0dB_codes", f_huffman_env_3_0dB_codes)
    T("t_huffman_env_bal_3_0dB_bits", t_huffman_env_bal_3_0dB_bits) T("t_huffman_env_bal_3_0dB_codes", t_huffman_env_bal_3_0dB_codes)
    T("f_huffman_env_bal_3_0dB_bits", f_huffman_env_bal_3_0dB_bits) T("f_huffman_env_bal_3_0dB_codes", f_huffman_env_bal_3_0dB_codes)
    T("t_huffman_noise_3_0dB_bits", t_huffman_noise_3_0dB_bits) T("t_huffman_noise_3_0dB_codes", t_huffman_noise_3_0dB_codes)
    T("t_huffman_noise_bal_3_0dB_bits", t_huffman_noise_bal_3_0dB_bits) T("t_huffman_noise_bal_3_0dB_codes", t_huffman_noise_bal_3_0dB_codes)
    /* PS huffman */
    T("huff_iid_df1_bits", huff_iid_df1_bits) T("huff_iid_df1_codes", huff_iid_df1_codes)
    T("huff_iid_dt1_bits", huff_iid_dt1_bits) T("huff_iid_dt1_codes", huff_iid_dt1_codes)
    T("huff_iid_df0_bits", huff_iid_df0_bits) T("huff_iid_df0_codes", huff_iid_df0_codes)
    T("huff_iid_dt0_bits", huff_iid_dt0_bits) T("huff_iid_dt0_codes", huff_iid_dt0_codes)
    T("huff_icc_df_bits", huff_icc_df_bits) T("huff_icc_df_codes", huff_icc_df_codes)
    T("huff_icc_dt_bits", huff_icc_dt_bits) T("huff_icc_dt_codes", huff_icc_dt_codes)
    T("huff_ipd_df_bits", huff_ipd_df_bits) T("huff_ipd_df_codes", huff_ipd_df_codes)
    T("huff_ipd_dt_bits", huff_ipd_dt_bits) T("huff_ipd_dt_codes", huff_ipd_dt_codes)
    T("huff_opd_df_bits", huff_opd_df_bits) T("huff_opd_df_codes", huff_opd_df_codes)
    T("huff_opd_dt_bits", huff_opd_dt_bits) T("huff_opd_dt_codes", huff_opd_dt_codes)
    T("ps_huff_offset", huff_offset)
    T("k_to_i_20", k_to_i_20) T("k_to_i_34", k_to_i_34) T("g1_Q2", g1_Q2)
    /* PS computed tables (aacps_tablegen.h:87-209) */
    T("pd_re_smooth", pd_re_smooth) T("pd_im_smooth", pd_im_smooth)
    T("HA", HA) T("HB", HB)
    T("f20_0_8", f20_0_8) T("f34_0_12", f34_0_12) T("f34_1_8", f34_1_8) T("f34_2_4", f34_2_4)
    T("Q_fract_allpass", Q_fract_allpass) T("phi_fract", phi_fract)
    if (!strncmp(name, "spectral_codes_", 15) || !strncmp(name, "spectral_bits_", 14) ||
        !strncmp(name, "codebook_vectors_", 17)) {
        int which = name[0] == 'c' ? 2 : name[9] == 'c' ? 0 : 1;
        int idx = atoi(strrchr(name, '_') + 1);
        if (idx < 0 || idx > 10) return -1;
        static const int dim[11] = { 4, 4, 4, 4, 2, 2, 2, 2, 2, 2, 2 };
        if (which == 0) { *p = ff_aac_spectral_codes[idx]; return ff_aac_spectral_sizes[idx] * 2; }
        if (which == 1) { *p = ff_aac_spectral_bits[idx];  return ff_aac_spectral_sizes[idx]; }
        *p = ff_aac_codebook_vectors[idx]; return ff_aac_spectral_sizes[idx] * dim[idx] * 4;
    }
    return -1;
}
