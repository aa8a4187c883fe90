/*
 * heaac_sbr_tables.c -- host export of the SBR frequency table derivation (heaac_sbr_tables_core.h):
 * what the reference's sbr_reset() computes when a header changes (aacsbr.c:1023-1034).
 */
#include "heaac_sbr_tables_core.h"

int heaac_b200_sbr_tables(int core_sample_rate, int bs_start_freq, int bs_stop_freq, int bs_xover_band,
                          int bs_freq_scale, int bs_alter_scale, int bs_noise_bands, int bs_limiter_bands,
                          HeaacSbrElem *out)
{
    return fe_sbr_tables(core_sample_rate, bs_start_freq, bs_stop_freq, bs_xover_band, bs_freq_scale, bs_alter_scale,
                         bs_noise_bands, bs_limiter_bands, out);
}
