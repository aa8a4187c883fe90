/*
 * heaac_sbr_tables_core.h -- the SBR frequency band tables of a header (ISO/IEC 14496-3 4.6.18.3.2), for two targets: the
 * host front end (heaac_sbr_tables.c, heaac_frontend.c) and the device front end (heaac_parse_gpu.cu, FE_DEVICE defined).
 *
 * Organisation (own): every table is a list of band EDGES; the standard builds each one from a list of band WIDTHS, so the
 * code is a handful of list operations (geometric widths, sort, accumulate, pick, merge-and-prune) composed per table:
 *
 *   master    widths on a linear or a geometric scale between k0 and k2 (one or two octave regions), accumulated from k0
 *   high      the master edges from the cross-over band up;   low = every other high edge counted from the top
 *   noise     n_q edges picked from the low table at near-equal strides
 *   patches   the low-band stretches copied upwards: for each patch the highest master edge the source range can still reach
 *   limiter   low edges and inner patch borders merged, pruned where neighbours are closer than the header's ratio
 *
 * The results must be the tables the reference's sbr_reset() chain derives (aacsbr.c:159-205, 269-285, 303-593), bit for bit,
 * for every header it accepts, and the same headers must be rejected.  Where an edge is the rounding of a float (geometric
 * scales, band counts from log2) the reference's numerical recipe decides borderline cases -- single precision, the running
 * product carried from band to band, powf / log2f / lrintf -- so those expressions are kept as they are there.  Checked against
 * the reference on random headers (tests/test_sbr_tables.py), against the previous implementation over the whole header space
 * (2.36 M headers x 9 core rates, tools/sbr_tables_sweep.c) and device against host over the same space (tests/test_gpu_parser.py).
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

#define SBT_MAX 50
typedef struct SbtList { int n; int v[SBT_MAX]; } SbtList;      /* band widths, or band edges (n + 1 values then) */

/* ---- list operations ------------------------------------------------------------------------------------------- */

/* n widths between the edges lo and hi on a geometric scale.  Edge k is the rounding of lo r^k, r = (hi / lo)^(1/n), with the
 * product carried in single precision from one edge to the next (aacsbr.c:269-285: the recipe is part of the result) */
FE_FN void sbt_geometric(SbtList *w, int lo, int hi, int n)
{
    const float ratio = powf((float)hi / lo, 1.0f / n);
    float running = lo;
    int below = lo;
    w->n = n;
    for (int k = 0; k + 1 < n; k++) {
        running *= ratio;
        const int edge = (int)lrintf(running);
        w->v[k] = edge - below;
        below = edge;
    }
    if (n > 0) w->v[n - 1] = hi - below;
}

FE_FN void sbt_sort(int *v, int n)                    /* ascending; the lists have at most 49 entries */
{
    for (int i = 1; i < n; i++) {
        const int x = v[i];
        int j = i;
        while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; }
        v[j] = x;
    }
}

/* edges[0] = start, edges[k] = edges[k - 1] + w[k - 1]; appends to `edges` (sharing the start edge with what is there already
 * when it is not empty).  Returns 0 when a width is not positive. */
FE_FN int sbt_accumulate(SbtList *edges, int start, const SbtList *w)
{
    if (edges->n == 0) edges->v[edges->n++] = start;
    for (int k = 0; k < w->n; k++) {
        if (w->v[k] <= 0 || edges->n >= SBT_MAX) return 0;
        edges->v[edges->n] = edges->v[edges->n - 1] + w->v[k];
        edges->n++;
    }
    return 1;
}

FE_FN int sbt_contains(const int *v, int n, int x)
{
    for (int i = 0; i < n; i++) if (v[i] == x) return 1;
    return 0;
}

/* ---- the header's frequency range ------------------------------------------------------------------------------- */

/* offsets of the start band per bs_start_freq, one row per sampling-rate class (aacsbrdata.h:260-267, Table 4.A.? of the standard) */
FE_TAB int8_t sbt_start_offset[6][16] = {
    { -8, -7, -6, -5, -4, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7 },   /* 16000 */
    { -5, -4, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13 },   /* 22050 */
    { -5, -3, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16 },   /* 24000 */
    { -6, -4, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16 },   /* 32000 */
    { -4, -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16, 20 },   /* 44100..64000 */
    { -2, -1,  0,  1,  2,  3,  4,  5,  6,  7,  9, 11, 13, 16, 20, 24 },   /* > 64000 */
};

FE_FN int sbt_rate_class(int sbr_rate)
{
    static const int rates[12] = { 16000, 22050, 24000, 32000, 44100, 48000, 64000, 88200, 96000, 128000, 176400, 192000 };
    static const int8_t cls[12] = { 0, 1, 2, 3, 4, 4, 4, 5, 5, 5, 5, 5 };
    for (int i = 0; i < 12; i++) if (rates[i] == sbr_rate) return cls[i];
    return -1;
}

/* k0 and k2 (first and one-past-last QMF band of the master table).  Returns 0 when the header is not acceptable. */
FE_FN int sbt_range(int sbr_rate, int start_freq, int stop_freq, int *k0_o, int *k2_o)
{
    const int cls = sbt_rate_class(sbr_rate);
    if (cls < 0) return 0;
    /* the band at 3 / 4 / 5 kHz (start) and at twice that (stop), rounded: NINT(f * 128 / rate), NINT(f * 256 / rate) */
    const unsigned f = sbr_rate < 32000 ? 3000 : sbr_rate < 64000 ? 4000 : 5000;
    const int start_min = (int)(((f << 7) + (unsigned)(sbr_rate >> 1)) / (unsigned)sbr_rate);
    const int stop_min  = (int)(((f << 8) + (unsigned)(sbr_rate >> 1)) / (unsigned)sbr_rate);
    const int k0 = start_min + sbt_start_offset[cls][start_freq];
    int k2;
    if (stop_freq < 14) {                            /* stop_min plus the stop_freq narrowest of 13 geometric bands up to 64 */
        SbtList w;
        sbt_geometric(&w, stop_min, 64, 13);
        sbt_sort(w.v, 13);
        k2 = stop_min;
        for (int k = 0; k < stop_freq; k++) k2 += w.v[k];
    } else
        k2 = (stop_freq == 14 ? 2 : 3) * k0;
    if (k2 > 64) k2 = 64;
    const int widest = sbr_rate <= 32000 ? 48 : sbr_rate == 44100 ? 35 : 32;
    if (k2 - k0 > widest) return 0;
    *k0_o = k0; *k2_o = k2;
    return 1;
}

/* ---- master table ----------------------------------------------------------------------------------------------- */
FE_FN int sbt_master(SbtList *master, int k0, int k2, int freq_scale, int alter_scale, int xover)
{
    master->n = 0;
    if (!freq_scale) {
        /* linear scale: an even number of bands of width 1 or 2; what does not divide goes to the last band (one too few)
         * or comes off the first two (up to two too many) */
        const int width = alter_scale + 1, span = k2 - k0;
        const int n = width == 1 ? (span & ~1) : 2 * ((span + 2) >> 2);
        if (n <= 0 || xover >= n || n >= SBT_MAX) return 0;
        SbtList w;
        w.n = n;
        for (int k = 0; k < n; k++) w.v[k] = width;
        const int slack = span - n * width;
        if (slack > 0) w.v[n - 1]++;
        if (slack < 0) { w.v[0]--; if (slack < -1 && n > 1) w.v[1]--; }
        /* (the reference accumulates without looking at the widths here) */
        master->v[0] = k0; master->n = 1;
        for (int k = 0; k < n; k++) { master->v[master->n] = master->v[master->n - 1] + w.v[k]; master->n++; }
        return 1;
    }
    /* geometric scale, 12 / 10 / 8 bands per octave; above 2.2449 octaves.. a ratio k2 / k0 above 2.2449 splits the range
     * at 2 k0 and the upper region may use a coarser scale (alter_scale) */
    const int half = 7 - freq_scale;
    const int split = 49 * k2 > 110 * k0;
    const int k1 = split ? 2 * k0 : k2;
    SbtList lower, upper;
    const int n_lower = (int)lrintf(half * log2f(k1 / (float)k0)) * 2;
    if (n_lower <= 0 || n_lower >= SBT_MAX) return 0;
    sbt_geometric(&lower, k0, k1, n_lower);
    sbt_sort(lower.v, n_lower);
    if (!sbt_accumulate(master, k0, &lower)) return 0;
    if (split) {
        const float warp = alter_scale ? 0.76923076923076923077f : 1.0f;        /* 1 / 1.3 */
        const int n_upper = (int)lrintf(half * warp * log2f(k2 / (float)k1)) * 2;
        if (n_upper <= 0 || n_upper >= SBT_MAX) return 0;
        sbt_geometric(&upper, k1, k2, n_upper);
        sbt_sort(upper.v, n_upper);
        /* no band of the upper region may be narrower than the widest of the lower one: the narrowest takes what it lacks
         * from the widest (at most half their difference) */
        const int widest_lower = lower.v[n_lower - 1];
        if (upper.v[0] < widest_lower) {
            const int lacks = widest_lower - upper.v[0], spare = (upper.v[n_upper - 1] - upper.v[0]) >> 1;
            const int moved = lacks < spare ? lacks : spare;
            upper.v[0] += moved;
            upper.v[n_upper - 1] -= moved;
            sbt_sort(upper.v, n_upper);
        }
        if (!sbt_accumulate(master, k1, &upper)) return 0;
    }
    const int n = master->n - 1;
    if (n <= 0 || xover >= n || n > 48) return 0;
    return 1;
}

/* ---- patches (aacsbr.c:490-539): which low bands are copied where --------------------------------------------------- */
FE_FN int sbt_patches(HeaacSbrElem *o, const SbtList *master, int k0, int kx, int m, int sbr_rate)
{
    const int n_master = master->n - 1, top = kx + m;
    /* patches should not start above 1 kHz worth of bands unless the range ends below it */
    const int goal = ((1000 << 11) + (sbr_rate >> 1)) / sbr_rate;
    int from;                                        /* index of the master edge the search for a patch top starts at */
    if (goal < top) { from = 0; while (master->v[from] < goal) from++; }
    else from = n_master;
    int count = 0, covered = kx, reach = k0, sb;
    uint8_t width[6] = { 0 }, source[6] = { 0 };
    do {
        /* the highest master edge, from `from` downwards, that the source range (bands 1 .. k0 - 1, shifted by an even amount)
         * can still reach from `reach` */
        int i = from, parity;
        for (;;) {
            sb = master->v[i];
            parity = (sb + k0) & 1;
            if (sb <= k0 - 1 + reach - parity || i == 0) break;
            i--;
        }
        if (count > 5) return 0;
        const int w = sb - covered > 0 ? sb - covered : 0;
        width[count] = (uint8_t)w;
        source[count] = (uint8_t)(k0 - parity - w);
        if (w > 0) { covered = sb; reach = sb; count++; }
        else reach = kx;
        if (master->v[from] - sb < 3) from = n_master;
    } while (sb != top);
    if (count > 1 && width[count - 1] < 3) count--;      /* a last patch of fewer than 3 bands is dropped */
    o->num_patches = (uint8_t)count;
    memcpy(o->patch_num_subbands, width, 6);
    memcpy(o->patch_start_subband, source, 6);
    return 1;
}

/* ---- limiter table (aacsbr.c:159-205) ------------------------------------------------------------------------------ */
FE_FN void sbt_limiter(HeaacSbrElem *o, int limiter_bands)
{
    int kept[32], n_kept = 0;
    memset(kept, 0, sizeof(kept));
    if (limiter_bands <= 0) {                        /* one band over the whole range */
        kept[0] = o->f_tablelow[0];
        kept[1] = o->f_tablelow[o->n[0]];
        n_kept = 2;
    } else {
        /* 2^(1 / 1.2), 2^(1 / 2), 2^(1 / 3) "warped": neighbours closer than this ratio are merged */
        const float ratio = limiter_bands == 1 ? 1.32715174233856803909f :
                            limiter_bands == 2 ? 1.18509277094158210129f : 1.11987160404675912501f;
        int border[7], cand[32], n_cand = 0;
        border[0] = o->kx[1];
        for (int k = 1; k <= o->num_patches; k++) border[k] = border[k - 1] + o->patch_num_subbands[k - 1];
        for (int k = 0; k <= o->n[0]; k++) cand[n_cand++] = o->f_tablelow[k];
        for (int k = 1; k < o->num_patches; k++) cand[n_cand++] = border[k];
        sbt_sort(cand, n_cand);
        kept[n_kept++] = cand[0];
        for (int c = 1; c < n_cand; c++) {
            const int prev = kept[n_kept - 1], cur = cand[c];
            if (cur >= prev * ratio) kept[n_kept++] = cur;                                   /* far enough: a new edge */
            else if (cur == prev || !sbt_contains(border, o->num_patches + 1, cur)) continue;   /* too close, not a patch border: dropped */
            else if (!sbt_contains(border, o->num_patches + 1, prev)) kept[n_kept - 1] = cur;   /* a patch border wins over a plain edge */
            else kept[n_kept++] = cur;                                                       /* two patch borders both stay */
        }
    }
    o->n_lim = (uint8_t)(n_kept - 1);
    for (int k = 0; k < 29; k++) o->f_tablelim[k] = (uint8_t)kept[k];
}

/*
 * Fill the table part of *out (k0, kx[1], m[1], n[], n_q, n_lim, f_table*, patches) for the given
 * SBR header.  core_sample_rate is the AAC-LC rate; SBR runs at twice that (aacsbr.c:1054-1055).
 * kx[0] / m[0] (previous frame's values) and all per-frame fields are left untouched.
 * Returns 0, or HEAAC_E_INVAL when the reference would reject the header ("pure upsampling").
 */
#ifndef FE_BIG
#ifdef FE_DEVICE
#define FE_BIG static __device__ __noinline__
#else
#define FE_BIG static
#endif
#endif
FE_BIG int fe_sbr_tables(int core_sample_rate, int bs_start_freq, int bs_stop_freq, int bs_xover_band,
                          int bs_freq_scale, int bs_alter_scale, int bs_noise_bands, int bs_limiter_bands,
                          HeaacSbrElem *out)
{
    if (!out || bs_start_freq < 0 || bs_start_freq > 15 || bs_stop_freq < 0 || bs_stop_freq > 15 ||
        bs_xover_band < 0 || bs_xover_band > 7 || bs_freq_scale < 0 || bs_freq_scale > 3 ||
        bs_noise_bands < 0 || bs_noise_bands > 3 || bs_limiter_bands < 0 || bs_limiter_bands > 3)
        return HEAAC_E_INVAL;
    const int sbr_rate = 2 * core_sample_rate;
    int k0, k2;
    SbtList master;
    if (!sbt_range(sbr_rate, bs_start_freq, bs_stop_freq, &k0, &k2)) return HEAAC_E_INVAL;
    if (!sbt_master(&master, k0, k2, bs_freq_scale, bs_alter_scale, bs_xover_band)) return HEAAC_E_INVAL;

    /* high: the master edges from the cross-over band up; low: every other high edge, counted down from the top edge
     * (aacsbr.c:541-593) */
    const int n_high = master.n - 1 - bs_xover_band, n_low = (n_high + 1) >> 1;
    const int *high = master.v + bs_xover_band;
    const int kx = high[0], m = high[n_high] - kx;
    if (kx + m > 64 || kx > 32) return HEAAC_E_INVAL;
    int low[25];
    low[0] = kx;
    for (int k = n_low, h = n_high; k >= 1; k--, h -= 2) low[k] = high[h];
    /* noise floor bands: about bs_noise_bands per octave, their edges spread evenly over the low table */
    const long per_range = lrintf(bs_noise_bands * log2f(k2 / (float)kx));
    const int n_q = per_range > 1 ? (int)per_range : 1;
    if (n_q > 5) return HEAAC_E_INVAL;

    memset(out->f_tablelow, 0, sizeof(out->f_tablelow));
    memset(out->f_tablehigh, 0, sizeof(out->f_tablehigh));
    memset(out->f_tablenoise, 0, sizeof(out->f_tablenoise));
    out->k0 = (uint8_t)k0; out->kx[1] = (uint8_t)kx; out->m[1] = (uint8_t)m;
    out->n[0] = (uint8_t)n_low; out->n[1] = (uint8_t)n_high; out->n_q = (uint8_t)n_q;
    for (int k = 0; k <= n_low; k++)  out->f_tablelow[k] = (uint8_t)low[k];
    for (int k = 0; k <= n_high; k++) out->f_tablehigh[k] = (uint8_t)high[k];
    out->f_tablenoise[0] = (uint8_t)kx;
    for (int k = 1, at = 0; k <= n_q; k++) {
        at += (n_low - at) / (n_q + 1 - k);
        out->f_tablenoise[k] = (uint8_t)low[at];
    }
    if (!sbt_patches(out, &master, k0, kx, m, sbr_rate)) return HEAAC_E_INVAL;
    sbt_limiter(out, bs_limiter_bands);
    return 0;
}

#endif
