/* Sweep of the whole SBR header space through fe_sbr_tables (csrc/heaac_sbr_tables_core.h, or the header named by -DHDR=...):
 * 16 start x 16 stop x 8 xover x 4 freq_scale x 2 alter_scale x 4 noise_bands x 4 limiter_bands = 262,144 headers... x 9 core
 * rates; prints per rate the number of accepted headers and a hash over every field the back end reads (table tails past
 * n / n_q / n_lim / num_patches excluded).  Two builds that print the same lines derive the same tables.
 *   gcc -O2 -o sweep tools/sbr_tables_sweep.c -lm && ./sweep
 *   gcc -O2 -DHDR='"/path/to/other/heaac_sbr_tables_core.h"' ... */
#include <stdio.h>
#include <stdint.h>
#ifndef HDR
#define HDR "../ffmpeg-heaac_b200/csrc/heaac_sbr_tables_core.h"
#endif
#include HDR

static uint64_t h;
static void mix(const void *p, int n) { const uint8_t *b = p; for (int i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } }

int main(void)
{
    static const int rates[9] = { 8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000 };
    for (int r = 0; r < 9; r++) {
        long ok = 0, total = 0;
        h = 1469598103934665603ull;
        for (int sf = 0; sf < 16; sf++) for (int st = 0; st < 16; st++) for (int xo = 0; xo < 8; xo++)
        for (int fs = 0; fs < 4; fs++) for (int al = 0; al < 2; al++) for (int nb = 0; nb < 4; nb++) for (int lb = 0; lb < 4; lb++) {
            HeaacSbrElem t;
            memset(&t, 0, sizeof(t));
            int rc = fe_sbr_tables(rates[r], sf, st, xo, fs, al, nb, lb, &t);
            total++;
            mix(&rc, sizeof(rc));
            if (rc) continue;
            ok++;
            mix(&t.k0, 1); mix(&t.kx[1], 1); mix(&t.m[1], 1); mix(t.n, 2); mix(&t.n_q, 1); mix(&t.n_lim, 1); mix(&t.num_patches, 1);
            mix(t.f_tablehigh, t.n[1] + 1); mix(t.f_tablelow, t.n[0] + 1); mix(t.f_tablenoise, t.n_q + 1); mix(t.f_tablelim, t.n_lim + 1);
            mix(t.patch_num_subbands, t.num_patches); mix(t.patch_start_subband, t.num_patches);
        }
        printf("core rate %5d: %ld of %ld headers accepted, hash %016llx\n", rates[r], ok, total, (unsigned long long)h);
    }
    return 0;
}
