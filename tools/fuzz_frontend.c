/* Fuzz harness for the host front end: corrupts the golden packets (byte overwrites, truncation, bit flips) and parses
 * them with both the one-pass and the two-phase parser; every buffer is an exact-size heap copy so AddressSanitizer
 * sees any over-read.  Packets come from cases.bin, written by tests/test_frontend.py::dump_fuzz_cases (or see the
 * python snippet in that test).  Build and run:
 *   gcc -O1 -g -fsanitize=address,undefined -Iinclude -o /tmp/fz tools/fuzz_frontend.c \
 *       ffmpeg-heaac_b200/csrc/heaac_frontend.c ffmpeg-heaac_b200/csrc/heaac_sbr_tables.c -lm -lpthread
 *   ASAN_OPTIONS=detect_leaks=0 /tmp/fz 400 /tmp/cases.bin
 * Round 1: 22,400 corrupted packets, no finding. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "heaac_b200.h"
static unsigned rs = 12345;
static unsigned rnd(void) { rs = rs * 1664525u + 1013904223u; return rs >> 8; }
int main(int argc, char **argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 200;
    FILE *f = fopen(argc > 2 ? argv[2] : "/tmp/cases.bin", "rb"); int nc;
    if (!f || fread(&nc, 4, 1, f) != 1) { printf("no cases file\n"); return 1; }
    long ok = 0, err = 0;
    for (int c = 0; c < nc; c++) {
        int h[4]; fread(h, 4, 4, f);
        unsigned char ed[64]; fread(ed, 1, h[0], f);
        int np = h[3]; unsigned char *pk[64]; int sz[64];
        for (int i = 0; i < np; i++) { fread(&sz[i], 4, 1, f); pk[i] = malloc(sz[i]); fread(pk[i], 1, sz[i], f); }
        float *co = malloc(16 * 1024 * 4); HeaacIcs ic[16]; HeaacTns tn[16]; HeaacSbrElem sb[8]; HeaacPsFrame ps;
        for (int it = 0; it < iters; it++) {
            HeaacFrontend *fe; if (heaac_fe_open(&fe, ed, h[0], h[1], h[2], 0)) { printf("open fail\n"); return 1; }
            for (int i = 0; i < np; i++) {
                int n = sz[i]; unsigned char *b = malloc(n);          /* exact-size heap copy: ASan sees any overread */
                memcpy(b, pk[i], n);
                int mode = rnd() % 4;
                if (mode == 0) { int k = 1 + rnd() % 6; while (k--) b[rnd() % n] = rnd(); }
                else if (mode == 1) { n = 1 + rnd() % n; }
                else if (mode == 2) { b[rnd() % n] ^= 1 << (rnd() % 8); }
                unsigned char *bb = malloc(n); memcpy(bb, b, n);
                int r = (it & 1) ? heaac_fe_parse_two_phase(fe, bb, n, co, ic, tn, sb, &ps) : heaac_fe_parse(fe, bb, n, co, ic, tn, sb, &ps);
                if (r < 0) err++; else ok++;
                free(bb); free(b);
            }
            heaac_fe_close(fe);
        }
        for (int i = 0; i < np; i++) free(pk[i]);
        free(co);
        printf("case %d ok %ld err %ld\n", c, ok, err); fflush(stdout);
    }
    return 0;
}
