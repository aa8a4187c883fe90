/* Fuzz harness for the host front end: corrupts the golden packets (byte overwrites, truncation, bit flips) and parses
 * them with both the one-pass and the two-phase parser; every buffer is an exact-size heap copy so AddressSanitizer
 * sees any over-read.  Packets come from cases.bin, written by tests/test_frontend.py::dump_fuzz_cases (or see the
 * python snippet in that test).  Build and run:
 *   gcc -O1 -g -fsanitize=address,undefined -Iinclude -o /tmp/fz tools/fuzz_frontend.c \
 *       ffmpeg-heaac_b200/csrc/heaac_frontend.c ffmpeg-heaac_b200/csrc/heaac_sbr_tables.c -lm -lpthread
 *   ASAN_OPTIONS=detect_leaks=0 /tmp/fz 400 /tmp/cases.bin
 * Round 1: 22,400 corrupted packets, no finding.  Round 2 (two-step SBR payload, chunk routine added to the harness): see DESIGN.md. */
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
            if (it % 3 == 0 && np <= 32) {                            /* the same kind of corruption through the chunk routine of the device front end */
                if (heaac_fe_open(&fe, ed, h[0], h[1], h[2], 0)) { printf("open fail\n"); return 1; }
                uint8_t ety[HEAAC_MAX_ELEMS]; int ncc = 0, nse = 0;
                heaac_fe_info(fe, &ncc, &nse, NULL, NULL, NULL, ety);
                if (ncc > 0) {
                    unsigned char *cp[32]; int cs[32], cons[32];
                    for (int i = 0; i < np; i++) {
                        int n = sz[i]; int mode = rnd() % 5;
                        if (mode == 1 && i) n = 1 + rnd() % n;
                        cp[i] = malloc(n); memcpy(cp[i], pk[i], n); cs[i] = n;
                        if (mode == 0 && i) { int k = 1 + rnd() % 6; while (k--) cp[i][rnd() % n] = rnd(); }
                        else if (mode == 2 && i) cp[i][rnd() % n] ^= 1 << (rnd() % 8);
                    }
                    float *cco = malloc((size_t)np * ncc * 1024 * 4);
                    HeaacIcs *cic = malloc(sizeof(HeaacIcs) * np * ncc); HeaacTns *ctn = malloc(sizeof(HeaacTns) * np * ncc);
                    HeaacSbrElem *csb = malloc(sizeof(HeaacSbrElem) * np * (nse > 0 ? nse : 1)); HeaacPsFrame *cps = malloc(sizeof(HeaacPsFrame) * np);
                    int r = heaac_fe_parse_chunk(fe, np, (const uint8_t *const *)cp, cs, cco, cic, ctn, csb, cps, cons);
                    if (r) { printf("chunk call failed %d\n", r); return 1; }
                    for (int i = 0; i < np; i++) { if (cons[i] < 0) err++; else ok++; free(cp[i]); }
                    free(cco); free(cic); free(ctn); free(csb); free(cps);
                }
                heaac_fe_close(fe);
            }
        }
        for (int i = 0; i < np; i++) free(pk[i]);
        free(co);
        printf("case %d ok %ld err %ld\n", c, ok, err); fflush(stdout);
    }
    return 0;
}
