/* TEST INFRASTRUCTURE ONLY.  Force-included (-include) when oracle/Makefile compiles the
 * reference's aacdec.c: declares the two hooks spliced around spectral_to_sample(ac):
 *   if (!heaac_ref_pre_hook(ac)) spectral_to_sample(ac); heaac_ref_post_hook(ac);
 * (aacdec.c:2078).  Defined in ref_harness.c. */
#ifndef ORACLE_REF_HOOKS_H
#define ORACLE_REF_HOOKS_H
struct AACContext_tag;
int  heaac_ref_pre_hook(void *ac);     /* != 0: the hook produced the frame's output itself (drop-in mode), skip spectral_to_sample */
void heaac_ref_post_hook(void *ac);
#endif
