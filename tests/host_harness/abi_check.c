/* Test harness (plain C, not part of libb200g16): include/b200g16.h must be valid C — it is what cgo
 * compiles — and the library must link, answer host-only calls, and REFUSE to start without a GPU
 * (no CPU fallback).  Built and run by tests/test_lib_cpu.py. */
#include <stdio.h>
#include <string.h>

#include "b200g16.h"

int main(void) {
  /* struct layouts cgo relies on */
  if (sizeof(b200g16_proof) != (8 + 16 + 8 + 8 + 8 + 8 + 8 + 16 + 8) * sizeof(uint64_t)) return 10;
  b200g16_pk_desc d;
  b200g16_vk_desc v;
  memset(&d, 0, sizeof d);
  memset(&v, 0, sizeof v);
  if (b200g16_version() < 120) return 11;
  /* host-only group helper: G + G = EIP-196 vector (Montgomery limbs from gnark-crypto: x = 1 -> "one") */
  uint64_t g[8] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full,
                   0xa6ba871b8b1e1b3aull, 0x14f1d651eb8e167bull, 0xccdd46def0f28c58ull, 0x1c14ef83340fbe5eull};
  uint64_t out[8];
  if (b200g16_g1_add(g, g, out) != 0) return 12;
  if (out[0] == g[0] && out[1] == g[1]) return 13; /* doubling must move the point */
  b200g16_ctx* ctx = NULL;
  int st = b200g16_init(0, &ctx);
  /* one-process multi-GPU surface: argument checks answer without a device, a group refuses to start without one */
  b200g16_group* grp = NULL;
  int devs[2] = {0, 0};
  if (b200g16_group_init(NULL, 1, &grp) != B200G16_ERR_ARG) return 16;
  if (b200g16_group_init(devs, 0, &grp) != B200G16_ERR_ARG) return 17;
  if (b200g16_group_size(NULL) != 0 || b200g16_group_ctx(NULL, 0) != NULL) return 18;
  if (b200g16_host_register(NULL, 0) != B200G16_ERR_ARG || b200g16_host_unregister(NULL) != B200G16_ERR_ARG) return 19;
  if (b200g16_group_msm_g1(NULL, NULL, NULL, 0, out) != B200G16_ERR_ARG) return 20;
  {
    int ticket = -1; /* asynchronous MSM: argument checks answer without a device */
    if (b200g16_msm_g1_begin(NULL, NULL, 0, NULL, 0, &ticket) != B200G16_ERR_ARG) return 24;
    if (b200g16_msm_g1_begin_dev(NULL, NULL, 0, NULL, 0, &ticket) != B200G16_ERR_ARG) return 25;
    if (b200g16_msm_g1_end(NULL, 0, out) != B200G16_ERR_ARG) return 26;
  }
  if (b200g16_group_pk_upload(NULL, &d, NULL) != B200G16_ERR_ARG) return 21;
  if (b200g16_group_prove(NULL, NULL, NULL, 0, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL) != B200G16_ERR_ARG) return 22;
  b200g16_group_bases_free(NULL);
  b200g16_group_pk_free(NULL);
  b200g16_group_destroy(NULL);
  if (st == 0) { /* a GPU is present: a two-shard group on device 0 comes up and goes away */
    b200g16_destroy(ctx);
    if (b200g16_group_init(devs, 2, &grp) != 0 || b200g16_group_size(grp) != 2 || !b200g16_group_ctx(grp, 1)) return 23;
    b200g16_group_destroy(grp);
    printf("gpu\n");
    return 0;
  }
  if (b200g16_group_init(devs, 2, &grp) == 0) return 24;
  if (st != B200G16_ERR_NO_DEVICE && st != B200G16_ERR_CUDA) return 14;
  if (strlen(b200g16_last_error()) == 0) return 15;
  printf("nogpu: %s\n", b200g16_last_error());
  return 0;
}
