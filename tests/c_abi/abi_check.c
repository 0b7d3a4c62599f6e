/* Plain-C consumer of include/fdsr.h: proves the boundary is a C ABI (no C++ / torch types) and that the library
 * fails loudly — never falls back — when no sm_100 GPU is visible.  Built and run by tests/test_host.py with gcc.
 * Exit code: 0 = behaved as specified on this machine (with or without a GPU), non-zero = contract violated. */
#include <stdio.h>
#include <string.h>
#include "fdsr.h"

int main(void) {
  fdsr_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.in_channel = 6;
  cfg.out_channel = 3;
  cfg.inner_channel = 64;
  cfg.norm_groups = 32;
  cfg.n_levels = 4;
  cfg.channel_mults[0] = 1;
  cfg.channel_mults[1] = 2;
  cfg.channel_mults[2] = 4;
  cfg.channel_mults[3] = 4;
  cfg.res_blocks = 2;
  cfg.dtype = FDSR_DTYPE_FP16;
  cfg.model = FDSR_MODEL_FASTDIFFSR;
  fdsr_ctx* ctx = NULL;
  int rc = fdsr_create(&cfg, &ctx);
  if (rc == FDSR_OK) {
    /* a B200 is present: the call order contract must hold */
    float dummy = 0.f;
    int rc2 = fdsr_unet_forward(ctx, &dummy, &dummy, 0, &dummy, 1, 64, 64, NULL);
    printf("create ok; forward before weights -> %d (%s)\n", rc2, fdsr_last_error(ctx));
    if (rc2 != FDSR_E_STATE) return 2;
    if (fdsr_launch_count(ctx) != 0) return 3;
    fdsr_destroy(ctx);
    return 0;
  }
  printf("create -> %d (%s)\n", rc, fdsr_global_error());
  if (rc != FDSR_E_CUDA || ctx != NULL) return 4;           /* no GPU: a CUDA error, not a fallback */
  if (strlen(fdsr_global_error()) == 0) return 5;
  cfg.dtype = 99;                                            /* argument errors are reported the same way */
  if (fdsr_create(NULL, &ctx) != FDSR_E_INVALID) return 6;
  if (fdsr_trace_frames(NULL) != 0 || fdsr_workspace_bytes(NULL) != 0) return 7;
  return 0;
}
