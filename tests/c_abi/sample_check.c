/* A plain-C consumer of libfdsr that SAMPLES on the GPU (the C program only checked the no-GPU error path):
 *   fdsr_create -> fdsr_load_weights -> fdsr_set_schedule -> fdsr_sample (injected noise) -> compare with the reference's
 *   own output (tests/golden/unet64_default.npz, exported to raw files by tests/test_gpu_cabi.py).
 * Only C types cross the boundary; device memory comes from the CUDA runtime's C API.
 * usage: sample_check <dir>    (dir holds manifest.txt, weights.bin, betas.bin, cond.bin, noise.bin, sr_ref.bin) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "fdsr.h"

static void* slurp(const char* dir, const char* name, size_t* bytes) {
  char path[1024];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc((size_t)n);
  if (fread(p, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read %s\n", path); exit(2); }
  fclose(f);
  if (bytes) *bytes = (size_t)n;
  return p;
}

#define CHECK(call)                                                                         \
  do {                                                                                      \
    int rc_ = (call);                                                                       \
    if (rc_ != FDSR_OK) {                                                                   \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? fdsr_last_error(ctx) : fdsr_global_error()); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const char* dir = argv[1];
  fdsr_ctx* ctx = NULL;
  fdsr_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.in_channel = 6; cfg.out_channel = 3; cfg.inner_channel = 64; cfg.norm_groups = 32;
  cfg.n_levels = 4; cfg.channel_mults[0] = 1; cfg.channel_mults[1] = 2; cfg.channel_mults[2] = 4; cfg.channel_mults[3] = 4;
  cfg.res_blocks = 2; cfg.dtype = FDSR_DTYPE_FP16; cfg.model = FDSR_MODEL_FASTDIFFSR;
  CHECK(fdsr_create(&cfg, &ctx));

  /* state_dict: manifest lines "<name> <numel>", tensors back to back in weights.bin */
  size_t wbytes = 0, mbytes = 0;
  float* weights = (float*)slurp(dir, "weights.bin", &wbytes);
  char* manifest = (char*)slurp(dir, "manifest.txt", &mbytes);
  int n = 0;
  for (size_t i = 0; i < mbytes; ++i) n += manifest[i] == '\n';
  const char** names = (const char**)malloc(sizeof(char*) * (size_t)n);
  const float** ptrs = (const float**)malloc(sizeof(float*) * (size_t)n);
  int64_t* numels = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  size_t off = 0;
  char* line = manifest;
  for (int i = 0; i < n; ++i) {
    char* nl = memchr(line, '\n', (size_t)(manifest + mbytes - line));
    *nl = 0;
    char* sp = strchr(line, ' ');
    *sp = 0;
    names[i] = line;
    numels[i] = atoll(sp + 1);
    ptrs[i] = weights + off;
    off += (size_t)numels[i];
    line = nl + 1;
  }
  if (off * 4 != wbytes) { fprintf(stderr, "manifest / weights.bin mismatch\n"); return 2; }
  CHECK(fdsr_load_weights(ctx, names, ptrs, numels, n));
  size_t bbytes = 0;
  double* betas = (double*)slurp(dir, "betas.bin", &bbytes);
  const int T = (int)(bbytes / 8);
  CHECK(fdsr_set_schedule(ctx, betas, T));
  double c1[64];
  if (fdsr_get_table(ctx, "posterior_mean_coef1", c1, 64) != T || fabs(c1[0] - 1.0) > 1e-12) { fprintf(stderr, "table\n"); return 1; }

  const int B = 1, H = 64, W = 64;
  const size_t img = (size_t)B * 3 * H * W * 4;
  size_t cb = 0, nb = 0, rb = 0;
  float* cond = (float*)slurp(dir, "cond.bin", &cb);
  float* noise = (float*)slurp(dir, "noise.bin", &nb);
  float* ref = (float*)slurp(dir, "sr_ref.bin", &rb);
  if (cb != img || rb != img || nb != img * (size_t)T) { fprintf(stderr, "input sizes\n"); return 2; }
  float *d_cond, *d_noise, *d_sr;
  if (cudaMalloc((void**)&d_cond, img) || cudaMalloc((void**)&d_noise, nb) || cudaMalloc((void**)&d_sr, img)) return 3;
  cudaMemcpy(d_cond, cond, img, cudaMemcpyHostToDevice);
  cudaMemcpy(d_noise, noise, nb, cudaMemcpyHostToDevice);
  CHECK(fdsr_sample(ctx, d_cond, d_noise, 0, d_sr, NULL, B, H, W, NULL));
  CHECK(fdsr_check_overflow(ctx, NULL));
  float* sr = (float*)malloc(img);
  if (cudaMemcpy(sr, d_sr, img, cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
  double num = 0, den = 0;
  for (size_t i = 0; i < img / 4; ++i) {
    const double d = (double)sr[i] - ref[i];
    num += d * d;
    den += (double)ref[i] * ref[i];
  }
  const double rel = sqrt(num / den);
  /* a second call with the built-in generator: seeded, reproducible, different from the injected-noise result */
  float* sr2 = (float*)malloc(img);
  float* sr3 = (float*)malloc(img);
  CHECK(fdsr_sample(ctx, d_cond, NULL, 42, d_sr, NULL, B, H, W, NULL));
  cudaMemcpy(sr2, d_sr, img, cudaMemcpyDeviceToHost);
  CHECK(fdsr_sample(ctx, d_cond, NULL, 42, d_sr, NULL, B, H, W, NULL));
  cudaMemcpy(sr3, d_sr, img, cudaMemcpyDeviceToHost);
  const int repro = memcmp(sr2, sr3, img) == 0, differs = memcmp(sr2, sr, img) != 0;
  printf("plain-C fdsr_sample: rel-L2 vs the reference's output %.3e, launches %lld, seeded-reproducible %d, differs %d\n",
         rel, (long long)fdsr_launch_count(ctx), repro, differs);
  CHECK(fdsr_destroy(ctx));
  cudaFree(d_cond); cudaFree(d_noise); cudaFree(d_sr);
  return (rel <= 1e-2 && repro && differs) ? 0 : 1;
}
