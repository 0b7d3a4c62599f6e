// fp32 parity mode (FDSR_DTYPE_FP32): the same fused layer plan as conv_gemm_kernel — GroupNorm + Swish on
// the input, virtual concat, nearest-upsample / stride-2 gathers, 1x1 residual chunks, bias + FiLM, identity
// residual, GroupNorm pair statistics of the output — evaluated with fp32 activations, fp32 weights and fp32
// FMA accumulation on the CUDA cores.  It exists to show that the plan itself (everything except 16-bit
// rounding) reproduces the reference to <= 1e-4 relative L2 (BASELINE north_star, "fp32 mode"); it is a
// correctness mode (~60x slower than the tensor-core path), never the benchmarked one.
//
// Reference ops covered: the same as conv_kernel.cuh (unet.py:66-120, 257-258, 297, 317-321).
#pragma once
#include "conv_desc.h"

namespace fdsr {

constexpr int kF32Tile = 8;     // 8x8 output pixels per block
constexpr int kF32NBlk = 64;    // output channels per block
constexpr int kF32KBlk = 16;    // input channels staged per step
constexpr int kF32Patch = kF32Tile + 2;

struct F32Chunk {
  const float* src;   // NHWC fp32 source tensor
  int32_t C, H, W;    // source channels / spatial size
  int32_t c0, nch;    // channels [c0, c0 + nch) of the source feed this chunk (nch <= 64)
  int32_t gn, vc0;    // GroupNorm (+ Swish when gn = 1) with table entries [vc0, vc0 + nch)
  int32_t pa, pb;     // space-to-depth parity plane (kModeS2D)
  int32_t ntaps;
  int8_t dy[kMaxTaps], dx[kMaxTaps];  // tap offsets in the conv's input space (block space for kModeS2D)
  int32_t w_off;      // float offset of this chunk's weights: [tap][ci < nch][N]
};

struct F32Layer {
  F32Chunk chunk[kMaxChunks];
  int32_t nchunks, mode;
  int32_t B, H, W, N;          // output size, N = padded output channels (multiple of 64, or 16)
  int32_t gn_C;
  const float2* gn_tab;        // [B][gn_C] (scale, shift) of this launch's GroupNorm
  const float* weights;
  const float* bias;           // [T][N]
  int32_t bias_tstride;
  const float* resid;          // identity residual NHWC fp32 [B][H][W][N] or null
  float* out;                  // NHWC fp32 [B][H][W][N] | NCHW fp32 [B][out_c][H][W]
  unsigned long long* out_stats;
  double out_sq_scale;         // fixed-point scale of the output's sums of squares (stat_sq_scale)
  int32_t out_mode, out_c;
};

// (scale, shift) per (sample, virtual channel) from the fixed-point pair statistics of up to two sources
struct F32GnArgs {
  const unsigned long long* stats[2];
  int32_t C[2];
  int32_t gn_C, groups, HW;
  double sq_scale;             // fixed-point scale of the sources' sums of squares
  float eps;
  const float* gamma;
  const float* beta;
  float2* tab;
};
__global__ void f32_gn_table_kernel(F32GnArgs a) {
  __shared__ float2 gs[64];
  const int b = blockIdx.x;
  const int cpg = a.gn_C / a.groups;
  if (int(threadIdx.x) < a.groups) {
    const int ppg = cpg >> 1, p0 = a.C[0] >> 1;
    long long Si = 0, Qi = 0;
    for (int pv = threadIdx.x * ppg; pv < int(threadIdx.x + 1) * ppg; ++pv) {
      const int si = pv < p0 ? 0 : 1;
      const int pl = pv < p0 ? pv : pv - p0;
      const long long* st = reinterpret_cast<const long long*>(a.stats[si]) + (size_t(b) * (a.C[si] >> 1) + pl) * 2;
      Si += st[0];
      Qi += st[1];
    }
    const double n = double(cpg) * a.HW;
    const double mean = double(Si) * (1.0 / 16777216.0) / n;
    double var = double(Qi) / a.sq_scale / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    gs[threadIdx.x] = make_float2(float(mean), float(1.0 / sqrt(var + double(a.eps))));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.gn_C; c += blockDim.x) {
    const float2 g = gs[c / cpg];
    const float sc = a.gamma[c] * g.y;
    a.tab[size_t(b) * a.gn_C + c] = make_float2(sc, a.beta[c] - g.x * sc);
  }
}

__global__ void __launch_bounds__(256) conv_f32_kernel(const __grid_constant__ F32Layer L, int t_step) {
  __shared__ float Ps[kF32KBlk][kF32Patch * kF32Patch + 4];
  __shared__ __align__(16) float Ws[kMaxTaps][kF32KBlk][kF32NBlk];
  const int tid = threadIdx.x;
  const int tiles_x = (L.W + kF32Tile - 1) / kF32Tile, tiles_y = (L.H + kF32Tile - 1) / kF32Tile;
  int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  tile -= b * tiles_x * tiles_y;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * kF32Tile, x0 = tx * kF32Tile;
  const int n0 = blockIdx.y * kF32NBlk;
  const int nblk = L.N - n0 < kF32NBlk ? L.N - n0 : kF32NBlk;  // 64, or 16 for the final conv
  // thread -> 4 pixels (one row, 4 consecutive columns) x 4 output channels
  const int pg = tid & 15, cgp = tid >> 4;
  const int pr = pg >> 1, pc0 = (pg & 1) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c = 0; c < L.nchunks; ++c) {
    const F32Chunk& ck = L.chunk[c];
    for (int ci0 = 0; ci0 < ck.nch; ci0 += kF32KBlk) {
      const int kb = ck.nch - ci0 < kF32KBlk ? ck.nch - ci0 : kF32KBlk;
      __syncthreads();
      // ---- input patch (conv-input space), GroupNorm + Swish applied, zero outside the image
      for (int e = tid; e < kF32KBlk * kF32Patch * kF32Patch; e += 256) {
        const int ci = e % kF32KBlk, pos = e / kF32KBlk;
        const int py = pos / kF32Patch, px = pos - py * kF32Patch;
        const int y = y0 + py - 1, x = x0 + px - 1;
        float v = 0.f;
        if (ci < kb && y >= 0 && y < L.H && x >= 0 && x < L.W) {
          int sy = y, sx = x;
          if (L.mode == kModeUp2x) {
            sy >>= 1;
            sx >>= 1;
          } else if (L.mode == kModeS2D) {
            sy = 2 * y + ck.pa;
            sx = 2 * x + ck.pb;
          }
          v = ck.src[((size_t(b) * ck.H + sy) * ck.W + sx) * ck.C + ck.c0 + ci0 + ci];
          if (ck.gn) {
            const float2 g = L.gn_tab[size_t(b) * L.gn_C + ck.vc0 + ci0 + ci];
            const float u = fmaf(v, g.x, g.y);
            v = ck.gn == 2 ? u : u / (1.0f + expf(-u));  // gn = 2: GroupNorm without activation (attention norm)
          }
        }
        Ps[ci][pos] = v;
      }
      // ---- weights of these input channels: [tap][ci][n]
      const float* w = L.weights + ck.w_off;
      for (int e = tid; e < ck.ntaps * kF32KBlk * kF32NBlk; e += 256) {
        const int n = e % kF32NBlk, r = e / kF32NBlk;
        const int ci = r % kF32KBlk, tp = r / kF32KBlk;
        Ws[tp][ci][n] = (ci < kb && n < nblk) ? w[(size_t(tp) * ck.nch + ci0 + ci) * L.N + n0 + n] : 0.f;
      }
      __syncthreads();
      for (int tp = 0; tp < ck.ntaps; ++tp) {
        const int base = (pr + 1 + ck.dy[tp]) * kF32Patch + pc0 + 1 + ck.dx[tp];
#pragma unroll 4
        for (int ci = 0; ci < kF32KBlk; ++ci) {
          const float4 wv = *reinterpret_cast<const float4*>(&Ws[tp][ci][cgp * 4]);
          float a[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = Ps[ci][base + i];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(a[i], wv.x, acc[i][0]);
            acc[i][1] = fmaf(a[i], wv.y, acc[i][1]);
            acc[i][2] = fmaf(a[i], wv.z, acc[i][2]);
            acc[i][3] = fmaf(a[i], wv.w, acc[i][3]);
          }
        }
      }
    }
  }

  // ---- epilogue
  const int n = n0 + cgp * 4;
  const bool nvalid = cgp * 4 < nblk;
  const float* bias = L.bias + size_t(t_step) * L.bias_tstride;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const int y = y0 + pr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + pc0 + i;
    if (!nvalid || y >= L.H || x >= L.W) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[n + j];
    const size_t pix = (size_t(b) * L.H + y) * L.W + x;
    if (L.out_mode == kOutAct) {
      if (L.resid != nullptr) {
        const float4 r = *reinterpret_cast<const float4*>(L.resid + pix * L.N + n);
        v[0] += r.x;
        v[1] += r.y;
        v[2] += r.z;
        v[3] += r.w;
      }
      *reinterpret_cast<float4*>(L.out + pix * L.N + n) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += v[j];
        q[j] = fmaf(v[j], v[j], q[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < L.out_c) L.out[((size_t(b) * L.out_c + n + j) * L.H + y) * L.W + x] = v[j];
    }
  }
  if (L.out_mode == kOutAct && L.out_stats != nullptr) {
    // the 16 pixel groups of one channel group are 16 consecutive lanes
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 8; o; o >>= 1) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
      }
    if (pg == 0 && nvalid) {
      unsigned long long* st = L.out_stats + size_t(b) * L.N + n;  // pair entries: (sum, sum of squares)
      atomicAdd(st + 0, static_cast<unsigned long long>(__double2ll_rn((double(s[0]) + double(s[1])) * 16777216.0)));
      atomicAdd(st + 1, static_cast<unsigned long long>(__double2ll_rn((double(q[0]) + double(q[1])) * L.out_sq_scale)));
      atomicAdd(st + 2, static_cast<unsigned long long>(__double2ll_rn((double(s[2]) + double(s[3])) * 16777216.0)));
      atomicAdd(st + 3, static_cast<unsigned long long>(__double2ll_rn((double(q[2]) + double(q[3])) * L.out_sq_scale)));
    }
  }
}

}  // namespace fdsr
