// SelfAttention core of the SR3 baseline (which_model_G = "ddpm"):
//   o[b, i, :] = sum_j softmax_j(q[b, i, :] . k[b, j, :] / sqrt(C)) * v[b, j, :]
// over the H*W tokens of one image, n_head = 1 (model/ddpm_modules/unet.py:100-131; note the reference
// divides by sqrt(channel)).  q, k, v come from three fused conv launches (GroupNorm without activation
// in the prologue, 1x1 weights = row blocks of attn.qkv.weight); the output projection + residual is a
// fourth conv launch.  Tokens are NHWC rows, so q/k/v/o are plain [B][HW][C] matrices.
//
//   attn_core_kernel<T, C>     16-bit modes: tcgen05 tensor cores, accumulators in TMEM, operands by TMA
//   attn_core_ref_kernel<T>    fp32 parity mode (and FDSR_ATTN_REF=1): one warp per query on the CUDA cores
#pragma once
#include "aux_kernels.cuh"

namespace fdsr {

// ------------------------------------------------------------------------------------------
// CUDA-core version: one warp per query, online softmax in fp32.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_core_ref_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                            const T* __restrict__ v, T* __restrict__ o, int HW, int C,
                                                            float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + warp;
  if (i >= HW) return;
  const int np = C >> 6;  // channel pairs per lane (C = 64, 128, 256 -> 1, 2, 4)
  const size_t base = size_t(b) * HW * C;
  const int c0 = lane * 2 * np;
  float2 qv[4], acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    acc[e] = make_float2(0.f, 0.f);
    qv[e] = e < np ? ld_pair(q + base + size_t(i) * C + c0 + 2 * e) : make_float2(0.f, 0.f);
  }
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < HW; ++j) {
    const T* kp = k + base + size_t(j) * C + c0;
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < np) {
        const float2 kv = ld_pair(kp + 2 * e);
        d = fmaf(qv[e].x, kv.x, d);
        d = fmaf(qv[e].y, kv.y, d);
      }
#pragma unroll
    for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    const float s = d * scale;
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), p = expf(s - mn);
    l = fmaf(l, corr, p);
    const T* vp = v + base + size_t(j) * C + c0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < np) {
        const float2 vv = ld_pair(vp + 2 * e);
        acc[e].x = fmaf(acc[e].x, corr, p * vv.x);
        acc[e].y = fmaf(acc[e].y, corr, p * vv.y);
      }
    m = mn;
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (e < np) st_pair(o + base + size_t(i) * C + c0 + 2 * e, acc[e].x * inv, acc[e].y * inv);
}

// ------------------------------------------------------------------------------------------
// Tensor-core version.  One CTA = 128 queries of one image (thread r owns query row r = TMEM lane r).
//   S = Q K^T      tcgen05.mma M=128, N=64 keys, K = C: Q and K tiles are K-major 128B-swizzled rows of
//                  64 channels, exactly what a SWIZZLE_128B tensor load of the [HW][C] matrices leaves
//   softmax        two passes over the key blocks (row maximum first, then exp2 with the exact maximum: no
//                  rescaling of the TMEM-resident output, and the same arithmetic as softmax(); with a
//                  single key block the first pass is skipped); thread r reads its S row with tcgen05.ld
//   O += P V       P (16-bit) is written by the softmax threads into a K-major 128B-swizzled A tile; V is
//                  used as it lies in memory ([keys][C], channel-contiguous) as an MN-major B operand
//                  (LBO = next 64-channel block, SBO = next 8 keys)
// Key/value blocks are double-buffered: the next block's tensor loads are in flight while the current one
// is multiplied and exponentiated.  Out-of-range keys (TMA zero fill) are masked to p = 0.
// ------------------------------------------------------------------------------------------
constexpr int kAttnQ = 128;   // queries per CTA
constexpr int kAttnKB = 64;   // keys per block

struct AttnParams {
  alignas(64) CUtensorMap q_map;  // {C, HW, B}, box {64, 128, 1}, 128B swizzle
  CUtensorMap k_map, v_map;       // {C, HW, B}, box {64, 64, 1}, 128B swizzle
  void* out;                      // [B][HW][C] 16-bit
  int32_t HW;
  float scale_log2;               // log2(e) / sqrt(C)
};

template <int C>
struct AttnCfg {
  static constexpr int kChunks = C / 64;
  static constexpr int kQBytes = kChunks * kAttnQ * 128;
  static constexpr int kKVBytes = kChunks * kAttnKB * 128;  // one K (or V) block
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;             // 2 stages
  static constexpr int kOffV = kOffK + 2 * kKVBytes;        // 2 stages
  static constexpr int kOffP = kOffV + 2 * kKVBytes;
  static constexpr int kOffBar = kOffP + kAttnQ * 128;
  static constexpr int kSmemBytes = kOffBar + 64 + 1024;    // + slack for the 1024-byte alignment of the base
  static constexpr int kTmemCols = C + 64 <= 128 ? 128 : (C + 64 <= 256 ? 256 : 512);
  static constexpr int kColO = 64;                          // S: columns [0, 64), O: [64, 64 + C)
};

__device__ __forceinline__ void tma_load_3d(const void* tmap, uint32_t dst_smem, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          dst_smem),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 128B-swizzled shared-memory matrix descriptor (version 1), explicit LBO / SBO in bytes
__device__ __forceinline__ uint64_t make_desc_sw128_lbo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <typename T, int C>
__global__ void __launch_bounds__(kAttnQ, 1) attn_core_kernel(const __grid_constant__ AttnParams P) {
  using Cfg = AttnCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, q0 = blockIdx.x * kAttnQ;
  const int HW = P.HW;
  const int nkb = (HW + kAttnKB - 1) / kAttnKB;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP);
  const uint32_t bar0 = smem_u32(smem + Cfg::kOffBar);
  const uint32_t bar_q = bar0, bar_s = bar0 + 8, bar_pv = bar0 + 16;
  auto bar_kv = [&](int s) { return bar0 + 24 + 8u * s; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffBar + 40);

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_pv, 1);
    mbar_init(bar_kv(0), 1);
    mbar_init(bar_kv(1), 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);

  constexpr uint32_t idesc_s = make_idesc_f16(128, kAttnKB, Cvt<T>::kFmt);
  constexpr uint32_t idesc_o = make_idesc_f16(128, C, Cvt<T>::kFmt) | (1u << 16);  // B (= V) is MN-major

  auto load_kv = [&](int kb, int stage, bool with_v) {  // thread 0 only
    const uint32_t bar = bar_kv(stage);
    mbar_arrive_expect_tx(bar, uint32_t(with_v ? 2 : 1) * Cfg::kKVBytes);
    for (int cc = 0; cc < Cfg::kChunks; ++cc) {
      tma_load_3d(&P.k_map, sK + stage * Cfg::kKVBytes + cc * (kAttnKB * 128), bar, cc * 64, kb * kAttnKB, b);
      if (with_v) tma_load_3d(&P.v_map, sV + stage * Cfg::kKVBytes + cc * (kAttnKB * 128), bar, cc * 64, kb * kAttnKB, b);
    }
  };
  auto mma_s = [&](int stage) {  // thread 0 only: S[128 x 64] = Q K^T
    tc_fence_after();
    for (int cc = 0; cc < Cfg::kChunks; ++cc)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_desc_sw128_lbo(sQ + cc * (kAttnQ * 128) + ks * 32, 16, 1024);
        const uint64_t bd = make_desc_sw128_lbo(sK + stage * Cfg::kKVBytes + cc * (kAttnKB * 128) + ks * 32, 16, 1024);
        umma_f16(tmem, ad, bd, idesc_s, (cc | ks) ? 1u : 0u);
      }
    umma_commit(bar_s);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, Cfg::kQBytes);
    for (int cc = 0; cc < Cfg::kChunks; ++cc) tma_load_3d(&P.q_map, sQ + cc * (kAttnQ * 128), bar_q, cc * 64, q0, b);
  }
  int g = 0;       // key blocks consumed so far (both passes): stage = g & 1, phase = (g >> 1) & 1
  int s_uses = 0;  // completed phases of bar_s
  float mrow = -INFINITY;

  // ---------------------------------------------------------------- pass 1: row maxima
  if (nkb > 1) {
    if (tid == 0) load_kv(0, g & 1, false);
    mbar_wait(bar_q, 0);
    for (int kb = 0; kb < nkb; ++kb, ++g) {
      const int stage = g & 1;
      if (tid == 0 && kb + 1 < nkb) load_kv(kb + 1, stage ^ 1, false);  // (its previous reader finished: bar_s waited)
      mbar_wait(bar_kv(stage), (g >> 1) & 1);
      if (tid == 0) mma_s(stage);
      mbar_wait(bar_s, s_uses & 1);
      ++s_uses;
      tc_fence_after();
      const int nvalid = HW - kb * kAttnKB;  // keys of this block inside the image
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t raw[32];
        tmem_ld32(lane_base + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (h * 32 + j < nvalid) mrow = fmaxf(mrow, __uint_as_float(raw[j]));
      }
      tc_fence_before();
      __syncthreads();  // everyone has read S before the next block's MMA overwrites it
    }
  } else {
    mbar_wait(bar_q, 0);
  }

  // ---------------------------------------------------------------- pass 2: P = exp2(..), O += P V
  float lsum = 0.f;
  if (tid == 0) load_kv(0, g & 1, true);
  for (int kb = 0; kb < nkb; ++kb, ++g) {
    const int stage = g & 1;
    if (kb >= 1) mbar_wait(bar_pv, (kb - 1) & 1);  // block kb-1's P V has read its stage and the P tile
    if (tid == 0 && kb + 1 < nkb) load_kv(kb + 1, stage ^ 1, true);
    mbar_wait(bar_kv(stage), (g >> 1) & 1);
    if (tid == 0) mma_s(stage);
    mbar_wait(bar_s, s_uses & 1);
    ++s_uses;
    tc_fence_after();
    const int nvalid = HW - kb * kAttnKB;
    uint32_t raw[2][32];
    tmem_ld32(lane_base, raw[0]);
    tmem_ld32(lane_base + 32, raw[1]);
    tmem_ld_wait();
    if (nkb == 1) {
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (j < nvalid) mrow = fmaxf(mrow, __uint_as_float(raw[j >> 5][j & 31]));
    }
    const float mb = mrow * P.scale_log2;
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // 8 keys = one 16-byte unit of the P row
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = u * 8 + e * 2;
        float p0 = exp2f(fmaf(__uint_as_float(raw[j >> 5][j & 31]), P.scale_log2, -mb));
        float p1 = exp2f(fmaf(__uint_as_float(raw[(j + 1) >> 5][(j + 1) & 31]), P.scale_log2, -mb));
        p0 = j < nvalid ? p0 : 0.f;
        p1 = j + 1 < nvalid ? p1 : 0.f;
        pk[e] = Cvt<T>::pack(p0, p1);
        const float2 r = Cvt<T>::unpack(pk[e]);  // normalise with the values the tensor cores will see
        lsum += r.x + r.y;
      }
      sts128(sP + uint32_t(tid) * 128 + uint32_t((u ^ (tid & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_desc_sw128_lbo(sP + ks * 32, 16, 1024);
        const uint64_t bd = make_desc_sw128_lbo(sV + stage * Cfg::kKVBytes + ks * (16 * 128), kAttnKB * 128, 1024);
        umma_f16(tmem + Cfg::kColO, ad, bd, idesc_o, (kb | ks) ? 1u : 0u);
      }
      umma_commit(bar_pv);
    }
  }
  mbar_wait(bar_pv, (nkb - 1) & 1);
  tc_fence_after();

  // ---------------------------------------------------------------- O / l -> 16-bit rows
  const int row = q0 + tid;
  const float inv = 1.0f / lsum;
  uint8_t* orow = reinterpret_cast<uint8_t*>(P.out) + (size_t(b) * HW + row) * C * 2;
#pragma unroll 1
  for (int cb = 0; cb < C / 32; ++cb) {
    uint32_t raw[32];
    tmem_ld32(lane_base + Cfg::kColO + cb * 32, raw);
    tmem_ld_wait();
    if (row < HW) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        uint4 u;
        u.x = Cvt<T>::pack(__uint_as_float(raw[k4 * 8 + 0]) * inv, __uint_as_float(raw[k4 * 8 + 1]) * inv);
        u.y = Cvt<T>::pack(__uint_as_float(raw[k4 * 8 + 2]) * inv, __uint_as_float(raw[k4 * 8 + 3]) * inv);
        u.z = Cvt<T>::pack(__uint_as_float(raw[k4 * 8 + 4]) * inv, __uint_as_float(raw[k4 * 8 + 5]) * inv);
        u.w = Cvt<T>::pack(__uint_as_float(raw[k4 * 8 + 6]) * inv, __uint_as_float(raw[k4 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + cb * 64 + k4 * 16) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::kTmemCols>(tmem);
}

}  // namespace fdsr
