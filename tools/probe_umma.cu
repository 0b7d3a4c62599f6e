// Bring-up probe for the tcgen05 conventions the conv kernel relies on (run on a B200 via gpurun):
//   1. no-swizzle K-major descriptors with an arbitrary 16-byte start offset and SBO = 160 B
//      (a 10-position-pitch input patch viewed as a 128-row A operand),
//   2. idesc encoding for fp16 / bf16, N = 64 / 128 / 256,
//   3. TMEM accumulator layout as read back by tcgen05.ld.32x32b,
//   4. issue-rate of back-to-back MMAs for the no-swizzle layout vs the 128B-swizzle layout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_umma tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../fastdiffsr_b200/csrc/ptx.cuh"

using namespace fdsr;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int KTOT = 64;       // one 64-channel chunk
constexpr int APOS = 181;      // positions per channel-group plane (odd => conflict-free STS)
constexpr int PITCH = 10;      // patch pitch (8 output px + 2 halo)
constexpr int BASEPOS = 11;    // tap (dy=1,dx=1)

struct Params {
  const uint16_t* a_img;  // smem image of A
  const uint16_t* b_img;  // smem image of B
  float* d;               // [128][N]
  long long* cycles;
  int a_bytes, b_bytes;
  int N;
  int fmt;     // 0 fp16, 1 bf16
  int layout;  // 0 nosw pitch10 shifted, 1 nosw contiguous (pitch 8, no shift), 2 sw128,
               // 3 sw128 pixel-major patch (128 B per position, pitch 10, shifted start, SBO = 1280 B) with
               //   no-swizzle B, 4 = 3 with the patch base 512 B off the 1024-B swizzle period
  int a_shift; // byte offset of the A image inside shared memory (layout 4)
  int reps;    // MMA repetitions (timing)
  int two_acc; // alternate two accumulators sharing B (timing)
  int ws;      // 1: tcgen05.mma.ws with the B operand latched in collector b0 (fill on the first accumulator, lastuse on
               //    the second); 2: .ws fill on every MMA (no reuse)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem + p.a_shift;
  uint8_t* sB = smem + ((p.a_shift + p.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < p.a_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(p.a_img)[i];
  for (int i = tid; i < p.b_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.b_img)[i];
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, p.N, p.fmt);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    // descriptors precomputed so the timed loop is MMA issue only
    uint64_t ad[KTOT / 16], bd[KTOT / 16], ad2[KTOT / 16];
#pragma unroll
    for (int ks = 0; ks < KTOT / 16; ++ks) {
      if (p.layout == 0) {
        ad[ks] = make_desc_nosw(a0 + (2 * ks) * APOS * 16 + BASEPOS * 16, APOS * 16, PITCH * 16);
        ad2[ks] = make_desc_nosw(a0 + (2 * ks) * APOS * 16 + (BASEPOS - 1) * 16, APOS * 16, PITCH * 16);
        bd[ks] = make_desc_nosw(b0 + (2 * ks) * p.N * 16, p.N * 16, 128);
      } else if (p.layout == 1) {
        ad[ks] = make_desc_nosw(a0 + (2 * ks) * 128 * 16, 128 * 16, 128);
        ad2[ks] = ad[ks];
        bd[ks] = make_desc_nosw(b0 + (2 * ks) * p.N * 16, p.N * 16, 128);
      } else if (p.layout == 2) {
        ad[ks] = make_desc_sw128(a0 + ks * 32, 1024);
        ad2[ks] = ad[ks];
        bd[ks] = make_desc_sw128(b0 + ks * 32, 1024);
      } else {
        ad[ks] = make_desc_sw128(a0 + BASEPOS * 128 + ks * 32, PITCH * 128);
        ad2[ks] = make_desc_sw128(a0 + (BASEPOS - 1) * 128 + ks * 32, PITCH * 128);
        bd[ks] = make_desc_nosw(b0 + (2 * ks) * p.N * 16, p.N * 16, 128);
      }
    }
    const uint32_t tmem2 = ((p.two_acc || p.ws) && p.N <= 128) ? tmem + 128 : tmem;
    t0 = clock64();
    if (p.reps == 1 && p.ws) {
#pragma unroll
      for (int ks = 0; ks < KTOT / 16; ++ks) {  // both accumulators get the same product (checked below on tmem)
        umma_f16_ws<0>(tmem2, ad[ks], bd[ks], idesc, ks != 0);
        umma_f16_ws<1>(tmem, ad[ks], bd[ks], idesc, ks != 0);
      }
    } else if (p.reps == 1) {
#pragma unroll
      for (int ks = 0; ks < KTOT / 16; ++ks) umma_f16(tmem, ad[ks], bd[ks], idesc, ks != 0);
    } else if (p.ws) {
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < KTOT / 16; ++ks) {
          umma_f16_ws<0>(tmem, ad[ks], bd[ks], idesc, 1);
          if (p.ws == 1) umma_f16_ws<1>(tmem2, ad2[ks], bd[ks], idesc, 1);
          else umma_f16_ws<0>(tmem2, ad2[ks], bd[ks], idesc, 1);
        }
      }
    } else {
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < KTOT / 16; ++ks) {
          umma_f16(tmem, ad[ks], bd[ks], idesc, 1);
          if (p.two_acc) umma_f16(tmem2, ad2[ks], bd[ks], idesc, 1);
        }
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (tid == 0) {
    t1 = clock64();
    p.cycles[0] = t1 - t0;
  }
  // epilogue: warp w owns lanes 32w..32w+31
  for (int c = 0; c < p.N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[(size_t)tid * p.N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

static uint16_t f2h(float f, int fmt) {
  if (fmt == 0) {
    __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static float h2f(uint16_t u, int fmt) {
  if (fmt == 0) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
  }
  __nv_bfloat16 h;
  memcpy(&h, &u, 2);
  return __bfloat162float(h);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair probe (tcgen05.mma.cta_group::2, M = 256): cluster of two CTAs, CTA r holds its own 128 rows of A and
// columns [N/2 r, N/2 r + N/2) of B (no-swizzle layouts 0 of the single-CTA probe); the leader issues.  Checks which
// B half lands in which accumulator columns and measures the issue rate with one / two accumulators sharing B.
struct PairParams {
  const uint16_t* a_img;  // [2][a_bytes] smem images of A (one per CTA)
  const uint16_t* b_img;  // [2][b_bytes] smem images of the B halves
  float* d;               // [2][128][N]
  long long* cycles;
  int a_bytes, b_bytes, N, fmt, reps, two_acc;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair_kernel(PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.a_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(p.a_img + size_t(rank) * p.a_bytes / 2)[i];
  for (int i = tid; i < p.b_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.b_img + size_t(rank) * p.b_bytes / 2)[i];
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<256, true>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int NH = p.N / 2;
  if (tid == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, p.N, p.fmt);
    const uint32_t a0 = smem_u32(sA) & 0x3FFFFu, b0 = smem_u32(sB) & 0x3FFFFu;
    uint64_t ad[KTOT / 16], bd[KTOT / 16], ad2[KTOT / 16];
#pragma unroll
    for (int ks = 0; ks < KTOT / 16; ++ks) {
      ad[ks] = make_desc_nosw(a0 + (2 * ks) * APOS * 16 + BASEPOS * 16, APOS * 16, PITCH * 16);
      ad2[ks] = make_desc_nosw(a0 + (2 * ks) * APOS * 16 + (BASEPOS - 1) * 16, APOS * 16, PITCH * 16);
      bd[ks] = make_desc_nosw(b0 + (2 * ks) * NH * 16, NH * 16, 128);
    }
    const uint32_t tmem2 = (p.two_acc && p.N <= 128) ? tmem + 128 : tmem;
    const long long t0 = clock64();
    if (p.reps == 1) {
#pragma unroll
      for (int ks = 0; ks < KTOT / 16; ++ks) umma_f16_pair(tmem, ad[ks], bd[ks], idesc, ks != 0);
    } else {
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < KTOT / 16; ++ks) {
          umma_f16_pair(tmem, ad[ks], bd[ks], idesc, 1);
          if (p.two_acc) umma_f16_pair(tmem2, ad2[ks], bd[ks], idesc, 1);
        }
      }
    }
    umma_commit_pair(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    p.cycles[0] = clock64() - t0;
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < p.N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[(size_t(rank) * 128 + tid) * p.N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc<256, true>(tmem);
}

static int run_pair_probe() {
  int fails = 0;
  CK(cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int fmt = 0; fmt < 2; ++fmt)
    for (int N : {64, 128, 256}) {
      std::vector<float> A(256 * KTOT), B((size_t)N * KTOT);
      srand(4321 + N + fmt);
      for (auto& x : A) x = h2f(f2h((rand() % 2001 - 1000) / 1000.0f, fmt), fmt);
      for (auto& x : B) x = h2f(f2h((rand() % 2001 - 1000) / 1000.0f, fmt), fmt);
      const int a_bytes = 8 * APOS * 16, NH = N / 2, b_bytes = NH * 128;
      std::vector<uint16_t> aimg(a_bytes, f2h(77.0f, fmt)), bimg(b_bytes, 0);  // two images each (bytes/2 elements x 2)
      for (int r = 0; r < 2; ++r) {
        for (int m = 0; m < 128; ++m) {
          const int pos = BASEPOS + (m / 8) * PITCH + (m % 8);
          for (int k = 0; k < KTOT; ++k)
            aimg[size_t(r) * a_bytes / 2 + ((k / 8) * APOS + pos) * 8 + (k % 8)] = f2h(A[(r * 128 + m) * KTOT + k], fmt);
        }
        for (int n = 0; n < NH; ++n)
          for (int k = 0; k < KTOT; ++k)
            bimg[size_t(r) * b_bytes / 2 + ((k / 8) * NH + n) * 8 + (k % 8)] = f2h(B[(size_t)(r * NH + n) * KTOT + k], fmt);
      }
      uint16_t *da, *db;
      float* dd;
      long long* dc;
      CK(cudaMalloc(&da, 2 * a_bytes));
      CK(cudaMalloc(&db, 2 * b_bytes));
      CK(cudaMalloc(&dd, 2 * 128 * N * 4));
      CK(cudaMalloc(&dc, 8));
      CK(cudaMemcpy(da, aimg.data(), 2 * a_bytes, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(db, bimg.data(), 2 * b_bytes, cudaMemcpyHostToDevice));
      PairParams p{da, db, dd, dc, a_bytes, b_bytes, N, fmt, 1, 0};
      const size_t smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
      probe_pair_kernel<<<2, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      std::vector<float> D(2 * 128 * N);
      CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < KTOT; ++k) ref += (double)A[m * KTOT + k] * B[(size_t)n * KTOT + k];
          maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        }
      long long cyc, cyc2;
      p.reps = 256;
      probe_pair_kernel<<<2, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
      p.two_acc = 1;
      probe_pair_kernel<<<2, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&cyc2, dc, 8, cudaMemcpyDeviceToHost));
      const bool ok = maxerr < 1e-3;
      if (!ok) ++fails;
      printf("pair (cta_group::2, M=256) fmt=%s N=%3d maxerr=%.3e %s  cycles/MMA(K=16)=%.1f two-acc=%.1f (ideal %d)\n",
             fmt ? "bf16" : "fp16", N, maxerr, ok ? "OK" : "FAIL", cyc / (256.0 * 4), cyc2 / (256.0 * 8), N / 2);
      cudaFree(da);
      cudaFree(db);
      cudaFree(dd);
      cudaFree(dc);
    }
  return fails;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  int fails = run_pair_probe();
  if (argc > 1 && !strcmp(argv[1], "pair")) {
    printf("pair probe %s (%d failures)\n", fails ? "FAILED" : "PASSED", fails);
    return fails ? 1 : 0;
  }
  for (int layout = 0; layout < 5; ++layout)
    for (int fmt = 0; fmt < 2; ++fmt)
      for (int N : {64, 128, 256}) {
        // logical operands
        std::vector<float> A(128 * KTOT), B((size_t)N * KTOT);
        srand(1234 + N + fmt);
        for (auto& x : A) x = h2f(f2h((rand() % 2001 - 1000) / 1000.0f, fmt), fmt);
        for (auto& x : B) x = h2f(f2h((rand() % 2001 - 1000) / 1000.0f, fmt), fmt);
        int a_bytes, b_bytes = N * 128, a_shift = 0;
        std::vector<uint16_t> aimg, bimg(b_bytes / 2, 0);
        if (layout == 0) {
          a_bytes = 8 * APOS * 16;
          aimg.assign(a_bytes / 2, f2h(77.0f, fmt));  // poison: wrong addressing shows up loudly
          for (int m = 0; m < 128; ++m) {
            int pos = BASEPOS + (m / 8) * PITCH + (m % 8);
            for (int k = 0; k < KTOT; ++k)
              aimg[((k / 8) * APOS + pos) * 8 + (k % 8)] = f2h(A[m * KTOT + k], fmt);
          }
        } else if (layout == 1) {
          a_bytes = 8 * 128 * 16;
          aimg.assign(a_bytes / 2, 0);
          for (int m = 0; m < 128; ++m)
            for (int k = 0; k < KTOT; ++k)
              aimg[((k / 8) * 128 + m) * 8 + (k % 8)] = f2h(A[m * KTOT + k], fmt);
        } else if (layout == 2) {
          a_bytes = 128 * 128;
          aimg.assign(a_bytes / 2, 0);
          for (int m = 0; m < 128; ++m)
            for (int k = 0; k < KTOT; ++k)
              aimg[m * 64 + (((k / 8) ^ (m & 7)) * 8) + (k % 8)] = f2h(A[m * KTOT + k], fmt);
        } else {
          // what a SWIZZLE_128B TMA load of the raw NHWC patch would leave in shared memory: position p
          // at byte p*128, its 16-byte channel groups XOR-ed with bits [7:9] of the absolute address
          a_shift = layout == 4 ? 512 : 0;
          a_bytes = APOS * 128;
          aimg.assign(a_bytes / 2, f2h(77.0f, fmt));
          for (int m = 0; m < 128; ++m) {
            int pos = BASEPOS + (m / 8) * PITCH + (m % 8);
            int phase = (pos + a_shift / 128) & 7;
            for (int k = 0; k < KTOT; ++k)
              aimg[pos * 64 + (((k / 8) ^ phase) * 8) + (k % 8)] = f2h(A[m * KTOT + k], fmt);
          }
        }
        if (layout != 2) {
          for (int n = 0; n < N; ++n)
            for (int k = 0; k < KTOT; ++k)
              bimg[((k / 8) * N + n) * 8 + (k % 8)] = f2h(B[(size_t)n * KTOT + k], fmt);
        } else {
          for (int n = 0; n < N; ++n)
            for (int k = 0; k < KTOT; ++k)
              bimg[n * 64 + (((k / 8) ^ (n & 7)) * 8) + (k % 8)] = f2h(B[(size_t)n * KTOT + k], fmt);
        }
        uint16_t *da, *db;
        float* dd;
        long long* dc;
        CK(cudaMalloc(&da, a_bytes));
        CK(cudaMalloc(&db, b_bytes));
        CK(cudaMalloc(&dd, 128 * N * 4));
        CK(cudaMalloc(&dc, 8));
        CK(cudaMemcpy(da, aimg.data(), a_bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bimg.data(), b_bytes, cudaMemcpyHostToDevice));
        Params p{da, db, dd, dc, a_bytes, b_bytes, N, fmt, layout, a_shift, 1, 0, 0};
        size_t smem = ((a_shift + a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
        probe_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * N);
        CK(cudaMemcpy(D.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KTOT; ++k) ref += (double)A[m * KTOT + k] * B[(size_t)n * KTOT + k];
            maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
          }
        // timing
        p.reps = 256;
        probe_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        long long cyc, cyc2;
        CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
        p.two_acc = 1;
        probe_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&cyc2, dc, 8, cudaMemcpyDeviceToHost));
        bool ok = maxerr < 1e-3;
        if (!ok) ++fails;
        // weight-stationary form (N <= 128 so that two accumulators fit the 256 allocated columns)
        double wserr = -1.0;
        long long cyc3 = 0, cyc4 = 0;
        if (N <= 128 && fmt == 0) {
          p.reps = 1;
          p.two_acc = 0;
          p.ws = 1;
          probe_kernel<<<1, 128, smem>>>(p);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(D.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost));
          wserr = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
              double ref = 0;
              for (int k = 0; k < KTOT; ++k) ref += (double)A[m * KTOT + k] * B[(size_t)n * KTOT + k];
              wserr = fmax(wserr, fabs(ref - D[m * N + n]));
            }
          p.reps = 256;
          probe_kernel<<<1, 128, smem>>>(p);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(&cyc3, dc, 8, cudaMemcpyDeviceToHost));
          p.ws = 2;
          probe_kernel<<<1, 128, smem>>>(p);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(&cyc4, dc, 8, cudaMemcpyDeviceToHost));
          if (wserr > 1e-3) ++fails;
        }
        printf("layout=%d fmt=%s N=%3d maxerr=%.3e %s  cycles/MMA(K=16)=%.1f two-acc=%.1f (ideal %d)  ws: err=%.3e "
               "two-acc fill+lastuse=%.1f fill+fill=%.1f\n", layout,
               fmt ? "bf16" : "fp16", N, maxerr, ok ? "OK" : "FAIL", cyc / (256.0 * 4),
               cyc2 / (256.0 * 8), N / 2, wserr, cyc3 / (256.0 * 8), cyc4 / (256.0 * 8));
        cudaFree(da);
        cudaFree(db);
        cudaFree(dd);
        cudaFree(dc);
      }
  printf("probe %s (%d failures)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
