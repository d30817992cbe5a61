// Self-test of the tcgen05 / TMEM / bulk-copy building blocks with the production tile layout.
//   variant 0:  S[128,128] = A[128,256] * C[128,256]^T      (both operands K-major, like sweep MMA 1)
//   variant 1:  V[128,256] = E[128,128] * C[128,256]        (A K-major from thread-written smem,
//                                                             B = the same C tile read MN-major)
//   variant 2:  same product, but E is written to TENSOR MEMORY with tcgen05.st and the MMA reads its A
//               operand from there (the production path of the sweeps)
// Used by tests/test_gpu_umma.py; not on the product path.
#include "umma.cuh"
#include "../../include/ucd_b200_debug.h"

#include <stdlib.h>
#include <vector>

namespace ucd {

constexpr uint32_t ST_OFF_A = 0, ST_OFF_C = 65536, ST_OFF_E = 131072, ST_OFF_BAR = 163840, ST_SMEM = 163840 + 64;

__global__ void __launch_bounds__(128, 1)
selftest_kernel(int variant, const __nv_bfloat16* __restrict__ a_tile, const __nv_bfloat16* __restrict__ c_tile,
                const float* __restrict__ e_rows /*[128][128] fp32*/, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_ld = sbase + ST_OFF_BAR, bar_mma = sbase + ST_OFF_BAR + 8;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_ld, 2 * 65536u);
    for (int q = 0; q < 4; ++q) {
      bulk_g2s(sbase + ST_OFF_A + q * 16384u, reinterpret_cast<const uint8_t*>(a_tile) + q * 16384u, 16384u, bar_ld);
      bulk_g2s(sbase + ST_OFF_C + q * 16384u, reinterpret_cast<const uint8_t*>(c_tile) + q * 16384u, 16384u, bar_ld);
    }
  }
  if (variant == 1) {
    // every thread writes its row of E as bf16 into the K-major chunk layout, exactly like the sweep epilogue
    const int r = threadIdx.x;
    for (int kc = 0; kc < 16; ++kc) {
      uint32_t pk[4];
      for (int u = 0; u < 4; ++u) {
        __nv_bfloat162 v = __floats2bfloat162_rn(e_rows[r * 128 + kc * 8 + 2 * u], e_rows[r * 128 + kc * 8 + 2 * u + 1]);
        pk[u] = *reinterpret_cast<uint32_t*>(&v);
      }
      *reinterpret_cast<uint4*>(smem + ST_OFF_E + kc * 2048 + r * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    fence_proxy_async();
  }
  if (variant == 2) {
    // row r of E as packed bf16 pairs -> 64 TMEM columns at [256, 320) of this thread's lane
    const int r = threadIdx.x;
    for (int c16 = 0; c16 < 4; ++c16) {
      uint32_t pk[16];
      for (int u = 0; u < 16; ++u) {
        __nv_bfloat162 v = __floats2bfloat162_rn(e_rows[r * 128 + c16 * 32 + 2 * u], e_rows[r * 128 + c16 * 32 + 2 * u + 1]);
        pk[u] = *reinterpret_cast<uint32_t*>(&v);
      }
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c16 * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    if (variant == 0) {
      constexpr uint32_t idesc = umma_idesc(128, 128, 0, 0);
      for (int ks = 0; ks < 16; ++ks)
        umma_bf16(tmem, umma_desc(sbase + ST_OFF_A + ks * 4096, 2048, 128),
                  umma_desc(sbase + ST_OFF_C + ks * 4096, 2048, 128), idesc, ks > 0);
    } else if (variant == 1) {
      constexpr uint32_t idesc = umma_idesc(128, 256, 0, 1);
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(tmem, umma_desc(sbase + ST_OFF_E + kk * 4096, 2048, 128),
                  umma_desc(sbase + ST_OFF_C + kk * 16 * 16, 128, 2048), idesc, kk > 0);
    } else {
      constexpr uint32_t idesc = umma_idesc(128, 256, 0, 1);
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16_ts(tmem, tmem + 256 + kk * 8, umma_desc(sbase + ST_OFF_C + kk * 16 * 16, 128, 2048), idesc, kk > 0);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  const int ncols = variant == 0 ? 128 : 256;
  const int r = warp * 32 + lane;
  for (int cc = 0; cc < ncols / 32; ++cc) {
    uint32_t rv[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, rv);
    tmem_ld_wait();
    tmem_ld_fence(rv);
    for (int j = 0; j < 32; ++j) out[r * ncols + cc * 32 + j] = __uint_as_float(rv[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// MMA issue-rate probe: `iters` back-to-back tcgen05.mma of one shape onto one accumulator (K-step chain, like
// the sweeps), timed with clock64 from the first issue to the commit's arrival.
//   mode: 0 SS N=64 | 1 SS N=128 | 2 SS N=256 (B MN-major) | 3 TS N=64 | 4 TS N=128 | 5 TS N=256 (B MN-major)
//         6 SS N=256 K-major B | 7 TS N=128 alternating two accumulators | 8 SS N=128 alternating two accumulators
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + ST_OFF_BAR;
  for (int i = threadIdx.x; i < (int)(ST_OFF_E / 4); i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0 && mode <= 8) {
    const bool ts = (mode >= 3 && mode <= 5) || mode == 7;
    const int N = (mode == 0 || mode == 3) ? 64 : ((mode == 1 || mode == 4 || mode == 7 || mode == 8) ? 128 : 256);
    const bool mn = (mode == 2 || mode == 5);
    const uint32_t idesc = umma_idesc(128, N, 0, mn ? 1 : 0);
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 15;
      const uint64_t bdesc = mn ? umma_desc(sbase + ST_OFF_C + (ks & 7) * 256, 128, 2048)
                                : umma_desc(sbase + ST_OFF_C + ks * 4096, 2048, 128);
      const uint32_t d = tmem + 256 + (((mode == 7 || mode == 8) && (i & 1)) ? 128 : 0);
      if (ts)
        umma_bf16_ts(d, tmem + ks * 8, bdesc, idesc, 1u);
      else
        umma_bf16(d, umma_desc(sbase + ST_OFF_A + ks * 4096, 2048, 128), bdesc, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles[0] = clock64() - c0;
  }
  // modes 12..15: ONE warp, warp-uniform loop, 16 K steps per elect_one_sync() block (straight-line UTCHMMA in SASS):
  //   12 SS N=128 | 13 TS N=128 | 14 SS N=64 | 15 TS N=256 (B MN-major)
  if (mode >= 12 && mode <= 15 && warp == 0) {
    const bool ts = mode == 13 || mode == 15;
    const int N = mode == 14 ? 64 : (mode == 15 ? 256 : 128);
    const bool mn = mode == 15;
    const uint32_t idesc = umma_idesc(128, N, 0, mn ? 1 : 0);
    const uint64_t a0 = umma_desc(sbase + ST_OFF_A, 2048, 128);
    const uint64_t b0 = mn ? umma_desc(sbase + ST_OFF_C, 128, 2048) : umma_desc(sbase + ST_OFF_C, 2048, 128);
    const long long c0 = clock64();
    for (int i = 0; i < iters; i += 16) {
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t bdesc = umma_desc_adv(b0, mn ? (ks & 7) * 256 : ks * 4096);
          if (ts)
            umma_bf16_ts(tmem + 256, tmem + ks * 8, bdesc, idesc, 1u);
          else
            umma_bf16(tmem + 256, umma_desc_adv(a0, ks * 4096), bdesc, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) cycles[0] = clock64() - c0;
  }
  // modes 16..19: the sweep-1 mix from TWO warps, no data dependence between them; cycles per tile / 16 is reported
  // (iters = S MMAs).  warp 0 issues the 16 S K-steps of a tile (N=128), warp 1 the V K-steps:
  //   16: S = SS, V = 8 x TS N=256           17: S = SS, V = 16 x TS N=128 (feature halves, as the ring kernel does)
  //   18: S = TS (anchor tile in TMEM), V = 16 x TS N=128      19: S = 8 TS + 8 SS (half the anchor tile in TMEM), V as 17
  if (mode >= 16 && mode <= 19 && warp < 2) {
    const uint32_t mybar = bar + (warp == 1 ? 8 : 0);
    const uint32_t idesc_s = umma_idesc(128, 128, 0, 0), idesc_v = umma_idesc(128, 256, 0, 1),
                   idesc_h = umma_idesc(128, 128, 0, 1);
    const uint64_t a0 = umma_desc(sbase + ST_OFF_A, 2048, 128);
    const uint64_t bk = umma_desc(sbase + ST_OFF_C, 2048, 128), bm = umma_desc(sbase + ST_OFF_C, 128, 2048);
    const long long c0 = clock64();
    for (int i = 0; i < iters; i += 16) {
      if (elect_one_sync()) {
        if (warp == 0) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const bool ts = mode == 18 || (mode == 19 && ks < 8);
            if (ts)
              umma_bf16_ts(tmem + 128, tmem + ks * 8, umma_desc_adv(bk, ks * 4096), idesc_s, 1u);
            else
              umma_bf16(tmem + 128, umma_desc_adv(a0, ks * 4096), umma_desc_adv(bk, ks * 4096), idesc_s, 1u);
          }
        } else if (mode == 16) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(tmem + 256, tmem + 192 + kk * 8, umma_desc_adv(bm, kk * 256), idesc_v, 1u);
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            umma_bf16_ts(tmem + 256, tmem + 192 + kk * 8, umma_desc_adv(bm, kk * 256), idesc_h, 1u);
            umma_bf16_ts(tmem + 384, tmem + 192 + kk * 8, umma_desc_adv(bm, 32768 + kk * 256), idesc_h, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(mybar);
    __syncwarp();
    mbar_wait(mybar, 0);
    __syncwarp();
    if (threadIdx.x == 0) cycles[0] = clock64() - c0;
    if (threadIdx.x == 32) cycles[1] = clock64() - c0;
  }
  // modes 9..11: TWO issuing threads (warps 0 and 1), each its own accumulator: 9 SS N=128 | 10 TS N=128 | 11 SS N=64
  if (mode >= 9 && mode <= 11 && (threadIdx.x == 0 || threadIdx.x == 32)) {
    const bool ts = mode == 10;
    const int N = mode == 11 ? 64 : 128;
    const uint32_t idesc = umma_idesc(128, N, 0, 0);
    const uint32_t d = tmem + 256 + (threadIdx.x == 32 ? 128 : 0);
    const uint32_t mybar = bar + (threadIdx.x == 32 ? 8 : 0);
    const long long c0 = clock64();
    for (int i = 0; i < iters / 2; ++i) {
      const int ks = i & 15;
      const uint64_t bdesc = umma_desc(sbase + ST_OFF_C + ks * 4096, 2048, 128);
      if (ts)
        umma_bf16_ts(d, tmem + ks * 8, bdesc, idesc, 1u);
      else
        umma_bf16(d, umma_desc(sbase + ST_OFF_A + ks * 4096, 2048, 128), bdesc, idesc, 1u);
    }
    umma_commit(mybar);
    mbar_wait(mybar, 0);
    if (threadIdx.x == 0) cycles[0] = clock64() - c0;
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Two-warp mix probe with explicit tensor-memory placement: warp 0 issues the 16 S K-steps of a tile (N=128; SS, or TS
// with the A operand at column s_a), accumulating alternately at columns s_acc0 / s_acc1 per tile; warp 1 issues the V
// K-steps (TS, A operand at column v_a: 16 x N=128 into v_acc and v_acc+128, or 8 x N=256 into v_acc).
__global__ void __launch_bounds__(128, 1) mma_mix_kernel(int s_ts, int s_a, int s_acc0, int s_acc1, int v_n256, int v_a,
                                                         int v_acc, int tiles, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + ST_OFF_BAR;
  for (int i = threadIdx.x; i < (int)(ST_OFF_E / 4); i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp < 2) {
    const uint32_t mybar = bar + (warp == 1 ? 8 : 0);
    const uint32_t idesc_s = umma_idesc(128, 128, 0, 0), idesc_v = umma_idesc(128, 256, 0, 1),
                   idesc_h = umma_idesc(128, 128, 0, 1);
    const uint64_t a0 = umma_desc(sbase + ST_OFF_A, 2048, 128);
    const uint64_t bk = umma_desc(sbase + ST_OFF_C, 2048, 128), bm = umma_desc(sbase + ST_OFF_C, 128, 2048);
    const long long c0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      if (elect_one_sync()) {
        if (warp == 0) {
          const uint32_t d = tmem + ((t & 1) ? s_acc1 : s_acc0);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            if (s_ts)
              umma_bf16_ts(d, tmem + s_a + ks * 8, umma_desc_adv(bk, ks * 4096), idesc_s, ks > 0);
            else
              umma_bf16(d, umma_desc_adv(a0, ks * 4096), umma_desc_adv(bk, ks * 4096), idesc_s, ks > 0);
          }
        } else if (v_n256) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(tmem + v_acc, tmem + v_a + kk * 8, umma_desc_adv(bm, kk * 256), idesc_v, 1u);
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            umma_bf16_ts(tmem + v_acc, tmem + v_a + kk * 8, umma_desc_adv(bm, kk * 256), idesc_h, 1u);
            umma_bf16_ts(tmem + v_acc + 128, tmem + v_a + kk * 8, umma_desc_adv(bm, 32768 + kk * 256), idesc_h, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(mybar);
    __syncwarp();
    mbar_wait(mybar, 0);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) cycles[warp] = clock64() - c0;
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace ucd

using namespace ucd;

extern "C" int ucd_selftest_umma(int variant, float* max_err_host) {
  UCD_CHECK_ARG(variant >= 0 && variant <= 2, "ucd_selftest_umma: variant must be 0, 1 or 2");
  UCD_CHECK_ARG(max_err_host, "ucd_selftest_umma: null pointer");
  const int M = 128, D = 256;
  std::vector<float> A(M * D), C(M * D), E(M * 128);
  unsigned s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((s >> 8) & 0xffff) / 65536.f - 0.5f;
  };
  for (auto& v : A) v = bf16_round(rnd());
  for (auto& v : C) v = bf16_round(rnd());
  for (auto& v : E) v = bf16_round(rnd() * 4.f);
  std::vector<__nv_bfloat16> At(M * D), Ct(M * D);
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < D; ++k) {
      const size_t o = ((size_t)(k / 8) * 128 + r) * 8 + k % 8;
      At[o] = __float2bfloat16_rn(A[r * D + k]);
      Ct[o] = __float2bfloat16_rn(C[r * D + k]);
    }
  const int ncols = variant == 0 ? 128 : 256;
  std::vector<float> ref((size_t)M * ncols, 0.f), got((size_t)M * ncols, 0.f);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < ncols; ++j) {
      double acc = 0;
      if (variant == 0)
        for (int k = 0; k < D; ++k) acc += (double)A[i * D + k] * C[j * D + k];
      else
        for (int k = 0; k < 128; ++k) acc += (double)E[i * 128 + k] * C[k * D + j];
      ref[(size_t)i * ncols + j] = (float)acc;
    }
  __nv_bfloat16 *dA = nullptr, *dC = nullptr;
  float *dE = nullptr, *dO = nullptr;
  cudaError_t e;
#define ST_TRY(x)                                   \
  if ((e = (x)) != cudaSuccess) {                   \
    cudaFree(dA), cudaFree(dC), cudaFree(dE), cudaFree(dO); \
    return cuda_fail(e, #x);                        \
  }
  ST_TRY(cudaMalloc(&dA, At.size() * 2));
  ST_TRY(cudaMalloc(&dC, Ct.size() * 2));
  ST_TRY(cudaMalloc(&dE, E.size() * 4));
  ST_TRY(cudaMalloc(&dO, got.size() * 4));
  ST_TRY(cudaMemcpy(dA, At.data(), At.size() * 2, cudaMemcpyHostToDevice));
  ST_TRY(cudaMemcpy(dC, Ct.data(), Ct.size() * 2, cudaMemcpyHostToDevice));
  ST_TRY(cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice));
  ST_TRY(cudaMemset(dO, 0, got.size() * 4));
  ST_TRY(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
  selftest_kernel<<<1, 128, ST_SMEM>>>(variant, dA, dC, dE, dO);
  ST_TRY(cudaGetLastError());
  ST_TRY(cudaDeviceSynchronize());
  ST_TRY(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
#undef ST_TRY
  cudaFree(dA), cudaFree(dC), cudaFree(dE), cudaFree(dO);
  float mx = 0.f;
  for (size_t i = 0; i < ref.size(); ++i) mx = fmaxf(mx, fabsf(ref[i] - got[i]));
  *max_err_host = mx;
  return UCD_OK;
}

extern "C" int ucd_selftest_mma_rate(int mode, int iters, float* cycles_per_instr_host) {
  UCD_CHECK_ARG(mode >= 0 && mode <= 19 && iters > 0 && cycles_per_instr_host, "ucd_selftest_mma_rate: bad argument");
  long long* d = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d, 16)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaMemset(d, 0, 16);
  if ((e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM)) != cudaSuccess)
    return cuda_fail(e, "cudaFuncSetAttribute(mma_rate_kernel)");
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {  // second run is warm
    mma_rate_kernel<<<1, 128, ST_SMEM>>>(mode, iters, d);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) {
      cudaFree(d);
      return cuda_fail(e, "mma_rate_kernel");
    }
  }
  long long h2[2] = {0, 0};
  cudaMemcpy(h2, d, 16, cudaMemcpyDeviceToHost);
  h = h2[0] > h2[1] ? h2[0] : h2[1];
  cudaFree(d);
  *cycles_per_instr_host = (float)h / (float)iters;
  return UCD_OK;
}

// ---- CUDA-core pipe probe for the sweep epilogue: which instructions share the MUFU's issue rate? ----
// One block of `warps` warps, 8 independent values per thread; cycles per loop iteration are reported.
//   mode 0: 8 ex2            1: 4 cvt.rn.bf16x2.f32 (+8 FADD)     2: 8 ex2 + 4 cvt
//   mode 3: 8 ex2 + integer round-half-up pack (8 IADD + 4 PRMT)  4: 8 FADD only (loop overhead reference)
namespace ucd {
__global__ void pipe_rate_kernel(int mode, int iters, long long* out, unsigned* sink_out) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (float)(threadIdx.x + i);
  unsigned sink = 0;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0 || mode == 2 || mode == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ex2f(x[i] * -0.75f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += 1.0f;
    }
    if (mode == 1 || mode == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        __nv_bfloat162 v = __floats2bfloat162_rn(x[i], x[i + 1]);
        sink ^= *reinterpret_cast<unsigned*>(&v);
      }
    }
    if (mode == 3) {
#pragma unroll
      for (int i = 0; i < 8; i += 2)
        sink ^= __byte_perm(__float_as_uint(x[i]) + 0x8000u, __float_as_uint(x[i + 1]) + 0x8000u, 0x7632);
    }
  }
  const long long c1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i];
  if (acc == 123.456f) sink ^= 1u;
  sink_out[threadIdx.x] = sink;
  if (threadIdx.x == 0) out[0] = c1 - c0;
}
}  // namespace ucd

extern "C" int ucd_selftest_pipe_rate(int mode, int warps, int iters, float* cycles_per_iter_host) {
  UCD_CHECK_ARG(mode >= 0 && mode <= 4 && warps >= 1 && warps <= 32 && iters > 0 && cycles_per_iter_host,
                "ucd_selftest_pipe_rate: bad argument");
  long long* d = nullptr;
  unsigned* sink = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d, 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&sink, 4096)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  for (int rep = 0; rep < 2; ++rep) {
    pipe_rate_kernel<<<1, warps * 32>>>(mode, iters, d, sink);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) {
      cudaFree(d), cudaFree(sink);
      return cuda_fail(e, "pipe_rate_kernel");
    }
  }
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d), cudaFree(sink);
  *cycles_per_iter_host = (float)h / (float)iters;
  return UCD_OK;
}


extern "C" int ucd_selftest_mma_mix(int s_ts, int s_a, int s_acc0, int s_acc1, int v_n256, int v_a, int v_acc, int tiles,
                                    float* cycles_per_tile_host) {
  UCD_CHECK_ARG(tiles > 0 && cycles_per_tile_host, "ucd_selftest_mma_mix: bad argument");
  long long* d = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d, 16)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaMemset(d, 0, 16);
  if ((e = cudaFuncSetAttribute(mma_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM)) != cudaSuccess)
    return cuda_fail(e, "cudaFuncSetAttribute(mma_mix_kernel)");
  for (int rep = 0; rep < 2; ++rep) {
    mma_mix_kernel<<<1, 128, ST_SMEM>>>(s_ts, s_a, s_acc0, s_acc1, v_n256, v_a, v_acc, tiles, d);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) {
      cudaFree(d);
      return cuda_fail(e, "mma_mix_kernel");
    }
  }
  long long h2[2] = {0, 0};
  cudaMemcpy(h2, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  *cycles_per_tile_host = (float)(h2[0] > h2[1] ? h2[0] : h2[1]) / (float)tiles;
  return UCD_OK;
}


// ---- HBM read probe: what a read-only streaming kernel can reach on this part (the denominator of the read-bound
// kernels: UNCE / UNKD forward, upsample backward).  mode 0: UN 128-bit register loads in flight per thread, batch by
// batch; mode 1: cp.async ring of UN 16-byte copies per thread (rolling); mode 2: planes walked with a stride like
// the NCHW loss kernels (17 streams per thread block).  Returns microseconds per pass (best of `reps`).
namespace ucd {
template <int UN>
__global__ void __launch_bounds__(256) read_probe_regs(const float4* __restrict__ p, long long n4, float* sink) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UN - 1) * stride < n4; i += UN * stride) {
    float4 v[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) v[j] = ldg_stream4(reinterpret_cast<const float*>(p + i + j * stride));
#pragma unroll
    for (int j = 0; j < UN; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  for (; i < n4; i += stride) {
    const float4 v = ldg_stream4(reinterpret_cast<const float*>(p + i));
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1.2345e-30f) *sink = acc;
}
template <int UN>
__global__ void __launch_bounds__(256) read_probe_async(const float4* __restrict__ p, long long n4, float* sink) {
  extern __shared__ __align__(16) float4 ring[];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cnt = i0 < n4 ? (n4 - i0 + stride - 1) / stride : 0;
  auto fetch = [&](long long k) {
    if (k < cnt) cp_async16(ring + (size_t)(k % UN) * blockDim.x + threadIdx.x, p + i0 + k * stride);
    cp_async_commit();
  };
#pragma unroll
  for (int j = 0; j < UN - 1; ++j) fetch(j);
  for (long long k = 0; k < cnt; ++k) {
    fetch(k + UN - 1);
    cp_async_wait<UN - 1>();
    const float4 v = ring[(size_t)(k % UN) * blockDim.x + threadIdx.x];
    acc += v.x + v.y + v.z + v.w;
  }
  cp_async_wait<0>();
  if (acc == 1.2345e-30f) *sink = acc;
}
// every block streams its OWN contiguous region front to back (the access pattern of a per-block row sweep)
template <int UN>
__global__ void __launch_bounds__(256) read_probe_regions(const float4* __restrict__ p, long long n4, float* sink) {
  float acc = 0.f;
  const long long lo = n4 * (long long)blockIdx.x / gridDim.x, hi = n4 * ((long long)blockIdx.x + 1) / gridDim.x;
  long long i = lo + threadIdx.x;
  for (; i + (UN - 1) * (long long)blockDim.x < hi; i += UN * (long long)blockDim.x) {
    float4 v[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) v[j] = ldg_stream4(reinterpret_cast<const float*>(p + i + j * (long long)blockDim.x));
#pragma unroll
    for (int j = 0; j < UN; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  for (; i < hi; i += blockDim.x) {
    const float4 v = ldg_stream4(reinterpret_cast<const float*>(p + i));
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1.2345e-30f) *sink = acc;
}
// NCHW walk: thread = 4 adjacent pixels, C planes of HW floats each, CH channels in flight
template <int CH>
__global__ void __launch_bounds__(256) read_probe_planes(const float* __restrict__ p, int B, int C, long long HW,
                                                         float* sink) {
  float acc = 0.f;
  const long long gpi = HW / 4, n_groups = gpi * B;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups;
       g += (long long)gridDim.x * blockDim.x) {
    const long long b = g / gpi, px = (g - b * gpi) * 4;
    const float* xp = p + (b * C) * HW + px;
    for (int c = 0; c < C; c += CH) {
      float4 v[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k)
        v[k] = (c + k < C) ? ldg_stream4(xp + (long long)(c + k) * HW) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CH; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  }
  if (acc == 1.2345e-30f) *sink = acc;
}
}  // namespace ucd

extern "C" int ucd_selftest_read_probe(const void* buf, long long bytes, int mode, int un, int blocks_per_sm, int reps,
                                       float* us_host) {
  using namespace ucd;
  UCD_CHECK_ARG(buf && us_host && bytes >= (1 << 20) && reps >= 1, "ucd_selftest_read_probe: bad argument");
  float* sink = nullptr;
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const long long n4 = bytes / 16;
  const int grid = kNumSMs * blocks_per_sm;
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0);
    if (mode == 0) {
      if (un == 4) read_probe_regs<4><<<grid, 256>>>((const float4*)buf, n4, sink);
      else if (un == 8) read_probe_regs<8><<<grid, 256>>>((const float4*)buf, n4, sink);
      else read_probe_regs<16><<<grid, 256>>>((const float4*)buf, n4, sink);
    } else if (mode == 1) {
      if (un == 4) read_probe_async<4><<<grid, 256, 4 * 256 * 16>>>((const float4*)buf, n4, sink);
      else if (un == 8) read_probe_async<8><<<grid, 256, 8 * 256 * 16>>>((const float4*)buf, n4, sink);
      else {
        cudaFuncSetAttribute(read_probe_async<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 256 * 16);
        read_probe_async<16><<<grid, 256, 16 * 256 * 16>>>((const float4*)buf, n4, sink);
      }
    } else if (mode == 3) {
      if (un == 4) read_probe_regions<4><<<grid, 128>>>((const float4*)buf, n4, sink);
      else if (un == 8) read_probe_regions<8><<<grid, 128>>>((const float4*)buf, n4, sink);
      else read_probe_regions<16><<<grid, 128>>>((const float4*)buf, n4, sink);
    } else {
      const int C = 17, B = 24;
      const long long HW = bytes / 4 / C / B / 4 * 4;
      if (un == 4) read_probe_planes<4><<<grid, 256>>>((const float*)buf, B, C, HW, sink);
      else if (un == 8) read_probe_planes<8><<<grid, 256>>>((const float*)buf, B, C, HW, sink);
      else read_probe_planes<17><<<grid, 256>>>((const float*)buf, B, C, HW, sink);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return cuda_fail(e, "read_probe");
  *us_host = best * 1e3f;
  return UCD_OK;
}
