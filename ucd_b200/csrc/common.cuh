// Shared helpers for the ucd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ucd_b200.h"

namespace ucd {

// ---- thread-local error string ------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define UCD_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::ucd::set_error(__VA_ARGS__);        \
      return UCD_EINVAL;                    \
    }                                       \
  } while (0)

#define UCD_CHECK_LAUNCH(what)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::ucd::cuda_fail(e__, what);  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs (grid sizing only: any SM count runs correctly)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- device math --------------------------------------------------------------------------
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Online log-sum-exp in the log2 domain with ONE exp per element:
//   state (m, s) represents log2-sum = m + log2(s);  v is the new element already scaled by log2(e).
__device__ __forceinline__ void lse_push(float& m, float& s, float v) {
  float d = v - m;
  float e = ex2f(-fabsf(d));
  s = (d > 0.f) ? fmaf(s, e, 1.f) : (s + e);
  m = fmaxf(m, v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of `v` (all threads must call; blockDim.x multiple of 32, <= 1024).
// Deterministic: fixed shuffle tree + fixed warp order.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = 0.f;
  if (wid == 0) {
    r = (lane < nw) ? smem32[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L1 bypass) and its group bookkeeping
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- bilinear taps, ATen semantics (align_corners=False), exact fp32 op order -------------
struct Tap {
  int i0, i1;
  float w0, w1;
};
// scale = float(in)/float(out) computed by the caller (one fp32 division, like ATen)
__device__ __forceinline__ Tap bilinear_tap(int dst, float scale, int in_size, int out_size) {
  Tap t;
  if (in_size == out_size) {
    t.i0 = t.i1 = dst;
    t.w0 = 1.f;
    t.w1 = 0.f;
    return t;
  }
  float src = __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);  // ATen's x86 build contracts this to one fma
  src = src < 0.f ? 0.f : src;
  int i0 = (int)floorf(src);
  i0 = i0 < in_size - 1 ? i0 : in_size - 1;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  float l1 = __fsub_rn(src, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  t.w1 = l1;
  t.w0 = __fsub_rn(1.f, l1);
  return t;
}

// 4-tap blend with ATen's CPU operation order (oracle/ucd_oracle.py::_bilinear_eval_f32):
//   small outputs (out_h + out_w <= 128, ATen's "vectorized" kernel): weight products first, then an fma chain
//   otherwise (generic N-d kernel): blend along x, then along y
__device__ __forceinline__ float bilinear_blend(bool small_out, float v00, float v01, float v10, float v11,
                                                float h0, float h1, float w0, float w1) {
  if (small_out) {
    float acc = __fmul_rn(__fmul_rn(h0, w1), v01);
    acc = __fmaf_rn(__fmul_rn(h0, w0), v00, acc);
    acc = __fmaf_rn(__fmul_rn(h1, w0), v10, acc);
    return __fmaf_rn(__fmul_rn(h1, w1), v11, acc);
  }
  const float t0 = __fmaf_rn(v00, w0, __fmul_rn(v01, w1));
  const float t1 = __fmaf_rn(v10, w0, __fmul_rn(v11, w1));
  return __fmaf_rn(t0, h0, __fmul_rn(t1, h1));
}

}  // namespace ucd
