// Bilinear logit upsample (align_corners=False) and its adjoint.
// Reference: segmentation_module.py:133  F.interpolate(sem_logits, size=out_size, mode="bilinear").
//
// Forward is write-bound (4 B per output element; every low-res value is reused ~(H/h)*(W/w) times
// and stays in L1/L2): a thread owns 4 consecutive output x of one output row and loops over the
// planes of its image block, so tap indices/weights are computed once and every store is a
// coalesced 128-bit streaming store.
// Backward is read-bound (4 B per full-res gradient element): a block owns RY low-res rows of one
// plane, streams the full-res rows that touch them once (coalesced), reduces them along y in
// registers, then along x out of shared memory.  No atomics; summation order is fixed.
#include "common.cuh"

namespace ucd {

constexpr int kUpThreads = 256;

// grid: x = ceil(W/(4*kUpThreads_x)) ... we flatten (Y, X4) into one index; blockIdx.y = plane group
template <int VEC>
__global__ void __launch_bounds__(kUpThreads)
upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int h, int w, int H,
                    int W, float scale_h, float scale_w, int planes_per_block) {
  const int wv = (W + VEC - 1) / VEC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)H * wv) return;
  const int Y = (int)(idx / wv);
  const int X = (int)(idx - (long long)Y * wv) * VEC;
  const Tap ty = bilinear_tap(Y, scale_h, h, H);
  int o00[VEC], o01[VEC], o10[VEC], o11[VEC];
  float wx0[VEC], wx1[VEC];
  const bool small_out = H + W <= 128;  // selects ATen's operation order (bit-exact parity with the CPU reference)
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int xx = (X + i < W) ? X + i : W - 1;
    const Tap tx = bilinear_tap(xx, scale_w, w, W);
    o00[i] = ty.i0 * w + tx.i0;
    o01[i] = ty.i0 * w + tx.i1;
    o10[i] = ty.i1 * w + tx.i0;
    o11[i] = ty.i1 * w + tx.i1;
    wx0[i] = tx.w0;
    wx1[i] = tx.w1;
  }
  const long long p0 = (long long)blockIdx.y * planes_per_block;
  const long long p1 = (p0 + planes_per_block < planes) ? p0 + planes_per_block : planes;
  const size_t in_plane = (size_t)h * w, out_plane = (size_t)H * W;
  // Fast path (any upscale factor >= ~4): the VEC outputs draw on at most 3 adjacent source columns,
  // so 6 loads per plane instead of 4*VEC keep the LSU below the store rate.
  const int xmin = o00[0] - ty.i0 * w;
  bool narrow = true;
#pragma unroll
  for (int i = 0; i < VEC; ++i) narrow = narrow && (o01[i] - ty.i0 * w - xmin <= 2) && (o00[i] - ty.i0 * w >= xmin);
  if (narrow) {
    int d0[VEC], d1[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      d0[i] = o00[i] - ty.i0 * w - xmin;
      d1[i] = o01[i] - ty.i0 * w - xmin;
    }
    const int c0 = xmin, c1 = min(xmin + 1, w - 1), c2 = min(xmin + 2, w - 1);
    const int ra = ty.i0 * w, rb = ty.i1 * w;
    for (long long p = p0; p < p1; ++p) {
      const float* src = in + p * in_plane;
      const float a0 = __ldg(src + ra + c0), a1 = __ldg(src + ra + c1), a2 = __ldg(src + ra + c2);
      const float b0 = __ldg(src + rb + c0), b1 = __ldg(src + rb + c1), b2 = __ldg(src + rb + c2);
      float r[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float v00 = d0[i] == 0 ? a0 : (d0[i] == 1 ? a1 : a2);
        const float v01 = d1[i] == 0 ? a0 : (d1[i] == 1 ? a1 : a2);
        const float v10 = d0[i] == 0 ? b0 : (d0[i] == 1 ? b1 : b2);
        const float v11 = d1[i] == 0 ? b0 : (d1[i] == 1 ? b1 : b2);
        r[i] = bilinear_blend(small_out, v00, v01, v10, v11, ty.w0, ty.w1, wx0[i], wx1[i]);
      }
      float* dst = out + p * out_plane + (size_t)Y * W + X;
      if (VEC == 4) {
        stg_stream4(dst, make_float4(r[0], r[1], r[2], r[3]));
      } else {
        stg_stream1(dst, r[0]);
      }
    }
    return;
  }
  for (long long p = p0; p < p1; ++p) {
    const float* src = in + p * in_plane;
    float r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      r[i] = bilinear_blend(small_out, __ldg(src + o00[i]), __ldg(src + o01[i]), __ldg(src + o10[i]),
                            __ldg(src + o11[i]), ty.w0, ty.w1, wx0[i], wx1[i]);
    float* dst = out + p * out_plane + (size_t)Y * W + X;
    if (VEC == 4) {
      stg_stream4(dst, make_float4(r[0], r[1], r[2], r[3]));
    } else {
      stg_stream1(dst, r[0]);
    }
  }
}

// Adjoint.  block = (plane, group of RY low-res rows).  smem: colsum[RY][W].
template <int RY>
__global__ void __launch_bounds__(kUpThreads)
upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int h, int w, int H, int W,
                    float scale_h, float scale_w, float inv_scale_h, float inv_scale_w) {
  extern __shared__ float colsum[];  // [RY][W]
  const long long plane = blockIdx.y;
  const int ybase = blockIdx.x * RY;
  const int ylast = min(ybase + RY, h) - 1;
  const float* g = gout + (size_t)plane * H * W;
  // conservative full-res row range touching low-res rows [ybase, ylast]; exact membership is decided per row
  int Ylo = (int)floorf(((float)ybase - 0.5f) * inv_scale_h - 0.5f) - 2;
  int Yhi = (int)ceilf(((float)ylast + 1.5f) * inv_scale_h - 0.5f) + 2;
  Ylo = max(Ylo, 0);
  Yhi = min(Yhi, H - 1);
  if (h == H) {
    Ylo = ybase;
    Yhi = ylast;
  }
  constexpr int KX = 4;  // columns per thread per pass
  for (int x0 = 0; x0 < W; x0 += KX * kUpThreads) {
    float acc[RY][KX];
#pragma unroll
    for (int r = 0; r < RY; ++r)
#pragma unroll
      for (int k = 0; k < KX; ++k) acc[r][k] = 0.f;
#pragma unroll 2
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      const Tap ty = bilinear_tap(Y, scale_h, h, H);
      const int r0 = ty.i0 - ybase, r1 = ty.i1 - ybase;
      const bool in0 = (r0 >= 0 && r0 < RY), in1 = (r1 >= 0 && r1 < RY);
      if (!in0 && !in1) continue;
      float v[KX];
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int X = x0 + k * kUpThreads + threadIdx.x;
        v[k] = (X < W) ? ldg_stream1(g + (size_t)Y * W + X) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        float wgt = 0.f;
        if (in0 && r == r0) wgt += ty.w0;
        if (in1 && r == r1) wgt += ty.w1;
#pragma unroll
        for (int k = 0; k < KX; ++k) acc[r][k] = fmaf(wgt, v[k], acc[r][k]);
      }
    }
#pragma unroll
    for (int r = 0; r < RY; ++r)
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int X = x0 + k * kUpThreads + threadIdx.x;
        if (X < W) colsum[r * W + X] = acc[r][k];
      }
  }
  __syncthreads();
  // reduce along x: out[y][x] = sum_X wx(X,x) colsum[y][X]
  const int n_out = RY * w;
  for (int o = threadIdx.x; o < n_out; o += kUpThreads) {
    const int r = o / w, x = o - r * w;
    if (ybase + r >= h) continue;
    int Xlo = (int)floorf(((float)x - 0.5f) * inv_scale_w - 0.5f) - 2;
    int Xhi = (int)ceilf(((float)x + 1.5f) * inv_scale_w - 0.5f) + 2;
    Xlo = max(Xlo, 0);
    Xhi = min(Xhi, W - 1);
    if (w == W) Xlo = Xhi = x;
    float acc = 0.f;
    for (int X = Xlo; X <= Xhi; ++X) {
      const Tap tx = bilinear_tap(X, scale_w, w, W);
      float wgt = 0.f;
      if (tx.i0 == x) wgt += tx.w0;
      if (tx.i1 == x) wgt += tx.w1;
      acc = fmaf(wgt, colsum[r * W + X], acc);
    }
    gin[((size_t)plane * h + ybase + r) * w + x] = acc;
  }
}

}  // namespace ucd

using namespace ucd;

extern "C" int ucd_upsample_bilinear_fwd(const float* in, float* out, int64_t planes, int h, int w, int H, int W,
                                         void* stream) {
  UCD_CHECK_ARG(in && out, "ucd_upsample_bilinear_fwd: null pointer");
  UCD_CHECK_ARG(planes > 0 && h > 0 && w > 0 && H > 0 && W > 0, "ucd_upsample_bilinear_fwd: bad shape");
  UCD_CHECK_ARG((long long)h * w < (1ll << 30), "ucd_upsample_bilinear_fwd: source plane too large");
  cudaStream_t st = (cudaStream_t)stream;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const bool v4 = (W % 4 == 0) && aligned16(out);
  const int wv = v4 ? W / 4 : W;
  const long long work = (long long)H * wv;
  const int gx = (int)((work + kUpThreads - 1) / kUpThreads);
  // enough blocks in y to fill the machine a few times over, while amortising the tap computation
  long long want_y = (4ll * kNumSMs + gx - 1) / gx;
  if (want_y < 1) want_y = 1;
  if (want_y > planes) want_y = planes;
  const int ppb = (int)((planes + want_y - 1) / want_y);
  const int gy = (int)((planes + ppb - 1) / ppb);
  UCD_CHECK_ARG(gy <= 65535, "ucd_upsample_bilinear_fwd: too many planes");
  dim3 grid(gx, gy);
  if (v4)
    upsample_fwd_kernel<4><<<grid, kUpThreads, 0, st>>>(in, out, planes, h, w, H, W, sh, sw, ppb);
  else
    upsample_fwd_kernel<1><<<grid, kUpThreads, 0, st>>>(in, out, planes, h, w, H, W, sh, sw, ppb);
  UCD_CHECK_LAUNCH("upsample_fwd_kernel");
  return UCD_OK;
}

extern "C" int ucd_upsample_bilinear_bwd(const float* gout, float* gin, int64_t planes, int h, int w, int H, int W,
                                         void* stream) {
  UCD_CHECK_ARG(gout && gin, "ucd_upsample_bilinear_bwd: null pointer");
  UCD_CHECK_ARG(planes > 0 && h > 0 && w > 0 && H > 0 && W > 0, "ucd_upsample_bilinear_bwd: bad shape");
  UCD_CHECK_ARG(planes <= 65535ll * 32768, "ucd_upsample_bilinear_bwd: too many planes");
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int RY = 4;
  const size_t smem = (size_t)RY * W * sizeof(float);
  UCD_CHECK_ARG(smem <= 200 * 1024, "ucd_upsample_bilinear_bwd: W=%d too wide", W);
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(upsample_bwd_kernel<RY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(upsample_bwd)");
  }
  // planes on grid.y is limited to 65535: fold the excess into grid.z-free loop by chunking launches
  const int gyb = (h + RY - 1) / RY;
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    const int np = (int)((planes - p0 < 65535) ? planes - p0 : 65535);
    dim3 grid(gyb, np);
    upsample_bwd_kernel<RY><<<grid, kUpThreads, smem, st>>>(gout + (size_t)p0 * H * W, gin + (size_t)p0 * h * w, h,
                                                             w, H, W, sh, sw, (float)H / (float)h,
                                                             (float)W / (float)w);
    UCD_CHECK_LAUNCH("upsample_bwd_kernel");
  }
  return UCD_OK;
}
