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

#include <stdlib.h>

namespace ucd {

constexpr int kUpThreads = 256;

// Forward.  A thread owns VEC consecutive output x of one output row and loops over the planes of its
// plane group.  blockIdx.x covers (Y, X/VEC) flattened, blockIdx.y the plane group.
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
  int x0[VEC], x1[VEC];
  float wx0[VEC], wx1[VEC];
  const bool small_out = H + W <= 128;  // selects ATen's operation order (bit-exact parity with the CPU reference)
  bool uniform = true;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const Tap tx = bilinear_tap(X + i, scale_w, w, W);
    x0[i] = tx.i0, x1[i] = tx.i1, wx0[i] = tx.w0, wx1[i] = tx.w1;
    uniform = uniform && tx.i0 == x0[0] && tx.i1 == x1[0];
  }
  const long long p0 = (long long)blockIdx.y * planes_per_block;
  const long long p1 = (p0 + planes_per_block < planes) ? p0 + planes_per_block : planes;
  const size_t in_plane = (size_t)h * w, out_plane = (size_t)H * W;
  float* dst = out + p0 * out_plane + (size_t)Y * W + X;
  if (uniform) {
    // All VEC outputs read the same 2x2 source cell (always the case for upscale factors that are a multiple of
    // 2*VEC, e.g. the 16x of DeepLab): 4 loads per plane, UP planes in flight.
    const float* s00 = in + p0 * in_plane + ty.i0 * w + x0[0];
    const int d01 = x1[0] - x0[0], d10 = (ty.i1 - ty.i0) * w;
    constexpr int UP = 4;
    long long p = p0;
    for (; p + UP <= p1; p += UP) {
      float v[UP][4];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const float* s = s00 + u * in_plane;
        v[u][0] = __ldg(s), v[u][1] = __ldg(s + d01), v[u][2] = __ldg(s + d10), v[u][3] = __ldg(s + d10 + d01);
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        float r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          r[i] = bilinear_blend(small_out, v[u][0], v[u][1], v[u][2], v[u][3], ty.w0, ty.w1, wx0[i], wx1[i]);
        if (VEC == 4)  // streaming (evict-first) stores: 5.1 TB/s vs 3.3 TB/s with default stores (scripts/up_probe.py)
          stg_stream4(dst + u * out_plane, make_float4(r[0], r[1], r[2], r[3]));
        else
          stg_stream1(dst + u * out_plane, r[0]);
      }
      s00 += UP * in_plane;
      dst += UP * out_plane;
    }
    for (; p < p1; ++p) {
      const float a = __ldg(s00), b = __ldg(s00 + d01), c = __ldg(s00 + d10), d = __ldg(s00 + d10 + d01);
      float r[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) r[i] = bilinear_blend(small_out, a, b, c, d, ty.w0, ty.w1, wx0[i], wx1[i]);
      if (VEC == 4)
        stg_stream4(dst, make_float4(r[0], r[1], r[2], r[3]));
      else
        stg_stream1(dst, r[0]);
      s00 += in_plane;
      dst += out_plane;
    }
    return;
  }
  for (long long p = p0; p < p1; ++p) {
    const float* src = in + p * in_plane;
    float r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      r[i] = bilinear_blend(small_out, __ldg(src + ty.i0 * w + x0[i]), __ldg(src + ty.i0 * w + x1[i]),
                            __ldg(src + ty.i1 * w + x0[i]), __ldg(src + ty.i1 * w + x1[i]), ty.w0, ty.w1, wx0[i],
                            wx1[i]);
    if (VEC == 4)
      stg_stream4(dst, make_float4(r[0], r[1], r[2], r[3]));
    else
      stg_stream1(dst, r[0]);
    dst += out_plane;
  }
}

// Forward, interval form (the one the training shapes take).  block = (plane, low-res row interval k, column tile):
// all output rows whose upper tap is source row k are contiguous in memory and share both source rows, so
//   out(Y, X) = fma(t0(X), h0(Y), t1(X) * h1(Y))   with  t_r(X) = fma(v_r0, w0, v_r1 * w1)
// (exactly ATen's generic-kernel operation order) costs one FMUL + one FMA per output after 4 loads per thread.
// Consecutive blocks write consecutive ~32 KB runs: the whole grid is one sequential write stream.
constexpr int kUpMaxIvRows = 96;  // candidate rows of one interval (2*scale + 5): scale factors up to 40
template <int VEC>
__global__ void __launch_bounds__(kUpThreads)
upsample_fwd_interval_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int H, int W,
                             float scale_h, float scale_w, int tw, int n_col_tiles) {
  __shared__ float s_h0[kUpMaxIvRows], s_h1[kUpMaxIvRows];
  __shared__ int s_lo[3], s_hi[3];
  const int wv = (W + VEC - 1) / VEC;
  const int rp = kUpThreads / tw;  // rows per pass
  const int per_plane = h * n_col_tiles;
  const long long plane = blockIdx.x / per_plane;
  const int rem = (int)(blockIdx.x - plane * per_plane);
  const int k = rem / n_col_tiles, col_tile = rem - k * n_col_tiles;
  // rows of this interval: all Y with tap.i0 == k (contiguous); candidates evaluated in parallel by 3 warps
  const float inv = (float)H / (float)h;
  int lo = (int)floorf(((float)k - 0.5f) * inv) - 2, hi = (int)ceilf(((float)k + 1.5f) * inv) + 2;
  lo = max(lo, 0), hi = min(hi, H - 1);
  if (threadIdx.x < kUpMaxIvRows) {
    const int Y = lo + (int)threadIdx.x;
    bool match = false;
    if (Y <= hi) {
      const Tap ty = bilinear_tap(Y, scale_h, h, H);
      match = ty.i0 == k;
      s_h0[threadIdx.x] = ty.w0, s_h1[threadIdx.x] = ty.w1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, match);
    if ((threadIdx.x & 31) == 0) {
      const int wq = threadIdx.x >> 5;
      s_lo[wq] = bal ? wq * 32 + __ffs(bal) - 1 : 1 << 30;
      s_hi[wq] = bal ? wq * 32 + 31 - __clz(bal) : -1;
    }
  }
  __syncthreads();
  const int ia = min(s_lo[0], min(s_lo[1], s_lo[2])), ib = max(s_hi[0], max(s_hi[1], s_hi[2]));
  const int nrow = ib - ia + 1;
  if (nrow <= 0) return;
  const int Ya = lo + ia;
  const int xg = col_tile * tw + (int)threadIdx.x % tw, rphase = (int)threadIdx.x / tw;
  if (xg >= wv || rphase >= rp) return;
  const int X = xg * VEC;
  const int k1 = min(k + 1, h - 1);
  const float* r0 = in + (size_t)plane * h * w + (size_t)k * w;
  const float* r1 = in + (size_t)plane * h * w + (size_t)k1 * w;
  float t0[VEC], t1[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const Tap tx = bilinear_tap(min(X + i, W - 1), scale_w, w, W);
    t0[i] = __fmaf_rn(__ldg(r0 + tx.i0), tx.w0, __fmul_rn(__ldg(r0 + tx.i1), tx.w1));
    t1[i] = __fmaf_rn(__ldg(r1 + tx.i0), tx.w0, __fmul_rn(__ldg(r1 + tx.i1), tx.w1));
  }
  float* dst = out + (size_t)plane * H * W + (size_t)(Ya + rphase) * W + X;
  const size_t step = (size_t)rp * W;
#pragma unroll 4
  for (int r = rphase; r < nrow; r += rp) {
    const float h0 = s_h0[ia + r], h1 = s_h1[ia + r];
    float o[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] = __fmaf_rn(t0[i], h0, __fmul_rn(t1[i], h1));
    if (VEC == 4)
      stg_stream4(dst, make_float4(o[0], o[1], o[2], o[3]));
    else
      stg_stream1(dst, o[0]);
    dst += step;
  }
}

// Adjoint.  block = (plane, group of RY low-res rows).  Dynamic smem: colsum[RY][W] | wrow[ny_cap][RY].
//   step 0: per full-res row Y of the group's footprint, its weight towards each of the RY low-res rows
//   step 1: stream the footprint rows once (VEC=4: 128-bit loads), reduce along y in registers -> colsum
//   step 2: reduce colsum along x into the RY*w outputs
template <int RY, int VEC>
__global__ void __launch_bounds__(kUpThreads)
upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int h, int w, int H, int W,
                    float scale_h, float scale_w, float inv_scale_h, float inv_scale_w, int ny_cap) {
  extern __shared__ __align__(16) float up_smem[];
  float* colsum = up_smem;                 // [RY][W]
  float* wrow = up_smem + (size_t)RY * W;  // [ny_cap][RY]
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
  Yhi = min(Yhi, Ylo + ny_cap - 1);  // ny_cap is sized by the host to cover the footprint
  const int ny = Yhi - Ylo + 1;
  for (int i = threadIdx.x; i < ny; i += kUpThreads) {
    const Tap ty = bilinear_tap(Ylo + i, scale_h, h, H);
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      float wgt = 0.f;
      if (ty.i0 - ybase == r) wgt += ty.w0;
      if (ty.i1 - ybase == r) wgt += ty.w1;
      wrow[i * RY + r] = wgt;
    }
  }
  for (int i = threadIdx.x; i < RY * W; i += kUpThreads) colsum[i] = 0.f;
  __syncthreads();
  const int ngroups = (W + VEC - 1) / VEC;                  // column groups per row
  const int nsplit = max(1, min(kUpThreads / ngroups, 8));  // row-interleaved thread teams per column group
  const int team = threadIdx.x / ngroups;
  const bool working = team < nsplit || ngroups >= kUpThreads;
  for (int g0 = 0; g0 < ngroups; g0 += kUpThreads) {
    const int xg = g0 + (ngroups >= kUpThreads ? threadIdx.x : threadIdx.x - team * ngroups);
    float acc[RY][VEC];
#pragma unroll
    for (int r = 0; r < RY; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    if (working && xg < ngroups) {
      const int ystep = ngroups >= kUpThreads ? 1 : nsplit;
      const float* gp = g + (size_t)Ylo * W + (size_t)xg * VEC;
#pragma unroll 4
      for (int i = (ngroups >= kUpThreads ? 0 : team); i < ny; i += ystep) {
        float v[VEC];
        if (VEC == 4) {
          const float4 t = ldg_stream4(gp + (size_t)i * W);
          v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
        } else {
          v[0] = ldg_stream1(gp + (size_t)i * W);
        }
        float wr[RY];
        if (RY == 4) {
          const float4 t = *reinterpret_cast<const float4*>(wrow + i * RY);
          wr[0] = t.x, wr[1] = t.y, wr[2] = t.z, wr[3] = t.w;
        } else {
#pragma unroll
          for (int r = 0; r < RY; ++r) wr[r] = wrow[i * RY + r];
        }
#pragma unroll
        for (int r = 0; r < RY; ++r)
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[r][k] = fmaf(wr[r], v[k], acc[r][k]);
      }
    }
    // fixed-order combination of the teams (deterministic)
    for (int s = 0; s < (ngroups >= kUpThreads ? 1 : nsplit); ++s) {
      if (working && xg < ngroups && (ngroups >= kUpThreads || team == s)) {
#pragma unroll
        for (int r = 0; r < RY; ++r)
#pragma unroll
          for (int k = 0; k < VEC; ++k) colsum[r * W + xg * VEC + k] += acc[r][k];
      }
      __syncthreads();
    }
  }
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

// Adjoint, sweep form (integer upscale factors that are a multiple of 8 along x, rows of <= 1024 elements: the
// training shapes).  block = (plane, segment of seg_rows low-res rows); thread = 4 consecutive X, which then share
// one pair of x taps.  The block walks the full-res rows of its segment top-down exactly once (plus the one interval
// above it, whose lower-tap part belongs to the segment's first row): per 128-bit load 8 FMAs fold the x weights,
// 4 more the y weights; a low-res row is emitted when the sweep leaves its interval (shared-memory gather over the
// threads whose taps hit each cell, fixed order).  No atomics, no re-reads beyond 1/seg_rows.
constexpr int kUpSweepCap = 640;  // tap-table rows: (seg_rows + 2) * scale + 8 must fit (footprint + margins)
template <int UN>
__global__ void __launch_bounds__(256)
upsample_bwd_sweep_kernel(const float* __restrict__ gout, float* __restrict__ gin, int h, int w, int H, int W,
                          float scale_h, float scale_w, int nseg, int seg_rows) {
  __shared__ int s_k[kUpSweepCap];
  __shared__ float s_h0[kUpSweepCap], s_h1[kUpSweepCap];
  __shared__ float s_p0[256], s_p1[256];
  __shared__ int s_x0[256], s_x1[256];
  const int wv = blockDim.x;  // == W / 4
  const long long plane = blockIdx.x / nseg;
  const int seg = (int)(blockIdx.x - plane * nseg);
  const int y0 = seg * seg_rows, ylast = min(y0 + seg_rows, h) - 1;
  const int kfirst = max(y0 - 1, 0);  // first interval read (for y0 > 0 only its lower-tap part is used)
  const float inv = (float)H / (float)h;
  int Ylo = (int)floorf(((float)kfirst - 0.5f) * inv) - 2, Yhi = (int)ceilf(((float)ylast + 1.5f) * inv) + 2;
  Ylo = max(Ylo, 0), Yhi = min(Yhi, H - 1);
  const int ny = min(Yhi - Ylo + 1, kUpSweepCap);
  for (int i = threadIdx.x; i < ny; i += blockDim.x) {
    const Tap ty = bilinear_tap(Ylo + i, scale_h, h, H);
    s_k[i] = ty.i0;
    const bool clamped = ty.i1 == ty.i0;  // bottom border: both taps hit the same low-res row
    s_h0[i] = clamped ? ty.w0 + ty.w1 : ty.w0;
    s_h1[i] = clamped ? 0.f : ty.w1;
  }
  const int X = (int)threadIdx.x * 4;
  float wx0[4], wx1[4];
  {
    const Tap t0 = bilinear_tap(X, scale_w, w, W);
    s_x0[threadIdx.x] = t0.i0, s_x1[threadIdx.x] = t0.i1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const Tap tx = bilinear_tap(X + i, scale_w, w, W);  // same i0 / i1 as t0 (host-checked precondition)
      wx0[i] = tx.w0, wx1[i] = tx.w1;
    }
  }
  __syncthreads();
  const float* g = gout + (size_t)plane * H * W + (size_t)Ylo * W + X;
  float* out = gin + (size_t)plane * h * w;
  const int sc4 = (W / w) / 4;  // threads per low-res column
  float a0p0 = 0.f, a0p1 = 0.f, a1p0 = 0.f, a1p1 = 0.f, c0 = 0.f, c1 = 0.f;  // current interval, carry from the one above
  int cur = -1;
  // leave interval `cur`: emit low-res row cur (if it belongs to this segment), its lower-tap sums become the carry
  auto flush = [&]() {
    if (cur >= y0 && cur <= ylast) {
      s_p0[threadIdx.x] = a0p0 + c0;
      s_p1[threadIdx.x] = a0p1 + c1;
      __syncthreads();
      if ((int)threadIdx.x < w) {
        const int x = threadIdx.x;
        const int j0 = max((x - 2) * sc4 - 2, 0), j1 = min((x + 2) * sc4 + 2, wv - 1);
        float acc = 0.f;
        for (int j = j0; j <= j1; ++j) {
          if (s_x0[j] == x) acc += s_p0[j];
          if (s_x1[j] == x) acc += s_p1[j];
        }
        out[(size_t)cur * w + x] = acc;
      }
      __syncthreads();
    }
    c0 = a1p0, c1 = a1p1;
    a0p0 = a0p1 = a1p0 = a1p1 = 0.f;
  };
  for (int base = 0; base < ny; base += UN) {
    float4 v[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const int i = base + j;
      const bool use = i < ny && s_k[min(i, ny - 1)] >= kfirst && s_k[min(i, ny - 1)] <= ylast;
      v[j] = use ? ldg_stream4(g + (size_t)i * W) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const int i = base + j;
      if (i < ny) {
        const int k = s_k[i];
        if (k >= kfirst && k <= ylast) {  // uniform over the block
          if (k != cur) {
            if (cur >= 0) flush();
            cur = k;
          }
          const float p0 = fmaf(v[j].w, wx0[3], fmaf(v[j].z, wx0[2], fmaf(v[j].y, wx0[1], v[j].x * wx0[0])));
          const float p1 = fmaf(v[j].w, wx1[3], fmaf(v[j].z, wx1[2], fmaf(v[j].y, wx1[1], v[j].x * wx1[0])));
          const float h0 = s_h0[i], h1 = s_h1[i];
          a0p0 = fmaf(h0, p0, a0p0), a0p1 = fmaf(h0, p1, a0p1);
          a1p0 = fmaf(h1, p0, a1p0), a1p1 = fmaf(h1, p1, a1p1);
        }
      }
    }
  }
  if (cur >= 0) flush();
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
  // upsampling beyond ATen's small-output regime (every training shape): interval kernel
  if (H + W > 128 && H >= h && W >= w && (double)H / h <= 40.0 /* candidate rows per interval: 2*scale + 5 <= kUpMaxIvRows */ && planes * (long long)h < (1ll << 30) / 64) {
    const int tw = wv < kUpThreads ? wv : kUpThreads;
    const int nct = (wv + tw - 1) / tw;
    const unsigned nblk = (unsigned)(planes * h * nct);
    if (v4)
      upsample_fwd_interval_kernel<4><<<nblk, kUpThreads, 0, st>>>(in, out, h, w, H, W, sh, sw, tw, nct);
    else
      upsample_fwd_interval_kernel<1><<<nblk, kUpThreads, 0, st>>>(in, out, h, w, H, W, sh, sw, tw, nct);
    UCD_CHECK_LAUNCH("upsample_fwd_interval_kernel");
    return UCD_OK;
  }
  const long long work = (long long)H * wv;
  const int gx = (int)((work + kUpThreads - 1) / kUpThreads);
  // planes per block: enough to amortise the tap computation (>= 16 planes when there are that many), few enough
  // that the grid is many waves deep (blocks finish at different times; a 2-3 wave grid left SMs idle)
  long long want_y = (32ll * kNumSMs + gx - 1) / gx;
  if (want_y < 1) want_y = 1;
  if (want_y > (planes + 15) / 16) want_y = (planes + 15) / 16;
#ifdef UCD_DEBUG_KNOBS
  if (const char* e = getenv("UCD_UP_GY")) want_y = atoi(e) > 0 ? atoi(e) : want_y;  // tuning knob
#endif
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
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int RY = 4;
  // full-res rows that can touch RY consecutive low-res rows (+ the slack the kernel adds around its estimate)
  const int ny_cap = (h == H) ? RY : (int)((double)(RY + 2) * H / h) + 8;
  const size_t smem = ((size_t)RY * W + (size_t)ny_cap * RY) * sizeof(float);
  UCD_CHECK_ARG(smem <= 200 * 1024, "ucd_upsample_bilinear_bwd: W=%d / scale too large for one block", W);
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const bool v4 = (W % 4 == 0) && aligned16(gout);
  // sweep form: integer scales, a multiple of 8 along x (4 consecutive X share their taps), one block row = W/4 threads
  // low-res rows per block: 16 measured best (98 us vs 106 / 113 us for 8 / 4 at 24x17x512x512: re-reading the
  // interval above the segment is not free), reduced until the tap table fits; 8 rows of loads in flight per thread
  int seg_rows = 16;
  while (seg_rows > 1 && (seg_rows + 2) * (H / (h > 0 ? h : 1)) + 8 > kUpSweepCap) seg_rows >>= 1;
  if (v4 && H % h == 0 && W % w == 0 && (W / w) % 8 == 0 && W / 4 <= 256 && (W / 4) % 32 == 0 && w <= W / 4 &&
      (seg_rows + 2) * (H / h) + 8 <= kUpSweepCap) {
    const int nseg = (h + seg_rows - 1) / seg_rows;
    UCD_CHECK_ARG(planes * nseg < (1ll << 31), "ucd_upsample_bilinear_bwd: too many blocks");
    const unsigned nb = (unsigned)(planes * nseg);
    upsample_bwd_sweep_kernel<8><<<nb, W / 4, 0, st>>>(gout, gin, h, w, H, W, sh, sw, nseg, seg_rows);
    UCD_CHECK_LAUNCH("upsample_bwd_sweep_kernel");
    return UCD_OK;
  }
  auto kern = v4 ? upsample_bwd_kernel<RY, 4> : upsample_bwd_kernel<RY, 1>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(upsample_bwd)");
  }
  const int gyb = (h + RY - 1) / RY;
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {  // grid.y limit
    const int np = (int)((planes - p0 < 65535) ? planes - p0 : 65535);
    dim3 grid(gyb, np);
    kern<<<grid, kUpThreads, smem, st>>>(gout + (size_t)p0 * H * W, gin + (size_t)p0 * h * w, h, w, H, W, sh, sw,
                                         (float)H / (float)h, (float)W / (float)w, ny_cap);
    UCD_CHECK_LAUNCH("upsample_bwd_kernel");
  }
  return UCD_OK;
}
