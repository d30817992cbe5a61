// Contrastive prep: label downsample / pseudo-label mixing / compaction / normalise / bf16 tile pack.
// Reference: utils/loss.py:258-395 (pre_contrastive_pixel, v2 branch) == utils/utils.py:256-397;
// exact semantics restated in SURVEY.md Appendix A.1 and oracle/ucd_oracle.py.
//
// Everything here is O(N_px * D) streaming work (HBM/L2 bound, tiny next to the N_a x N_c sweeps).
// No host synchronisation: counts, ranks and offsets stay on the device.
#include "common.cuh"

namespace ucd {

constexpr int kPrepBlock = 256;  // pixels per counting block

// Per-pixel metadata planes (int32 [n_px] each) and per-block offset tables.
//   px_meta: 0 label_n | 1 mix | 2 flags | 3 rank_a | 4 rank_o | 5 srank_a | 6 srank_o
//     flags: bit0 anchor (mix>0), bit1 pseudo (anchor & !GT-new), bit2 GT-new (label_n>0)
//     rank_*: rank among the block's anchors / pseudos in pixel order (reference row order)
//     srank_*: rank among the block's anchors / pseudos OF THE SAME LABEL (class-sorted tile order)
//   blk_meta: ref_off[2][nblk] | cls_off[2][nb][nblk]   (counts after the label kernel, exclusive offsets after the scan)
enum { PX_LABEL_N = 0, PX_MIX, PX_FLAGS, PX_RANK_A, PX_RANK_O, PX_SRANK_A, PX_SRANK_O, PX_PLANES };

struct BlkMeta {
  int* ref;  // [2][nblk]
  int* cls;  // [2][nb][nblk]
  int nblk, nb;
  __host__ __device__ BlkMeta(int* base, int nblk_, int nb_) : ref(base), cls(base + 2 * nblk_), nblk(nblk_), nb(nb_) {}
  __device__ int& ref_at(int set, int blk) const { return ref[set * nblk + blk]; }
  __device__ int& cls_at(int set, int c, int blk) const { return cls[((size_t)set * nb + c) * nblk + blk]; }
};

__global__ void __launch_bounds__(kPrepBlock)
prep_labels_kernel(const long long* __restrict__ labels, const float* __restrict__ l_po, int B, int C_old, int h,
                   int w, int H, int W, int max_label, float scale_h, float scale_w, int* __restrict__ px_meta,
                   int* __restrict__ blk_base, int nb, int* __restrict__ counts) {
  extern __shared__ int wcnt[];  // [2][8 warps][nb] per-warp class counts
  const int n_px = B * h * w;
  const int nblk = gridDim.x;
  const BlkMeta bm(blk_base, nblk, nb);
  const int p = blockIdx.x * kPrepBlock + threadIdx.x;
  for (int i = threadIdx.x; i < 2 * 8 * nb; i += kPrepBlock) wcnt[i] = 0;
  int is_a = 0, is_o = 0, m = 0;
  if (p < n_px) {
    const int hw = h * w;
    const int b = p / hw, q = p - b * hw;
    const int y = q / w, x = q - y * w;
    // --- bilinear label downsample, bit-exact with ATen's CPU kernel (oracle/_bilinear_eval_f32) ---
    const Tap ty = bilinear_tap(y, scale_h, H, h);
    const Tap tx = bilinear_tap(x, scale_w, W, w);
    const long long* lb = labels + (size_t)b * H * W;
    const float v00 = (float)lb[(size_t)ty.i0 * W + tx.i0], v01 = (float)lb[(size_t)ty.i0 * W + tx.i1];
    const float v10 = (float)lb[(size_t)ty.i1 * W + tx.i0], v11 = (float)lb[(size_t)ty.i1 * W + tx.i1];
    const float acc = bilinear_blend(h + w <= 128, v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1);
    int g = (int)truncf(acc);                 // .type(int8) + the two masked fills (loss.py:262,269-270)
    if (g < 0 || g > max_label) g = 0;
    // --- pseudo label: first maximal channel of the old model's low-res logits (loss.py:357) ---
    const float* lp = l_po + (size_t)b * C_old * hw + q;
    float best = lp[0];
    int arg = 0;
    for (int c = 1; c < C_old; ++c) {
      const float v = lp[(size_t)c * hw];
      if (v > best) {
        best = v;
        arg = c;
      }
    }
    m = g > 0 ? g : arg;
    is_a = m > 0;
    is_o = is_a && !(g > 0);
    px_meta[(size_t)PX_LABEL_N * n_px + p] = g;
    px_meta[(size_t)PX_MIX * n_px + p] = m;
    px_meta[(size_t)PX_FLAGS * n_px + p] = is_a | (is_o << 1) | ((g > 0) << 2);
    if (g > 0) atomicMin(&counts[2], g);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  // pixel-order ranks (reference row order)
  __shared__ int wsum_a[kPrepBlock / 32], wsum_o[kPrepBlock / 32];
  const unsigned ba = __ballot_sync(0xffffffffu, is_a), bo = __ballot_sync(0xffffffffu, is_o);
  int ra = __popc(ba & lt), ro = __popc(bo & lt);
  // same-label ranks (class-sorted order): lanes holding the same label form a match group
  const unsigned ma = __match_any_sync(0xffffffffu, is_a ? m : (0x40000000 | lane));
  const unsigned mo = __match_any_sync(0xffffffffu, is_o ? m : (0x40000000 | lane));
  int sra = __popc(ma & lt), sro = __popc(mo & lt);
  __syncthreads();  // wcnt zeroed
  if (lane == 0) {
    wsum_a[wid] = __popc(ba);
    wsum_o[wid] = __popc(bo);
  }
  if (is_a && sra == 0) wcnt[(0 * 8 + wid) * nb + m] = __popc(ma);
  if (is_o && sro == 0) wcnt[(1 * 8 + wid) * nb + m] = __popc(mo);
  __syncthreads();
  int tot_a = 0, tot_o = 0;
  for (int i = 0; i < kPrepBlock / 32; ++i) {
    if (i < wid) {
      ra += wsum_a[i];
      ro += wsum_o[i];
      if (is_a) sra += wcnt[(0 * 8 + i) * nb + m];
      if (is_o) sro += wcnt[(1 * 8 + i) * nb + m];
    }
    tot_a += wsum_a[i];
    tot_o += wsum_o[i];
  }
  if (p < n_px) {
    px_meta[(size_t)PX_RANK_A * n_px + p] = ra;
    px_meta[(size_t)PX_RANK_O * n_px + p] = ro;
    px_meta[(size_t)PX_SRANK_A * n_px + p] = sra;
    px_meta[(size_t)PX_SRANK_O * n_px + p] = sro;
  }
  if (threadIdx.x == 0) {
    bm.ref_at(0, blockIdx.x) = tot_a;
    bm.ref_at(1, blockIdx.x) = tot_o;
  }
  for (int i = threadIdx.x; i < 2 * nb; i += kPrepBlock) {
    const int set = i / nb, c = i - set * nb;
    int t = 0;
    for (int k = 0; k < 8; ++k) t += wcnt[(set * 8 + k) * nb + c];
    bm.cls_at(set, c, blockIdx.x) = t;
  }
}

// in-place exclusive scan of data[0..n) by one 1024-thread block; returns the total (valid in all threads)
__device__ int block_excl_scan(int* __restrict__ data, int n, int* sh /*[33]*/) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + 1023) / 1024;
  const int lo = min(t * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += data[i];
  int inc = sum;  // inclusive scan of the per-thread sums: shuffles within a warp, then over the 32 warp totals
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += v;
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = sh[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += v;
    }
    sh[lane] = w;                // inclusive totals of warps 0..lane
    if (lane == 31) sh[32] = w;  // grand total
  }
  __syncthreads();
  int run = inc - sum + (wid > 0 ? sh[wid - 1] : 0);
  const int total = sh[32];
  for (int i = lo; i < hi; ++i) {
    const int v = data[i];
    data[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// counts -> exclusive offsets for the four tables, one block per table; totals into counts[0..1]
__global__ void __launch_bounds__(1024)
prep_scan_kernel(int* __restrict__ blk_base, int nblk, int nb, int* __restrict__ counts, int n_px,
                 int* __restrict__ counts_host) {
  __shared__ int sh[33];
  const BlkMeta bm(blk_base, nblk, nb);
  volatile int* hp = counts_host;  // mapped pinned host memory: the host reads it after an event, no copy engine involved
  if (blockIdx.x == 0) {
    const int ta = block_excl_scan(bm.ref, nblk, sh);
    if (threadIdx.x == 0) {
      counts[0] = ta;
      counts[3] = n_px;
      if (hp != nullptr) hp[0] = ta, hp[2] = counts[2], hp[3] = n_px;
    }
  } else if (blockIdx.x == 1) {
    const int to = block_excl_scan(bm.ref + nblk, nblk, sh);
    if (threadIdx.x == 0) {
      counts[1] = to;
      if (hp != nullptr) hp[1] = to;
    }
  } else {  // class-major: sorted by label, then by pixel
    block_excl_scan(bm.cls + (size_t)(blockIdx.x - 2) * nb * nblk, nb * nblk, sh);
  }
  if (threadIdx.x == 0 && hp != nullptr) __threadfence_system();
}

// The tiles are written row by row by the pack kernels; what they do not write of the tiles the sweeps read is the
// padding of the LAST column tile (rows n_c % 128 .. 127): zero features / probabilities, label -1.  (Tiles beyond
// ceil(n_c / 128) are never read: the sweeps bound their tile loops by the counts.  Clearing the worst-case buffers
// with three memsets cost 17 us per step.)
__global__ void __launch_bounds__(128)
prep_pad_kernel(const int* __restrict__ counts, int kpad, long long max_tiles, __nv_bfloat16* __restrict__ feat_tiles,
                __nv_bfloat16* __restrict__ prob_tiles, int* __restrict__ lab_tiles) {
  const int n_c = counts[0] + counts[1];
  const long long T = n_c >> 7;
  const int r = threadIdx.x;
  if (T >= max_tiles || r < (n_c & 127)) return;
  if ((n_c & 127) == 0 && n_c > 0) return;  // the last tile is full (n_c == 0: tile 0 is all padding)
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int ch = 0; ch < 32; ++ch) *reinterpret_cast<uint4*>(feat_tiles + (((size_t)T * 32 + ch) * 128 + r) * 8) = z;
  for (int ch = 0; ch < kpad / 8; ++ch)
    *reinterpret_cast<uint4*>(prob_tiles + (((size_t)T * (kpad / 8) + ch) * 128 + r) * 8) = z;
  lab_tiles[T * 128 + r] = -1;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// tile layout: element (row r, feature k) of tile T at  ((T*(K/8) + k/8)*128 + r)*8 + k%8
__device__ __forceinline__ size_t tile_chunk_off(long long tile, int chunks_per_tile, int chunk, int r) {
  return (((size_t)tile * chunks_per_tile + chunk) * 128 + r) * 8;
}

// Pack kernel: block = 32 consecutive pixels, blockIdx.y = source (0: f_n -> anchors, 1: f_o -> pseudo columns).
// fp32 rows / labels go to the REFERENCE slot (pixel order); bf16 tiles go to the CLASS-SORTED slot.
// FT = element type of the NCHW head features: float, or __nv_bfloat16 when the head runs in bf16 (the
// features are consumed as they leave the head, no fp32 round trip; everything downstream is identical).
__device__ __forceinline__ float feat_ld(const float* p) { return *p; }
__device__ __forceinline__ float feat_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void feat_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void feat_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename FT>
__global__ void __launch_bounds__(128)
prep_pack_kernel(const FT* __restrict__ f_n, const FT* __restrict__ f_o, const int* __restrict__ px_meta,
                 int* __restrict__ blk_base, int nblk, int nb, const int* __restrict__ counts, int n_px, int hw,
                 float* __restrict__ anchor_f32, float* __restrict__ contrast_f32, void* __restrict__ la,
                 void* __restrict__ lc, int label_bytes, __nv_bfloat16* __restrict__ feat_tiles,
                 int* __restrict__ lab_tiles,
                 int* __restrict__ row_ref, float* __restrict__ inv_norm) {
  __shared__ float tile[256][33];
  __shared__ float ss[4][32];
  __shared__ int slot_s[32];
  __shared__ float inv_s[32];
  const BlkMeta bm(blk_base, nblk, nb);
  const int src = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  const int n_a = counts[0];
  int slot = -1, sslot = -1;  // reference slot, class-sorted slot (within the anchor / pseudo section)
  if (p < n_px) {
    const int f = px_meta[(size_t)PX_FLAGS * n_px + p];
    const int blk = p / kPrepBlock;
    if ((src == 0 && (f & 1)) || (src == 1 && (f & 2))) {
      const int m = px_meta[(size_t)PX_MIX * n_px + p];
      slot = bm.ref_at(src, blk) + px_meta[(size_t)(src == 0 ? PX_RANK_A : PX_RANK_O) * n_px + p];
      sslot = bm.cls_at(src, m, blk) + px_meta[(size_t)(src == 0 ? PX_SRANK_A : PX_SRANK_O) * n_px + p];
    }
  }
  if (!__syncthreads_or(slot >= 0)) return;
  const FT* f = src == 0 ? f_n : f_o;
  float part = 0.f;
  if (p < n_px) {
    const int b = p / hw, q = p - b * hw;
    const FT* fp = f + ((size_t)b * 256 + warp * 64) * hw + q;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float v = feat_ld(fp + (size_t)c * hw);
      tile[warp * 64 + c][lane] = v;
      part = fmaf(v, v, part);
    }
  } else {
    for (int c = 0; c < 64; ++c) tile[warp * 64 + c][lane] = 0.f;
  }
  ss[warp][lane] = part;
  __syncthreads();
  if (warp == 0) {
    const float tot = ss[0][lane] + ss[1][lane] + ss[2][lane] + ss[3][lane];
    const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize eps
    inv_s[lane] = inv;
    slot_s[lane] = slot;
    if (slot >= 0) {
      const int m = px_meta[(size_t)PX_MIX * n_px + p];
      // the tuple's label vectors in the element type the reference hands out (int8 on VOC / Cityscapes)
      const int cslot = src == 0 ? slot : n_a + slot;
      if (label_bytes == 1)
        static_cast<signed char*>(lc)[cslot] = (signed char)m;
      else
        static_cast<int*>(lc)[cslot] = m;
      lab_tiles[src == 0 ? sslot : n_a + sslot] = m;
      if (src == 0) {
        if (label_bytes == 1)
          static_cast<signed char*>(la)[slot] = (signed char)m;
        else
          static_cast<int*>(la)[slot] = m;
        row_ref[sslot] = slot;
        inv_norm[slot] = inv;
      }
    }
  }
  __syncthreads();
  // (a) fp32 rows in reference order: warp per pixel, lanes over channels (coalesced 128 B stores)
  for (int pi = warp; pi < 32; pi += 4) {
    const int s = slot_s[pi];
    if (s < 0) continue;
    const float inv = inv_s[pi];
    const size_t crow = (size_t)(src == 0 ? s : n_a + s) * 256;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      const float v = tile[c][pi] * inv;
      contrast_f32[crow + c] = v;
      if (src == 0) anchor_f32[(size_t)s * 256 + c] = v;
    }
  }
  // (b) bf16 tiles in class-sorted order: lane = pixel, each warp writes 8 of the 32 k-chunks (16 B per lane)
  if (slot >= 0) {
    const int cs = src == 0 ? sslot : n_a + sslot;
    const long long T = cs >> 7;
    const int r = cs & 127;
    const float inv = inv_s[lane];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int chunk = warp * 8 + k;
      uint4 o;
      o.x = pack_bf16x2(tile[chunk * 8 + 0][lane] * inv, tile[chunk * 8 + 1][lane] * inv);
      o.y = pack_bf16x2(tile[chunk * 8 + 2][lane] * inv, tile[chunk * 8 + 3][lane] * inv);
      o.z = pack_bf16x2(tile[chunk * 8 + 4][lane] * inv, tile[chunk * 8 + 5][lane] * inv);
      o.w = pack_bf16x2(tile[chunk * 8 + 6][lane] * inv, tile[chunk * 8 + 7][lane] * inv);
      *reinterpret_cast<uint4*>(feat_tiles + tile_chunk_off(T, 32, chunk, r)) = o;
    }
  }
}

// softmax(l_po) per pixel -> bf16 prob tiles at the pixel's (class-sorted) anchor slot and, if pseudo, pseudo slot
__global__ void __launch_bounds__(256)
prep_prob_kernel(const float* __restrict__ l_po, const int* __restrict__ px_meta, int* __restrict__ blk_base, int nblk,
                 int nb, const int* __restrict__ counts, int n_px, int hw, int C_old, int kpad,
                 __nv_bfloat16* __restrict__ prob_tiles) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const int f = px_meta[(size_t)PX_FLAGS * n_px + p];
  if (!(f & 1)) return;
  const BlkMeta bm(blk_base, nblk, nb);
  const int n_a = counts[0];
  const int blk = p / kPrepBlock;
  const int mlab = px_meta[(size_t)PX_MIX * n_px + p];
  const int b = p / hw, q = p - b * hw;
  const float* lp = l_po + (size_t)b * C_old * hw + q;
  float m = lp[0];
  for (int c = 1; c < C_old; ++c) m = fmaxf(m, lp[(size_t)c * hw]);
  float s = 0.f;
  for (int c = 0; c < C_old; ++c) s += __expf(lp[(size_t)c * hw] - m);
  const float inv = 1.f / s;
  const int sa = bm.cls_at(0, mlab, blk) + px_meta[(size_t)PX_SRANK_A * n_px + p];
  const int so = (f & 2) ? n_a + bm.cls_at(1, mlab, blk) + px_meta[(size_t)PX_SRANK_O * n_px + p] : -1;
  const int chunks = kpad / 8;
  for (int ch = 0; ch < chunks; ++ch) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ch * 8 + e;
      v[e] = c < C_old ? __expf(lp[(size_t)c * hw] - m) * inv : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(prob_tiles + tile_chunk_off(sa >> 7, chunks, ch, sa & 127)) = o;
    if (so >= 0) *reinterpret_cast<uint4*>(prob_tiles + tile_chunk_off(so >> 7, chunks, ch, so & 127)) = o;
  }
}

// label range of every tile (valid labels are >= 0): lets sweep 2 skip tiles that cannot hold a positive pair.
// Tiles [0, n_tiles) -> range (entries below *n_limit only, if given); a second set of n_tiles2 tiles of the same
// label array (the row tiles: entries below *n_limit2) -> range2, in the same launch.
__global__ void __launch_bounds__(256)
tile_range_kernel(const int* __restrict__ lab_tiles, long long n_tiles, const int* __restrict__ n_limit,
                  int* __restrict__ range /*[n_tiles][2]*/, long long n_tiles2, const int* __restrict__ n_limit2,
                  int* __restrict__ range2) {
  long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= n_tiles) {  // second set
    t -= n_tiles;
    if (t >= n_tiles2) return;
    n_limit = n_limit2, range = range2;
  }
  const long long limit = n_limit ? (long long)*n_limit : (1ll << 62);  // only entries [0, limit) count
  const int4 v = reinterpret_cast<const int4*>(lab_tiles + t * 128)[lane];
  int lo = 0x7fffffff, hi = -1;
  const int ls[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (ls[i] >= 0 && t * 128 + lane * 4 + i < limit) {
      lo = min(lo, ls[i]);
      hi = max(hi, ls[i]);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    range[2 * t] = lo;
    range[2 * t + 1] = hi;
  }
}

// adjoint of gather + normalize; block = 32 consecutive pixels, writes all 256 channels (zeros for non-anchors)
template <typename FT>
__global__ void __launch_bounds__(128)
prep_bwd_kernel(const float* __restrict__ g_anchor, const float* __restrict__ anchor_f32,
                const float* __restrict__ inv_norm, const int* __restrict__ px_meta, const int* __restrict__ blk_base,
                FT* __restrict__ df_n, int n_px, int hw) {
  __shared__ float tile[256][33];
  __shared__ int slot_s[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  int slot = -1;
  if (p < n_px && (px_meta[(size_t)PX_FLAGS * n_px + p] & 1))
    slot = blk_base[p / kPrepBlock] + px_meta[(size_t)PX_RANK_A * n_px + p];  // ref_off[0][blk]
  if (warp == 0) slot_s[lane] = slot;
  __syncthreads();
  // two pixels per pass: the rows of both are requested before either dot product is reduced (a warp working on one
  // pixel at a time had 2 KB in flight and a shuffle reduction between loads: latency-bound at 3.7 TB/s)
  for (int pi = warp; pi < 32; pi += 8) {
    const int s0 = slot_s[pi], s1 = slot_s[pi + 4];
    float g[2][8], a[2][8], dot[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int s = j ? s1 : s0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[j][k] = s >= 0 ? g_anchor[(size_t)s * 256 + lane + 32 * k] : 0.f;
        a[j][k] = s >= 0 ? anchor_f32[(size_t)s * 256 + lane + 32 * k] : 0.f;
      }
    }
    const float inv0 = s0 >= 0 ? inv_norm[s0] : 0.f, inv1 = s1 >= 0 ? inv_norm[s1] : 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dot[j] = fmaf(g[j][k], a[j][k], dot[j]);
      dot[j] = warp_sum(dot[j]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      tile[lane + 32 * k][pi] = (g[0][k] - dot[0] * a[0][k]) * inv0;       // non-anchor pixel: zeros
      tile[lane + 32 * k][pi + 4] = (g[1][k] - dot[1] * a[1][k]) * inv1;
    }
  }
  __syncthreads();
  if (p < n_px) {
    const int b = p / hw, q = p - b * hw;
    FT* dp = df_n + ((size_t)b * 256 + warp * 64) * hw + q;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) feat_st(dp + (size_t)c * hw, tile[warp * 64 + c][lane]);
  }
}

// Pixel-to-pixel branches of pre_contrastive_pixel (utils/loss.py:273-289): EVERY pixel becomes a unit-norm row
// [n_px, 256] in (b, y, x) order.  block = 32 consecutive pixels, coalesced NCHW reads -> shared tile -> row writes.
__global__ void __launch_bounds__(128)
rows_normalize_kernel(const float* __restrict__ f, float* __restrict__ rows, float* __restrict__ inv_norm, int n_px,
                      int hw) {
  __shared__ float tile[256][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  if (p < n_px) {
    const int b = p / hw, q = p - b * hw;
    const float* src = f + ((size_t)b * 256 + warp * 64) * hw + q;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) tile[warp * 64 + c][lane] = src[(size_t)c * hw];
  }
  __syncthreads();
  for (int pi = warp; pi < 32; pi += 4) {
    const int pp = blockIdx.x * 32 + pi;
    if (pp >= n_px) break;
    float v[8], ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = tile[lane + 32 * k][pi];
      ss = fmaf(v[k], v[k], ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
#pragma unroll
    for (int k = 0; k < 8; ++k) rows[(size_t)pp * 256 + lane + 32 * k] = v[k] * inv;
    if (lane == 0) inv_norm[pp] = inv;
  }
}

// adjoint: df[b,:,y,x] = (g - (g.a) a) * inv_norm for every pixel
__global__ void __launch_bounds__(128)
rows_normalize_bwd_kernel(const float* __restrict__ g_rows, const float* __restrict__ rows,
                          const float* __restrict__ inv_norm, float* __restrict__ df, int n_px, int hw) {
  __shared__ float tile[256][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pi = warp; pi < 32; pi += 4) {
    const int pp = blockIdx.x * 32 + pi;
    if (pp >= n_px) break;
    float g[8], a[8], dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g[k] = g_rows[(size_t)pp * 256 + lane + 32 * k];
      a[k] = rows[(size_t)pp * 256 + lane + 32 * k];
      dot = fmaf(g[k], a[k], dot);
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[pp];
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[lane + 32 * k][pi] = (g[k] - dot * a[k]) * inv;
  }
  __syncthreads();
  const int p = blockIdx.x * 32 + lane;
  if (p < n_px) {
    const int b = p / hw, q = p - b * hw;
    float* dp = df + ((size_t)b * 256 + warp * 64) * hw + q;
#pragma unroll 8
    for (int c = 0; c < 64; ++c) dp[(size_t)c * hw] = tile[warp * 64 + c][lane];
  }
}

// compat path: caller-supplied fp32 rows [n,256] (+labels) -> bf16 tiles.  warp per row.
__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ rows, const int* __restrict__ labels, long long n,
                 __nv_bfloat16* __restrict__ feat_tiles, int* __restrict__ lab_tiles) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* src = reinterpret_cast<const float4*>(rows + row * 256 + lane * 8);
  const float4 a = src[0], b = src[1];
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y);
  o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(b.x, b.y);
  o.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(feat_tiles + tile_chunk_off(row >> 7, 32, lane, (int)(row & 127))) = o;
  if (lane == 0 && lab_tiles != nullptr) lab_tiles[row] = labels[row];
}

}  // namespace ucd

using namespace ucd;

extern "C" int64_t ucd_con_max_tiles(int64_t n_px) { return (2 * n_px + 127) / 128 + 1; }
extern "C" int ucd_con_prob_kpad(int C_old) { return (C_old + 15) / 16 * 16; }
extern "C" int ucd_con_num_bins(int max_label, int C_old) { return (max_label > C_old - 1 ? max_label : C_old - 1) + 1; }
extern "C" int64_t ucd_con_px_meta_ints(int64_t n_px) { return (int64_t)PX_PLANES * n_px; }
extern "C" int64_t ucd_con_blk_meta_ints(int64_t n_px, int nb) {
  const int64_t nblk = (n_px + kPrepBlock - 1) / kPrepBlock;
  return 2 * nblk * (1 + (int64_t)nb);
}

extern "C" int ucd_con_prep_labels(const int64_t* labels, const float* l_po, int B, int C_old, int h, int w, int H,
                                   int W, int max_label, int32_t* px_meta, int32_t* blk_meta, int32_t* counts,
                                   int32_t* counts_host, void* stream) {
  UCD_CHECK_ARG(labels && l_po && px_meta && blk_meta && counts, "ucd_con_prep_labels: null pointer");
  UCD_CHECK_ARG(B > 0 && C_old > 0 && h > 0 && w > 0 && H > 0 && W > 0, "ucd_con_prep_labels: bad shape");
  UCD_CHECK_ARG((long long)B * h * w < (1ll << 30), "ucd_con_prep_labels: too many pixels");
  const int nb = ucd_con_num_bins(max_label, C_old);
  UCD_CHECK_ARG(max_label >= 0 && nb <= 1024, "ucd_con_prep_labels: %d label bins not supported (max 1024)", nb);
  cudaStream_t st = (cudaStream_t)stream;
  const int n_px = B * h * w;
  const int nblk = (n_px + kPrepBlock - 1) / kPrepBlock;
  cudaError_t e = cudaMemsetAsync(counts, 0x7f, 4 * sizeof(int32_t), st);  // min_new starts at 0x7f7f7f7f
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(counts)");
  const size_t smem = (size_t)2 * 8 * nb * sizeof(int);
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(prep_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(prep_labels_kernel)");
  }
  prep_labels_kernel<<<nblk, kPrepBlock, smem, st>>>((const long long*)labels, l_po, B, C_old, h, w, H, W, max_label,
                                                     (float)H / (float)h, (float)W / (float)w, px_meta, blk_meta, nb,
                                                     counts);
  UCD_CHECK_LAUNCH("prep_labels_kernel");
  prep_scan_kernel<<<4, 1024, 0, st>>>(blk_meta, nblk, nb, counts, n_px, counts_host);
  UCD_CHECK_LAUNCH("prep_scan_kernel");
  return UCD_OK;
}

extern "C" int ucd_con_tile_ranges(const int32_t* lab_tiles, int64_t n_tiles, const int32_t* n_limit,
                                   int32_t* tile_range, void* stream) {
  UCD_CHECK_ARG(lab_tiles && tile_range && n_tiles >= 0, "ucd_con_tile_ranges: bad argument");
  UCD_CHECK_ARG(aligned16(lab_tiles), "ucd_con_tile_ranges: lab_tiles must be 16 B aligned");
  if (n_tiles == 0) return UCD_OK;
  tile_range_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, (cudaStream_t)stream>>>(lab_tiles, n_tiles, n_limit,
                                                                                      tile_range, 0, nullptr, nullptr);
  UCD_CHECK_LAUNCH("tile_range_kernel");
  return UCD_OK;
}

template <typename FT>
static int prep_pack_impl(const FT* f_n, const FT* f_o, const float* l_po, const int32_t* px_meta,
                          int32_t* blk_meta, const int32_t* counts, int B, int C_old, int h, int w,
                          int max_label, float* anchor_f32, float* contrast_f32, void* la, void* lc,
                          int label_bytes, void* feat_tiles, void* prob_tiles, int32_t* lab_tiles, int32_t* tile_range,
                          int32_t* row_range, int32_t* row_ref, float* inv_norm, int64_t max_tiles, void* stream) {
  UCD_CHECK_ARG(f_n && f_o && l_po && px_meta && blk_meta && counts && anchor_f32 && contrast_f32 && la && lc &&
                    feat_tiles && prob_tiles && lab_tiles && tile_range && row_range && row_ref && inv_norm,
                "ucd_con_prep_pack: null pointer");
  UCD_CHECK_ARG(aligned16(feat_tiles) && aligned16(prob_tiles) && aligned16(lab_tiles),
                "ucd_con_prep_pack: tiles must be 16 B aligned");
  // la / lc hold GT labels (<= max_label) AND pseudo labels (old-model argmax, <= C_old - 1): both must fit
  UCD_CHECK_ARG(label_bytes == 4 || (label_bytes == 1 && max_label <= 127 && C_old - 1 <= 127),
                "ucd_con_prep_pack: label_bytes must be 4, or 1 when max_label <= 127 and C_old <= 128");
  const int n_px = B * h * w;
  UCD_CHECK_ARG(max_tiles >= ucd_con_max_tiles(n_px), "ucd_con_prep_pack: max_tiles too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int kpad = ucd_con_prob_kpad(C_old);
  const int nb = ucd_con_num_bins(max_label, C_old);
  const int nblk = (n_px + kPrepBlock - 1) / kPrepBlock;
  prep_pad_kernel<<<1, 128, 0, st>>>(counts, kpad, max_tiles, (__nv_bfloat16*)feat_tiles, (__nv_bfloat16*)prob_tiles,
                                     lab_tiles);
  UCD_CHECK_LAUNCH("prep_pad_kernel");
  dim3 grid((n_px + 31) / 32, 2);
  prep_pack_kernel<FT><<<grid, 128, 0, st>>>(f_n, f_o, px_meta, blk_meta, nblk, nb, counts, n_px, h * w, anchor_f32,
                                             contrast_f32, la, lc, label_bytes, (__nv_bfloat16*)feat_tiles, lab_tiles,
                                             row_ref, inv_norm);
  UCD_CHECK_LAUNCH("prep_pack_kernel");
  prep_prob_kernel<<<(n_px + 255) / 256, 256, 0, st>>>(l_po, px_meta, blk_meta, nblk, nb, counts, n_px, h * w, C_old,
                                                       kpad, (__nv_bfloat16*)prob_tiles);
  UCD_CHECK_LAUNCH("prep_prob_kernel");
  // column tiles, and in the same launch the row tiles.  Rows are the first N_a (= counts[0]) columns: the last anchor
  // tile also holds pseudo columns, which must not widen the ROW range (it would make every column tile "active"
  // for that row block in sweep 2)
  const long long row_tiles = (n_px + 127) / 128;
  tile_range_kernel<<<(unsigned)((max_tiles + row_tiles + 7) / 8), 256, 0, st>>>(lab_tiles, max_tiles, nullptr,
                                                                                 tile_range, row_tiles, counts, row_range);
  UCD_CHECK_LAUNCH("tile_range_kernel");
  return UCD_OK;
}

extern "C" int ucd_con_prep_pack(const float* f_n, const float* f_o, const float* l_po, const int32_t* px_meta,
                                 int32_t* blk_meta, const int32_t* counts, int B, int C_old, int h, int w,
                                 int max_label, float* anchor_f32, float* contrast_f32, void* la, void* lc,
                                 int label_bytes, void* feat_tiles, void* prob_tiles, int32_t* lab_tiles, int32_t* tile_range,
                                 int32_t* row_range, int32_t* row_ref, float* inv_norm, int64_t max_tiles, void* stream) {
  return prep_pack_impl<float>(f_n, f_o, l_po, px_meta, blk_meta, counts, B, C_old, h, w, max_label, anchor_f32,
                               contrast_f32, la, lc, label_bytes, feat_tiles, prob_tiles, lab_tiles, tile_range,
                               row_range, row_ref, inv_norm, max_tiles, stream);
}

extern "C" int ucd_con_prep_pack_bf16(const void* f_n, const void* f_o, const float* l_po, const int32_t* px_meta,
                                      int32_t* blk_meta, const int32_t* counts, int B, int C_old, int h, int w,
                                      int max_label, float* anchor_f32, float* contrast_f32, void* la, void* lc,
                                      int label_bytes, void* feat_tiles, void* prob_tiles, int32_t* lab_tiles,
                                      int32_t* tile_range, int32_t* row_range, int32_t* row_ref, float* inv_norm,
                                      int64_t max_tiles, void* stream) {
  return prep_pack_impl<__nv_bfloat16>((const __nv_bfloat16*)f_n, (const __nv_bfloat16*)f_o, l_po, px_meta, blk_meta,
                                       counts, B, C_old, h, w, max_label, anchor_f32, contrast_f32, la, lc, label_bytes,
                                       feat_tiles, prob_tiles, lab_tiles, tile_range, row_range, row_ref, inv_norm,
                                       max_tiles, stream);
}

extern "C" int ucd_con_prep_bwd(const float* g_anchor, const float* anchor_f32, const float* inv_norm,
                                const int32_t* px_meta, const int32_t* blk_meta, float* df_n, int B, int h, int w,
                                void* stream) {
  UCD_CHECK_ARG(g_anchor && anchor_f32 && inv_norm && px_meta && blk_meta && df_n, "ucd_con_prep_bwd: null pointer");
  const int n_px = B * h * w;
  prep_bwd_kernel<float><<<(n_px + 31) / 32, 128, 0, (cudaStream_t)stream>>>(g_anchor, anchor_f32, inv_norm, px_meta,
                                                                              blk_meta, df_n, n_px, h * w);
  UCD_CHECK_LAUNCH("prep_bwd_kernel");
  return UCD_OK;
}

extern "C" int ucd_con_prep_bwd_bf16(const float* g_anchor, const float* anchor_f32, const float* inv_norm,
                                     const int32_t* px_meta, const int32_t* blk_meta, void* df_n, int B, int h, int w,
                                     void* stream) {
  UCD_CHECK_ARG(g_anchor && anchor_f32 && inv_norm && px_meta && blk_meta && df_n, "ucd_con_prep_bwd_bf16: null pointer");
  const int n_px = B * h * w;
  prep_bwd_kernel<__nv_bfloat16><<<(n_px + 31) / 32, 128, 0, (cudaStream_t)stream>>>(
      g_anchor, anchor_f32, inv_norm, px_meta, blk_meta, (__nv_bfloat16*)df_n, n_px, h * w);
  UCD_CHECK_LAUNCH("prep_bwd_kernel");
  return UCD_OK;
}

extern "C" int ucd_con_pack_rows(const float* rows, const int32_t* labels, int64_t n, void* feat_tiles,
                                 int32_t* lab_tiles, int64_t max_tiles, void* stream) {
  UCD_CHECK_ARG(rows && feat_tiles, "ucd_con_pack_rows: null pointer");
  UCD_CHECK_ARG(labels || !lab_tiles, "ucd_con_pack_rows: lab_tiles without labels");
  UCD_CHECK_ARG(aligned16(rows) && aligned16(feat_tiles), "ucd_con_pack_rows: 16 B alignment required");
  UCD_CHECK_ARG(n >= 0 && max_tiles * 128 >= n, "ucd_con_pack_rows: max_tiles too small");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(feat_tiles, 0, (size_t)max_tiles * 128 * 256 * 2, st);
  if (e == cudaSuccess && lab_tiles) e = cudaMemsetAsync(lab_tiles, 0xff, (size_t)max_tiles * 128 * 4, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(pack_rows)");
  if (n > 0) {
    pack_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(rows, labels, n, (__nv_bfloat16*)feat_tiles, lab_tiles);
    UCD_CHECK_LAUNCH("pack_rows_kernel");
  }
  return UCD_OK;
}

extern "C" int ucd_rows_normalize_fwd(const float* f, float* rows, float* inv_norm, int B, int h, int w, void* stream) {
  UCD_CHECK_ARG(f && rows && inv_norm, "ucd_rows_normalize_fwd: null pointer");
  UCD_CHECK_ARG(B > 0 && h > 0 && w > 0 && (long long)B * h * w < (1ll << 30), "ucd_rows_normalize_fwd: bad shape");
  const int n_px = B * h * w;
  rows_normalize_kernel<<<(n_px + 31) / 32, 128, 0, (cudaStream_t)stream>>>(f, rows, inv_norm, n_px, h * w);
  UCD_CHECK_LAUNCH("rows_normalize_kernel");
  return UCD_OK;
}

extern "C" int ucd_rows_normalize_bwd(const float* g_rows, const float* rows, const float* inv_norm, float* df, int B,
                                      int h, int w, void* stream) {
  UCD_CHECK_ARG(g_rows && rows && inv_norm && df, "ucd_rows_normalize_bwd: null pointer");
  UCD_CHECK_ARG(B > 0 && h > 0 && w > 0 && (long long)B * h * w < (1ll << 30), "ucd_rows_normalize_bwd: bad shape");
  const int n_px = B * h * w;
  rows_normalize_bwd_kernel<<<(n_px + 31) / 32, 128, 0, (cudaStream_t)stream>>>(g_rows, rows, inv_norm, df, n_px, h * w);
  UCD_CHECK_LAUNCH("rows_normalize_bwd_kernel");
  return UCD_OK;
}
