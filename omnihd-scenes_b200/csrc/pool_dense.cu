// Fused ("dense") bev_pool_v2 kernels for sm_100a: the forms the view-transform shim and
// bev_pool_v2() use. Same arithmetic as the reference-contract kernels in pool.cu, organised for
// the machine:
//
//   voxel_table_kernel     vox_pt[v] = number of sorted points whose voxel rank is < v (a lower-bound
//                          table written by gap filling, no scan, no atomics). It replaces the
//                          interval arrays on the device: voxel v owns points [vox_pt[v], vox_pt[v+1]).
//   pool_fwd_tile_kernel   one CTA per 32(x) x 4(rows) voxel tile. Warps walk the tile's points FLAT
//                          in sorted order (across interval boundaries) with 8 feature rows in flight
//                          and the per-point scalars prefetched two batches ahead; a running sum is
//                          flushed into a shared-memory tile whenever the voxel changes. Empty voxels
//                          are the tile's initial zeros, and the tile is written out with 128-byte
//                          rows in [B,C,Z,Y,X] — memset + kernel + permute of the reference
//                          (bev_pool.py:27,29,91) in one pass. The 2-D tile keeps the ~4x re-use of
//                          feature rows inside one SM's L1 (SURVEY.md §7 hard part 1).
//   pool_bwd_block_kernel  sort-free backward. One CTA per 8(w) x 4(h) pixel block: point_rank / depth
//                          columns are staged in shared memory with sector-sized coalesced loads, each
//                          warp walks the kept depth bins of a pixel 8 at a time (8 out_grad rows in
//                          flight), dot products are finished with a reduce-scatter butterfly, and both
//                          gradients are written densely (zeros included) with coalesced rows.
//                          Replaces argsort + where + 2x new_zeros + kernel (bev_pool.py:47-70).
//
// No atomics anywhere; every sum has a fixed order, so results are run-to-run deterministic.
#include "common.cuh"

namespace bevpool {

// ------------------------------------------------------------------------------------------ voxel table
__global__ void __launch_bounds__(256)
voxel_table_kernel(const int* __restrict__ keys, int64_t n_points, const int* __restrict__ counts_dev,
                   int64_t n_voxels_total, int* __restrict__ vox_pt) {
  if (counts_dev) n_points = counts_dev[0];
  const int lane = lane_id();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // one virtual point at index n_points closes the table (vox_pt[v] = n_points for v > last key)
  const int64_t n_iter = (n_points + 1 + nthreads - 1) / nthreads;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t p = it * nthreads + tid;
    int64_t lo = 1, hi = 0;  // empty
    if (p <= n_points) {
      const int64_t k = (p < n_points) ? (int64_t)keys[p] : n_voxels_total;
      const int64_t kp = (p > 0) ? (int64_t)keys[p - 1] : -1;
      lo = kp + 1;
      hi = k;  // voxels (kp, k] start at point p
    }
    const int64_t gap = hi - lo + 1;
    if (gap > 0 && gap <= 8)
      for (int64_t v = lo; v <= hi; ++v) vox_pt[v] = (int)p;
    // long gaps (sparse grids) are filled by the whole warp
    unsigned big = __ballot_sync(kFullMask, gap > 8);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const int64_t l = __shfl_sync(kFullMask, lo, src), h = __shfl_sync(kFullMask, hi, src);
      const int val = (int)__shfl_sync(kFullMask, p, src);
      for (int64_t v = l + lane; v <= h; v += 32) vox_pt[v] = val;
    }
  }
}

// ------------------------------------------------------------------------------------------ forward
constexpr int kTileX = 32;
constexpr int kTileRows = 4;
constexpr int kTileCols = kTileX * kTileRows;  // 128 voxel columns
constexpr int kTileStride = kTileCols + 1;     // +1: conflict-free row reads at write-out
constexpr int kFwdWarps = 8;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kGroup = 8;          // voxels per work item
constexpr int kLongVoxel = 512;    // voxels with more points are summed by the whole CTA

struct FwdParams {
  int c;                 // channels
  int x;                 // X
  int64_t rows;          // Z*Y rows per frame
  int64_t frames;
  int dhw, hw;           // to derive ranks_feat from ranks_depth when rf == nullptr
};

// Per-point scalars of one batch: one lane per point.
template <typename T>
__device__ __forceinline__ void load_point(const T* __restrict__ depth, const int* __restrict__ rf, int rd_val,
                                           int64_t p, const FwdParams& prm, int c4, int& off4, float& d) {
  // feature-row offset in float4 units; derived from ranks_depth when the caller has no ranks_feat
  const int f = rf ? ldg_stream_i32(rf + p) : (rd_val / prm.dhw) * prm.hw + rd_val % prm.hw;
  off4 = f * c4;
  d = Vec4<T>::load1(depth, rd_val);
}

// Sum points [p, pe) of tile row `row_pt` (shared-memory copy of vox_pt for that row, relative voxel
// index) starting in voxel `cur`; flush the running sum into column col0+cur of the tile each time the
// voxel changes. BOUNDS=false: single voxel slice, nothing is flushed, the sum is returned.
template <typename T, bool BOUNDS>
__device__ __forceinline__ float4 flat_sum(const T* __restrict__ depth, const T* __restrict__ feat,
                                           const int* __restrict__ rd, const int* __restrict__ rf, int64_t p,
                                           int64_t pe, const FwdParams& prm, int c4, int cb, bool act,
                                           const int* row_pt, int cur, float* tile_col0) {
  const int lane = lane_id();
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = zero;
  if (p >= pe) return acc;
  const float4* __restrict__ feat4 = reinterpret_cast<const float4*>(feat);
  int64_t next_b = BOUNDS ? (int64_t)row_pt[cur + 1] : pe;

  // software pipeline: ranks_depth two batches ahead, (offset, depth) one batch ahead
  int rd1 = (p + lane < pe) ? ldg_stream_i32(rd + p + lane) : 0;
  int rd2 = (p + 32 + lane < pe) ? ldg_stream_i32(rd + p + 32 + lane) : 0;
  int off_n = 0;
  float d_n = 0.f;
  if (p + lane < pe) load_point<T>(depth, rf, rd1, p + lane, prm, c4, off_n, d_n);

  for (int64_t q = p; q < pe; q += 32) {
    const int my_off = off_n;
    const float my_d = d_n;
    // issue the prefetches for the following batches before touching this one
    const int rd_next = rd2;
    rd2 = (q + 64 + lane < pe) ? ldg_stream_i32(rd + q + 64 + lane) : 0;
    off_n = 0;
    d_n = 0.f;
    if (q + 32 + lane < pe) load_point<T>(depth, rf, rd_next, q + 32 + lane, prm, c4, off_n, d_n);

    const int n = (int)min((int64_t)32, pe - q);
    for (int i0 = 0; i0 < n; i0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int off = __shfl_sync(kFullMask, my_off, i0 + u);
        v[u] = (act && i0 + u < n) ? (sizeof(T) == 4 ? __ldg(feat4 + off + cb + lane)
                                                     : Vec4<T>::load(feat, ((int64_t)off + cb + lane) * 4))
                                   : zero;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
        if (BOUNDS) {
          // warp-uniform: leave the current voxel (and skip empty ones) before consuming point q+i0+u
          while (i0 + u < n && q + i0 + u == next_b) {
            if (act) {
              float* col = tile_col0 + cur;
              col[(4 * lane + 0) * kTileStride] = acc.x;
              col[(4 * lane + 1) * kTileStride] = acc.y;
              col[(4 * lane + 2) * kTileStride] = acc.z;
              col[(4 * lane + 3) * kTileStride] = acc.w;
            }
            acc = zero;
            do { ++cur; next_b = row_pt[cur + 1]; } while (next_b == q + i0 + u);
          }
        }
        acc = fma4(v[u], dd, acc);  // masked points carry v == 0
      }
    }
  }
  if (BOUNDS && act) {
    float* col = tile_col0 + cur;
    col[(4 * lane + 0) * kTileStride] = acc.x;
    col[(4 * lane + 1) * kTileStride] = acc.y;
    col[(4 * lane + 2) * kTileStride] = acc.z;
    col[(4 * lane + 3) * kTileStride] = acc.w;
  }
  return acc;
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(kFwdThreads)
pool_fwd_tile_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out,
                     const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ vox_pt,
                     FwdParams prm, int tiles_x, int64_t tiles_per_frame) {
  extern __shared__ float tile[];                       // [cw][kTileStride]
  __shared__ int s_pt[kTileRows][kTileX + 1];           // vox_pt of the tile rows (+ closing entry)
  __shared__ float4 s_partial[kFwdWarps][32];
  __shared__ int s_next;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int c4 = prm.c >> 2;

  const int64_t t = blockIdx.x;
  const int64_t frame = t / tiles_per_frame;
  const int64_t tr = (t % tiles_per_frame) / tiles_x;   // tile row
  const int tx = (int)(t % tiles_x);
  const int x0 = tx * kTileX;
  const int w = min(kTileX, prm.x - x0);                                  // valid columns per row
  const int64_t row0 = tr * kTileRows;
  const int nrows = (int)min((int64_t)kTileRows, prm.rows - row0);
  const int64_t vpf = prm.rows * prm.x;
  const int64_t rank0 = frame * vpf + row0 * prm.x + x0;                  // rank of (row0, x0)

  for (int i = threadIdx.x; i < kTileRows * (kTileX + 1); i += kFwdThreads) {
    const int r = i / (kTileX + 1), xx = i % (kTileX + 1);
    int v = 0;
    if (r < nrows) v = __ldg(vox_pt + rank0 + (int64_t)r * prm.x + min(xx, w));
    s_pt[r][xx] = v;
  }

  for (int cb = 0; cb < c4; cb += 32) {   // one sweep when C <= 128
    const bool act = cb + lane < c4;
    const int cw = min(prm.c - 4 * cb, 128);
    for (int i = threadIdx.x; i < cw * kTileStride; i += kFwdThreads) tile[i] = 0.f;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();

    // ---- phase 1: 8-voxel groups, grabbed dynamically; long voxels are left for phase 2
    bool any_long = false;
    const int groups_per_row = (w + kGroup - 1) / kGroup;
    const int n_items = nrows * groups_per_row;
    for (;;) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&s_next, 1);
      item = __shfl_sync(kFullMask, item, 0);
      if (item >= n_items) break;
      const int r = item / groups_per_row, g = item % groups_per_row;
      const int xb = g * kGroup, xe = min(w, xb + kGroup);
      const int* row_pt = s_pt[r];
      int xv = xb;
      while (xv < xe) {
        if (row_pt[xv + 1] - row_pt[xv] > kLongVoxel) { any_long = true; ++xv; continue; }
        const int xs = xv;
        while (xv < xe && row_pt[xv + 1] - row_pt[xv] <= kLongVoxel) ++xv;
        // skip leading empty voxels so `cur` always names the voxel of the first point
        int cur = xs;
        while (cur < xv && row_pt[cur + 1] == row_pt[xs]) ++cur;
        if (cur < xv)
          flat_sum<T, true>(depth, feat, rd, rf, row_pt[xs], row_pt[xv], prm, c4, cb, act, row_pt, cur,
                            tile + r * kTileX);
      }
    }
    // ---- phase 2: long voxels, all warps split the point range; partials combined in warp order
    if (__syncthreads_or(any_long)) {
      for (int r = 0; r < nrows; ++r)
        for (int xv = 0; xv < w; ++xv) {
          const int s = s_pt[r][xv], len = s_pt[r][xv + 1] - s;
          if (len <= kLongVoxel) continue;
          const int per = ((len + kFwdWarps - 1) / kFwdWarps + 31) & ~31;
          const int b = min(len, warp * per), e = min(len, b + per);
          s_partial[warp][lane] = flat_sum<T, false>(depth, feat, rd, rf, (int64_t)s + b, (int64_t)s + e, prm, c4, cb,
                                                     act, nullptr, 0, nullptr);
          __syncthreads();
          if (warp == 0 && act) {
            float4 acc = s_partial[0][lane];
#pragma unroll
            for (int k = 1; k < kFwdWarps; ++k) {
              const float4 q = s_partial[k][lane];
              acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            float* col = tile + r * kTileX + xv;
            col[(4 * lane + 0) * kTileStride] = acc.x;
            col[(4 * lane + 1) * kTileStride] = acc.y;
            col[(4 * lane + 2) * kTileStride] = acc.z;
            col[(4 * lane + 3) * kTileStride] = acc.w;
          }
          __syncthreads();
        }
    }
    __syncthreads();

    // ---- write-out
    if (LAYOUT == BEVPOOL_LAYOUT_BCZYX) {
      // out[((frame*C + ch)*rows + row0 + r)*X + x0 + lane]: one 128-byte row per warp store
      const int64_t base = frame * prm.c * vpf + row0 * prm.x + x0 + lane;
      for (int i = warp; i < cw * nrows; i += kFwdWarps) {
        const int cc = i / nrows, r = i % nrows;
        if (lane < w)
          Vec4<T>::store1s(out, base + (int64_t)(4 * cb + cc) * vpf + (int64_t)r * prm.x,
                           tile[cc * kTileStride + r * kTileX + lane]);
      }
    } else {
      // out[(rank)*C + ch]: consecutive threads take consecutive channels of one voxel
      for (int i = threadIdx.x; i < nrows * w * cw; i += kFwdThreads) {
        const int cc = i % cw, col = i / cw;
        const int r = col / w, xx = col % w;
        Vec4<T>::store1s(out, (rank0 + (int64_t)r * prm.x + xx) * prm.c + 4 * cb + cc,
                         tile[cc * kTileStride + r * kTileX + xx]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ backward
constexpr int kPixW = 8, kPixH = 4, kPixBlock = kPixW * kPixH;   // 32 pixels per CTA
constexpr int kBwdWarps = 8;
constexpr int kBwdThreads = kBwdWarps * 32;

struct BwdParams {
  int c, d, h, w;        // channels, depth bins, feature map H x W
  int bn;                // B*N camera images
  int blocks_w, blocks_h;
  int feat_grad_nchw;    // 0: feat_grad is [BN,H,W,C] (bev_pool_v2 contract); 1: [BN,C,H,W]
};

template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
pool_bwd_block_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                      const int* __restrict__ point_rank, BwdParams prm, T* __restrict__ depth_grad,
                      T* __restrict__ feat_grad) {
  extern __shared__ unsigned char smem_raw[];
  int* s_rank = reinterpret_cast<int*>(smem_raw);                                  // [d][32]
  float* s_depth = reinterpret_cast<float*>(smem_raw) + (size_t)prm.d * kPixBlock;  // [d][32]
  float* s_dg = s_depth + (size_t)prm.d * kPixBlock;                                // [d][32]
  float* s_fg = s_dg + (size_t)prm.d * kPixBlock;                                   // [cw][33] (NCHW output only)
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int c4 = prm.c >> 2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  const int blk = blockIdx.x;
  const int bn = blk / (prm.blocks_w * prm.blocks_h);
  const int bh = (blk / prm.blocks_w) % prm.blocks_h, bw = blk % prm.blocks_w;
  const int h0 = bh * kPixH, w0 = bw * kPixW;
  const int64_t hw = (int64_t)prm.h * prm.w;
  const int64_t img_base = (int64_t)bn * prm.d * hw;   // + d*hw + h*W + w

  // ---- stage point_rank / depth of the block: 8 consecutive w = one 32-byte sector per (d, h)
  for (int i = threadIdx.x; i < prm.d * kPixBlock; i += kBwdThreads) {
    const int dd = i / kPixBlock, px = i % kPixBlock;
    const int hh = h0 + px / kPixW, ww = w0 + px % kPixW;
    int r = -1;
    float dv = 0.f;
    if (hh < prm.h && ww < prm.w) {
      const int64_t o = img_base + (int64_t)dd * hw + (int64_t)hh * prm.w + ww;
      r = ldg_stream_i32(point_rank + o);
      if (r >= 0) dv = Vec4<T>::load1(depth, o);
    }
    s_rank[i] = r;
    s_depth[i] = dv;
    s_dg[i] = 0.f;
  }
  __syncthreads();

  for (int cb = 0; cb < c4; cb += 32) {
    const bool act = cb + lane < c4;
    const int cw = min(prm.c - 4 * cb, 128);
    for (int px = warp; px < kPixBlock; px += kBwdWarps) {
      const int hh = h0 + px / kPixW, ww = w0 + px % kPixW;
      if (hh >= prm.h || ww >= prm.w) continue;   // warp-uniform
      const int64_t pix = (int64_t)bn * hw + (int64_t)hh * prm.w + ww;
      const float4 fv = act ? Vec4<T>::load(feat, pix * prm.c + 4 * (cb + lane)) : zero;
      float4 fg = zero;
      for (int d0 = 0; d0 < prm.d; d0 += 32) {
        const int my_r = (d0 + lane < prm.d) ? s_rank[(d0 + lane) * kPixBlock + px] : -1;
        unsigned live = __ballot_sync(kFullMask, my_r >= 0);
        while (live) {
          // next (up to) 8 kept depth bins of this pixel
          int dsel[8];
          float4 g[8];
          float pr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            dsel[u] = live ? (__ffs(live) - 1) : -1;
            live &= live - 1;   // 0 stays 0
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = __shfl_sync(kFullMask, my_r, dsel[u] < 0 ? 0 : dsel[u]);
            g[u] = (act && dsel[u] >= 0) ? Vec4<T>::load(og, (int64_t)r * prm.c + 4 * (cb + lane)) : zero;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float dv = dsel[u] >= 0 ? s_depth[(d0 + dsel[u]) * kPixBlock + px] : 0.f;
            fg = fma4(g[u], dv, fg);
            pr[u] = dot4(g[u], fv, 0.f);
          }
          // reduce-scatter over lane bits 2,1,0, then finish over bits 3,4: lane l ends up with the
          // complete dot product of point (l & 7)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float mine = (lane & 4) ? pr[u + 4] : pr[u];
            const float send = (lane & 4) ? pr[u] : pr[u + 4];
            pr[u] = mine + __shfl_xor_sync(kFullMask, send, 4);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float mine = (lane & 2) ? pr[u + 2] : pr[u];
            const float send = (lane & 2) ? pr[u] : pr[u + 2];
            pr[u] = mine + __shfl_xor_sync(kFullMask, send, 2);
          }
          {
            const float mine = (lane & 1) ? pr[1] : pr[0];
            const float send = (lane & 1) ? pr[0] : pr[1];
            pr[0] = mine + __shfl_xor_sync(kFullMask, send, 1);
          }
          pr[0] += __shfl_xor_sync(kFullMask, pr[0], 8);
          pr[0] += __shfl_xor_sync(kFullMask, pr[0], 16);
          // lane u (< 8) publishes point u
          int my_sel = dsel[0];
#pragma unroll
          for (int u = 1; u < 8; ++u) my_sel = (lane == u) ? dsel[u] : my_sel;
          if (lane < 8 && my_sel >= 0) {
            float* slot = s_dg + (d0 + my_sel) * kPixBlock + px;
            *slot = (cb == 0) ? pr[0] : *slot + pr[0];
          }
        }
      }
      if (prm.feat_grad_nchw) {
        if (act) {
          s_fg[(4 * lane + 0) * (kPixBlock + 1) + px] = fg.x;
          s_fg[(4 * lane + 1) * (kPixBlock + 1) + px] = fg.y;
          s_fg[(4 * lane + 2) * (kPixBlock + 1) + px] = fg.z;
          s_fg[(4 * lane + 3) * (kPixBlock + 1) + px] = fg.w;
        }
      } else if (act) {
        Vec4<T>::store(feat_grad, pix * prm.c + 4 * (cb + lane), fg);
      }
    }
    if (prm.feat_grad_nchw) {
      __syncthreads();
      // feat_grad[bn][ch][h][w]: 8 consecutive w per (ch, h) = one sector
      for (int i = threadIdx.x; i < cw * kPixBlock; i += kBwdThreads) {
        const int cc = i / kPixBlock, px = i % kPixBlock;
        const int hh = h0 + px / kPixW, ww = w0 + px % kPixW;
        if (hh < prm.h && ww < prm.w)
          Vec4<T>::store1s(feat_grad, ((int64_t)bn * prm.c + 4 * cb + cc) * hw + (int64_t)hh * prm.w + ww,
                           s_fg[cc * (kPixBlock + 1) + px]);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- depth_grad of the block, zeros for dropped points included
  for (int i = threadIdx.x; i < prm.d * kPixBlock; i += kBwdThreads) {
    const int dd = i / kPixBlock, px = i % kPixBlock;
    const int hh = h0 + px / kPixW, ww = w0 + px % kPixW;
    if (hh < prm.h && ww < prm.w)
      Vec4<T>::store1s(depth_grad, img_base + (int64_t)dd * hw + (int64_t)hh * prm.w + ww, s_dg[i]);
  }
}

// ------------------------------------------------------------------------------------------ host side
template <typename T, int LAYOUT>
static int forward_tile_t(const void* depth, const void* feat, void* out, const int* rd, const int* rf,
                          const int* vox_pt, const FwdParams& prm, cudaStream_t st) {
  const int tiles_x = (prm.x + kTileX - 1) / kTileX;
  const int64_t tiles_per_frame = (int64_t)tiles_x * ((prm.rows + kTileRows - 1) / kTileRows);
  const int64_t n_tiles = tiles_per_frame * prm.frames;
  if (n_tiles == 0) return 0;
  if (n_tiles > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const int cw = prm.c < 128 ? prm.c : 128;
  const size_t smem = sizeof(float) * (size_t)cw * kTileStride;
  auto kern = pool_fwd_tile_kernel<T, LAYOUT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * 128 * kTileStride));
    attr_set = true;
  }
  kern<<<(unsigned)n_tiles, kFwdThreads, smem, st>>>((const T*)depth, (const T*)feat, (T*)out, rd, rf, vox_pt, prm,
                                                      tiles_x, tiles_per_frame);
  count_launch();
  return launch_status();
}

template <typename T>
static int backward_block_t(const void* og, void* dg, void* fg, const void* depth, const void* feat,
                            const int* point_rank, BwdParams prm, cudaStream_t st) {
  prm.blocks_w = (prm.w + kPixW - 1) / kPixW;
  prm.blocks_h = (prm.h + kPixH - 1) / kPixH;
  const int64_t n_blocks = (int64_t)prm.bn * prm.blocks_w * prm.blocks_h;
  if (n_blocks == 0) return 0;
  if (n_blocks > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const int cw = prm.c < 128 ? prm.c : 128;
  const size_t smem = sizeof(float) * ((size_t)3 * prm.d * kPixBlock + (prm.feat_grad_nchw ? (size_t)cw * (kPixBlock + 1) : 0));
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_ARG;   // D > ~500 depth bins
  auto kern = pool_bwd_block_kernel<T>;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  kern<<<(unsigned)n_blocks, kBwdThreads, smem, st>>>((const T*)og, (const T*)depth, (const T*)feat, point_rank, prm,
                                                       (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_voxel_table(const int32_t* ranks_bev_sorted, int64_t n_points, const int32_t* counts_dev,
                                   int64_t n_voxels_total, int32_t* vox_pt, void* stream) {
  if (n_points < 0 || n_voxels_total < 0 || n_voxels_total >= INT32_MAX) return BEVPOOL_ERR_BAD_ARG;
  if (!vox_pt || ((n_points > 0 || counts_dev) && !ranks_bev_sorted)) return BEVPOOL_ERR_BAD_ARG;
  int64_t blocks = (n_points + 1 + 255) / 256;   // n_points is an upper bound when the count lives on the device
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  voxel_table_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ranks_bev_sorted, n_points, counts_dev,
                                                                        n_voxels_total, vox_pt);
  count_launch();
  return launch_status();
}

extern "C" int bevpool_v2_forward_dense(const void* depth, const void* feat, void* out, const int32_t* ranks_depth,
                                        const int32_t* ranks_feat, const int32_t* vox_pt, int c, int64_t n_frames,
                                        int64_t rows_per_frame, int x, int dhw, int hw, int layout, int dtype,
                                        void* stream) {
  if (n_frames < 0 || rows_per_frame < 0 || x < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_frames * rows_per_frame * x == 0) return BEVPOOL_OK;
  if (n_frames * rows_per_frame * x >= INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  if (!depth || !feat || !out || !ranks_depth || !vox_pt) return BEVPOOL_ERR_BAD_ARG;
  if (!ranks_feat && (dhw <= 0 || hw <= 0)) return BEVPOOL_ERR_BAD_ARG;
  if (layout != BEVPOOL_LAYOUT_BZYXC && layout != BEVPOOL_LAYOUT_BCZYX) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)feat % 16) return BEVPOOL_ERR_BAD_ARG;
  FwdParams prm;
  prm.c = c;
  prm.x = x;
  prm.rows = rows_per_frame;
  prm.frames = n_frames;
  prm.dhw = dhw;
  prm.hw = hw;
  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH(T, L) return forward_tile_t<T, L>(depth, feat, out, ranks_depth, ranks_feat, vox_pt, prm, st)
  if (dtype == BEVPOOL_F32) {
    if (layout == BEVPOOL_LAYOUT_BCZYX) DISPATCH(float, BEVPOOL_LAYOUT_BCZYX);
    DISPATCH(float, BEVPOOL_LAYOUT_BZYXC);
  }
  if (dtype == BEVPOOL_BF16) {
    if (layout == BEVPOOL_LAYOUT_BCZYX) DISPATCH(__nv_bfloat16, BEVPOOL_LAYOUT_BCZYX);
    DISPATCH(__nv_bfloat16, BEVPOOL_LAYOUT_BZYXC);
  }
#undef DISPATCH
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" int bevpool_v2_backward_dense(const void* out_grad, void* depth_grad, void* feat_grad, const void* depth,
                                         const void* feat, const int32_t* point_rank, int bn, int d, int h, int w,
                                         int c, int feat_grad_nchw, int dtype, void* stream) {
  if (bn < 0 || d <= 0 || h < 0 || w < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4) return BEVPOOL_ERR_BAD_CHANNELS;
  if ((int64_t)bn * h * w == 0) return BEVPOOL_OK;
  if (!out_grad || !depth_grad || !feat_grad || !depth || !feat || !point_rank) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)feat % 16 || (uintptr_t)out_grad % 16 || (uintptr_t)feat_grad % 16) return BEVPOOL_ERR_BAD_ARG;
  BwdParams prm;
  prm.c = c;
  prm.d = d;
  prm.h = h;
  prm.w = w;
  prm.bn = bn;
  prm.blocks_w = prm.blocks_h = 0;
  prm.feat_grad_nchw = feat_grad_nchw ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    return backward_block_t<float>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st);
  if (dtype == BEVPOOL_BF16)
    return backward_block_t<__nv_bfloat16>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st);
  return BEVPOOL_ERR_BAD_ARG;
}
