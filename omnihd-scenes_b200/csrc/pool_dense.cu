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
#include <stdlib.h>

#include "common.cuh"

namespace bevpool {

// ------------------------------------------------------------------------------------------ voxel table
__global__ void __launch_bounds__(256)
voxel_table_kernel(const int* __restrict__ keys, int64_t n_points, const int* __restrict__ counts_dev,
                   int64_t n_voxels_total, int* __restrict__ vox_pt) {
  pdl_wait();
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
constexpr int kFwdMaxWarps = 16;
constexpr int kMaxItems = 2 * kFwdMaxWarps + kTileRows;   // chunks per tile (upper bound)

struct FwdParams {
  int c;                 // channels
  int x;                 // X
  int64_t rows;          // Z*Y rows per frame
  int64_t frames;
  int chunks_per_warp;   // target work items per warp and tile
  int hw;                // H*W            } to derive ranks_feat from ranks_depth when rf == nullptr:
  uint32_t dhw_mul;      // magic multiplier } rd / (D*H*W) == (rd * dhw_mul) >> dhw_shift   (rd < 2^31)
  int dhw_shift, dhw;
};

// Per-point scalars of one batch, one lane per point: feature-row index, depth weight and the tile
// column of the point's voxel.
template <typename T>
__device__ __forceinline__ void load_point(const T* __restrict__ depth, const int* __restrict__ rf,
                                           const int* __restrict__ rb, int rd_val, int p, const FwdParams& prm,
                                           int rank_col0, int& f, float& d, int& col) {
  if (rf) {
    f = ldg_stream_i32(rf + p);
  } else {
    const int cam = (int)(((uint64_t)(uint32_t)rd_val * prm.dhw_mul) >> prm.dhw_shift);   // rd / DHW
    const int rem = rd_val - cam * prm.dhw;                                                // d*HW + hw
    f = cam * prm.hw + rem % prm.hw;
  }
  f = (int)((unsigned)f * (unsigned)prm.c);   // element offset of the feature row as u32 (F*C < 2^32 elements)
  d = Vec4<T>::load1(depth, rd_val);
  col = ldg_stream_i32(rb + p) - rank_col0;
}

__device__ __forceinline__ void flush_column(float* tile_lane, int col, float4 acc) {
  float* c = tile_lane + col;
  c[0 * kTileStride] = acc.x;
  c[1 * kTileStride] = acc.y;
  c[2 * kTileStride] = acc.z;
  c[3 * kTileStride] = acc.w;
}

// One work item: sorted points [p, pe) of one tile row, walked FLAT with 8 feature rows in flight.
// A running sum is flushed each time the voxel changes: complete voxels go straight into their tile
// column; the first / last voxel of the chunk, when cut by the chunk boundary, go to the item's two
// partial slots (part_lane = &part[item][0][4*lane], second slot at +cw) and are added to the tile in
// item order afterwards, so the summation order is fixed.
// feat_lane = feat + this lane's channel chunk; idle lanes alias the last chunk (never stored).
template <typename T>
__device__ __forceinline__ void chunk_sum(const T* __restrict__ depth, const T* __restrict__ feat_lane,
                                          const int* __restrict__ rd, const int* __restrict__ rf,
                                          const int* __restrict__ rb, int p, int pe, bool head_partial,
                                          bool tail_partial, const FwdParams& prm, int rank_col0, float* tile_lane,
                                          float* part_lane, int* part_col, int cw, bool act) {
  const int lane = lane_id();
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = zero;

  // software pipeline: ranks_depth two batches ahead, (row, depth, column) one batch ahead
  int rd2 = (p + 32 + lane < pe) ? ldg_stream_i32(rd + p + 32 + lane) : 0;
  int f_n = 0, col_n = 0;
  float d_n = 0.f;
  if (p + lane < pe)
    load_point<T>(depth, rf, rb, ldg_stream_i32(rd + p + lane), p + lane, prm, rank_col0, f_n, d_n, col_n);
  int cur_col = __shfl_sync(kFullMask, col_n, 0);   // column of the first point
  int carry = cur_col;
  bool first_seg = true;

  auto flush = [&](bool last) {
    const bool to_part = (first_seg && head_partial) || (last && tail_partial);
    if (to_part) {
      const int slot = first_seg ? 0 : 1;   // a chunk inside one voxel uses the head slot only
      if (act) *reinterpret_cast<float4*>(part_lane + slot * cw) = acc;
      if (lane == 0) part_col[slot] = cur_col;
    } else if (act) {
      flush_column(tile_lane, cur_col, acc);
    }
    first_seg = false;
  };

  for (int q = p; q < pe; q += 32) {
    const int my_f = f_n, my_col = col_n;
    const float my_d = d_n;
    const int rd_next = rd2;
    rd2 = (q + 64 + lane < pe) ? ldg_stream_i32(rd + q + 64 + lane) : 0;
    f_n = 0; col_n = 0; d_n = 0.f;
    if (q + 32 + lane < pe)
      load_point<T>(depth, rf, rb, rd_next, q + 32 + lane, prm, rank_col0, f_n, d_n, col_n);

    const int n = min(32, pe - q);
    int prev = __shfl_up_sync(kFullMask, my_col, 1);
    if (lane == 0) prev = carry;
    const unsigned bmask = __ballot_sync(kFullMask, lane < n && my_col != prev);   // bit i: point i opens a voxel
    carry = __shfl_sync(kFullMask, my_col, n - 1);
    if (n == 32) {
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = Vec4<T>::load(feat_lane, (unsigned)__shfl_sync(kFullMask, my_f, i0 + u));
        const unsigned m8 = bmask >> i0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
          if (m8 & (1u << u)) {   // warp-uniform
            flush(false);
            cur_col = __shfl_sync(kFullMask, my_col, i0 + u);
            acc = zero;
          }
          acc = fma4(v[u], dd, acc);
        }
      }
    } else {
      for (int i0 = 0; i0 < n; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = __shfl_sync(kFullMask, my_f, i0 + u);
          v[u] = (i0 + u < n) ? Vec4<T>::load(feat_lane, (unsigned)f) : zero;
        }
        const unsigned m8 = bmask >> i0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
          if (m8 & (1u << u)) {
            flush(false);
            cur_col = __shfl_sync(kFullMask, my_col, i0 + u);
            acc = zero;
          }
          acc = fma4(v[u], dd, acc);   // masked points carry v == 0 and d == 0
        }
      }
    }
  }
  flush(true);
}

// optional per-CTA timeline (debug builds of the bench only): {tile, smid, t_start, t_end} in ns
__device__ unsigned long long g_fwd_timeline[8 * 8192];
__device__ int g_fwd_timeline_on = 0;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <typename T, int LAYOUT, int MINB>
__global__ void __launch_bounds__(kFwdMaxWarps * 32, MINB)
pool_fwd_tile_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out,
                     const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                     const int* __restrict__ vox_pt, FwdParams prm, int tiles_x, int64_t tiles_per_frame) {
  extern __shared__ float smem[];                 // tile [cw][kTileStride], then partial sums [kMaxItems][2][cw]
  __shared__ int s_col[kMaxItems][2];             // tile column of each partial (-1: none)
  __shared__ int s_row_lo[kTileRows + 1], s_row_hi[kTileRows], s_item0[kTileRows + 1];
  __shared__ int s_next, s_chunk;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int n_threads = blockDim.x, n_warps = blockDim.x >> 5;
  const int c4 = prm.c >> 2;

  const int64_t t = blockIdx.x;
  unsigned long long t_start = 0;
  if (g_fwd_timeline_on && threadIdx.x == 0) t_start = globaltimer_ns();
  const int64_t frame = t / tiles_per_frame;
  const int64_t tr = (t % tiles_per_frame) / tiles_x;   // tile row
  const int tx = (int)(t % tiles_x);
  const int x0 = tx * kTileX;
  const int w = min(kTileX, prm.x - x0);                                  // valid columns per row
  const int64_t row0 = tr * kTileRows;
  const int nrows = (int)min((int64_t)kTileRows, prm.rows - row0);
  const int64_t vpf = prm.rows * prm.x;
  const int64_t rank0 = frame * vpf + row0 * prm.x + x0;                  // rank of (row0, x0)

  // sorted-point range of each tile row, then equal-size chunks (multiples of 32 points)
  if (threadIdx.x < kTileRows) {
    const int r = threadIdx.x;
    int lo = 0, hi = 0;
    if (r < nrows) {
      lo = __ldg(vox_pt + rank0 + (int64_t)r * prm.x);
      hi = __ldg(vox_pt + rank0 + (int64_t)r * prm.x + w);
    }
    s_row_lo[r] = lo;
    s_row_hi[r] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int r = 0; r < kTileRows; ++r) total += s_row_hi[r] - s_row_lo[r];
    const int target = max(1, n_warps * prm.chunks_per_warp);
    int chunk = ((total + target - 1) / target + 31) & ~31;
    if (chunk < 32) chunk = 32;
    int items = 0;
    for (int r = 0; r < kTileRows; ++r) {
      s_item0[r] = items;
      items += (s_row_hi[r] - s_row_lo[r] + chunk - 1) / chunk;
    }
    s_item0[kTileRows] = items;
    s_chunk = chunk;
  }
  __syncthreads();
  const int chunk = s_chunk, n_items = s_item0[kTileRows];
  unsigned long long t_a = 0, t_b = 0, t_c = 0, t_d = 0;
  if (g_fwd_timeline_on && threadIdx.x == 0) t_a = globaltimer_ns();

  for (int cb = 0; cb < c4; cb += 32) {   // one sweep when C <= 128
    const bool act = cb + lane < c4;
    const int cw = min(prm.c - 4 * cb, 128);
    float* tile = smem;
    float* part = smem + cw * kTileStride;
    const T* feat_lane = feat + 4 * (cb + min(lane, c4 - 1 - cb));
    for (int i = threadIdx.x; i < cw * kTileStride; i += n_threads) tile[i] = 0.f;
    for (int i = threadIdx.x; i < 2 * kMaxItems; i += n_threads) (&s_col[0][0])[i] = -1;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    if (g_fwd_timeline_on && threadIdx.x == 0) t_b = globaltimer_ns();

    // ---- phase 1: chunks grabbed dynamically
    for (;;) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&s_next, 1);
      item = __shfl_sync(kFullMask, item, 0);
      if (item >= n_items) break;
      int r = 0;
      while (item >= s_item0[r + 1]) ++r;
      const int lo = s_row_lo[r], hi = s_row_hi[r];
      const int p = lo + (item - s_item0[r]) * chunk, pe = min(hi, p + chunk);
      // is the first / last voxel of the chunk shared with a neighbouring chunk?
      const bool head_partial = p > lo && __ldg(rb + p - 1) == __ldg(rb + p);
      const bool tail_partial = pe < hi && __ldg(rb + pe) == __ldg(rb + pe - 1);
      chunk_sum<T>(depth, feat_lane, rd, rf, rb, p, pe, head_partial, tail_partial, prm,
                   (int)(rank0 + (int64_t)r * prm.x), tile + (4 * lane) * kTileStride + r * kTileX,
                   part + (size_t)item * 2 * cw + 4 * lane, s_col[item], cw, act);
    }
    __syncthreads();
    if (g_fwd_timeline_on && threadIdx.x == 0) t_c = globaltimer_ns();
    // ---- phase 2: add the cut voxels' partial sums in item order (fixed summation order)
    for (int r = warp; r < nrows; r += n_warps) {
      float* tile_lane = tile + (4 * lane) * kTileStride + r * kTileX;
      for (int item = s_item0[r]; item < s_item0[r + 1]; ++item)
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          const int col = s_col[item][slot];
          if (col >= 0 && act) {
            const float4 pv = *reinterpret_cast<const float4*>(part + ((size_t)item * 2 + slot) * cw + 4 * lane);
            float* c = tile_lane + col;
            c[0 * kTileStride] += pv.x;
            c[1 * kTileStride] += pv.y;
            c[2 * kTileStride] += pv.z;
            c[3 * kTileStride] += pv.w;
          }
        }
    }
    __syncthreads();
    if (g_fwd_timeline_on && threadIdx.x == 0) t_d = globaltimer_ns();

    // ---- write-out (no integer divisions here: they were a third of the kernel's instructions)
    if (LAYOUT == BEVPOOL_LAYOUT_BCZYX) {
      // out[((frame*C + ch)*rows + row0 + r)*X + x0 + lane]: one 128-byte row per warp store
      if (lane < w) {
        for (int cc = warp; cc < cw; cc += n_warps) {
          const int64_t o = frame * prm.c * vpf + (int64_t)(4 * cb + cc) * vpf + row0 * prm.x + x0 + lane;
          const float* src = tile + cc * kTileStride + lane;
#pragma unroll
          for (int r = 0; r < kTileRows; ++r)
            if (r < nrows) Vec4<T>::store1s(out, o + (int64_t)r * prm.x, src[r * kTileX]);
        }
      }
    } else {
      // out[(rank)*C + ch]: a warp writes the channels of one voxel (coalesced), voxels striped over warps
      for (int r = 0; r < nrows; ++r)
        for (int xx = warp; xx < w; xx += n_warps) {
          const int64_t o = (rank0 + (int64_t)r * prm.x + xx) * prm.c + 4 * cb;
          for (int cc = lane; cc < cw; cc += 32) Vec4<T>::store1s(out, o + cc, tile[cc * kTileStride + r * kTileX + xx]);
        }
    }
    __syncthreads();
  }
  if (g_fwd_timeline_on && threadIdx.x == 0 && t < 8192) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    g_fwd_timeline[8 * t + 0] = (unsigned long long)(s_row_hi[0] - s_row_lo[0] + s_row_hi[1] - s_row_lo[1] +
                                                     s_row_hi[2] - s_row_lo[2] + s_row_hi[3] - s_row_lo[3]);
    g_fwd_timeline[8 * t + 1] = smid;
    g_fwd_timeline[8 * t + 2] = t_start;
    g_fwd_timeline[8 * t + 3] = globaltimer_ns();
    g_fwd_timeline[8 * t + 4] = t_a;
    g_fwd_timeline[8 * t + 5] = t_b;
    g_fwd_timeline[8 * t + 6] = t_c;
    g_fwd_timeline[8 * t + 7] = t_d;
  }
}

// ------------------------------------------------------------------------------------------ forward, streaming
// The default forward. The grid unit is a fixed-size chunk of the SORTED POINT LIST (perfect balance, work
// proportional to points, no per-tile cost — the tile kernel above pays ~5 us of latency chains per tile and
// its heavy tiles near the cameras set the kernel's tail; on sparse occupancy grids that was fatal). One warp
// per chunk, no barriers: each complete voxel is written as one coalesced channels-last row straight to global
// memory; voxels cut by a chunk border go through two scratch rows per chunk and are finished in chunk order by
// chunk_fixup_kernel (fixed summation order, no atomics). Empty voxels are not touched here: the layout pass
// that follows (cl_to_bczyx_zero_fill_kernel) knows them from vox_pt and writes zeros while it transposes.
constexpr int kTcCols = 64;

// chunk_sum with global row stores. part_lane = &s_part[item][0][4*lane]; s_col[0..1]: voxel rank of the slots
// (-1 none; s_col[1] == -2: the chunk lies inside one voxel that continues into the next chunk).
template <typename T>
__device__ __forceinline__ void chunk_stream(const T* __restrict__ depth, const T* __restrict__ feat_lane,
                                             const int* __restrict__ rd, const int* __restrict__ rf,
                                             const int* __restrict__ rb, int p, int pe, bool head_partial,
                                             bool tail_partial, const FwdParams& prm, T* __restrict__ out_lane,
                                             float* part_lane, int* s_col, int cw, bool act) {
  const int lane = lane_id();
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = zero;
  int rd2 = (p + 32 + lane < pe) ? ldg_stream_i32(rd + p + 32 + lane) : 0;
  int f_n = 0, col_n = 0;
  float d_n = 0.f;
  if (p + lane < pe)
    load_point<T>(depth, rf, rb, ldg_stream_i32(rd + p + lane), p + lane, prm, 0, f_n, d_n, col_n);   // col = voxel rank
  int cur = __shfl_sync(kFullMask, col_n, 0);
  int carry = cur;
  bool first_seg = true;

  auto flush = [&](bool last) {
    const bool to_part = (first_seg && head_partial) || (last && tail_partial);
    if (to_part) {
      // slot 0: continuation of a voxel opened by an earlier chunk; slot 1: voxel opened here, continues later
      const bool cont = first_seg && head_partial;
      const int slot = cont ? 0 : 1;
      if (act) *reinterpret_cast<float4*>(part_lane + slot * cw) = acc;
      if (lane == 0) {
        s_col[slot] = cur;
        if (cont && last && tail_partial) s_col[1] = -2;   // the voxel stays open across this whole chunk
      }
    } else if (act) {
      Vec4<T>::store_keep(out_lane, (int64_t)cur * prm.c, acc);
    }
    first_seg = false;
  };

  for (int q = p; q < pe; q += 32) {
    const int my_f = f_n, my_col = col_n;
    const float my_d = d_n;
    const int rd_next = rd2;
    rd2 = (q + 64 + lane < pe) ? ldg_stream_i32(rd + q + 64 + lane) : 0;
    f_n = 0; col_n = 0; d_n = 0.f;
    if (q + 32 + lane < pe) load_point<T>(depth, rf, rb, rd_next, q + 32 + lane, prm, 0, f_n, d_n, col_n);

    const int n = min(32, pe - q);
    int prev = __shfl_up_sync(kFullMask, my_col, 1);
    if (lane == 0) prev = carry;
    const unsigned bmask = __ballot_sync(kFullMask, lane < n && my_col != prev);
    carry = __shfl_sync(kFullMask, my_col, n - 1);
    if (n == 32) {
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = Vec4<T>::load(feat_lane, (unsigned)__shfl_sync(kFullMask, my_f, i0 + u));
        const unsigned m8 = (bmask >> i0) & 0xffu;
        if (m8 == 0) {   // common: the 8 points stay inside the current voxel
#pragma unroll
          for (int u = 0; u < 8; ++u) acc = fma4(v[u], __shfl_sync(kFullMask, my_d, i0 + u), acc);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
            if (m8 & (1u << u)) {
              flush(false);
              cur = __shfl_sync(kFullMask, my_col, i0 + u);
              acc = zero;
            }
            acc = fma4(v[u], dd, acc);
          }
        }
      }
    } else {
      for (int i0 = 0; i0 < n; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = __shfl_sync(kFullMask, my_f, i0 + u);
          v[u] = (i0 + u < n) ? Vec4<T>::load(feat_lane, (unsigned)f) : zero;
        }
        const unsigned m8 = bmask >> i0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
          if (m8 & (1u << u)) {
            flush(false);
            cur = __shfl_sync(kFullMask, my_col, i0 + u);
            acc = zero;
          }
          acc = fma4(v[u], dd, acc);
        }
      }
    }
  }
  flush(true);
}

template <typename T, int MINB>
__global__ void __launch_bounds__(256, MINB)
pool_fwd_chunk_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out_cl,
                      const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                      const int* __restrict__ counts_dev, int n_points, int chunk, FwdParams prm,
                      float* __restrict__ part /*[chunks][2][c]*/, int* __restrict__ part_rank /*[chunks][2]*/) {
  const int lane = lane_id();
  const int c4 = prm.c >> 2;
  pdl_wait();
  if (counts_dev) n_points = counts_dev[0];
  const int64_t cidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t p64 = cidx * chunk;
  if (p64 >= n_points) return;
  const int p = (int)p64, pe = min(n_points, p + chunk);
  const bool head_partial = p > 0 && __ldg(rb + p - 1) == __ldg(rb + p);
  const bool tail_partial = pe < n_points && __ldg(rb + pe) == __ldg(rb + pe - 1);
  int* my_rank = part_rank + cidx * 2;
  if (lane < 2) my_rank[lane] = -1;
  __syncwarp();
  const bool act = lane < c4;
  const int lane_c = 4 * min(lane, c4 - 1);
  chunk_stream<T>(depth, feat + lane_c, rd, rf, rb, p, pe, head_partial, tail_partial, prm, out_cl + lane_c,
                  part + cidx * 2 * prm.c + 4 * lane, my_rank, prm.c, act);
}

// One warp per chunk whose tail slot opens a cut voxel: adds the following chunks' head slots in order.
template <typename T>
__global__ void __launch_bounds__(256)
chunk_fixup_kernel(const float* __restrict__ part, const int* __restrict__ part_rank, const int* __restrict__ counts_dev,
                   int n_points, int chunk, int c, T* __restrict__ out_cl) {
  pdl_wait();
  if (counts_dev) n_points = counts_dev[0];
  const int64_t n_chunks = ((int64_t)n_points + chunk - 1) / chunk;
  const int lane = lane_id();
  const int64_t cidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (cidx >= n_chunks) return;
  const int open = part_rank[cidx * 2 + 1];
  if (open < 0) return;
  for (int ch = 4 * lane; ch < c; ch += 128) {
    float4 acc = *reinterpret_cast<const float4*>(part + (cidx * 2 + 1) * c + ch);
    for (int64_t j = cidx + 1; j < n_chunks; ++j) {
      const float4 pv = *reinterpret_cast<const float4*>(part + (j * 2) * c + ch);
      acc.x += pv.x; acc.y += pv.y; acc.z += pv.z; acc.w += pv.w;
      if (part_rank[j * 2 + 1] != -2) break;   // the voxel ends inside chunk j
    }
    Vec4<T>::store_keep(out_cl, (int64_t)open * c + ch, acc);
  }
}

// [B][V][C] channels-last rows of the NON-EMPTY voxels -> [B][C][V], zeros for empty voxels (known from
// vox_pt). A CTA moves all channels of 64 consecutive voxels: 128-bit coalesced on both sides.
template <typename T>
__global__ void __launch_bounds__(256)
cl_to_bczyx_zero_fill_kernel(const T* __restrict__ src_cl, const int* __restrict__ vox_pt, T* __restrict__ dst,
                             int c, int64_t vpf, int64_t tiles_per_frame) {
  extern __shared__ float t[];            // [c][kTcCols + 1]
  __shared__ int s_pt[kTcCols + 1];
  pdl_wait();
  const int64_t b = blockIdx.x / tiles_per_frame;
  const int64_t v0 = (blockIdx.x % tiles_per_frame) * kTcCols;
  const int ncol = (int)min((int64_t)kTcCols, vpf - v0);
  const int64_t rank0 = b * vpf + v0;
  for (int i = threadIdx.x; i <= ncol; i += 256) s_pt[i] = __ldg(vox_pt + rank0 + i);
  __syncthreads();
  const int c4 = c >> 2;
  for (int i = threadIdx.x; i < ncol * c4; i += 256) {
    const int col = i / c4, q = i - col * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s_pt[col + 1] > s_pt[col]) v = Vec4<T>::load_stream(src_cl, (rank0 + col) * c + 4 * q);
    float* o = t + (4 * q) * (kTcCols + 1) + col;
    o[0] = v.x; o[kTcCols + 1] = v.y; o[2 * (kTcCols + 1)] = v.z; o[3 * (kTcCols + 1)] = v.w;
  }
  __syncthreads();
  T* d = dst + (b * c) * vpf + v0;
  const bool vec = ncol == kTcCols && (vpf & 3) == 0 && ((((uintptr_t)dst) & 15) == 0);
  if (vec) {
    for (int i = threadIdx.x; i < c * (kTcCols / 4); i += 256) {
      const int ch = i / (kTcCols / 4), q = i % (kTcCols / 4);
      const float* p = t + ch * (kTcCols + 1) + 4 * q;
      Vec4<T>::store(d, (int64_t)ch * vpf + 4 * q, make_float4(p[0], p[1], p[2], p[3]));
    }
  } else {
    for (int i = threadIdx.x; i < c * kTcCols; i += 256) {
      const int ch = i / kTcCols, q = i % kTcCols;
      if (q < ncol) Vec4<T>::store1s(d, (int64_t)ch * vpf + q, t[ch * (kTcCols + 1) + q]);
    }
  }
}

// channels-last final layout: rows of empty voxels are zeroed (the streaming forward writes the others)
template <typename T>
__global__ void __launch_bounds__(256)
zero_empty_rows_kernel(T* __restrict__ out_cl, const int* __restrict__ vox_pt, int c, int64_t n_voxels) {
  const int c4 = c >> 2;
  const int64_t total = n_voxels * c4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t v = i / c4;
    if (__ldg(vox_pt + v + 1) == __ldg(vox_pt + v)) Vec4<T>::store(out_cl, i * 4, make_float4(0.f, 0.f, 0.f, 0.f));
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
  // [d][32] each: voxel rank, depth, depth_grad of the block; then per-warp compacted point lists
  int* s_rank = reinterpret_cast<int*>(smem_raw);
  float* s_depth = reinterpret_cast<float*>(smem_raw) + (size_t)prm.d * kPixBlock;
  float* s_dg = s_depth + (size_t)prm.d * kPixBlock;
  int4* s_list = reinterpret_cast<int4*>(s_dg + (size_t)prm.d * kPixBlock);            // [8 warps][d] {rank, depth, dbin, -}
  float* s_fg = reinterpret_cast<float*>(s_list + (size_t)kBwdWarps * prm.d);           // [cw][33] (NCHW output only)
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int c4 = prm.c >> 2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  const int blk = blockIdx.x;
  const int per_img = prm.blocks_w * prm.blocks_h;
  const int bn = blk / per_img;
  const int brem = blk - bn * per_img;
  const int bh = brem / prm.blocks_w, bw = brem - bh * prm.blocks_w;
  const int h0 = bh * kPixH, w0 = bw * kPixW;
  const int64_t hw = (int64_t)prm.h * prm.w;
  const int64_t img_base = (int64_t)bn * prm.d * hw;   // + d*hw + h*W + w

  // ---- stage point_rank / depth of the block: 8 consecutive w = one 32-byte sector per (d, h)
  {
    const int px = threadIdx.x & 31;
    const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
    const bool in = hh < prm.h && ww < prm.w;
    const int64_t o0 = img_base + (int64_t)hh * prm.w + ww;
    // four bins per warp in flight, rank and depth loaded independently (one dependent pair per bin and iteration
    // cost two memory round trips per bin: D / 8 * 2 serialized latencies per CTA)
    for (int dd0 = threadIdx.x >> 5; dd0 < prm.d; dd0 += 4 * kBwdWarps) {
      int r[4];
      float dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = dd0 + k * kBwdWarps;
        r[k] = -1;
        dv[k] = 0.f;
        if (in && dd < prm.d) {
          r[k] = ldg_stream_i32(point_rank + o0 + dd * hw);
          dv[k] = Vec4<T>::load1(depth, o0 + dd * hw);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = dd0 + k * kBwdWarps;
        if (dd < prm.d) {
          s_rank[dd * kPixBlock + px] = r[k];
          s_depth[dd * kPixBlock + px] = r[k] >= 0 ? dv[k] : 0.f;
          s_dg[dd * kPixBlock + px] = 0.f;
        }
      }
    }
  }
  __syncthreads();

  int4* my_list = s_list + (size_t)warp * prm.d;
  for (int cb = 0; cb < c4; cb += 32) {
    const bool act = cb + lane < c4;
    const int cw = min(prm.c - 4 * cb, 128);
    const int lane_c = 4 * (cb + min(lane, c4 - 1 - cb));     // idle lanes alias the last chunk; never stored
    const T* og_lane = og + lane_c;
    for (int px = warp; px < kPixBlock; px += kBwdWarps) {
      const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
      if (hh >= prm.h || ww >= prm.w) continue;   // warp-uniform
      const int64_t pix = (int64_t)bn * hw + (int64_t)hh * prm.w + ww;
      // compact the kept depth bins of this pixel into the warp's list
      int n_kept = 0;
      for (int d0 = 0; d0 < prm.d; d0 += 32) {
        const int dd = d0 + lane;
        const int r = dd < prm.d ? s_rank[dd * kPixBlock + px] : -1;
        const unsigned live = __ballot_sync(kFullMask, r >= 0);
        if (r >= 0)
          my_list[n_kept + __popc(live & ((1u << lane) - 1u))] =
              make_int4(r, __float_as_int(s_depth[dd * kPixBlock + px]), dd, 0);
        n_kept += __popc(live);
      }
      __syncwarp();
      const float4 fv = Vec4<T>::load(feat, pix * prm.c + lane_c);
      float4 fg = zero;
      for (int b = 0; b < n_kept; b += 8) {
        float4 g[8];
        float dv[8], pr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (b + u < n_kept) {                     // warp-uniform
            const int4 e = my_list[b + u];          // broadcast read
            g[u] = Vec4<T>::load(og_lane, (int64_t)e.x * prm.c);
            dv[u] = __int_as_float(e.y);
          } else {
            g[u] = zero;
            dv[u] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          fg = fma4(g[u], dv[u], fg);
          pr[u] = act ? dot4_packed(g[u], fv) : 0.f;
        }
        // reduce-scatter over lane bits 2,1,0, then finish over bits 3,4: lane l ends up with the
        // complete dot product of point b + (l & 7)
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
        if (lane < 8 && b + lane < n_kept) {
          float* slot = s_dg + my_list[b + lane].z * kPixBlock + px;
          *slot = (cb == 0) ? pr[0] : *slot + pr[0];
        }
      }
      __syncwarp();
      if (prm.feat_grad_nchw) {
        if (act) {
          float* c = s_fg + (4 * lane) * (kPixBlock + 1) + px;
          c[0 * (kPixBlock + 1)] = fg.x;
          c[1 * (kPixBlock + 1)] = fg.y;
          c[2 * (kPixBlock + 1)] = fg.z;
          c[3 * (kPixBlock + 1)] = fg.w;
        }
      } else if (act) {
        Vec4<T>::store(feat_grad, pix * prm.c + lane_c, fg);
      }
    }
    if (prm.feat_grad_nchw) {
      __syncthreads();
      // feat_grad[bn][ch][h][w]: 8 consecutive w per (ch, h) = one sector
      const int px = threadIdx.x & 31;
      const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
      if (hh < prm.h && ww < prm.w) {
        const int64_t o0 = ((int64_t)bn * prm.c + 4 * cb) * hw + (int64_t)hh * prm.w + ww;
        for (int cc = threadIdx.x >> 5; cc < cw; cc += kBwdWarps)
          Vec4<T>::store1s(feat_grad, o0 + cc * hw, s_fg[cc * (kPixBlock + 1) + px]);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- depth_grad of the block, zeros for dropped points included
  {
    const int px = threadIdx.x & 31;
    const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
    if (hh < prm.h && ww < prm.w) {
      const int64_t o0 = img_base + (int64_t)hh * prm.w + ww;
      for (int dd = threadIdx.x >> 5; dd < prm.d; dd += kBwdWarps)
        Vec4<T>::store1s(depth_grad, o0 + dd * hw, s_dg[dd * kPixBlock + px]);
    }
  }
}

// Same contract for C <= 64: a row is at most 16 lanes wide, so each HALF-warp takes its own pixel (own compacted
// list, own out_grad rows, own 16-lane reduce-scatter) and a warp finishes two pixels per pass — twice the lane
// utilisation of the kernel above on the OmniHD (C = 64) and occupancy (C = 32) shapes.
template <typename T, int CC>     // CC = C when known at compile time (32, 64), 0 = runtime
__global__ void __launch_bounds__(kBwdThreads)
pool_bwd_block_half_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                           const int* __restrict__ point_rank, BwdParams prm, T* __restrict__ depth_grad,
                           T* __restrict__ feat_grad) {
  extern __shared__ unsigned char smem_raw[];
  int* s_rank = reinterpret_cast<int*>(smem_raw);
  float* s_depth = reinterpret_cast<float*>(smem_raw) + (size_t)prm.d * kPixBlock;
  float* s_dg = s_depth + (size_t)prm.d * kPixBlock;
  int4* s_list = reinterpret_cast<int4*>(s_dg + (size_t)prm.d * kPixBlock);            // [8 warps][2 halves][d]
  float* s_fg = reinterpret_cast<float*>(s_list + (size_t)kBwdWarps * 2 * prm.d);       // [c][33] (NCHW output only)
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  const int C = CC ? CC : prm.c;   // the ncu capture showed 40 IMAD + 9 LDC per batch of row-address arithmetic with a runtime C
  const int c4 = C >> 2;   // <= 16
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  const int blk = blockIdx.x;
  const int per_img = prm.blocks_w * prm.blocks_h;
  const int bn = blk / per_img;
  const int brem = blk - bn * per_img;
  const int bh = brem / prm.blocks_w, bw = brem - bh * prm.blocks_w;
  const int h0 = bh * kPixH, w0 = bw * kPixW;
  const int64_t hw = (int64_t)prm.h * prm.w;
  const int64_t img_base = (int64_t)bn * prm.d * hw;
  pdl_wait();
  {
    const int px = threadIdx.x & 31;
    const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
    const bool in = hh < prm.h && ww < prm.w;
    const int64_t o0 = img_base + (int64_t)hh * prm.w + ww;
    // four bins per warp in flight, rank and depth loaded independently (one dependent pair per bin and iteration
    // cost two memory round trips per bin: D / 8 * 2 serialized latencies per CTA)
    for (int dd0 = threadIdx.x >> 5; dd0 < prm.d; dd0 += 4 * kBwdWarps) {
      int r[4];
      float dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = dd0 + k * kBwdWarps;
        r[k] = -1;
        dv[k] = 0.f;
        if (in && dd < prm.d) {
          r[k] = ldg_stream_i32(point_rank + o0 + dd * hw);
          dv[k] = Vec4<T>::load1(depth, o0 + dd * hw);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = dd0 + k * kBwdWarps;
        if (dd < prm.d) {
          s_rank[dd * kPixBlock + px] = r[k];
          s_depth[dd * kPixBlock + px] = r[k] >= 0 ? dv[k] : 0.f;
          s_dg[dd * kPixBlock + px] = 0.f;
        }
      }
    }
  }
  __syncthreads();

  int4* my_list = s_list + ((size_t)warp * 2 + half) * prm.d;
  const bool act = hl < c4;
  const int lane_c = 4 * min(hl, c4 - 1);     // idle lanes alias the last chunk; never stored
  const T* og_lane = og + lane_c;
  const unsigned hshift = 16u * half;
  for (int it = 0; it < kPixBlock / (2 * kBwdWarps); ++it) {
    const int px = it * 2 * kBwdWarps + 2 * warp + half;     // this half-warp's pixel
    const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
    const bool pin = hh < prm.h && ww < prm.w;
    const int64_t pix = (int64_t)bn * hw + (int64_t)hh * prm.w + ww;
    int n_kept = 0;
    for (int d0 = 0; d0 < prm.d; d0 += 16) {
      const int dd = d0 + hl;
      const int r = (pin && dd < prm.d) ? s_rank[dd * kPixBlock + px] : -1;
      const unsigned live = (__ballot_sync(kFullMask, r >= 0) >> hshift) & 0xffffu;
      if (r >= 0)
        my_list[n_kept + __popc(live & ((1u << hl) - 1u))] = make_int4(r, __float_as_int(s_depth[dd * kPixBlock + px]), dd, 0);
      n_kept += __popc(live);
    }
    if (n_kept == 0 && hl == 0) my_list[0] = make_int4(0, 0, 0, 0);   // so that the clamped re-read below is always valid
    __syncwarp();
    const int n_max = max(__shfl_sync(kFullMask, n_kept, 0), __shfl_sync(kFullMask, n_kept, 16));
    const float4 fv = pin ? Vec4<T>::load(feat, pix * C + lane_c) : zero;
    float4 fg = zero;
    // batches of 4 points, software-pipelined: the out_grad rows of the next batch are requested before the current
    // one is consumed (the loop is bound by the latency of these gathers, not by their bandwidth)
    float4 g_n[4];
    float d_n[4];
    auto request = [&](int b0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // past the end of the list: re-read the pixel's last item with weight 0 (no zero-filled registers, no branch;
        // its dot product is never stored)
        const int4 e = my_list[max(min(b0 + u, n_kept - 1), 0)];
        g_n[u] = Vec4<T>::load(og_lane, (int64_t)e.x * C);
        d_n[u] = b0 + u < n_kept ? __int_as_float(e.y) : 0.f;
      }
    };
    request(0);
    for (int b = 0; b < n_max; b += 4) {
      float4 g[4];
      float dv[4], pr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] = g_n[u];
        dv[u] = d_n[u];
      }
      request(b + 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fg = fma4(g[u], dv[u], fg);
        pr[u] = act ? dot4_packed(g[u], fv) : 0.f;
      }
      // reduce-scatter over lane bits 1,0, then bits 2,3 (inside the half-warp): lane hl ends up with the complete
      // dot product of point b + (hl & 3)
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
      pr[0] += __shfl_xor_sync(kFullMask, pr[0], 4);
      pr[0] += __shfl_xor_sync(kFullMask, pr[0], 8);
      if (hl < 4 && b + hl < n_kept) s_dg[my_list[b + hl].z * kPixBlock + px] = pr[0];
    }
    __syncwarp();
    if (prm.feat_grad_nchw) {
      if (act && pin) {
        float* c = s_fg + (4 * hl) * (kPixBlock + 1) + px;
        c[0 * (kPixBlock + 1)] = fg.x;
        c[1 * (kPixBlock + 1)] = fg.y;
        c[2 * (kPixBlock + 1)] = fg.z;
        c[3 * (kPixBlock + 1)] = fg.w;
      }
    } else if (act && pin) {
      Vec4<T>::store(feat_grad, pix * C + lane_c, fg);
    }
  }
  __syncthreads();
  const int px = threadIdx.x & 31;
  const int hh = h0 + (px >> 3), ww = w0 + (px & 7);
  if (hh < prm.h && ww < prm.w) {
    if (prm.feat_grad_nchw) {
      const int64_t o0 = (int64_t)bn * C * hw + (int64_t)hh * prm.w + ww;
      for (int cc = threadIdx.x >> 5; cc < C; cc += kBwdWarps)
        Vec4<T>::store1s(feat_grad, o0 + cc * hw, s_fg[cc * (kPixBlock + 1) + px]);
    }
    const int64_t o0 = img_base + (int64_t)hh * prm.w + ww;
    for (int dd = threadIdx.x >> 5; dd < prm.d; dd += kBwdWarps)
      Vec4<T>::store1s(depth_grad, o0 + dd * hw, s_dg[dd * kPixBlock + px]);
  }
}

// ------------------------------------------------------------------------------------------ backward, joint columns
// Same contract as pool_bwd_block_kernel, for grids where the pixels of one image column mostly land in the
// same voxel at a given depth bin (Z == 1 BEV grids: the 4 pixels (h0..h0+3, w) of a warp's column differ only
// in height, so they share the voxel or fall out of the z-range). The warp walks its 4 pixels JOINTLY, bin by
// bin: ONE out_grad row is loaded per bin and applied to every kept pixel (4 dot products + 4 feat_grad
// updates per row instead of one row load per point). Bins whose kept pixels do NOT share one voxel take a
// per-pixel slow path, so the result is the same on any grid — only the speed differs.
//
// The kernel is issue-bound, so the inner loop is kept minimal: the staged arrays are [d][w] vectors over the
// 4 rows (one broadcast 128-bit shared load gives all 4 ranks / depths of a bin), there are no lists and no
// compaction, the channel count is a template constant (CH4 = C/4 active lanes, immediate offsets), and the
// cross-lane sums of the dot products are done 8 bins at a time by a transposed shared-memory read (each lane
// parks its 4 partials in a [bin][lane] slot; lane (bin, pixel) then adds the CH4 partials) instead of
// shuffle trees.
constexpr int kJointPad = kPixBlock + 1;   // row stride of the [c][pixel] feat_grad transpose tile
constexpr int kJointBins = 8;              // bins per reduce group

template <typename T, int CH4>
__global__ void __launch_bounds__(kBwdThreads, 2)
pool_bwd_joint_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                      const int* __restrict__ point_rank, BwdParams prm, T* __restrict__ depth_grad,
                      T* __restrict__ feat_grad) {
  constexpr int C = CH4 * 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d_pad = (prm.d + kJointBins - 1) / kJointBins * kJointBins;             // pad bins: lead = -1
  int4* s_rank4 = reinterpret_cast<int4*>(smem_raw);                               // [d_pad][8]: ranks of rows h0..h0+3
  float4* s_depth4 = reinterpret_cast<float4*>(s_rank4 + (size_t)d_pad * kPixW);   // [d_pad][8]: 0 where dropped
  float4* s_dg4 = s_depth4 + (size_t)d_pad * kPixW;                                // [d_pad][8]
  int* s_lead = reinterpret_cast<int*>(s_dg4 + (size_t)d_pad * kPixW);             // [d_pad][8] column summaries
  float* s_part = reinterpret_cast<float*>(s_lead + (size_t)d_pad * kPixW);        // [8 warps][8 bins][C]; later [C][33]
  float* s_fg = s_part;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  // 3-D grid (w-block, image, row band): no integer divisions to find the block. Bands are dispatched from the bottom
  // of the image up, all images of a band together: bands differ a lot in kept points (upper rows leave the z-range
  // early) and with the lightest band last the tail of the last wave is short.
  const int bw = blockIdx.x, bn = blockIdx.y, bh = prm.blocks_h - 1 - (int)blockIdx.z;
  const int blk = (bh * prm.bn + bn) * prm.blocks_w + bw;   // timeline slot (debug hook only)
  const int h0 = bh * kPixH, w0 = bw * kPixW;
  const int hw = prm.h * prm.w;
  const int64_t img_base = (int64_t)bn * prm.d * hw;
  unsigned long long t_s = 0, t_a = 0, t_b = 0;
  pdl_wait();
  if (g_fwd_timeline_on && threadIdx.x == 0) t_s = globaltimer_ns();

  // ---- stage point_rank / depth of the block: 8 consecutive w = one 32-byte sector per (d, h);
  //      8 bins per thread are loaded before anything is stored (one memory latency, not eight)
  if (threadIdx.x < (d_pad - prm.d) * kPixW) {   // pad bins: empty, weight 0 (they ride through the straight-line path)
    s_lead[prm.d * kPixW + threadIdx.x] = -1;
    s_depth4[prm.d * kPixW + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {
    const int hl = lane >> 3, wl = lane & 7;
    const bool in = h0 + hl < prm.h && w0 + wl < prm.w;
    const int64_t o0 = img_base + (h0 + hl) * prm.w + w0 + wl;
    for (int d0 = 0; d0 < prm.d; d0 += 8 * kBwdWarps) {
      int r[8];
      float dv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dd = d0 + k * kBwdWarps + warp;
        r[k] = -1;
        dv[k] = 0.f;
        if (in && dd < prm.d) {
          r[k] = ldg_stream_i32(point_rank + o0 + (int64_t)dd * hw);
          dv[k] = Vec4<T>::load1(depth, o0 + (int64_t)dd * hw);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dd = d0 + k * kBwdWarps + warp;
        // column summary: -1 nothing kept; rank >= 0: every kept row of the column sits in that voxel;
        // -2 - rank: mixed column whose largest rank is `rank`
        int lead = max(r[k], __shfl_xor_sync(kFullMask, r[k], 8));
        lead = max(lead, __shfl_xor_sync(kFullMask, lead, 16));
        const unsigned agree = __ballot_sync(kFullMask, r[k] < 0 || r[k] == lead);
        const bool shared = ((agree >> wl) & 0x01010101u) == 0x01010101u;
        if (dd < prm.d) {
          reinterpret_cast<int*>(s_rank4 + dd * kPixW + wl)[hl] = r[k];
          reinterpret_cast<float*>(s_depth4 + dd * kPixW + wl)[hl] = r[k] >= 0 ? dv[k] : 0.f;
          if (hl == 0) s_lead[dd * kPixW + wl] = lead < 0 ? -1 : (shared ? lead : -2 - lead);
        }
      }
    }
  }
  __syncthreads();
  if (g_fwd_timeline_on && threadIdx.x == 0) t_a = globaltimer_ns();

  const int ww = w0 + warp;   // this warp's image column
  // C = 80 uses the 4 + 1 channel mapping (Frag<true>: lane sl holds channels 4sl..4sl+3 and 64+sl), so a row is 16
  // lanes wide and the warp splits into two lane groups that take alternating bins of the 8-bin window.
  constexpr bool X = CH4 == 20;
  constexpr int G = X ? 2 : (CH4 <= 8 ? 4 : (CH4 <= 16 ? 2 : 1));   // narrow rows: more groups
  constexpr int RL = X ? 16 : CH4;          // lanes per row
  constexpr int NU = 4 / G;                 // bins per group and pass
  const int grp = lane / (32 / G), sl = lane % (32 / G);
  const bool act = sl < RL;
  const int sc = act ? sl : RL - 1;         // idle lanes alias the last slice; never stored
  constexpr int PS = 4 * RL + 4;   // partial-row stride: (4k + p + 4l) mod 32 is conflict-free for the transposed read
  float* my_part = s_part + (size_t)warp * kJointBins * PS + 4 * sc;
  Frag<X> fv[kPixH], fg[kPixH];
#pragma unroll
  for (int p = 0; p < kPixH; ++p) {
    fg[p] = frag_zero<X>();
    fv[p] = (ww < prm.w && h0 + p < prm.h) ? frag_load<T, X>(feat + ((int64_t)bn * hw + (h0 + p) * prm.w + ww) * C, sc)
                                           : frag_zero<X>();
  }
  if (ww < prm.w) {   // warp-uniform
    const int4* rank_col = s_rank4 + warp;
    const float4* depth_col = s_depth4 + warp;
    const int* lead_col = s_lead + warp;
    // Software pipeline over passes of 4 bins: the out_grad rows of pass t + 1 are requested before pass t is
    // consumed, so the L2 round trip of the gather overlaps this warp's own arithmetic (with 4 warps per scheduler
    // the other warps alone do not cover it).
    int code_n[NU];
    Frag<X> g_n[NU];
    auto request = [&](int dbase) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        code_n[u] = dbase < d_pad ? lead_col[(dbase + G * u + grp) * kPixW] : -1;   // broadcast LDS (per lane group)
        g_n[u] = frag_zero<X>();
        if (code_n[u] != -1) g_n[u] = frag_load<T, X>(og + (int64_t)(code_n[u] >= 0 ? code_n[u] : -2 - code_n[u]) * C, sc);
      }
    };
    request(0);
    for (int d0 = 0; d0 < prm.d; d0 += kJointBins) {
#pragma unroll
      for (int k0 = 0; k0 < kJointBins; k0 += 4) {
        // NU bins per lane group at a time (bins k0 + G*u + grp)
        int code[NU];
        Frag<X> g[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          code[u] = code_n[u];
          g[u] = g_n[u];
        }
        request(d0 + k0 + 4);
        int cmax = code[0], cmin = code[0];
#pragma unroll
        for (int u = 1; u < NU; ++u) {
          cmax = max(cmax, code[u]);
          cmin = min(cmin, code[u]);
        }
        // group-uniform: ALL bins of the pass are empty (dropped points come in long runs). Mixed bins carry codes
        // <= -2, so the maximum alone does not tell (round 1 tested `cmax == -1` and skipped a mixed bin that shared its
        // pass with an empty one: wrong gradients under camera roll; caught by test_rolled_cameras_mixed_column_bins)
        if (cmax == -1 && cmin == -1) continue;
        if (cmin >= -1) {
          // no mixed column among the bins: straight-line code, no per-bin branches. Empty bins ride along with
          // g = 0 and weights 0 (their partials are discarded by the reducing lane below).
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int k = k0 + G * u + grp;
            const float4 dp = depth_col[(d0 + k) * kPixW];
            const float dw[kPixH] = {dp.x, dp.y, dp.z, dp.w};
            float dt[kPixH];
#pragma unroll
            for (int p = 0; p < kPixH; ++p) {
              fg[p] = frag_fma<X>(g[u], dw[p], fg[p]);
              dt[p] = frag_dot<X>(g[u], fv[p]);
            }
            if (act) *reinterpret_cast<float4*>(my_part + k * PS) = make_float4(dt[0], dt[1], dt[2], dt[3]);
          }
          continue;
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          if (code[u] == -1) continue;
          const int k = k0 + G * u + grp;
          const int dd = d0 + k;
          const float4 dp = depth_col[dd * kPixW];
          const float dw[kPixH] = {dp.x, dp.y, dp.z, dp.w};
          float dt[kPixH];
          if (code[u] >= 0) {   // every kept pixel of the column sits in the lead voxel; dropped rows weigh 0
#pragma unroll
            for (int p = 0; p < kPixH; ++p) {
              fg[p] = frag_fma<X>(g[u], dw[p], fg[p]);
              dt[p] = frag_dot<X>(g[u], fv[p]);   // dropped rows: discarded by the reducing lane below
            }
          } else {
            // mixed column: rows of one voxel are adjacent (height is monotonic along the column), so a row is
            // loaded once per RUN of equal ranks (the largest rank's row is already in g[u]); all loads of the bin
            // are issued before the first use
            const int lead = -2 - code[u];
            const int4 r4 = rank_col[dd * kPixW];
            const int rr[kPixH] = {r4.x, r4.y, r4.z, r4.w};
            Frag<X> gp[kPixH];
#pragma unroll
            for (int p = 0; p < kPixH; ++p) {
              if (rr[p] < 0 || rr[p] == lead) gp[p] = g[u];
              else if (p > 0 && rr[p] == rr[p - 1]) gp[p] = gp[p - 1];
              else gp[p] = frag_load<T, X>(og + (int64_t)rr[p] * C, sc);
            }
#pragma unroll
            for (int p = 0; p < kPixH; ++p) {
              dt[p] = 0.f;
              if (rr[p] >= 0) {
                fg[p] = frag_fma<X>(gp[p], dw[p], fg[p]);
                dt[p] = frag_dot<X>(gp[p], fv[p]);
              }
            }
          }
          if (act) *reinterpret_cast<float4*>(my_part + k * PS) = make_float4(dt[0], dt[1], dt[2], dt[3]);
        }
      }
      __syncwarp();
      // ---- lane (k, p) sums the RL partial dot products of bin d0 + k, pixel p
      {
        const int k = lane >> 2, pz = lane & 3;
        const int dd = d0 + k;
        if (dd < prm.d) {
          const int mine = reinterpret_cast<const int*>(rank_col + dd * kPixW)[pz];
          float sum = 0.f;
          if (mine >= 0) {
            const float* src = s_part + (size_t)warp * kJointBins * PS + k * PS + pz;
#pragma unroll
            for (int l = 0; l < RL; ++l) sum += src[4 * l];
          }
          reinterpret_cast<float*>(s_dg4 + dd * kPixW + warp)[pz] = sum;
        }
      }
      __syncwarp();
    }
  } else {
    for (int d = lane; d < prm.d; d += 32) s_dg4[d * kPixW + warp] = zero;
  }
  // the lane groups hold feat_grad of interleaved bins: add them up, group 0 stores
#pragma unroll
  for (int o = 16; o >= 32 / G && G > 1; o >>= 1) {
#pragma unroll
    for (int p = 0; p < kPixH; ++p) {
      fg[p].v.x += __shfl_xor_sync(kFullMask, fg[p].v.x, o);
      fg[p].v.y += __shfl_xor_sync(kFullMask, fg[p].v.y, o);
      fg[p].v.z += __shfl_xor_sync(kFullMask, fg[p].v.z, o);
      fg[p].v.w += __shfl_xor_sync(kFullMask, fg[p].v.w, o);
      if (X) fg[p].s += __shfl_xor_sync(kFullMask, fg[p].s, o);
    }
  }
  const bool writer = act && grp == 0;
  // ---- feat_grad of the 4 pixels
  if (prm.feat_grad_nchw) {
    __syncthreads();   // every warp is done with its partial buffer: re-use it as the transpose tile
    if (ww < prm.w && writer) {
#pragma unroll
      for (int p = 0; p < kPixH; ++p) {
        float* c = s_fg + (4 * sc) * kJointPad + p * kPixW + warp;
        c[0 * kJointPad] = fg[p].v.x;
        c[1 * kJointPad] = fg[p].v.y;
        c[2 * kJointPad] = fg[p].v.z;
        c[3 * kJointPad] = fg[p].v.w;
        if (X) s_fg[(64 + sc) * kJointPad + p * kPixW + warp] = fg[p].s;
      }
    }
    __syncthreads();
    const int hl = lane >> 3, wl = lane & 7;
    if (h0 + hl < prm.h && w0 + wl < prm.w) {
      const int64_t o0 = (int64_t)bn * C * hw + (h0 + hl) * prm.w + w0 + wl;
      for (int cc = warp; cc < C; cc += kBwdWarps)
        Vec4<T>::store1s(feat_grad, o0 + (int64_t)cc * hw, s_fg[cc * kJointPad + lane]);
    }
  } else if (ww < prm.w && writer) {
#pragma unroll
    for (int p = 0; p < kPixH; ++p)
      if (h0 + p < prm.h) {
        const int64_t o = ((int64_t)bn * hw + (h0 + p) * prm.w + ww) * C;
        Vec4<T>::store(feat_grad, o + 4 * sc, fg[p].v);
        if (X) Vec4<T>::store1(feat_grad, o + 64 + sc, fg[p].s);
      }
  }
  __syncthreads();
  if (g_fwd_timeline_on && threadIdx.x == 0) t_b = globaltimer_ns();
  // ---- depth_grad of the block, zeros for dropped points included
  {
    const int hl = lane >> 3, wl = lane & 7;
    if (h0 + hl < prm.h && w0 + wl < prm.w) {
      const int64_t o0 = img_base + (h0 + hl) * prm.w + w0 + wl;
      for (int dd = warp; dd < prm.d; dd += kBwdWarps)
        Vec4<T>::store1s(depth_grad, o0 + (int64_t)dd * hw, reinterpret_cast<const float*>(s_dg4 + dd * kPixW + wl)[hl]);
    }
  }
  if (g_fwd_timeline_on && threadIdx.x == 0 && blk < 8192) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    g_fwd_timeline[8 * blk + 0] = 0;
    g_fwd_timeline[8 * blk + 1] = smid;
    g_fwd_timeline[8 * blk + 2] = t_s;
    g_fwd_timeline[8 * blk + 3] = globaltimer_ns();
    g_fwd_timeline[8 * blk + 4] = t_a;
    g_fwd_timeline[8 * blk + 5] = t_b;
  }
}

// ------------------------------------------------------------------------------------------ host side
template <typename T, int LAYOUT>
static int forward_tile_t(const void* depth, const void* feat, void* out, const int* rd, const int* rf,
                          const int* rb, const int* vox_pt, const FwdParams& prm, cudaStream_t st) {
  const int tiles_x = (prm.x + kTileX - 1) / kTileX;
  const int64_t tiles_per_frame = (int64_t)tiles_x * ((prm.rows + kTileRows - 1) / kTileRows);
  const int64_t n_tiles = tiles_per_frame * prm.frames;
  if (n_tiles == 0) return 0;
  if (n_tiles > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const int cw = prm.c < 128 ? prm.c : 128;
  static int minb = 0;
  if (!minb) {
    const char* e = getenv("BEVPOOL_FWD_MINB");
    minb = e ? atoi(e) : 2;
  }
  auto kern = minb >= 2 ? pool_fwd_tile_kernel<T, LAYOUT, 2> : pool_fwd_tile_kernel<T, LAYOUT, 1>;
  {
    const size_t max_smem = sizeof(float) * (128 * kTileStride + kMaxItems * 2 * 128);
    if (int rc = ensure_dynamic_smem(kern, max_smem)) return rc;
  }
  static int warps = 0, cpw = 0;
  if (!warps) {  // tuning knobs (defaults are the measured best on B200)
    const char* e;
    warps = (e = getenv("BEVPOOL_FWD_WARPS")) ? atoi(e) : 16;
    cpw = (e = getenv("BEVPOOL_FWD_CPW")) ? atoi(e) : 1;
    if (warps < 1 || warps > kFwdMaxWarps) warps = 8;
    if (cpw < 1 || cpw > 2) cpw = 2;
  }
  FwdParams prm2 = prm;
  prm2.chunks_per_warp = cpw;
  const size_t smem = sizeof(float) * ((size_t)cw * kTileStride + (size_t)(warps * cpw + kTileRows) * 2 * cw);
  const int kFwdThreads = warps * 32;
  kern<<<(unsigned)n_tiles, kFwdThreads, smem, st>>>((const T*)depth, (const T*)feat, (T*)out, rd, rf, rb, vox_pt, prm2,
                                                      tiles_x, tiles_per_frame);
  count_launch();
  return launch_status();
}

static const int kFwdChunkMin = 64;   // smallest chunk the scratch is sized for

// Streaming forward + layout pass. `scratch` (n_voxels*c elements of T) holds the channels-last rows when the
// requested layout is BCZYX; for BZYXC the rows go straight to `out` and empty rows are zeroed.
template <typename T>
static int forward_stream_t(const void* depth, const void* feat, void* out, const int* rd, const int* rf,
                            const int* rb, const int* vox_pt, const FwdParams& prm, int layout, void* scratch,
                            const int* counts_dev, int64_t n_points_upper, float* part, int* part_rank,
                            cudaStream_t st) {
  static int chunk_env = -1, minb_env = 3;
  if (chunk_env < 0) {
    const char* e;
    chunk_env = (e = getenv("BEVPOOL_FWD_CHUNK")) ? atoi(e) : 128;
    if (chunk_env > 0 && chunk_env < kFwdChunkMin) chunk_env = kFwdChunkMin;
    minb_env = (e = getenv("BEVPOOL_FWD_MINB")) ? atoi(e) : 3;
  }
  T* rows_dst0 = (T*)(layout == BEVPOOL_LAYOUT_BCZYX ? scratch : out);
  if (chunk_env <= 0) chunk_env = 128;
  {
    const int64_t n_chunks = ((int64_t)n_points_upper + chunk_env - 1) / chunk_env;
    if (n_chunks > 0) {
      const unsigned blocks = (unsigned)((n_chunks + 7) / 8);
      if (minb_env >= 4)
        launch_pdl(pool_fwd_chunk_kernel<T, 4>, dim3(blocks), dim3(256), 0, st, (const T*)depth, (const T*)feat, rows_dst0, rd, rf,
                   rb, counts_dev, (int)n_points_upper, chunk_env, prm, part, part_rank);
      else
        launch_pdl(pool_fwd_chunk_kernel<T, 3>, dim3(blocks), dim3(256), 0, st, (const T*)depth, (const T*)feat, rows_dst0, rd, rf,
                   rb, counts_dev, (int)n_points_upper, chunk_env, prm, part, part_rank);
      launch_pdl(chunk_fixup_kernel<T>, dim3(blocks), dim3(256), 0, st, (const float*)part, (const int*)part_rank, counts_dev,
                 (int)n_points_upper, chunk_env, prm.c, rows_dst0);
      count_launch(2);
    }
  }
  const int64_t vpf = prm.rows * prm.x, n_vox = vpf * prm.frames;
  if (layout == BEVPOOL_LAYOUT_BCZYX) {
    const int64_t tpf = (vpf + kTcCols - 1) / kTcCols;
    const size_t smem2 = sizeof(float) * (size_t)prm.c * (kTcCols + 1);
    if (int rc = ensure_dynamic_smem(cl_to_bczyx_zero_fill_kernel<T>, smem2)) return rc;
    launch_pdl(cl_to_bczyx_zero_fill_kernel<T>, dim3((unsigned)(tpf * prm.frames)), dim3(256), smem2, st, (const T*)scratch,
               vox_pt, (T*)out, prm.c, vpf, tpf);
  } else {
    int64_t blocks = (n_vox * (prm.c >> 2) + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    zero_empty_rows_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((T*)out, vox_pt, prm.c, n_vox);
  }
  count_launch();
  return launch_status();
}

template <typename T, int CH4>
static int backward_joint_launch(const void* og, void* dg, void* fg, const void* depth, const void* feat,
                                 const int* point_rank, const BwdParams& prm, int64_t n_blocks, cudaStream_t st) {
  constexpr int C = CH4 * 4;
  size_t part_bytes = sizeof(float) * (size_t)kBwdWarps * kJointBins * (C + 4);   // >= 4 * RL + 4 per bin
  const size_t fg_bytes = prm.feat_grad_nchw ? sizeof(float) * (size_t)C * kJointPad : 0;
  if (fg_bytes > part_bytes) part_bytes = fg_bytes;
  const size_t d_pad = (size_t)(prm.d + kJointBins - 1) / kJointBins * kJointBins;
  const size_t smem = d_pad * kPixW * (3 * 16 + 4) + part_bytes;
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_ARG;
  auto kern = pool_bwd_joint_kernel<T, CH4>;
  if (int rc = ensure_dynamic_smem(kern, smem)) return rc;
  (void)n_blocks;
  if (prm.bn > 65535 || prm.blocks_h > 65535) return BEVPOOL_ERR_OVERFLOW;
  launch_pdl(kern, dim3((unsigned)prm.blocks_w, (unsigned)prm.bn, (unsigned)prm.blocks_h), dim3(kBwdThreads), smem, st,
             (const T*)og, (const T*)depth, (const T*)feat, point_rank, prm, (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

// Channel counts with a specialised joint kernel (CH4 = C/4 <= 32). Others use the generic block kernel.
template <typename T>
static int backward_joint_t(const void* og, void* dg, void* fg, const void* depth, const void* feat,
                            const int* point_rank, BwdParams prm, cudaStream_t st, bool* handled) {
  prm.blocks_w = (prm.w + kPixW - 1) / kPixW;
  prm.blocks_h = (prm.h + kPixH - 1) / kPixH;
  const int64_t n_blocks = (int64_t)prm.bn * prm.blocks_w * prm.blocks_h;
  *handled = true;
  if (n_blocks == 0) return 0;
  if (n_blocks > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  switch (prm.c) {
    case 32: return backward_joint_launch<T, 8>(og, dg, fg, depth, feat, point_rank, prm, n_blocks, st);
    case 64: return backward_joint_launch<T, 16>(og, dg, fg, depth, feat, point_rank, prm, n_blocks, st);
    case 80: return backward_joint_launch<T, 20>(og, dg, fg, depth, feat, point_rank, prm, n_blocks, st);
    case 128: return backward_joint_launch<T, 32>(og, dg, fg, depth, feat, point_rank, prm, n_blocks, st);
    default: break;
  }
  *handled = false;
  return 0;
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
  const bool half = prm.c <= 64;     // two pixels per warp
  const size_t smem = sizeof(float) * ((size_t)3 * prm.d * kPixBlock + (prm.feat_grad_nchw ? (size_t)cw * (kPixBlock + 1) : 0)) +
                      sizeof(int4) * (size_t)kBwdWarps * prm.d * (half ? 2 : 1);
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_ARG;   // D > ~500 depth bins
  auto kern = !half ? pool_bwd_block_kernel<T>
                    : (prm.c == 64 ? pool_bwd_block_half_kernel<T, 64>
                                   : (prm.c == 32 ? pool_bwd_block_half_kernel<T, 32> : pool_bwd_block_half_kernel<T, 0>));
  if (int rc = ensure_dynamic_smem(kern, smem)) return rc;
  kern<<<(unsigned)n_blocks, kBwdThreads, smem, st>>>((const T*)og, (const T*)depth, (const T*)feat, point_rank, prm,
                                                       (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

int backward_column(const void* og, void* dg, void* fg, const void* depth, const void* feat, const int* point_rank,
                    int bn, int d, int h, int w, int c, int feat_grad_nchw, int dtype, cudaStream_t st, bool* handled);   // pool_column.cu

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_voxel_table(const int32_t* ranks_bev_sorted, int64_t n_points, const int32_t* counts_dev,
                                   int64_t n_voxels_total, int32_t* vox_pt, void* stream) {
  if (n_points < 0 || n_voxels_total < 0 || n_voxels_total >= INT32_MAX) return BEVPOOL_ERR_BAD_ARG;
  if (!vox_pt || ((n_points > 0 || counts_dev) && !ranks_bev_sorted)) return BEVPOOL_ERR_BAD_ARG;
  int64_t blocks = (n_points + 1 + 255) / 256;   // n_points is an upper bound when the count lives on the device
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  launch_pdl(voxel_table_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const int*)ranks_bev_sorted,
             n_points, (const int*)counts_dev, n_voxels_total, (int*)vox_pt);
  count_launch();
  return launch_status();
}

extern "C" size_t bevpool_v2_forward_dense_scratch_bytes(int64_t n_points, int64_t n_voxels, int c, int layout,
                                                         int dtype) {
  if (n_points < 0 || n_voxels < 0 || c <= 0) return 0;
  const size_t esz = dtype == BEVPOOL_BF16 ? 2 : 4;
  size_t rows = layout == BEVPOOL_LAYOUT_BCZYX ? (size_t)n_voxels * c * esz : 0;
  rows = (rows + 255) / 256 * 256;
  const size_t n_chunks = (size_t)(n_points + kFwdChunkMin - 1) / kFwdChunkMin + 1;
  return rows + n_chunks * 2 * (size_t)c * sizeof(float) + n_chunks * 2 * sizeof(int) + 256;
}

extern "C" int bevpool_v2_forward_dense(const void* depth, const void* feat, void* out, const int32_t* ranks_depth,
                                        const int32_t* ranks_feat, const int32_t* ranks_bev, const int32_t* vox_pt,
                                        int64_t n_points, const int32_t* counts_dev, int c, int64_t n_frames,
                                        int64_t rows_per_frame, int x, int dhw, int hw, int layout, int dtype,
                                        void* scratch, size_t scratch_bytes, void* stream) {
  if (n_frames < 0 || rows_per_frame < 0 || x < 0 || n_points < 0 || n_points >= INT32_MAX) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4) return BEVPOOL_ERR_BAD_CHANNELS;
  const int64_t n_vox = n_frames * rows_per_frame * x;
  if (n_vox == 0) return BEVPOOL_OK;
  if (n_vox >= INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  if (!depth || !feat || !out || !ranks_depth || !ranks_bev || !vox_pt) return BEVPOOL_ERR_BAD_ARG;
  if (!ranks_feat && (dhw <= 0 || hw <= 0)) return BEVPOOL_ERR_BAD_ARG;
  if (layout != BEVPOOL_LAYOUT_BZYXC && layout != BEVPOOL_LAYOUT_BCZYX) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)feat % 16) return BEVPOOL_ERR_BAD_ARG;
  FwdParams prm;
  prm.c = c;
  prm.x = x;
  prm.rows = rows_per_frame;
  prm.frames = n_frames;
  prm.chunks_per_warp = 1;
  prm.dhw = dhw > 0 ? dhw : 1;
  prm.hw = hw > 0 ? hw : 1;
  {  // rd / dhw == (rd * mul) >> shift for 0 <= rd < 2^31 (round-up magic number)
    int s = 0;
    while (((int64_t)1 << s) < prm.dhw) ++s;
    prm.dhw_shift = 31 + s;
    prm.dhw_mul = (uint32_t)((((uint64_t)1 << (31 + s)) / (uint64_t)prm.dhw) + 1);
  }
  cudaStream_t st = (cudaStream_t)stream;
  static int use_stream = -1;
  if (use_stream < 0) {
    const char* e = getenv("BEVPOOL_FWD_KERNEL");   // "tile" selects the shared-memory tile kernel (A/B testing)
    use_stream = !(e && e[0] == 't');
  }
  const size_t need = bevpool_v2_forward_dense_scratch_bytes(n_points, n_vox, c, layout, dtype);
  if (use_stream && c <= 128 && scratch && scratch_bytes >= need && ((uintptr_t)scratch % 256) == 0) {
    const size_t esz = dtype == BEVPOOL_BF16 ? 2 : 4;
    size_t rows = layout == BEVPOOL_LAYOUT_BCZYX ? (size_t)n_vox * c * esz : 0;
    rows = (rows + 255) / 256 * 256;
    const size_t n_chunks = (size_t)(n_points + kFwdChunkMin - 1) / kFwdChunkMin + 1;
    float* part = (float*)((char*)scratch + rows);
    int* part_rank = (int*)((char*)scratch + rows + n_chunks * 2 * (size_t)c * sizeof(float));
    if (dtype == BEVPOOL_F32)
      return forward_stream_t<float>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, vox_pt, prm, layout, scratch,
                                     counts_dev, n_points, part, part_rank, st);
    if (dtype == BEVPOOL_BF16)
      return forward_stream_t<__nv_bfloat16>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, vox_pt, prm, layout,
                                             scratch, counts_dev, n_points, part, part_rank, st);
    return BEVPOOL_ERR_BAD_ARG;
  }
#define DISPATCH(T, L) return forward_tile_t<T, L>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, vox_pt, prm, st)
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
                                         int c, int feat_grad_nchw, int column_hint, int dtype, void* stream) {
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
  if (d >= 65536) return BEVPOOL_ERR_BAD_ARG;
  if (dtype != BEVPOOL_F32 && dtype != BEVPOOL_BF16) return BEVPOOL_ERR_BAD_ARG;
  if (column_hint) {
    bool handled = false;
    // column-GEMM kernel (pool_column.cu); BEVPOOL_BWD_COLUMN=0 keeps the round-1 joint kernel (A/B measurement only)
    static int use_column = -1;
    if (use_column < 0) {
      const char* e = getenv("BEVPOOL_BWD_COLUMN");
      use_column = !(e && e[0] == '0');
    }
    if (use_column) {
      const int rc = backward_column(out_grad, depth_grad, feat_grad, depth, feat, point_rank, bn, d, h, w, c,
                                     prm.feat_grad_nchw, dtype, st, &handled);
      if (handled) return rc;
    }
    const int rc = dtype == BEVPOOL_F32
                       ? backward_joint_t<float>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st, &handled)
                       : backward_joint_t<__nv_bfloat16>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st,
                                                         &handled);
    if (handled) return rc;
  }
  if (dtype == BEVPOOL_F32)
    return backward_block_t<float>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st);
  if (dtype == BEVPOOL_BF16)
    return backward_block_t<__nv_bfloat16>(out_grad, depth_grad, feat_grad, depth, feat, point_rank, prm, st);
  return BEVPOOL_ERR_BAD_ARG;
}

// debug: enable / fetch the forward kernel's per-CTA timeline (not part of the public header)
extern "C" int bevpool_debug_fwd_timeline(int enable, unsigned long long* host_out, int n_tiles) {
  cudaMemcpyToSymbol(g_fwd_timeline_on, &enable, sizeof(int));
  if (host_out && n_tiles > 0)
    cudaMemcpyFromSymbol(host_out, g_fwd_timeline, sizeof(unsigned long long) * 8 * (n_tiles < 8192 ? n_tiles : 8192));
  return (int)cudaGetLastError();
}
