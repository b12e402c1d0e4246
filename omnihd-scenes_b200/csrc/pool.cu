// bev_pool_v2 forward / backward kernels for sm_100a and their C-ABI entry points.
//
// The op is a gather-weighted segmented reduction (0.15-8.6 flop/B): HBM/L2-bound SIMT work,
// no tensor cores. Design rules applied throughout:
//   * one warp per interval (forward) / per feature pixel (backward); lanes span channels as
//     128-bit vectors, so every feature / grad row is fetched with one coalesced LDG.128 per lane;
//   * the per-point scalars (indices, depth) are loaded once per warp, 32 points at a time,
//     one lane per point, and broadcast by shuffle — never re-read per channel;
//   * no atomics: every output element has exactly one writer; sums run in sorted point order
//     with FMA, which makes the fp32 forward bit-identical to the reference kernel;
//   * streaming data (indices, depth, outputs) uses L1::no_allocate / st.cs so that L1 and L2
//     keep the gathered rows, which are re-used ~D*(P/P0) times.
#include "common.cuh"

namespace bevpool {

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;

// ------------------------------------------------------------------------------------------
// Accumulate `len` points starting at sorted position s into acc (one float4 channel chunk per
// lane). Points are consumed in order; 8 row loads are kept in flight.
template <typename T>
__device__ __forceinline__ float4 gather_weighted_sum(const T* __restrict__ depth, const T* __restrict__ feat,
                                                      const int* __restrict__ rd, const int* __restrict__ rf,
                                                      int64_t s, int len, int c, int ch, bool act, float4 acc) {
  const int lane = lane_id();
  for (int base = 0; base < len; base += 32) {
    int my_rf = 0;
    float my_d = 0.f;
    if (base + lane < len) {
      my_rf = ldg_stream_i32(rf + s + base + lane);
      my_d = Vec4<T>::load1(depth, ldg_stream_i32(rd + s + base + lane));
    }
    const int n = min(32, len - base);
    int i = 0;
    for (; i + 8 <= n; i += 8) {
      float4 v[8];
      float d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int f = __shfl_sync(kFullMask, my_rf, i + u);
        d[u] = __shfl_sync(kFullMask, my_d, i + u);
        v[u] = act ? Vec4<T>::load(feat, (int64_t)f * c + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fma4(v[u], d[u], acc);
    }
    for (; i < n; ++i) {
      const int f = __shfl_sync(kFullMask, my_rf, i);
      const float d = __shfl_sync(kFullMask, my_d, i);
      if (act) acc = fma4(Vec4<T>::load(feat, (int64_t)f * c + 4 * ch), d, acc);
    }
  }
  return acc;
}

// ------------------------------------------------------------------------------------------
// Reference-contract forward: out is channels-last and pre-zeroed; one warp per interval.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_fwd_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out,
                const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                const int* __restrict__ starts, const int* __restrict__ lengths, int64_t n_intervals, int c) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int c4 = c >> 2;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = __ldg(starts + k), len = __ldg(lengths + k);
    const int64_t obase = (int64_t)__ldg(rb + s) * c;
    for (int cb = 0; cb < c4; cb += 32) {
      const int ch = cb + lane;
      const bool act = ch < c4;
      float4 acc = gather_weighted_sum<T>(depth, feat, rd, rf, s, len, c, ch, act, make_float4(0.f, 0.f, 0.f, 0.f));
      if (act) Vec4<T>::store(out, obase + 4 * ch, acc);
    }
  }
}

// Any channel count (the reference's KAT uses C=2): one lane per channel, 32 channels per sweep.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_fwd_scalar_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out,
                       const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                       const int* __restrict__ starts, const int* __restrict__ lengths, int64_t n_intervals, int c) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = starts[k], len = lengths[k];
    const int64_t obase = (int64_t)rb[s] * c;
    for (int cb = 0; cb < c; cb += 32) {
      const int ch = cb + lane;
      if (ch >= c) continue;
      float acc = 0.f;
      for (int i = 0; i < len; ++i)
        acc = fmaf(Vec4<T>::load1(feat, (int64_t)rf[s + i] * c + ch), Vec4<T>::load1(depth, rd[s + i]), acc);
      Vec4<T>::store1(out, obase + ch, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Reference-contract backward: intervals are grouped by feature pixel; one warp per pixel.
//   depth_grad[rd_i] = <og[rb_i,:], feat[f,:]>      (one butterfly reduction per point)
//   feat_grad[f,:]   = sum_i depth[rd_i] * og[rb_i,:]
// The feature row stays in registers for the whole interval; og rows are read exactly once.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_bwd_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                const int* __restrict__ starts, const int* __restrict__ lengths, int64_t n_intervals, int c,
                T* __restrict__ depth_grad, T* __restrict__ feat_grad) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int c4 = c >> 2;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = __ldg(starts + k), len = __ldg(lengths + k);
    const int64_t fbase = (int64_t)__ldg(rf + s) * c;
    for (int cb = 0; cb < c4; cb += 32) {
      const int ch = cb + lane;
      const bool act = ch < c4;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 fv = act ? Vec4<T>::load(feat, fbase + 4 * ch) : zero;
      float4 fg = zero;
      for (int base = 0; base < len; base += 32) {
        int my_rb = 0, my_rd = 0;
        float my_d = 0.f, my_dg = 0.f;
        if (base + lane < len) {
          my_rb = ldg_stream_i32(rb + s + base + lane);
          my_rd = ldg_stream_i32(rd + s + base + lane);
          my_d = Vec4<T>::load1(depth, my_rd);
        }
        const int n = min(32, len - base);
        int i = 0;
        for (; i + 4 <= n; i += 4) {
          float4 g[4];
          float d[4], p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int b = __shfl_sync(kFullMask, my_rb, i + u);
            d[u] = __shfl_sync(kFullMask, my_d, i + u);
            g[u] = act ? Vec4<T>::load(og, (int64_t)b * c + 4 * ch) : zero;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            fg = fma4(g[u], d[u], fg);
            p[u] = warp_sum(dot4(g[u], fv, 0.f));
            if (lane == i + u) my_dg = p[u];
          }
        }
        for (; i < n; ++i) {
          const int b = __shfl_sync(kFullMask, my_rb, i);
          const float d = __shfl_sync(kFullMask, my_d, i);
          const float4 g = act ? Vec4<T>::load(og, (int64_t)b * c + 4 * ch) : zero;
          fg = fma4(g, d, fg);
          const float p = warp_sum(dot4(g, fv, 0.f));
          if (lane == i) my_dg = p;
        }
        if (base + lane < len) {
          // with more than 128 channels the dot product is completed over several sweeps
          if (cb == 0) Vec4<T>::store1(depth_grad, my_rd, my_dg);
          else Vec4<T>::store1(depth_grad, my_rd, Vec4<T>::load1(depth_grad, my_rd) + my_dg);
        }
      }
      if (act) Vec4<T>::store(feat_grad, fbase + 4 * ch, fg);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_bwd_scalar_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                       const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                       const int* __restrict__ starts, const int* __restrict__ lengths, int64_t n_intervals, int c,
                       T* __restrict__ depth_grad, T* __restrict__ feat_grad) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = starts[k], len = lengths[k];
    const int64_t fbase = (int64_t)rf[s] * c;
    for (int i = 0; i < len; ++i) {
      const int64_t gb = (int64_t)rb[s + i] * c;
      float p = 0.f;
      for (int ch = lane; ch < c; ch += 32) p = fmaf(Vec4<T>::load1(og, gb + ch), Vec4<T>::load1(feat, fbase + ch), p);
      p = warp_sum(p);
      if (lane == 0) Vec4<T>::store1(depth_grad, rd[s + i], p);
    }
    for (int ch = lane; ch < c; ch += 32) {
      float g = 0.f;
      for (int i = 0; i < len; ++i)
        g = fmaf(Vec4<T>::load1(og, (int64_t)rb[s + i] * c + ch), Vec4<T>::load1(depth, rd[s + i]), g);
      Vec4<T>::store1(feat_grad, fbase + ch, g);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dense forward. The voxel axis of one frame is cut into strips of 32 consecutive ranks; a CTA
// owns one strip, sums the strip's intervals (warp per interval; CTA-cooperative for long ones),
// parks the results in a shared-memory tile that starts out zero — so empty voxels cost nothing
// extra — and writes the tile with fully coalesced 128-byte rows in either layout. This fuses
// the reference's memset (bev_pool.py:27), kernel (:29) and permute (:91).
//
// strip_first[s] = index of the first interval whose voxel rank >= 32*s (table built below).
constexpr int kStrip = 32;
constexpr int kDenseWarps = 4;
constexpr int kDenseThreads = kDenseWarps * 32;
constexpr int kLongInterval = 96;  // intervals longer than this are split over the CTA's warps

__global__ void strip_table_kernel(const int* __restrict__ rb, const int* __restrict__ starts, int64_t n_intervals,
                                   const int* __restrict__ counts_dev, int64_t n_strips, int64_t strips_per_frame,
                                   int64_t voxels_per_frame, int* __restrict__ strip_first) {
  if (counts_dev) n_intervals = counts_dev[1];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // strip index of a voxel rank: strips never straddle frames
  auto strip_of = [&](int64_t v) { return (v / voxels_per_frame) * strips_per_frame + (v % voxels_per_frame) / kStrip; };
  auto strip_begin = [&](int64_t s) { return (s / strips_per_frame) * voxels_per_frame + (s % strips_per_frame) * kStrip; };
  for (int64_t j = tid; j <= n_intervals; j += nthreads) {
    // strips whose first voxel lies in (v_{j-1}, v_j] start at interval j
    const int64_t vprev = (j == 0) ? -1 : (int64_t)rb[starts[j - 1]];
    int64_t lo = (j == 0) ? 0 : strip_of(vprev);
    if (j > 0 && strip_begin(lo) <= vprev) lo += 1;  // first strip beginning after vprev
    int64_t hi;                                      // last strip beginning at or before v_j
    if (j == n_intervals) hi = n_strips;             // sentinel entry included
    else hi = strip_of((int64_t)rb[starts[j]]);
    for (int64_t s = lo; s <= hi; ++s) strip_first[s] = (int)j;
  }
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(kDenseThreads)
pool_fwd_dense_kernel(const T* __restrict__ depth, const T* __restrict__ feat, T* __restrict__ out,
                      const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                      const int* __restrict__ starts, const int* __restrict__ lengths,
                      const int* __restrict__ strip_first, int64_t n_strips, int64_t strips_per_frame,
                      int64_t voxels_per_frame, int c) {
  extern __shared__ float tile[];  // LAYOUT BCZYX: [c][33]; BZYXC: [32][c]
  __shared__ float4 partial[kDenseWarps][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int c4 = c >> 2;
  for (int64_t strip = blockIdx.x; strip < n_strips; strip += gridDim.x) {
    const int64_t frame = strip / strips_per_frame;
    const int64_t v0 = (strip % strips_per_frame) * kStrip;                 // first voxel within the frame
    const int nvox = (int)min((int64_t)kStrip, voxels_per_frame - v0);
    const int64_t rank0 = frame * voxels_per_frame + v0;
    const int j0 = __ldg(strip_first + strip), j1 = __ldg(strip_first + strip + 1);

    for (int cb = 0; cb < c4; cb += 32) {  // one sweep for C <= 128
      const int ch = cb + lane;
      const bool act = ch < c4;
      const int cw = min(c - 4 * cb, 128);  // channels covered by this sweep
      for (int i = threadIdx.x; i < (LAYOUT == BEVPOOL_LAYOUT_BCZYX ? cw * (kStrip + 1) : kStrip * cw); i += kDenseThreads)
        tile[i] = 0.f;
      __syncthreads();
      // short intervals: warp per interval
      bool any_long = false;
      for (int j = j0 + warp; j < j1; j += kDenseWarps) {
        const int s = __ldg(starts + j), len = __ldg(lengths + j);
        if (len > kLongInterval) { any_long = true; continue; }
        const int x = __ldg(rb + s) - (int)rank0;
        const float4 acc = gather_weighted_sum<T>(depth, feat, rd, rf, s, len, c, ch, act, make_float4(0.f, 0.f, 0.f, 0.f));
        if (act) {
          const int cl = 4 * lane;
          if (LAYOUT == BEVPOOL_LAYOUT_BCZYX) {
            tile[(cl + 0) * (kStrip + 1) + x] = acc.x;
            tile[(cl + 1) * (kStrip + 1) + x] = acc.y;
            tile[(cl + 2) * (kStrip + 1) + x] = acc.z;
            tile[(cl + 3) * (kStrip + 1) + x] = acc.w;
          } else {
            *reinterpret_cast<float4*>(tile + x * cw + cl) = acc;
          }
        }
      }
      // long intervals: all warps split the point range, partials are combined in warp order
      if (__syncthreads_or(any_long)) {
        for (int j = j0; j < j1; ++j) {
          const int s = __ldg(starts + j), len = __ldg(lengths + j);
          if (len <= kLongInterval) continue;
          const int per = (len + kDenseWarps - 1) / kDenseWarps;
          const int b = min(len, warp * per), e = min(len, b + per);
          partial[warp][lane] = gather_weighted_sum<T>(depth, feat, rd, rf, (int64_t)s + b, e - b, c, ch, act,
                                                       make_float4(0.f, 0.f, 0.f, 0.f));
          __syncthreads();
          if (warp == 0 && act) {
            float4 acc = partial[0][lane];
#pragma unroll
            for (int w = 1; w < kDenseWarps; ++w) {
              const float4 p = partial[w][lane];
              acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            }
            const int x = __ldg(rb + s) - (int)rank0;
            const int cl = 4 * lane;
            if (LAYOUT == BEVPOOL_LAYOUT_BCZYX) {
              tile[(cl + 0) * (kStrip + 1) + x] = acc.x;
              tile[(cl + 1) * (kStrip + 1) + x] = acc.y;
              tile[(cl + 2) * (kStrip + 1) + x] = acc.z;
              tile[(cl + 3) * (kStrip + 1) + x] = acc.w;
            } else {
              *reinterpret_cast<float4*>(tile + x * cw + cl) = acc;
            }
          }
          __syncthreads();
        }
      }
      __syncthreads();
      // coalesced tile write-out
      if (LAYOUT == BEVPOOL_LAYOUT_BCZYX) {
        // out[(frame*C + ch)*V + v0 + x]
        for (int cc = warp; cc < cw; cc += kDenseWarps) {
          if (lane < nvox) {
            const int64_t o = (frame * c + (4 * cb + cc)) * voxels_per_frame + v0 + lane;
            Vec4<T>::store1(out, o, tile[cc * (kStrip + 1) + lane]);
          }
        }
      } else {
        // out[(rank0 + x)*C + 4*cb + cc], contiguous over (x, cc) when one sweep covers C
        const int cw4 = cw >> 2;
        for (int i = threadIdx.x; i < nvox * cw4; i += kDenseThreads) {
          const int x = i / cw4, q = i % cw4;
          Vec4<T>::store(out, (rank0 + x) * c + 4 * cb + 4 * q, *reinterpret_cast<const float4*>(tile + x * cw + 4 * q));
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dense (sort-free) backward for ranks produced by bevpool_prepare_v2: the D depth bins of one
// feature pixel ARE its backward interval (ranks_feat is a function of ranks_depth), so the
// reference's per-step argsort (bev_pool.py:47) is replaced by a walk over point_rank[P0].
// One warp per pixel; 8 points per batch with 8 og rows in flight; dot products finished with a
// reduce-scatter butterfly (9 shuffles per 8 points instead of 40). Writes every element of both
// gradients, so no memset is needed either (bev_pool.py:67-68).
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_bwd_dense_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                      const int* __restrict__ point_rank, int64_t n_pixels, int d_bins, int hw, int c,
                      T* __restrict__ depth_grad, T* __restrict__ feat_grad) {
  const int lane = lane_id();
  const int c4 = c >> 2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // consecutive warps take consecutive pixels: the 32-byte sectors of point_rank / depth /
  // depth_grad (8 pixels wide) are shared through L1 by the 8 warps of a CTA
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pix = warp0; pix < n_pixels; pix += n_warps) {
    const int64_t bn = pix / hw, p = pix % hw;
    const int64_t pbase = bn * d_bins * (int64_t)hw + p;  // + d*hw
    for (int cb = 0; cb < c4; cb += 32) {
      const int ch = cb + lane;
      const bool act = ch < c4;
      const float4 fv = act ? Vec4<T>::load(feat, pix * c + 4 * ch) : zero;
      float4 fg = zero;
      for (int d0 = 0; d0 < d_bins; d0 += 32) {
        int my_r = -1;
        float my_d = 0.f, my_dg = 0.f;
        const bool have = d0 + lane < d_bins;
        if (have) {
          my_r = __ldg(point_rank + pbase + (int64_t)(d0 + lane) * hw);
          if (my_r >= 0) my_d = Vec4<T>::load1(depth, pbase + (int64_t)(d0 + lane) * hw);
        }
        const unsigned live = __ballot_sync(kFullMask, my_r >= 0);
        for (int i0 = 0; i0 < 32; i0 += 8) {
          if (((live >> i0) & 0xffu) == 0) continue;
          float4 g[8];
          float pr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = __shfl_sync(kFullMask, my_r, i0 + u);
            g[u] = (act && r >= 0) ? Vec4<T>::load(og, (int64_t)r * c + 4 * ch) : zero;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float dd = __shfl_sync(kFullMask, my_d, i0 + u);
            fg = fma4(g[u], dd, fg);
            pr[u] = dot4(g[u], fv, 0.f);
          }
          // reduce-scatter over lane bits 0..2, then finish over bits 3..4
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
          // lane l now holds the full dot product of point i0 + (l & 7)
          if ((lane >> 3) == (i0 >> 3)) my_dg = pr[0];  // lane i0+k takes point i0+k
        }
        if (have) {
          const int64_t o = pbase + (int64_t)(d0 + lane) * hw;
          if (cb == 0) Vec4<T>::store1(depth_grad, o, my_dg);
          else Vec4<T>::store1(depth_grad, o, Vec4<T>::load1(depth_grad, o) + my_dg);
        }
      }
      if (act) Vec4<T>::store(feat_grad, pix * c + 4 * ch, fg);
    }
  }
}

// ------------------------------------------------------------------------------------------
// [B,C,S] <-> [B,S,C] (S = Z*Y*X) through a 32x33 shared tile, coalesced on both sides.
template <typename T>
__global__ void __launch_bounds__(256)
grid_transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int rows, int64_t cols, int64_t tiles_c,
                      int64_t tiles_total) {
  // src is [B][rows][cols], dst is [B][cols][rows]
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t tiles_r = (rows + 31) / 32;
  for (int64_t tile_id = blockIdx.x; tile_id < tiles_total; tile_id += gridDim.x) {
    const int64_t b = tile_id / (tiles_r * tiles_c);
    const int64_t rem = tile_id % (tiles_r * tiles_c);
    const int64_t tr = rem / tiles_c, tc = rem % tiles_c;
    const T* s = src + b * rows * cols;
    T* d = dst + b * rows * cols;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = tr * 32 + ty + 8 * k, cc = tc * 32 + tx;
      if (r < rows && cc < cols) t[ty + 8 * k][tx] = Vec4<T>::load1(s, r * cols + cc);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t cc = tc * 32 + ty + 8 * k, r = tr * 32 + tx;
      if (r < rows && cc < cols) Vec4<T>::store1(d, cc * rows + r, t[tx][ty + 8 * k]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ host side
static inline int grid_for_warps(int64_t n_items, int warps_per_block, int max_blocks) {
  int64_t b = (n_items + warps_per_block - 1) / warps_per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

template <typename T>
static int forward_t(const void* depth, const void* feat, void* out, const int* rd, const int* rf, const int* rb,
                     const int* lengths, const int* starts, int64_t n_intervals, int c, cudaStream_t st) {
  if (n_intervals == 0) return 0;
  const int grid = grid_for_warps(n_intervals, kPoolWarps, kNumSMs * 32);
  const bool vec = (c % 4 == 0) && ((uintptr_t)feat % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (vec)
    pool_fwd_kernel<T><<<grid, kPoolThreads, 0, st>>>((const T*)depth, (const T*)feat, (T*)out, rd, rf, rb, starts,
                                                      lengths, n_intervals, c);
  else
    pool_fwd_scalar_kernel<T><<<grid, kPoolThreads, 0, st>>>((const T*)depth, (const T*)feat, (T*)out, rd, rf, rb,
                                                             starts, lengths, n_intervals, c);
  count_launch();
  return launch_status();
}

template <typename T>
static int backward_t(const void* og, void* dg, void* fg, const void* depth, const void* feat, const int* rd,
                      const int* rf, const int* rb, const int* lengths, const int* starts, int64_t n_intervals, int c,
                      cudaStream_t st) {
  if (n_intervals == 0) return 0;
  const int grid = grid_for_warps(n_intervals, kPoolWarps, kNumSMs * 32);
  const bool vec = (c % 4 == 0) && ((uintptr_t)feat % 16 == 0) && ((uintptr_t)og % 16 == 0) && ((uintptr_t)fg % 16 == 0);
  if (vec)
    pool_bwd_kernel<T><<<grid, kPoolThreads, 0, st>>>((const T*)og, (const T*)depth, (const T*)feat, rd, rf, rb,
                                                      starts, lengths, n_intervals, c, (T*)dg, (T*)fg);
  else
    pool_bwd_scalar_kernel<T><<<grid, kPoolThreads, 0, st>>>((const T*)og, (const T*)depth, (const T*)feat, rd, rf,
                                                             rb, starts, lengths, n_intervals, c, (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

template <typename T, int LAYOUT>
static int forward_dense_t(const void* depth, const void* feat, void* out, const int* rd, const int* rf, const int* rb,
                           const int* lengths, const int* starts, int64_t n_intervals, const int* counts_dev, int c,
                           int64_t n_voxels_total, int64_t vpf, int* strip_first, bool build_table, cudaStream_t st) {
  const int64_t frames = n_voxels_total / vpf;
  const int64_t spf = (vpf + kStrip - 1) / kStrip;
  const int64_t n_strips = frames * spf;
  if (build_table) {
    int64_t blocks = counts_dev ? (int64_t)kNumSMs * 8 : (n_intervals + 1 + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    strip_table_kernel<<<(int)blocks, 256, 0, st>>>(rb, starts, n_intervals, counts_dev, n_strips, spf, vpf, strip_first);
    count_launch();
  }
  const int cw = c < 128 ? c : 128;
  const size_t smem = sizeof(float) * (LAYOUT == BEVPOOL_LAYOUT_BCZYX ? (size_t)cw * (kStrip + 1) : (size_t)kStrip * cw);
  int grid = (int)(n_strips < (int64_t)kNumSMs * 64 ? n_strips : (int64_t)kNumSMs * 64);
  if (grid < 1) grid = 1;
  pool_fwd_dense_kernel<T, LAYOUT><<<grid, kDenseThreads, smem, st>>>(
      (const T*)depth, (const T*)feat, (T*)out, rd, rf, rb, starts, lengths, strip_first, n_strips, spf, vpf, c);
  count_launch();
  return launch_status();
}

template <typename T>
static int transpose_t(const void* src, void* dst, int b, int rows, int64_t cols, cudaStream_t st) {
  const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const int64_t total = (int64_t)b * tiles_c * tiles_r;
  if (total == 0) return 0;
  const int grid = (int)(total < (int64_t)kNumSMs * 64 ? total : (int64_t)kNumSMs * 64);
  grid_transpose_kernel<T><<<grid, 256, 0, st>>>((const T*)src, (T*)dst, rows, cols, tiles_c, total);
  count_launch();
  return launch_status();
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_v2_forward(const void* depth, const void* feat, void* out, const int32_t* ranks_depth,
                                  const int32_t* ranks_feat, const int32_t* ranks_bev,
                                  const int32_t* interval_lengths, const int32_t* interval_starts, int64_t n_points,
                                  int64_t n_intervals, int c, int dtype, void* stream) {
  if (n_intervals < 0 || n_points < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_intervals == 0) return BEVPOOL_OK;
  if (!depth || !feat || !out || !ranks_depth || !ranks_feat || !ranks_bev || !interval_lengths || !interval_starts)
    return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    return forward_t<float>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts,
                            n_intervals, c, st);
  if (dtype == BEVPOOL_BF16)
    return forward_t<__nv_bfloat16>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths,
                                    interval_starts, n_intervals, c, st);
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" int bevpool_v2_backward(const void* out_grad, void* depth_grad, void* feat_grad, const void* depth,
                                   const void* feat, const int32_t* ranks_depth, const int32_t* ranks_feat,
                                   const int32_t* ranks_bev, const int32_t* interval_lengths,
                                   const int32_t* interval_starts, int64_t n_points, int64_t n_intervals, int c,
                                   int dtype, void* stream) {
  if (n_intervals < 0 || n_points < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_intervals == 0) return BEVPOOL_OK;
  if (!out_grad || !depth_grad || !feat_grad || !depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev ||
      !interval_lengths || !interval_starts)
    return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    return backward_t<float>(out_grad, depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                             interval_lengths, interval_starts, n_intervals, c, st);
  if (dtype == BEVPOOL_BF16)
    return backward_t<__nv_bfloat16>(out_grad, depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                     interval_lengths, interval_starts, n_intervals, c, st);
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" size_t bevpool_v2_forward_dense_workspace_bytes(int64_t n_voxels_total, int64_t voxels_per_frame) {
  if (voxels_per_frame <= 0) return 0;
  const int64_t frames = n_voxels_total / voxels_per_frame;
  const int64_t spf = (voxels_per_frame + kStrip - 1) / kStrip;
  return (size_t)(frames * spf + 1) * sizeof(int32_t);
}

extern "C" int bevpool_v2_forward_dense(const void* depth, const void* feat, void* out, const int32_t* ranks_depth,
                                        const int32_t* ranks_feat, const int32_t* ranks_bev,
                                        const int32_t* interval_lengths, const int32_t* interval_starts,
                                        int64_t n_intervals, const int32_t* counts_dev, int c, int64_t n_voxels_total,
                                        int64_t voxels_per_frame, int layout, int dtype, void* workspace,
                                        size_t workspace_bytes, int build_table, void* stream) {
  if (n_intervals < 0 || n_voxels_total < 0 || voxels_per_frame <= 0 || n_voxels_total % voxels_per_frame)
    return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_voxels_total == 0) return BEVPOOL_OK;
  if (!depth || !feat || !out || !workspace) return BEVPOOL_ERR_BAD_ARG;
  if ((n_intervals > 0 || counts_dev) &&
      (!ranks_depth || !ranks_feat || !ranks_bev || !interval_lengths || !interval_starts))
    return BEVPOOL_ERR_BAD_ARG;
  if (workspace_bytes < bevpool_v2_forward_dense_workspace_bytes(n_voxels_total, voxels_per_frame))
    return BEVPOOL_ERR_WORKSPACE;
  if (layout != BEVPOOL_LAYOUT_BZYXC && layout != BEVPOOL_LAYOUT_BCZYX) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int* tbl = (int*)workspace;
#define DISPATCH(T, L)                                                                                           \
  return forward_dense_t<T, L>(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths,           \
                               interval_starts, n_intervals, counts_dev, c, n_voxels_total, voxels_per_frame, tbl,     \
                               build_table != 0, st)
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
                                         const void* feat, const int32_t* point_rank, int bn, int d, int hw, int c,
                                         int dtype, void* stream) {
  if (bn < 0 || d <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4) return BEVPOOL_ERR_BAD_CHANNELS;
  const int64_t n_pixels = (int64_t)bn * hw;
  if (n_pixels == 0) return BEVPOOL_OK;
  if (!out_grad || !depth_grad || !feat_grad || !depth || !feat || !point_rank) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_warps(n_pixels, kPoolWarps, kNumSMs * 32);
  if (dtype == BEVPOOL_F32)
    pool_bwd_dense_kernel<float><<<grid, kPoolThreads, 0, st>>>((const float*)out_grad, (const float*)depth,
                                                                (const float*)feat, point_rank, n_pixels, d, hw, c,
                                                                (float*)depth_grad, (float*)feat_grad);
  else if (dtype == BEVPOOL_BF16)
    pool_bwd_dense_kernel<__nv_bfloat16><<<grid, kPoolThreads, 0, st>>>(
        (const __nv_bfloat16*)out_grad, (const __nv_bfloat16*)depth, (const __nv_bfloat16*)feat, point_rank, n_pixels,
        d, hw, c, (__nv_bfloat16*)depth_grad, (__nv_bfloat16*)feat_grad);
  else
    return BEVPOOL_ERR_BAD_ARG;
  count_launch();
  return launch_status();
}

extern "C" int bevpool_grid_transpose(const void* src, void* dst, int b, int c, int64_t zyx, int to_channels_last,
                                      int dtype, void* stream) {
  if (b < 0 || c <= 0 || zyx < 0) return BEVPOOL_ERR_BAD_ARG;
  if ((int64_t)b * zyx == 0) return BEVPOOL_OK;
  if (!src || !dst) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // BCZYX is [B][c][zyx]; BZYXC is [B][zyx][c]
  if (dtype == BEVPOOL_F32) {
    if (to_channels_last) return transpose_t<float>(src, dst, b, c, zyx, st);
    if (zyx > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
    return transpose_t<float>(src, dst, b, (int)zyx, c, st);
  }
  if (dtype == BEVPOOL_BF16) {
    if (to_channels_last) return transpose_t<__nv_bfloat16>(src, dst, b, c, zyx, st);
    if (zyx > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
    return transpose_t<__nv_bfloat16>(src, dst, b, (int)zyx, c, st);
  }
  return BEVPOOL_ERR_BAD_ARG;
}
