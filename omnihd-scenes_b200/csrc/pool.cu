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
// S = channel sweeps of 128 (C <= 128 * S): the pixel's whole feature row and feat_grad row live in registers, so a
// point's dot product is completed in ONE pass and depth_grad is written exactly once (never re-read, never
// rounded to the io type between sweeps).
template <typename T, int S>
__global__ void __launch_bounds__(kPoolThreads)
pool_bwd_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                const int* __restrict__ rd, const int* __restrict__ rf, const int* __restrict__ rb,
                const int* __restrict__ starts, const int* __restrict__ lengths, int64_t n_intervals, int c,
                T* __restrict__ depth_grad, T* __restrict__ feat_grad) {
  constexpr int U = S == 1 ? 4 : (S == 2 ? 2 : 1);   // points in flight per step (register budget)
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int c4 = c >> 2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = __ldg(starts + k), len = __ldg(lengths + k);
    const int64_t fbase = (int64_t)__ldg(rf + s) * c;
    float4 fv[S], fg[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      fv[q] = (32 * q + lane < c4) ? Vec4<T>::load(feat, fbase + 4 * (32 * q + lane)) : zero;
      fg[q] = zero;
    }
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
      for (; i + U <= n; i += U) {
        float4 g[U][S];
        float d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int b = __shfl_sync(kFullMask, my_rb, i + u);
          d[u] = __shfl_sync(kFullMask, my_d, i + u);
#pragma unroll
          for (int q = 0; q < S; ++q)
            g[u][q] = (32 * q + lane < c4) ? Vec4<T>::load(og, (int64_t)b * c + 4 * (32 * q + lane)) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float dotp = 0.f;
#pragma unroll
          for (int q = 0; q < S; ++q) {
            fg[q] = fma4(g[u][q], d[u], fg[q]);
            dotp = dot4(g[u][q], fv[q], dotp);
          }
          const float pz = warp_sum(dotp);
          if (lane == i + u) my_dg = pz;
        }
      }
      for (; i < n; ++i) {
        const int b = __shfl_sync(kFullMask, my_rb, i);
        const float d = __shfl_sync(kFullMask, my_d, i);
        float dotp = 0.f;
#pragma unroll
        for (int q = 0; q < S; ++q) {
          const float4 g = (32 * q + lane < c4) ? Vec4<T>::load(og, (int64_t)b * c + 4 * (32 * q + lane)) : zero;
          fg[q] = fma4(g, d, fg[q]);
          dotp = dot4(g, fv[q], dotp);
        }
        const float pz = warp_sum(dotp);
        if (lane == i) my_dg = pz;
      }
      if (base + lane < len) Vec4<T>::store1(depth_grad, my_rd, my_dg);
    }
#pragma unroll
    for (int q = 0; q < S; ++q)
      if (32 * q + lane < c4) Vec4<T>::store(feat_grad, fbase + 4 * (32 * q + lane), fg[q]);
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

// [B][C][S] -> [B][S][C] for C % 4 == 0, C <= 128: a CTA moves ALL channels of 64 consecutive positions, so
// both sides are 128-bit and fully coalesced (reads: 256-byte row pieces, writes: 64*C*4 contiguous bytes).
constexpr int kTcCols = 64;
template <typename T>
__global__ void __launch_bounds__(256)
transpose_to_channels_last_kernel(const T* __restrict__ src, T* __restrict__ dst, int c, int64_t cols,
                                  int64_t tiles_per_batch) {
  extern __shared__ float t[];   // [c][kTcCols + 1]
  pdl_wait();
  const int64_t b = blockIdx.x / tiles_per_batch;
  const int64_t col0 = (blockIdx.x % tiles_per_batch) * kTcCols;
  const int ncol = (int)min((int64_t)kTcCols, cols - col0);
  const T* s = src + (b * c) * cols + col0;
  T* d = dst + (b * cols + col0) * c;
  const bool vec = (ncol == kTcCols) && ((cols & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
  if (vec) {
    for (int i = threadIdx.x; i < c * (kTcCols / 4); i += 256) {
      const int ch = i / (kTcCols / 4), q = i % (kTcCols / 4);
      const float4 v = Vec4<T>::load_stream(s, (int64_t)ch * cols + 4 * q);
      float* o = t + ch * (kTcCols + 1) + 4 * q;
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < c * kTcCols; i += 256) {
      const int ch = i / kTcCols, q = i % kTcCols;
      if (q < ncol) t[ch * (kTcCols + 1) + q] = Vec4<T>::load1(s, (int64_t)ch * cols + q);
    }
  }
  __syncthreads();
  const int c4 = c >> 2;
  for (int i = threadIdx.x; i < ncol * c4; i += 256) {
    const int col = i / c4, q = i % c4;
    const float* p = t + (4 * q) * (kTcCols + 1) + col;
    Vec4<T>::store_keep(d, (int64_t)col * c + 4 * q,
                        make_float4(p[0], p[kTcCols + 1], p[2 * (kTcCols + 1)], p[3 * (kTcCols + 1)]));
  }
}

// ------------------------------------------------------------------------------------------ v1 bev_pool
// ops/bev_pool (MIT-BEVFusion style): features are already multiplied by depth and SORTED by voxel rank, so an
// interval is a run of consecutive rows of x. out[b][d][h][w][:] = sum of the interval's rows, with
// (h, w, d, b) = geom_feats[start] (note the index order of the reference kernel, bev_pool_cuda.cu:36-38).
// One warp per interval, lanes span channels as 128-bit vectors, rows are streamed (each is read exactly once).
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_v1_fwd_kernel(const T* __restrict__ x, const int* __restrict__ geom, const int* __restrict__ starts,
                   const int* __restrict__ lengths, int64_t n_intervals, int d, int h, int w, int c, T* __restrict__ out) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (c & 3) == 0;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = __ldg(starts + k), len = __ldg(lengths + k);
    const int4 g = *reinterpret_cast<const int4*>(geom + 4 * (int64_t)s);
    const int64_t obase = ((((int64_t)g.w * d + g.z) * h + g.x) * w + g.y) * c;
    if (vec) {
      for (int ch = 4 * lane; ch < c; ch += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int i = 0;
        for (; i + 4 <= len; i += 4) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = Vec4<T>::load_stream(x, (int64_t)(s + i + u) * c + ch);
#pragma unroll
          for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; i < len; ++i) {
          const float4 v = Vec4<T>::load_stream(x, (int64_t)(s + i) * c + ch);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        Vec4<T>::store(out, obase + ch, acc);
      }
    } else {
      for (int ch = lane; ch < c; ch += 32) {
        float acc = 0.f;
        for (int i = 0; i < len; ++i) acc += Vec4<T>::load1(x, (int64_t)(s + i) * c + ch);
        Vec4<T>::store1(out, obase + ch, acc);
      }
    }
  }
}

// x_grad[row] = out_grad[voxel of the row's interval]: one row load, len row stores.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_v1_bwd_kernel(const T* __restrict__ og, const int* __restrict__ geom, const int* __restrict__ starts,
                   const int* __restrict__ lengths, int64_t n_intervals, int d, int h, int w, int c,
                   T* __restrict__ x_grad) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (c & 3) == 0;
  for (int64_t k = warp0; k < n_intervals; k += n_warps) {
    const int s = __ldg(starts + k), len = __ldg(lengths + k);
    const int4 g = *reinterpret_cast<const int4*>(geom + 4 * (int64_t)s);
    const int64_t obase = ((((int64_t)g.w * d + g.z) * h + g.x) * w + g.y) * c;
    if (vec) {
      for (int ch = 4 * lane; ch < c; ch += 128) {
        const float4 v = Vec4<T>::load(og, obase + ch);
        for (int i = 0; i < len; ++i) Vec4<T>::store(x_grad, (int64_t)(s + i) * c + ch, v);
      }
    } else {
      for (int ch = lane; ch < c; ch += 32) {
        const float v = Vec4<T>::load1(og, obase + ch);
        for (int i = 0; i < len; ++i) Vec4<T>::store1(x_grad, (int64_t)(s + i) * c + ch, v);
      }
    }
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
  if (vec && c <= 1024) {
    auto kern = c <= 128 ? pool_bwd_kernel<T, 1> : (c <= 256 ? pool_bwd_kernel<T, 2> : (c <= 512 ? pool_bwd_kernel<T, 4> : pool_bwd_kernel<T, 8>));
    kern<<<grid, kPoolThreads, 0, st>>>((const T*)og, (const T*)depth, (const T*)feat, rd, rf, rb,
                                       starts, lengths, n_intervals, c, (T*)dg, (T*)fg);
  }
  else
    pool_bwd_scalar_kernel<T><<<grid, kPoolThreads, 0, st>>>((const T*)og, (const T*)depth, (const T*)feat, rd, rf,
                                                             rb, starts, lengths, n_intervals, c, (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

template <typename T>
static int transpose_cl_t(const void* src, void* dst, int b, int c, int64_t cols, cudaStream_t st) {
  const int64_t tiles_per_batch = (cols + kTcCols - 1) / kTcCols;
  const int64_t total = (int64_t)b * tiles_per_batch;
  if (total == 0) return 0;
  if (total > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const size_t smem = sizeof(float) * (size_t)c * (kTcCols + 1);
  if (int rc = ensure_dynamic_smem(transpose_to_channels_last_kernel<T>, smem)) return rc;
  launch_pdl(transpose_to_channels_last_kernel<T>, dim3((unsigned)total), dim3(256), smem, st, (const T*)src, (T*)dst, c, cols,
             tiles_per_batch);
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

extern "C" int bevpool_grid_transpose(const void* src, void* dst, int b, int c, int64_t zyx, int to_channels_last,
                                      int dtype, void* stream) {
  if (b < 0 || c <= 0 || zyx < 0) return BEVPOOL_ERR_BAD_ARG;
  if ((int64_t)b * zyx == 0) return BEVPOOL_OK;
  if (!src || !dst) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // BCZYX is [B][c][zyx]; BZYXC is [B][zyx][c]
  const bool fast = to_channels_last && c % 4 == 0 && c <= 128 && (((uintptr_t)dst) & 15) == 0;
  if (dtype == BEVPOOL_F32) {
    if (fast) return transpose_cl_t<float>(src, dst, b, c, zyx, st);
    if (to_channels_last) return transpose_t<float>(src, dst, b, c, zyx, st);
    if (zyx > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
    return transpose_t<float>(src, dst, b, (int)zyx, c, st);
  }
  if (dtype == BEVPOOL_BF16) {
    if (fast && (zyx & 3) == 0) return transpose_cl_t<__nv_bfloat16>(src, dst, b, c, zyx, st);
    if (to_channels_last) return transpose_t<__nv_bfloat16>(src, dst, b, c, zyx, st);
    if (zyx > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
    return transpose_t<__nv_bfloat16>(src, dst, b, (int)zyx, c, st);
  }
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" int bevpool_v1_forward(const void* x, const int32_t* geom_feats, const int32_t* interval_lengths,
                                  const int32_t* interval_starts, void* out, int b, int d, int h, int w, int64_t n,
                                  int64_t n_intervals, int c, int dtype, void* stream) {
  if (b < 0 || d < 0 || h < 0 || w < 0 || n < 0 || n_intervals < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_intervals == 0) return BEVPOOL_OK;
  if (!x || !geom_feats || !interval_lengths || !interval_starts || !out) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)geom_feats % 16 || (c % 4 == 0 && ((uintptr_t)x % 16 || (uintptr_t)out % 16))) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_warps(n_intervals, kPoolWarps, kNumSMs * 32);
  if (dtype == BEVPOOL_F32)
    pool_v1_fwd_kernel<float><<<grid, kPoolThreads, 0, st>>>((const float*)x, geom_feats, interval_starts, interval_lengths,
                                                              n_intervals, d, h, w, c, (float*)out);
  else if (dtype == BEVPOOL_BF16)
    pool_v1_fwd_kernel<__nv_bfloat16><<<grid, kPoolThreads, 0, st>>>((const __nv_bfloat16*)x, geom_feats, interval_starts,
                                                                      interval_lengths, n_intervals, d, h, w, c,
                                                                      (__nv_bfloat16*)out);
  else
    return BEVPOOL_ERR_BAD_ARG;
  count_launch();
  return launch_status();
}

extern "C" int bevpool_v1_backward(const void* out_grad, const int32_t* geom_feats, const int32_t* interval_lengths,
                                   const int32_t* interval_starts, void* x_grad, int b, int d, int h, int w, int64_t n,
                                   int64_t n_intervals, int c, int dtype, void* stream) {
  if (b < 0 || d < 0 || h < 0 || w < 0 || n < 0 || n_intervals < 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0) return BEVPOOL_ERR_BAD_CHANNELS;
  if (n_intervals == 0) return BEVPOOL_OK;
  if (!out_grad || !geom_feats || !interval_lengths || !interval_starts || !x_grad) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)geom_feats % 16 || (c % 4 == 0 && ((uintptr_t)x_grad % 16 || (uintptr_t)out_grad % 16)))
    return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_warps(n_intervals, kPoolWarps, kNumSMs * 32);
  if (dtype == BEVPOOL_F32)
    pool_v1_bwd_kernel<float><<<grid, kPoolThreads, 0, st>>>((const float*)out_grad, geom_feats, interval_starts,
                                                              interval_lengths, n_intervals, d, h, w, c, (float*)x_grad);
  else if (dtype == BEVPOOL_BF16)
    pool_v1_bwd_kernel<__nv_bfloat16><<<grid, kPoolThreads, 0, st>>>((const __nv_bfloat16*)out_grad, geom_feats,
                                                                      interval_starts, interval_lengths, n_intervals, d, h,
                                                                      w, c, (__nv_bfloat16*)x_grad);
  else
    return BEVPOOL_ERR_BAD_ARG;
  count_launch();
  return launch_status();
}
