// Frustum geometry, voxel ranking, stable LSD radix sort and run-length segmentation for
// voxel_pooling_prepare_v2, plus the backward regrouping by ranks_feat — all on the device,
// no host synchronisation, bit-exact against the reference's CPU results.
//
// Pipeline of bevpool_prepare_v2 (one stream, 2 + n_passes + 2 kernels):
//   memset(zero region)                       histograms, look-back status words, tickets
//   point_rank_kernel      P0 threads         geometry (optional) -> voxel rank | -1, digit
//                                             histograms of every pass, kept-point count
//   radix_scatter_kernel   x n_passes         one-sweep stable scatter: warp-level match ranking,
//                                             decoupled look-back across tiles; the first pass
//                                             also compacts (dropped points are never written)
//   segment_heads_kernel   P threads          head flags (ballot) + scan -> interval_starts,
//                                             ranks_feat derived from ranks_depth
//   interval_lengths_kernel                   adjacent difference of starts
//
// Exactness (SURVEY.md §7 hard part 3): the voxel index is trunc((coor - lo) / dx) with an IEEE
// fp32 subtract and divide (__fsub_rn / __fdiv_rn; no reciprocal, no FMA), the geometry is
// r0*x + r1*y + r2*z + t with every product and sum rounded separately, ties keep ascending
// point index because every pass is stable.
#include "common.cuh"

namespace bevpool {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                          // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;   // keys per CTA
constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kMaxPasses = 4;

constexpr int kHeadThreads = 256;
constexpr int kHeadItems = 4;
constexpr int kHeadTile = kHeadThreads * kHeadItems;

struct SortPlan {
  int n_passes;
  int shift[kMaxPasses];
  int bits[kMaxPasses];
};

static SortPlan make_plan(int64_t max_key) {
  int total_bits = 1;
  while (total_bits < 31 && ((int64_t)1 << total_bits) <= max_key) ++total_bits;
  SortPlan p;
  p.n_passes = (total_bits + kRadixBits - 1) / kRadixBits;
  // spread the bits evenly over the passes (e.g. 17 bits -> 6+6+5)
  int left = total_bits, shift = 0;
  for (int i = 0; i < p.n_passes; ++i) {
    const int b = (left + (p.n_passes - i) - 1) / (p.n_passes - i);
    p.shift[i] = shift;
    p.bits[i] = b;
    shift += b;
    left -= b;
  }
  for (int i = p.n_passes; i < kMaxPasses; ++i) p.shift[i] = p.bits[i] = 0;
  return p;
}

// ------------------------------------------------------------------------------------------ block scan
// Exclusive scan of one value per thread over a 256-thread CTA. `total` is valid in all threads.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* warp_sums /*[8]*/, uint32_t* total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFullMask, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const uint32_t s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// ------------------------------------------------------------------------------------------ geometry
struct Cam {
  float r[9];
  float t[3];
};

__device__ __forceinline__ void cam_point(const float* __restrict__ frustum, const Cam& cam, int64_t dhw,
                                          float& x, float& y, float& z) {
  const float u = __ldg(frustum + 3 * dhw + 0), v = __ldg(frustum + 3 * dhw + 1), dd = __ldg(frustum + 3 * dhw + 2);
  const float px = __fmul_rn(u, dd), py = __fmul_rn(v, dd), pz = dd;
  x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam.r[0], px), __fmul_rn(cam.r[1], py)), __fmul_rn(cam.r[2], pz)), cam.t[0]);
  y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam.r[3], px), __fmul_rn(cam.r[4], py)), __fmul_rn(cam.r[5], pz)), cam.t[1]);
  z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam.r[6], px), __fmul_rn(cam.r[7], py)), __fmul_rn(cam.r[8], pz)), cam.t[2]);
}

__device__ __forceinline__ Cam load_cam(const float* __restrict__ rots, const float* __restrict__ trans, int64_t bn) {
  Cam c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.r[i] = __ldg(rots + bn * 9 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = __ldg(trans + bn * 3 + i);
  return c;
}

// grid: (ceil(DHW / 256), BN): every CTA works inside one camera, so R|t are CTA-uniform
__global__ void __launch_bounds__(256)
geometry_kernel(const float* __restrict__ frustum, const float* __restrict__ rots, const float* __restrict__ trans,
                float* __restrict__ coor, int64_t dhw_total) {
  const int64_t bn = blockIdx.y;
  const Cam cam = load_cam(rots, trans, bn);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dhw_total; i += (int64_t)gridDim.x * blockDim.x) {
    float x, y, z;
    cam_point(frustum, cam, i, x, y, z);
    float* o = coor + 3 * (bn * dhw_total + i);
    o[0] = x;
    o[1] = y;
    o[2] = z;
  }
}

// ------------------------------------------------------------------------------------------ voxel rank
struct GridDev {
  int64_t dhw;         // D*H*W
  int n_cams;          // N
  int nx, ny, nz;
  float lo[3], dx[3];
};

__device__ __forceinline__ bool voxel_index(float c, float lo, float dx, int n, int& v) {
  const float q = __fdiv_rn(__fsub_rn(c, lo), dx);
  // .long(): truncation toward zero; NaN / beyond int64 come out as INT64_MIN on the CPU -> dropped
  if (!(fabsf(q) < 9.0e18f)) return false;
  const long long t = (long long)q;
  v = (int)t;
  return t >= 0 && t < n;
}

template <bool FROM_COOR>
__global__ void __launch_bounds__(256)
point_rank_kernel(const float* __restrict__ coor, const float* __restrict__ frustum, const float* __restrict__ rots,
                  const float* __restrict__ trans, GridDev g, SortPlan plan, int* __restrict__ point_rank,
                  uint32_t* __restrict__ hist /*[kMaxPasses][256]*/, int* __restrict__ n_kept) {
  __shared__ uint32_t sh[kMaxPasses][kRadixBins];
  __shared__ uint32_t s_kept;
  for (int i = threadIdx.x; i < kMaxPasses * kRadixBins; i += blockDim.x) (&sh[0][0])[i] = 0;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  const int64_t bn = blockIdx.y;
  const int64_t frame = bn / g.n_cams;
  Cam cam;
  if (!FROM_COOR) cam = load_cam(rots, trans, bn);
  const int64_t vpf = (int64_t)g.nx * g.ny * g.nz;
  uint32_t kept = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.dhw; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = bn * g.dhw + i;
    float x, y, z;
    if (FROM_COOR) {
      x = ldg_stream_f32(coor + 3 * idx + 0);
      y = ldg_stream_f32(coor + 3 * idx + 1);
      z = ldg_stream_f32(coor + 3 * idx + 2);
    } else {
      cam_point(frustum, cam, i, x, y, z);
    }
    int vx, vy, vz;
    const bool ok = voxel_index(x, g.lo[0], g.dx[0], g.nx, vx) & voxel_index(y, g.lo[1], g.dx[1], g.ny, vy) &
                    voxel_index(z, g.lo[2], g.dx[2], g.nz, vz);
    int rank = -1;
    if (ok) {
      rank = (int)(frame * vpf + ((int64_t)vz * g.ny + vy) * g.nx + vx);
      ++kept;
#pragma unroll
      for (int p = 0; p < kMaxPasses; ++p)
        if (p < plan.n_passes) atomicAdd(&sh[p][(rank >> plan.shift[p]) & ((1 << plan.bits[p]) - 1)], 1u);
    }
    point_rank[idx] = rank;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(kFullMask, kept, o);
  if (lane_id() == 0 && kept) atomicAdd(&s_kept, kept);
  __syncthreads();
  for (int i = threadIdx.x; i < plan.n_passes * kRadixBins; i += blockDim.x) {
    const uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(hist + i, v);
  }
  if (threadIdx.x == 0 && s_kept) atomicAdd(n_kept, (int)s_kept);
}

// Digit histograms of every pass for an existing key array (backward regroup).
__global__ void __launch_bounds__(256)
key_hist_kernel(const int* __restrict__ keys, int64_t n, SortPlan plan, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kMaxPasses][kRadixBins];
  for (int i = threadIdx.x; i < kMaxPasses * kRadixBins; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = ldg_stream_i32(keys + i);
    if (k < 0) continue;
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p)
      if (p < plan.n_passes) atomicAdd(&sh[p][(k >> plan.shift[p]) & ((1 << plan.bits[p]) - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.n_passes * kRadixBins; i += blockDim.x) {
    const uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(hist + i, v);
  }
}

// ------------------------------------------------------------------------------------------ radix pass
// One stable LSD pass in a single sweep. Keys < 0 are dropped (first pass = compaction).
// vals_in == nullptr means "value = position" (first pass of an argsort).
// Tile order is handed out by an atomic ticket so look-back never waits on an unscheduled CTA.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int* __restrict__ keys_out,
                     int* __restrict__ vals_out, const uint32_t* __restrict__ digit_hist, uint32_t* __restrict__ status,
                     uint32_t* __restrict__ ticket, const int* __restrict__ n_dev, int64_t n_host, int shift, int bits) {
  __shared__ uint32_t warp_hist[kSortWarps][kRadixBins];
  __shared__ uint32_t digit_base[kRadixBins];
  __shared__ uint32_t scan_tmp[8];
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = 1 << bits;
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  for (int i = tid; i < kSortWarps * kRadixBins; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t tile_base = (int64_t)tile * kSortTile;
  if (tile_base >= n) return;

  int key[kSortItems], val[kSortItems];
  uint32_t rank[kSortItems];
  const int64_t base = tile_base + warp * (32 * kSortItems) + lane;
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t idx = base + j * 32;
    key[j] = idx < n ? ldg_stream_i32(keys_in + idx) : -1;
    val[j] = vals_in ? (idx < n ? ldg_stream_i32(vals_in + idx) : 0) : (int)idx;
  }
  // warp-level ranking: lanes with the same digit find each other with match.any
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const bool valid = key[j] >= 0;
    const int d = valid ? ((key[j] >> shift) & (bins - 1)) : bins;
    const unsigned peers = __match_any_sync(kFullMask, d);
    const int leader = __ffs(peers) - 1;
    uint32_t b = 0;
    if (valid && lane == leader) {
      b = warp_hist[warp][d];
      warp_hist[warp][d] = b + __popc(peers);
    }
    b = __shfl_sync(kFullMask, b, leader);
    rank[j] = b + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan of the global digit histogram (bins <= 256: one digit per thread)
  uint32_t total;
  const uint32_t digit_start = block_exclusive_scan_256(tid < bins ? __ldg(digit_hist + tid) : 0u, scan_tmp, &total);
  if (tid < bins) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = warp_hist[w][tid];
      warp_hist[w][tid] = run;
      run += t;
    }
    const uint32_t excl = lookback_exclusive(status + tid, bins, tile, run);
    digit_base[tid] = digit_start + excl;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    if (key[j] >= 0) {
      const int d = (key[j] >> shift) & (bins - 1);
      const uint32_t pos = digit_base[d] + warp_hist[warp][d] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
    }
  }
}

// ------------------------------------------------------------------------------------------ segmentation
// keys are sorted; a head is a position whose key differs from its predecessor. Heads are found per
// thread (4 consecutive keys, one 128-bit load), counted with a CTA scan, and given global slots by
// decoupled look-back. MODE 0 (prepare): also derives ranks_feat from ranks_depth.
// MODE 1 (regroup): vals are positions into the original arrays; gathers the three rank arrays.
template <int MODE>
__global__ void __launch_bounds__(kHeadThreads)
segment_heads_kernel(const int* __restrict__ keys, const int* __restrict__ vals, const int* __restrict__ n_dev,
                     int64_t n_host, int* __restrict__ starts, int* __restrict__ n_heads_out,
                     uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                     // MODE 0
                     int* __restrict__ ranks_feat, int dhw, int hw,
                     // MODE 1
                     const int* __restrict__ src_rd, const int* __restrict__ src_rb, int* __restrict__ dst_rd,
                     int* __restrict__ dst_rb) {
  __shared__ uint32_t scan_tmp[8];
  __shared__ int s_tile;
  __shared__ uint32_t s_prefix;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t tile_base = (int64_t)tile * kHeadTile;
  if (tile_base >= n) return;
  const int64_t i0 = tile_base + (int64_t)tid * kHeadItems;
  int k[kHeadItems], v[kHeadItems];
  const bool vec_ok = ((((uintptr_t)keys) | ((uintptr_t)vals)) & 15) == 0;   // caller buffers may be unaligned views
  if (vec_ok && i0 + kHeadItems <= n) {
    const int4 kk = *reinterpret_cast<const int4*>(keys + i0);
    const int4 vv = *reinterpret_cast<const int4*>(vals + i0);
    k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
    v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
  } else {
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j) {
      k[j] = (i0 + j < n) ? keys[i0 + j] : -1;
      v[j] = (i0 + j < n) ? vals[i0 + j] : 0;
    }
  }
  const int prev = (i0 > 0 && i0 < n) ? keys[i0 - 1] : -1;
  bool head[kHeadItems];
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kHeadItems; ++j) {
    head[j] = (i0 + j < n) && (k[j] != (j == 0 ? prev : k[j - 1]) || (i0 + j == 0));
    cnt += head[j];
  }
  uint32_t total;
  uint32_t local = block_exclusive_scan_256(cnt, scan_tmp, &total);
  if (tid == 0) {
    s_prefix = lookback_exclusive(status, 1, tile, total);
    if (tile_base + kHeadTile >= n) *n_heads_out = (int)(s_prefix + total);  // last tile publishes the count
  }
  __syncthreads();
  uint32_t slot = s_prefix + local;
#pragma unroll
  for (int j = 0; j < kHeadItems; ++j)
    if (head[j]) starts[slot++] = (int)(i0 + j);

  if (MODE == 0) {
    int f[kHeadItems];
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j) f[j] = (v[j] / dhw) * hw + v[j] % hw;
    if ((((uintptr_t)ranks_feat) & 15) == 0 && i0 + kHeadItems <= n) {
      *reinterpret_cast<int4*>(ranks_feat + i0) = make_int4(f[0], f[1], f[2], f[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kHeadItems; ++j)
        if (i0 + j < n) ranks_feat[i0 + j] = f[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j)
      if (i0 + j < n) {
        dst_rd[i0 + j] = __ldg(src_rd + v[j]);
        dst_rb[i0 + j] = __ldg(src_rb + v[j]);
      }
  }
}

__global__ void interval_lengths_kernel(const int* __restrict__ starts, const int* __restrict__ n_heads_dev,
                                        const int* __restrict__ n_dev, int64_t n_host, int* __restrict__ lengths) {
  const int n_heads = *n_heads_dev;
  const int n = n_dev ? *n_dev : (int)n_host;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_heads; j += (int64_t)gridDim.x * blockDim.x)
    lengths[j] = (j + 1 < n_heads ? starts[j + 1] : n) - starts[j];
}

// ------------------------------------------------------------------------------------------ host side
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SortWorkspace {
  // zeroed region
  uint32_t* hist;        // [kMaxPasses][256]
  uint32_t* tickets;     // [kMaxPasses + 1]
  uint32_t* status;      // [n_passes][tiles][256] then [head tiles]
  size_t zero_bytes;
  // scratch
  int* alt_keys;
  int* alt_vals;
  size_t total_bytes;
  int64_t sort_tiles, head_tiles;
};

static SortWorkspace carve(void* ws, int64_t n_max, int n_passes) {
  SortWorkspace w;
  w.sort_tiles = (n_max + kSortTile - 1) / kSortTile;
  w.head_tiles = (n_max + kHeadTile - 1) / kHeadTile;
  char* p = (char*)ws;
  size_t off = 0;
  w.hist = (uint32_t*)(p + off);
  off += sizeof(uint32_t) * kMaxPasses * kRadixBins;
  w.tickets = (uint32_t*)(p + off);
  off += sizeof(uint32_t) * 8;
  w.status = (uint32_t*)(p + off);
  off += sizeof(uint32_t) * ((size_t)n_passes * w.sort_tiles * kRadixBins + w.head_tiles);
  off = align_up(off, 256);
  w.zero_bytes = off;
  w.alt_keys = (int*)(p + off);
  off += align_up(sizeof(int) * (size_t)n_max, 256);
  w.alt_vals = (int*)(p + off);
  off += align_up(sizeof(int) * (size_t)n_max, 256);
  w.total_bytes = off;
  return w;
}

// Runs the scatter passes. Source of pass 0 is (keys0, vals0 or positions); the last pass lands in
// (out_keys, out_vals). n for pass 0 is n0 (host); later passes read the compacted count from n_dev
// (or n0 when n_dev is null).
static void run_passes(const SortPlan& plan, const SortWorkspace& w, const int* keys0, const int* vals0, int64_t n0,
                       const int* n_dev, int* out_keys, int* out_vals, cudaStream_t st) {
  const int* src_k = keys0;
  const int* src_v = vals0;
  for (int p = 0; p < plan.n_passes; ++p) {
    const bool to_out = ((plan.n_passes - 1 - p) % 2) == 0;
    int* dst_k = to_out ? out_keys : w.alt_keys;
    int* dst_v = to_out ? out_vals : w.alt_vals;
    radix_scatter_kernel<<<(unsigned)w.sort_tiles, kSortThreads, 0, st>>>(
        src_k, src_v, dst_k, dst_v, w.hist + p * kRadixBins, w.status + (size_t)p * w.sort_tiles * kRadixBins,
        w.tickets + p, p == 0 ? nullptr : n_dev, n0, plan.shift[p], plan.bits[p]);
    count_launch();
    src_k = dst_k;
    src_v = dst_v;
  }
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_geometry(const float* frustum, const float* rots, const float* trans, float* coor, int bn,
                                int d, int hw, void* stream) {
  if (bn < 0 || d < 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t dhw = (int64_t)d * hw;
  if (bn == 0 || dhw == 0) return BEVPOOL_OK;
  if (!frustum || !rots || !trans || !coor) return BEVPOOL_ERR_BAD_ARG;
  if (bn > 65535) return BEVPOOL_ERR_BAD_ARG;
  int bx = (int)((dhw + 255) / 256);
  if (bx > 4096) bx = 4096;
  geometry_kernel<<<dim3(bx, bn), 256, 0, (cudaStream_t)stream>>>(frustum, rots, trans, coor, dhw);
  count_launch();
  return launch_status();
}

static int check_grid(const bevpool_grid_t* g, int64_t* p0, int64_t* total_voxels) {
  if (!g) return BEVPOOL_ERR_BAD_ARG;
  if (g->b < 0 || g->n < 0 || g->d < 0 || g->h < 0 || g->w < 0) return BEVPOOL_ERR_BAD_ARG;
  if (g->nx[0] <= 0 || g->nx[1] <= 0 || g->nx[2] <= 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t n = (int64_t)g->b * g->n * g->d * g->h * g->w;
  const int64_t v = (int64_t)g->b * g->nx[0] * g->nx[1] * g->nx[2];
  if (n >= ((int64_t)1 << 30) || v >= ((int64_t)1 << 31) - 1) return BEVPOOL_ERR_OVERFLOW;  // int32 ranks, 30-bit look-back counters
  if ((int64_t)g->b * g->n > 65535) return BEVPOOL_ERR_BAD_ARG;
  *p0 = n;
  *total_voxels = v;
  return BEVPOOL_OK;
}

extern "C" size_t bevpool_prepare_v2_workspace_bytes(const bevpool_grid_t* g) {
  int64_t p0, v;
  if (check_grid(g, &p0, &v) != BEVPOOL_OK) return 0;
  if (p0 == 0) return 256;
  const SortPlan plan = make_plan(v > 0 ? v - 1 : 0);
  // + a private point_rank array in case the caller does not want one
  return carve(nullptr, p0, plan.n_passes).total_bytes + align_up(sizeof(int) * (size_t)p0, 256);
}

extern "C" int bevpool_prepare_v2(const float* coor, const float* frustum, const float* rots, const float* trans,
                                  const bevpool_grid_t* g, int32_t* ranks_bev, int32_t* ranks_depth,
                                  int32_t* ranks_feat, int32_t* interval_starts, int32_t* interval_lengths,
                                  int32_t* counts_dev, int32_t* point_rank, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  int64_t p0, v;
  int rc = check_grid(g, &p0, &v);
  if (rc != BEVPOOL_OK) return rc;
  if (!counts_dev) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int32_t), st);
  if (p0 == 0) return launch_status();
  if (!ranks_bev || !ranks_depth || !ranks_feat || !interval_starts || !interval_lengths || !workspace)
    return BEVPOOL_ERR_BAD_ARG;
  if (!coor && (!frustum || !rots || !trans)) return BEVPOOL_ERR_BAD_ARG;
  if (workspace_bytes < bevpool_prepare_v2_workspace_bytes(g)) return BEVPOOL_ERR_WORKSPACE;

  const SortPlan plan = make_plan(v - 1);
  const SortWorkspace w = carve(workspace, p0, plan.n_passes);
  if (!point_rank) point_rank = (int*)((char*)workspace + w.total_bytes);
  cudaMemsetAsync(workspace, 0, w.zero_bytes, st);

  GridDev gd;
  gd.dhw = (int64_t)g->d * g->h * g->w;
  gd.n_cams = g->n;
  gd.nx = g->nx[0];
  gd.ny = g->nx[1];
  gd.nz = g->nx[2];
  for (int a = 0; a < 3; ++a) {
    gd.lo[a] = g->lo[a];
    gd.dx[a] = g->dx[a];
  }
  const int bn = g->b * g->n;
  // ~2 CTAs per SM overall; each CTA stays inside one camera
  int bx = (int)((gd.dhw + 255) / 256);
  const int want = (kNumSMs * 8 + bn - 1) / bn;
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  if (coor)
    point_rank_kernel<true><<<dim3(bx, bn), 256, 0, st>>>(coor, frustum, rots, trans, gd, plan, point_rank, w.hist, counts_dev);
  else
    point_rank_kernel<false><<<dim3(bx, bn), 256, 0, st>>>(coor, frustum, rots, trans, gd, plan, point_rank, w.hist, counts_dev);
  count_launch();

  run_passes(plan, w, point_rank, nullptr, p0, counts_dev, ranks_bev, ranks_depth, st);

  segment_heads_kernel<0><<<(unsigned)w.head_tiles, kHeadThreads, 0, st>>>(
      ranks_bev, ranks_depth, counts_dev, p0, interval_starts, counts_dev + 1,
      w.status + (size_t)plan.n_passes * w.sort_tiles * kRadixBins, w.tickets + plan.n_passes, ranks_feat,
      (int)gd.dhw, g->h * g->w, nullptr, nullptr, nullptr, nullptr);
  count_launch();
  int lb = (int)((v < p0 ? v : p0) + 255) / 256;
  if (lb > kNumSMs * 8) lb = kNumSMs * 8;
  if (lb < 1) lb = 1;
  interval_lengths_kernel<<<lb, 256, 0, st>>>(interval_starts, counts_dev + 1, counts_dev, p0, interval_lengths);
  count_launch();
  return launch_status();
}

extern "C" size_t bevpool_v2_backward_regroup_workspace_bytes(int64_t n_points) {
  if (n_points <= 0) return 256;
  return carve(nullptr, n_points, kMaxPasses).total_bytes + align_up(sizeof(int) * (size_t)n_points, 256);
}

extern "C" int bevpool_v2_backward_regroup(const int32_t* ranks_depth, const int32_t* ranks_feat,
                                           const int32_t* ranks_bev, int64_t n_points, int32_t max_ranks_feat,
                                           int32_t* ranks_depth_bp, int32_t* ranks_feat_bp, int32_t* ranks_bev_bp,
                                           int32_t* interval_starts_bp, int32_t* interval_lengths_bp,
                                           int32_t* n_intervals_bp_dev, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  if (n_points < 0 || max_ranks_feat < 0 || !n_intervals_bp_dev) return BEVPOOL_ERR_BAD_ARG;
  if (n_points >= ((int64_t)1 << 30)) return BEVPOOL_ERR_OVERFLOW;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(n_intervals_bp_dev, 0, sizeof(int32_t), st);
  if (n_points == 0) return launch_status();
  if (!ranks_depth || !ranks_feat || !ranks_bev || !ranks_depth_bp || !ranks_feat_bp || !ranks_bev_bp ||
      !interval_starts_bp || !interval_lengths_bp || !workspace)
    return BEVPOOL_ERR_BAD_ARG;
  if (workspace_bytes < bevpool_v2_backward_regroup_workspace_bytes(n_points)) return BEVPOOL_ERR_WORKSPACE;
  const SortPlan plan = make_plan(max_ranks_feat);
  const SortWorkspace w = carve(workspace, n_points, plan.n_passes);
  int* order = (int*)((char*)workspace + carve(nullptr, n_points, kMaxPasses).total_bytes);
  cudaMemsetAsync(workspace, 0, w.zero_bytes, st);
  int hb = (int)((n_points + 2047) / 2048);
  if (hb > kNumSMs * 8) hb = kNumSMs * 8;
  key_hist_kernel<<<hb, 256, 0, st>>>(ranks_feat, n_points, plan, w.hist);
  count_launch();
  // argsort: sorted keys land in ranks_feat_bp, original positions in `order`
  run_passes(plan, w, ranks_feat, nullptr, n_points, nullptr, ranks_feat_bp, order, st);
  segment_heads_kernel<1><<<(unsigned)w.head_tiles, kHeadThreads, 0, st>>>(
      ranks_feat_bp, order, nullptr, n_points, interval_starts_bp, n_intervals_bp_dev,
      w.status + (size_t)plan.n_passes * w.sort_tiles * kRadixBins, w.tickets + plan.n_passes, nullptr, 1, 1,
      ranks_depth, ranks_bev, ranks_depth_bp, ranks_bev_bp);
  count_launch();
  int lb = (int)((n_points + 255) / 256);
  if (lb > kNumSMs * 8) lb = kNumSMs * 8;
  interval_lengths_kernel<<<lb, 256, 0, st>>>(interval_starts_bp, n_intervals_bp_dev, nullptr, n_points,
                                              interval_lengths_bp);
  count_launch();
  return launch_status();
}
