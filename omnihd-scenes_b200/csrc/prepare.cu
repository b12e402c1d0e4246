// Frustum geometry, voxel ranking, stable LSD radix sort and run-length segmentation for
// voxel_pooling_prepare_v2, plus the backward regrouping by ranks_feat — all on the device,
// no host synchronisation, bit-exact against the reference's CPU results.
//
// Pipeline of bevpool_prepare_v2 (one stream):
//   memset(zero region)                 digit totals, next-pass tile histograms
//   point_rank_kernel    P0 threads     geometry (optional) -> voxel rank | -1 per frustum point; the CTA
//                                       is one sort tile, so it also emits the tile's pass-0 digit
//                                       histogram, the digit totals of every pass and the kept count
//   per pass:  tile_scan_kernel         exclusive scan of the tile histograms along tiles (one CTA per digit)
//              radix_scatter_kernel     stable scatter: warp-level match ranking inside the tile + scanned
//                                       tile offsets; pass 0 also compacts (dropped points are never
//                                       written); each pass builds the next pass's tile histograms with
//                                       fire-and-forget REDs at the destination positions
//   (API path only) head_count / head_scan / head_write / interval_lengths: run-length segmentation by
//                                       counted head flags and a scan, ranks_feat derived from ranks_depth
//
// There is no inter-CTA waiting anywhere (an earlier decoupled-look-back version serialised on chains of
// ~500 predecessor tiles at these problem sizes: 23 us per pass; see profiles/). Sort tiles scale with the
// input (2048 * rounds keys) so the per-tile tables stay small at 3.7e8 points.
//
// Exactness (SURVEY.md §7 hard part 3): the voxel index is trunc((coor - lo) / dx) with an IEEE
// fp32 subtract and divide (__fsub_rn / __fdiv_rn; no reciprocal, no FMA), the geometry is
// r0*x + r1*y + r2*z + t with every product and sum rounded separately, ties keep ascending
// point index because every pass is stable.
#include "common.cuh"

namespace bevpool {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                          // keys per thread and round
constexpr int kSortRound = kSortThreads * kSortItems;  // keys per CTA and round
constexpr int kMaxRadixBits = 9;
constexpr int kMaxBins = 1 << kMaxRadixBits;
constexpr int kMaxPasses = 4;
constexpr int kMaxTiles = 4096;                        // tile tables are scanned by one CTA per digit

constexpr int kHeadThreads = 256;
constexpr int kHeadItems = 4;
constexpr int kHeadRound = kHeadThreads * kHeadItems;

struct SortPlan {
  int n_passes;
  int shift[kMaxPasses];
  int bits[kMaxPasses];
  int rounds;          // 2048-key rounds per sort tile
  int64_t n_tiles;     // tiles covering the (uncompacted) input
  int head_rounds;     // 1024-key rounds per segmentation tile
  int64_t head_tiles;
};

static SortPlan make_plan(int64_t max_key, int64_t n_max) {
  int total_bits = 1;
  while (total_bits < 31 && ((int64_t)1 << total_bits) <= max_key) ++total_bits;
  SortPlan p;
  p.n_passes = (total_bits + kMaxRadixBits - 1) / kMaxRadixBits;
  // spread the bits evenly over the passes (e.g. 17 bits -> 9+8)
  int left = total_bits, shift = 0;
  for (int i = 0; i < p.n_passes; ++i) {
    const int b = (left + (p.n_passes - i) - 1) / (p.n_passes - i);
    p.shift[i] = shift;
    p.bits[i] = b;
    shift += b;
    left -= b;
  }
  for (int i = p.n_passes; i < kMaxPasses; ++i) p.shift[i] = p.bits[i] = 0;
  const int64_t rounds_total = (n_max + kSortRound - 1) / kSortRound;
  p.rounds = (int)((rounds_total + kMaxTiles - 1) / kMaxTiles);
  if (p.rounds < 1) p.rounds = 1;
  p.n_tiles = (rounds_total + p.rounds - 1) / p.rounds;
  if (p.n_tiles < 1) p.n_tiles = 1;
  const int64_t hr = (n_max + kHeadRound - 1) / kHeadRound;
  p.head_rounds = (int)((hr + kMaxTiles - 1) / kMaxTiles);
  if (p.head_rounds < 1) p.head_rounds = 1;
  p.head_tiles = (hr + p.head_rounds - 1) / p.head_rounds;
  if (p.head_tiles < 1) p.head_tiles = 1;
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
// grid: (ceil(DHW / 256), BN): every CTA works inside one camera, so R|t are CTA-uniform
__global__ void __launch_bounds__(256)
geometry_kernel(const float* __restrict__ frustum, const float* __restrict__ rots, const float* __restrict__ trans,
                float* __restrict__ coor, int64_t dhw_total) {
  const int64_t bn = blockIdx.y;
  float cam[12];
#pragma unroll
  for (int i = 0; i < 9; ++i) cam[i] = __ldg(rots + bn * 9 + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) cam[9 + i] = __ldg(trans + bn * 3 + i);
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
  int64_t n_points;    // B*N*D*H*W
  int dhw;             // D*H*W
  uint32_t dhw_mul;    // idx / dhw == (idx * dhw_mul) >> dhw_shift for idx < 2^31
  int dhw_shift;
  int n_cams;          // N
  int bn;              // B*N
  int nx, ny, nz;
  float lo[3], dx[3], inv[3];   // inv: see voxel_index
};

// One CTA = one sort tile of the (uncompacted) point list: rank of every point, the tile's pass-0
// digit histogram (tile_hist0[tile][bins0]), digit totals of all passes, kept count.
template <bool FROM_COOR>
__global__ void __launch_bounds__(kSortThreads)
point_rank_kernel(const float* __restrict__ coor, const float* __restrict__ frustum, const float* __restrict__ rots,
                  const float* __restrict__ trans, GridDev g, SortPlan plan, int* __restrict__ point_rank,
                  uint32_t* __restrict__ tile_hist0, uint32_t* __restrict__ totals /*[kMaxPasses][kMaxBins]*/,
                  int* __restrict__ n_kept, uint8_t* __restrict__ occupied /* voxel byte map or nullptr */) {
  extern __shared__ float s_cam[];                       // [bn][12] (fused geometry only)
  __shared__ uint32_t sh[kMaxPasses][kMaxBins];
  __shared__ uint32_t s_kept;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kMaxPasses * kMaxBins; i += kSortThreads) (&sh[0][0])[i] = 0;
  if (tid == 0) s_kept = 0;
  if (!FROM_COOR)
    for (int i = tid; i < g.bn * 12; i += kSortThreads) {
      const int c = i / 12, k = i - c * 12;
      s_cam[i] = k < 9 ? __ldg(rots + c * 9 + k) : __ldg(trans + c * 3 + (k - 9));
    }
  __syncthreads();
  const int64_t vpf = (int64_t)g.nx * g.ny * g.nz;
  const int64_t tile_base = (int64_t)blockIdx.x * plan.rounds * kSortRound;
  uint32_t kept = 0;
  for (int r = 0; r < plan.rounds; ++r) {
    const int64_t base = tile_base + (int64_t)r * kSortRound + warp * (32 * kSortItems) + lane;
    // all loads of the round first: the index arithmetic below branches (guard band, range tests) and loads do not move
    // across branches — issued per point they cost one memory round trip each
    float px[kSortItems], py[kSortItems], pz[kSortItems];
    int cams[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int64_t idx = base + j * 32;
      px[j] = py[j] = pz[j] = 0.f;
      cams[j] = 0;
      if (idx < g.n_points) {
        cams[j] = (int)(((uint64_t)(uint32_t)idx * g.dhw_mul) >> g.dhw_shift);   // b*N + n
        if (FROM_COOR) {
          // (materialised coordinates are read where they are used, below: 24 stride-3 loads in flight per thread ran
          // 2.3x slower on a cold cache and only 4 % faster on a warm one)
        } else {
          const int64_t o = idx - (int64_t)cams[j] * g.dhw;
          px[j] = __ldg(frustum + 3 * o + 0);
          py[j] = __ldg(frustum + 3 * o + 1);
          pz[j] = __ldg(frustum + 3 * o + 2);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int64_t idx = base + j * 32;
      if (idx >= g.n_points) continue;
      const int cam = cams[j];
      float x, y, z;
      if (FROM_COOR) {
        x = ldg_stream_f32(coor + 3 * idx + 0);
        y = ldg_stream_f32(coor + 3 * idx + 1);
        z = ldg_stream_f32(coor + 3 * idx + 2);
      } else {
        cam_point_of(px[j], py[j], pz[j], s_cam + cam * 12, x, y, z);
      }
      int vx, vy, vz;
      const bool ok = voxel_index3(x, y, z, g.lo, g.dx, g.inv, g.nx, g.ny, g.nz, vx, vy, vz);
      int rank = -1;
      if (ok) {
        rank = (int)((int64_t)(cam / g.n_cams) * vpf + ((int64_t)vz * g.ny + vy) * g.nx + vx);
        ++kept;
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p)
          if (p < plan.n_passes) atomicAdd(&sh[p][(rank >> plan.shift[p]) & ((1 << plan.bits[p]) - 1)], 1u);
      }
      // voxel BYTE map for the early interval count: a plain store of 1 (idempotent, no atomics — 1.1 M single-bit atomicOr
      // into 32 K words, plain or warp-aggregated with match.any, doubled this kernel's time)
      if (ok && occupied) occupied[rank] = 1;
      point_rank[idx] = rank;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(kFullMask, kept, o);
  if (lane == 0 && kept) atomicAdd(&s_kept, kept);
  __syncthreads();
  const int bins0 = 1 << plan.bits[0];
  for (int d = tid; d < bins0; d += kSortThreads) tile_hist0[(int64_t)d * gridDim.x + blockIdx.x] = sh[0][d];
  for (int i = tid; i < plan.n_passes * kMaxBins; i += kSortThreads) {
    const uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(totals + i, v);
  }
  if (tid == 0 && s_kept) atomicAdd(n_kept, (int)s_kept);
}

// Same bookkeeping for an existing key array (backward regroup): tile histogram of pass 0 + totals.
__global__ void __launch_bounds__(kSortThreads)
key_hist_kernel(const int* __restrict__ keys, int64_t n, SortPlan plan, uint32_t* __restrict__ tile_hist0,
                uint32_t* __restrict__ totals) {
  __shared__ uint32_t sh[kMaxPasses][kMaxBins];
  const int tid = threadIdx.x;
  for (int i = tid; i < kMaxPasses * kMaxBins; i += kSortThreads) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile_base = (int64_t)blockIdx.x * plan.rounds * kSortRound;
  const int64_t tile_end = min(n, tile_base + (int64_t)plan.rounds * kSortRound);
  for (int64_t i = tile_base + tid; i < tile_end; i += kSortThreads) {
    const int k = ldg_stream_i32(keys + i);
    if (k < 0) continue;
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p)
      if (p < plan.n_passes) atomicAdd(&sh[p][(k >> plan.shift[p]) & ((1 << plan.bits[p]) - 1)], 1u);
  }
  __syncthreads();
  const int bins0 = 1 << plan.bits[0];
  for (int d = tid; d < bins0; d += kSortThreads) tile_hist0[(int64_t)d * gridDim.x + blockIdx.x] = sh[0][d];
  for (int i = tid; i < plan.n_passes * kMaxBins; i += kSortThreads) {
    const uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(totals + i, v);
  }
}

// Early counts for the host (bevpool_prepare_v2_counts): P is final once the rank kernel is done, and the number
// of intervals equals the number of occupied voxels — no need to wait for the sort and the segmentation.
// The last CTA to finish stores both straight into page-locked host memory (no copy-engine transfer: an 8-byte
// cudaMemcpyAsync would queue behind whatever bulk device->host copy another stream has in flight).
__global__ void __launch_bounds__(256)
early_counts_kernel(const uint32_t* __restrict__ occupied, int64_t words, const int* __restrict__ n_kept,
                    int* __restrict__ early /*[0] = occupied voxels, [1] = finished CTAs; zeroed*/,
                    volatile int32_t* __restrict__ host_counts /*[2], mapped pinned*/) {
  pdl_wait();
  __shared__ uint32_t s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < words; i += (int64_t)gridDim.x * 256)
    mine += __popc(occupied[i] & 0x01010101u);   // four map bytes per word, each 0 or 1
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFullMask, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_sum) atomicAdd(early, (int)s_sum);
    __threadfence();
    if (atomicAdd(early + 1, 1) == (int)gridDim.x - 1) {
      __threadfence();
      host_counts[0] = *n_kept;
      host_counts[1] = *(volatile int*)early;
      __threadfence_system();
    }
  }
}

// ------------------------------------------------------------------------------------------ tile scan
// hist is digit-major [bins][n_tiles]; CTA d turns row d into its exclusive prefix along tiles (in place,
// coalesced).
__global__ void __launch_bounds__(256)
tile_scan_kernel(uint32_t* __restrict__ hist, int64_t n_tiles, int bins) {
  __shared__ uint32_t scan_tmp[8];
  pdl_wait();
  const int d = blockIdx.x;
  uint32_t carry = 0;
  for (int64_t t0 = 0; t0 < n_tiles; t0 += 256) {
    const int64_t t = t0 + threadIdx.x;
    const uint32_t v = t < n_tiles ? hist[(int64_t)d * n_tiles + t] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(v, scan_tmp, &total);
    if (t < n_tiles) hist[(int64_t)d * n_tiles + t] = carry + ex;
    carry += total;
  }
}

// ------------------------------------------------------------------------------------------ radix pass
// One stable LSD pass. Keys < 0 are dropped (pass 0 = compaction). vals_in == nullptr means
// "value = position" (first pass of an argsort). tile_prefix[digit][tile] holds, per digit, the number
// of keys with that digit in earlier tiles. next_hist (optional): tile histograms of the NEXT pass,
// accumulated at the destination positions.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int* __restrict__ keys_out,
                     int* __restrict__ vals_out, const uint32_t* __restrict__ digit_totals,
                     const uint32_t* __restrict__ tile_prefix, uint32_t* __restrict__ next_hist,
                     const int* __restrict__ n_dev, int64_t n_host, int shift, int bits, int rounds, int next_shift,
                     int next_bits) {
  __shared__ uint32_t warp_hist[kSortWarps][kMaxBins];
  __shared__ uint32_t digit_base[kMaxBins];
  __shared__ uint32_t round_total[kMaxBins];
  __shared__ uint32_t scan_tmp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = 1 << bits;
  pdl_wait();
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t tile_keys = (int64_t)rounds * kSortRound;
  const int64_t tile_base = (int64_t)blockIdx.x * tile_keys;
  if (tile_base >= n) return;

  // the first round's keys are requested before the digit-base prologue (a scan with two barriers behind its own loads):
  // one memory round trip instead of two before the ranking can start
  int key[kSortItems], val[kSortItems];
  auto load_round = [&](int64_t round_base) {
    const int64_t base = round_base + warp * (32 * kSortItems) + lane;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int64_t idx = base + j * 32;
      key[j] = idx < n ? ldg_stream_i32(keys_in + idx) : -1;
      val[j] = vals_in ? (idx < n ? ldg_stream_i32(vals_in + idx) : 0) : (int)idx;
    }
  };
  load_round(tile_base);

  // digit_base = (keys with a smaller digit) + (keys with this digit in earlier tiles)
  {
    uint32_t carry = 0;
    for (int d0 = 0; d0 < bins; d0 += kSortThreads) {
      const int d = d0 + tid;
      uint32_t total;
      const uint32_t ex = block_exclusive_scan_256(d < bins ? __ldg(digit_totals + d) : 0u, scan_tmp, &total);
      if (d < bins) digit_base[d] = carry + ex + __ldg(tile_prefix + (int64_t)d * gridDim.x + blockIdx.x);
      carry += total;
    }
  }
  const uint32_t next_mask = (1u << next_bits) - 1u;
  for (int r = 0; r < rounds; ++r) {
    const int64_t round_base = tile_base + (int64_t)r * kSortRound;
    if (round_base >= n) break;   // CTA-uniform
    for (int w = 0; w < kSortWarps; ++w)
      for (int d = tid; d < bins; d += kSortThreads) warp_hist[w][d] = 0;
    __syncthreads();
    uint32_t rank[kSortItems];
    if (r > 0) load_round(round_base);
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
    // exclusive scan over the warps of this round
    for (int d = tid; d < bins; d += kSortThreads) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t t = warp_hist[w][d];
        warp_hist[w][d] = run;
        run += t;
      }
      round_total[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      if (key[j] >= 0) {
        const int d = (key[j] >> shift) & (bins - 1);
        const uint32_t pos = digit_base[d] + warp_hist[warp][d] + rank[j];
        keys_out[pos] = key[j];
        vals_out[pos] = val[j];
        if (next_hist)
          atomicAdd(next_hist + (int64_t)((key[j] >> next_shift) & next_mask) * gridDim.x + pos / (uint32_t)tile_keys, 1u);
      }
    }
    if (r + 1 < rounds) {   // the next round continues behind this round's keys
      __syncthreads();
      for (int d = tid; d < bins; d += kSortThreads) digit_base[d] += round_total[d];
    }
  }
}

// ------------------------------------------------------------------------------------------ segmentation
// keys are sorted; a head is a position whose key differs from its predecessor.
//   head_count_kernel : heads per tile; MODE 0 also derives ranks_feat from ranks_depth, MODE 1 gathers
//                       the regrouped rank arrays through the argsort permutation
//   head_scan_kernel  : exclusive scan of the tile counts (one CTA), total -> *n_heads_out
//   head_write_kernel : interval_starts
__device__ __forceinline__ void load_heads(const int* __restrict__ keys, int64_t i0, int64_t n, int k[kHeadItems],
                                           bool head[kHeadItems], uint32_t& cnt) {
  const bool vec_ok = (((uintptr_t)keys) & 15) == 0;   // caller buffers may be unaligned views
  if (vec_ok && i0 + kHeadItems <= n) {
    const int4 kk = *reinterpret_cast<const int4*>(keys + i0);
    k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
  } else {
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j) k[j] = (i0 + j < n) ? keys[i0 + j] : -1;
  }
  const int prev = (i0 > 0 && i0 < n) ? keys[i0 - 1] : -1;
  cnt = 0;
#pragma unroll
  for (int j = 0; j < kHeadItems; ++j) {
    head[j] = (i0 + j < n) && (k[j] != (j == 0 ? prev : k[j - 1]) || (i0 + j == 0));
    cnt += head[j];
  }
}

template <int MODE>
__global__ void __launch_bounds__(kHeadThreads)
head_count_kernel(const int* __restrict__ keys, const int* __restrict__ vals, const int* __restrict__ n_dev,
                  int64_t n_host, int rounds, uint32_t* __restrict__ tile_counts,
                  int* __restrict__ ranks_feat, int dhw, int hw,                                   // MODE 0
                  const int* __restrict__ src_rd, const int* __restrict__ src_rb, int* __restrict__ dst_rd,
                  int* __restrict__ dst_rb) {                                                      // MODE 1
  pdl_wait();
  __shared__ uint32_t s_cnt;
  const int tid = threadIdx.x;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t tile_base = (int64_t)blockIdx.x * rounds * kHeadRound;
  uint32_t mine = 0;
  for (int r = 0; r < rounds; ++r) {
    const int64_t i0 = tile_base + (int64_t)r * kHeadRound + (int64_t)tid * kHeadItems;
    if (i0 >= n) break;
    int k[kHeadItems];
    bool head[kHeadItems];
    uint32_t cnt;
    load_heads(keys, i0, n, k, head, cnt);
    mine += cnt;
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j)
      if (i0 + j < n) {
        const int v = vals[i0 + j];
        if (MODE == 0) {
          if (ranks_feat) ranks_feat[i0 + j] = (v / dhw) * hw + v % hw;
        } else {
          dst_rd[i0 + j] = __ldg(src_rd + v);
          dst_rb[i0 + j] = __ldg(src_rb + v);
        }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFullMask, mine, o);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (tid == 0) tile_counts[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(256)
head_scan_kernel(uint32_t* __restrict__ tile_counts, int64_t n_tiles, int* __restrict__ n_heads_out) {
  pdl_wait();
  __shared__ uint32_t scan_tmp[8];
  uint32_t carry = 0;
  for (int64_t t0 = 0; t0 < n_tiles; t0 += 256) {
    const int64_t t = t0 + threadIdx.x;
    const uint32_t v = t < n_tiles ? tile_counts[t] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(v, scan_tmp, &total);
    if (t < n_tiles) tile_counts[t] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *n_heads_out = (int)carry;
}

__global__ void __launch_bounds__(kHeadThreads)
head_write_kernel(const int* __restrict__ keys, const int* __restrict__ n_dev, int64_t n_host, int rounds,
                  const uint32_t* __restrict__ tile_prefix, int* __restrict__ starts) {
  pdl_wait();
  __shared__ uint32_t scan_tmp[8];
  const int tid = threadIdx.x;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t tile_base = (int64_t)blockIdx.x * rounds * kHeadRound;
  if (tile_base >= n) return;
  uint32_t slot0 = tile_prefix[blockIdx.x];
  for (int r = 0; r < rounds; ++r) {
    const int64_t round_base = tile_base + (int64_t)r * kHeadRound;
    if (round_base >= n) break;   // CTA-uniform
    const int64_t i0 = round_base + (int64_t)tid * kHeadItems;
    int k[kHeadItems];
    bool head[kHeadItems];
    uint32_t cnt;
    load_heads(keys, i0, n, k, head, cnt);
    uint32_t total;
    uint32_t slot = slot0 + block_exclusive_scan_256(cnt, scan_tmp, &total);
#pragma unroll
    for (int j = 0; j < kHeadItems; ++j)
      if (head[j]) starts[slot++] = (int)(i0 + j);
    slot0 += total;
  }
}

__global__ void interval_lengths_kernel(const int* __restrict__ starts, const int* __restrict__ n_heads_dev,
                                        const int* __restrict__ n_dev, int64_t n_host, int* __restrict__ lengths) {
  pdl_wait();
  const int n_heads = *n_heads_dev;
  const int n = n_dev ? *n_dev : (int)n_host;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_heads; j += (int64_t)gridDim.x * blockDim.x)
    lengths[j] = (j + 1 < n_heads ? starts[j + 1] : n) - starts[j];
}

// ------------------------------------------------------------------------------------------ host side
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SortWorkspace {
  // zeroed region
  uint32_t* totals;      // [kMaxPasses][kMaxBins]
  uint32_t* hist_odd;    // tile histograms built by REDs (passes 1, 3) : [n_tiles][bins]
  uint32_t* hist_even2;  // tile histograms built by REDs (pass 2)
  size_t zero_bytes;
  // scratch (not zeroed)
  uint32_t* hist0;       // pass-0 tile histograms, written in full by the rank / key_hist kernel
  uint32_t* head_counts; // [head_tiles]
  int* alt_keys;
  int* alt_vals;
  size_t total_bytes;
};

static SortWorkspace carve(void* ws, int64_t n_max, const SortPlan& plan) {
  SortWorkspace w;
  char* p = (char*)ws;
  size_t off = 0;
  const size_t hist_bytes = align_up(sizeof(uint32_t) * (size_t)plan.n_tiles * kMaxBins, 256);
  w.totals = (uint32_t*)(p + off);
  off += sizeof(uint32_t) * kMaxPasses * kMaxBins;
  w.hist_odd = (uint32_t*)(p + off);
  off += plan.n_passes > 1 ? hist_bytes : 0;
  w.hist_even2 = (uint32_t*)(p + off);
  off += plan.n_passes > 2 ? hist_bytes : 0;
  // a 4th pass re-uses hist_odd after the host re-zeroes it (see run_passes)
  off = align_up(off, 256);
  w.zero_bytes = off;
  w.hist0 = (uint32_t*)(p + off);
  off += hist_bytes;
  w.head_counts = (uint32_t*)(p + off);
  off += align_up(sizeof(uint32_t) * (size_t)plan.head_tiles, 256);
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
  uint32_t* hist = w.hist0;
  for (int p = 0; p < plan.n_passes; ++p) {
    const bool to_out = ((plan.n_passes - 1 - p) % 2) == 0;
    int* dst_k = to_out ? out_keys : w.alt_keys;
    int* dst_v = to_out ? out_vals : w.alt_vals;
    const int bins = 1 << plan.bits[p];
    uint32_t* next = nullptr;
    if (p + 1 < plan.n_passes) {
      next = (p + 1 == 2) ? w.hist_even2 : w.hist_odd;
      if (p + 1 == 3) cudaMemsetAsync(w.hist_odd, 0, sizeof(uint32_t) * (size_t)plan.n_tiles * kMaxBins, st);
    }
    launch_pdl(tile_scan_kernel, dim3(bins), dim3(256), 0, st, hist, plan.n_tiles, bins);
    count_launch();
    launch_pdl(radix_scatter_kernel, dim3((unsigned)plan.n_tiles), dim3(kSortThreads), 0, st, src_k, src_v, dst_k, dst_v,
               (const uint32_t*)(w.totals + p * kMaxBins), (const uint32_t*)hist, next, p == 0 ? (const int*)nullptr : n_dev,
               n0, plan.shift[p], plan.bits[p], plan.rounds, next ? plan.shift[p + 1] : 0, next ? plan.bits[p + 1] : 1);
    count_launch();
    src_k = dst_k;
    src_v = dst_v;
    hist = next;
  }
}

static void magic_div(int d, uint32_t* mul, int* shift) {
  int s = 0;
  while (((int64_t)1 << s) < d) ++s;
  *shift = 31 + s;
  *mul = (uint32_t)((((uint64_t)1 << (31 + s)) / (uint64_t)d) + 1);
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
  if (n >= ((int64_t)1 << 30) || v >= ((int64_t)1 << 31) - 1) return BEVPOOL_ERR_OVERFLOW;  // int32 ranks
  if ((int64_t)g->b * g->n > 4096) return BEVPOOL_ERR_BAD_ARG;   // camera table lives in shared memory
  *p0 = n;
  *total_voxels = v;
  return BEVPOOL_OK;
}

// early-count region of the prepare workspace: [2] counts + voxel byte map, behind the private point_rank array
static size_t early_region_bytes(int64_t v) { return 256 + align_up((size_t)v, 256); }

extern "C" size_t bevpool_prepare_v2_workspace_bytes(const bevpool_grid_t* g) {
  int64_t p0, v;
  if (check_grid(g, &p0, &v) != BEVPOOL_OK) return 0;
  if (p0 == 0) return 256;
  const SortPlan plan = make_plan(v > 0 ? v - 1 : 0, p0);
  // + a private point_rank array in case the caller does not want one + the early-count region
  return carve(nullptr, p0, plan).total_bytes + align_up(sizeof(int) * (size_t)p0, 256) + early_region_bytes(v);
}

static int prepare_impl(const float* coor, const float* frustum, const float* rots, const float* trans,
                        const bevpool_grid_t* g, int32_t* ranks_bev, int32_t* ranks_depth, int32_t* ranks_feat,
                        int32_t* interval_starts, int32_t* interval_lengths, int32_t* counts_dev, int32_t* point_rank,
                        void* workspace, size_t workspace_bytes, void* stream, int32_t* host_counts) {
  int64_t p0, v;
  int rc = check_grid(g, &p0, &v);
  if (rc != BEVPOOL_OK) return rc;
  if (!counts_dev) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (host_counts) {
    host_counts[0] = host_counts[1] = 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
      return BEVPOOL_ERR_BAD_ARG;                     // a host hand-back cannot be captured into a graph
  }
  cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int32_t), st);
  if (p0 == 0) return launch_status();
  if (!ranks_bev || !ranks_depth || !workspace) return BEVPOOL_ERR_BAD_ARG;
  if ((interval_starts == nullptr) != (interval_lengths == nullptr)) return BEVPOOL_ERR_BAD_ARG;
  if (!coor && (!frustum || !rots || !trans)) return BEVPOOL_ERR_BAD_ARG;
  if (workspace_bytes < bevpool_prepare_v2_workspace_bytes(g)) return BEVPOOL_ERR_WORKSPACE;

  const SortPlan plan = make_plan(v - 1, p0);
  const SortWorkspace w = carve(workspace, p0, plan);
  if (!point_rank) point_rank = (int*)((char*)workspace + w.total_bytes);
  cudaMemsetAsync(workspace, 0, w.zero_bytes, st);
  int* early = nullptr;
  uint8_t* occupied = nullptr;
  CountSlot* slot = nullptr;
  if (host_counts) {
    early = (int*)((char*)workspace + w.total_bytes + align_up(sizeof(int) * (size_t)p0, 256));
    occupied = (uint8_t*)((char*)early + 256);
    cudaMemsetAsync(early, 0, early_region_bytes(v), st);
    slot = acquire_count_slot();
    if (!slot) return (int)cudaErrorMemoryAllocation;
  }

  GridDev gd;
  gd.n_points = p0;
  gd.dhw = g->d * g->h * g->w;
  magic_div(gd.dhw, &gd.dhw_mul, &gd.dhw_shift);
  gd.n_cams = g->n;
  gd.bn = g->b * g->n;
  gd.nx = g->nx[0];
  gd.ny = g->nx[1];
  gd.nz = g->nx[2];
  for (int a = 0; a < 3; ++a) {
    gd.lo[a] = g->lo[a];
    gd.dx[a] = g->dx[a];
    gd.inv[a] = exact_reciprocal_or_zero(g->dx[a]);
  }
  if (coor) {
    point_rank_kernel<true><<<(unsigned)plan.n_tiles, kSortThreads, 0, st>>>(
        coor, frustum, rots, trans, gd, plan, point_rank, w.hist0, w.totals, counts_dev, occupied);
  } else {
    const size_t cam_smem = sizeof(float) * 12 * (size_t)gd.bn;
    // ~9 KB of static shared memory on top of the camera table: opt in as soon as the TOTAL passes 48 KB
    if (int rc = ensure_dynamic_smem(point_rank_kernel<false>, cam_smem + 9 * 1024)) {
      if (slot) release_count_slot(slot);
      return rc;
    }
    point_rank_kernel<false><<<(unsigned)plan.n_tiles, kSortThreads, cam_smem, st>>>(
        coor, frustum, rots, trans, gd, plan, point_rank, w.hist0, w.totals, counts_dev, occupied);
  }
  count_launch();
  if (slot) {
    // counts leave for the host now; everything below is queued behind them and the host waits for this kernel only
    const int64_t words = (v + 3) / 4;   // the region is zeroed up to a multiple of 256 bytes
    int eb = (int)((words + 255) / 256);
    if (eb > kNumSMs * 4) eb = kNumSMs * 4;
    early_counts_kernel<<<eb, 256, 0, st>>>((const uint32_t*)occupied, words, counts_dev, early, slot->host);
    count_launch();
    cudaEventRecord(slot->ev, st);
  }

  run_passes(plan, w, point_rank, nullptr, p0, counts_dev, ranks_bev, ranks_depth, st);

  if (interval_starts || ranks_feat) {
    launch_pdl(head_count_kernel<0>, dim3((unsigned)plan.head_tiles), dim3(kHeadThreads), 0, st, (const int*)ranks_bev,
               (const int*)ranks_depth, (const int*)counts_dev, (int64_t)p0, plan.head_rounds, w.head_counts, ranks_feat, gd.dhw,
               g->h * g->w, (const int*)nullptr, (const int*)nullptr, (int*)nullptr, (int*)nullptr);
    count_launch();
  }
  if (interval_starts) {
    launch_pdl(head_scan_kernel, dim3(1), dim3(256), 0, st, w.head_counts, (int64_t)plan.head_tiles, counts_dev + 1);
    launch_pdl(head_write_kernel, dim3((unsigned)plan.head_tiles), dim3(kHeadThreads), 0, st, (const int*)ranks_bev,
               (const int*)counts_dev, (int64_t)p0, plan.head_rounds, (const uint32_t*)w.head_counts, interval_starts);
    int lb = (int)(((v < p0 ? v : p0) + 255) / 256);
    if (lb > kNumSMs * 8) lb = kNumSMs * 8;
    if (lb < 1) lb = 1;
    launch_pdl(interval_lengths_kernel, dim3(lb), dim3(256), 0, st, (const int*)interval_starts, (const int*)(counts_dev + 1),
               (const int*)counts_dev, (int64_t)p0, interval_lengths);
    count_launch(3);
  }
  rc = launch_status();
  if (slot) {
    const cudaError_t e = cudaEventSynchronize(slot->ev);
    if (e == cudaSuccess) {
      host_counts[0] = slot->host[0];
      host_counts[1] = slot->host[1];
    } else if (rc == BEVPOOL_OK) {
      rc = (int)e;
    }
    release_count_slot(slot);
  }
  return rc;
}

extern "C" int bevpool_prepare_v2(const float* coor, const float* frustum, const float* rots, const float* trans,
                                  const bevpool_grid_t* g, int32_t* ranks_bev, int32_t* ranks_depth,
                                  int32_t* ranks_feat, int32_t* interval_starts, int32_t* interval_lengths,
                                  int32_t* counts_dev, int32_t* point_rank, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  return prepare_impl(coor, frustum, rots, trans, g, ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths,
                      counts_dev, point_rank, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int bevpool_prepare_v2_counts(const float* coor, const float* frustum, const float* rots, const float* trans,
                                         const bevpool_grid_t* g, int32_t* ranks_bev, int32_t* ranks_depth,
                                         int32_t* ranks_feat, int32_t* interval_starts, int32_t* interval_lengths,
                                         int32_t* counts_dev, int32_t* point_rank, void* workspace,
                                         size_t workspace_bytes, void* stream, int32_t* host_counts) {
  if (!host_counts) return BEVPOOL_ERR_BAD_ARG;
  return prepare_impl(coor, frustum, rots, trans, g, ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths,
                      counts_dev, point_rank, workspace, workspace_bytes, stream, host_counts);
}

extern "C" size_t bevpool_v2_backward_regroup_workspace_bytes(int64_t n_points) {
  if (n_points <= 0) return 256;
  SortPlan plan = make_plan(((int64_t)1 << 31) - 2, n_points);   // worst-case pass count
  return carve(nullptr, n_points, plan).total_bytes + align_up(sizeof(int) * (size_t)n_points, 256);
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
  const SortPlan plan = make_plan(max_ranks_feat, n_points);
  const SortWorkspace w = carve(workspace, n_points, plan);
  // the argsort permutation lives behind the largest possible carve so it never overlaps
  const SortPlan worst = make_plan(((int64_t)1 << 31) - 2, n_points);
  int* order = (int*)((char*)workspace + carve(nullptr, n_points, worst).total_bytes);
  cudaMemsetAsync(workspace, 0, w.zero_bytes, st);
  key_hist_kernel<<<(unsigned)plan.n_tiles, kSortThreads, 0, st>>>(ranks_feat, n_points, plan, w.hist0, w.totals);
  count_launch();
  // argsort: sorted keys land in ranks_feat_bp, original positions in `order`
  run_passes(plan, w, ranks_feat, nullptr, n_points, nullptr, ranks_feat_bp, order, st);
  head_count_kernel<1><<<(unsigned)plan.head_tiles, kHeadThreads, 0, st>>>(
      ranks_feat_bp, order, nullptr, n_points, plan.head_rounds, w.head_counts, nullptr, 1, 1, ranks_depth, ranks_bev,
      ranks_depth_bp, ranks_bev_bp);
  head_scan_kernel<<<1, 256, 0, st>>>(w.head_counts, plan.head_tiles, n_intervals_bp_dev);
  head_write_kernel<<<(unsigned)plan.head_tiles, kHeadThreads, 0, st>>>(ranks_feat_bp, nullptr, n_points,
                                                                       plan.head_rounds, w.head_counts,
                                                                       interval_starts_bp);
  int lb = (int)((n_points + 255) / 256);
  if (lb > kNumSMs * 8) lb = kNumSMs * 8;
  interval_lengths_kernel<<<lb, 256, 0, st>>>(interval_starts_bp, n_intervals_bp_dev, nullptr, n_points,
                                              interval_lengths_bp);
  count_launch(4);
  return launch_status();
}
