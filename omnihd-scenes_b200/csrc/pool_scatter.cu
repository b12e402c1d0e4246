// Sort-free fused view transform, forward direction (the default path of LSSViewTransform.forward):
//
//   memset(acc)                          fp32 accumulation grid [V][C], channels-last
//   view_fwd_scatter_kernel              geometry -> voxel rank (written to point_rank for the backward) ->
//                                        depth-weighted feature sums pushed into the grid with 128-bit REDs
//   acc_layout_kernel / acc_convert      [V][C] fp32 -> [B,C,Z,Y,X] (or s2c / channels-last) in the io dtype
//
// Why: in the sorted formulation (pool_dense.cu) every kept point gathers its C-channel feature row from L2
// (P*C*e bytes, 353 MB per step at cfg 2 — the forward's actual bound) and the voxel order has to be produced by
// a multi-pass radix sort first. Walking the frustum PIXEL-major instead keeps the feature rows of a pixel
// column in registers for all D bins; what leaves the SM is one row per RUN of consecutive points that share a
// voxel (all H rows of an image column at one depth bin usually do, SURVEY.md §8(d): mean interval = 15 points),
// i.e. ~P/10 rows of RED traffic, and no sort, no interval table, no ranks_* arrays exist at all.
//
// Price: contributions of different image columns / cameras to one voxel meet in L2 atomics, so the fp32
// summation ORDER across them is not fixed (voxels fed by one or two runs — the majority — are still
// bit-reproducible). torch.use_deterministic_algorithms(True) or LSSViewTransform(deterministic=True) selects
// the sorted path instead; both are inside the 1e-5 parity bar against the float64 oracle.
//
// Block shape and staging are those of the joint backward (pool_dense.cu): a CTA owns 8(w) x 4(h) pixels of
// one camera image, warp = image column, ranks and depths staged as [d][8] int4/float4 so one broadcast LDS.128
// yields the 4 rows of a bin. Each warp keeps FOUR open accumulators keyed by voxel rank (a 4-entry fully
// associative cache, all comparisons warp-uniform): a point joins the accumulator that already holds its
// voxel, otherwise it evicts (REDs out) the accumulator of its own row.
#include "common.cuh"

namespace bevpool {

constexpr int kScW = 8, kScH = 4, kScPix = kScW * kScH;
constexpr int kScWarps = 8, kScThreads = kScWarps * 32;

struct ScatterParams {
  int c, d, h, w;
  int bn, n_cams;
  int blocks_w, blocks_h;
  int nx, ny, nz;
  float lo[3], dx[3], inv[3];   // inv: see voxel_index
  int from_geometry;   // 1: compute ranks from frustum/rots/trans and write point_rank; 0: read point_rank
};

__device__ __forceinline__ void red_add_f32x4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}


#ifdef BEVPOOL_TIMELINE   // measurement builds only (profiles/timeline_scatter.py): per-CTA phase stamps in ns
__device__ unsigned long long g_sc_timeline[8 * 4096];
__device__ __forceinline__ unsigned long long sc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SC_STAMP(k) do { if (threadIdx.x == 0) tl[k] = sc_now(); } while (0)
#else
#define SC_STAMP(k) do { } while (0)
#endif

template <bool X>
__device__ __forceinline__ void frag_red(float* row, int sl, const Frag<X>& a) {
  red_add_f32x4(row + 4 * sl, a.v);
  if (X) red_add_f32(row + 64 + sl, a.s);
}

// HALVES = 2: the CTA covers 4 columns instead of 8 and two warps share a column, each walking half of the depth range
// (nothing to combine: both push REDs). Warps then live half as long, which halves the idle tail of the last wave —
// used when the launch is only a few waves long (cfg 2: 2.6 waves, 21 % of the SM-cycles were idle).
template <typename T, int CH4, int HALVES>
__global__ void __launch_bounds__(kScThreads, 3)
view_fwd_scatter_kernel(const T* __restrict__ depth, const T* __restrict__ feat, const float* __restrict__ frustum,
                        const float* __restrict__ rots, const float* __restrict__ trans, ScatterParams prm,
                        int* __restrict__ point_rank, float* __restrict__ acc_grid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // Narrow rows (C <= 64 / C <= 32) would leave half / three quarters of the lanes idle: the warp is cut into G
  // lane groups that walk the SAME column but interleaved depth bins (group g takes bins g, g + G, ...), each with
  // its own accumulators. Branches are then uniform per group, not per warp (groups usually take the same path).
  // C = 80 uses the 4 + 1 channel mapping (Frag<true>): 16 lanes cover a row, so it also gets two groups.
  constexpr bool X = CH4 == 20;
  constexpr int G = X ? 2 : ((CH4 == 0 || CH4 > 16) ? 1 : (CH4 > 8 ? 2 : 4));
  constexpr int LG = 32 / G;
  constexpr int WB = kScW / HALVES;         // columns per CTA
  constexpr int DQ = 4 * G * HALVES;        // bins are padded so every warp walks whole 4-bin groups per lane group
  const int d_pad = (prm.d + DQ - 1) / DQ * DQ;
  int4* s_rank4 = reinterpret_cast<int4*>(smem_raw);                             // [d_pad][WB]: ranks of rows h0..h0+3
  float4* s_depth4 = reinterpret_cast<float4*>(s_rank4 + (size_t)d_pad * WB);    // [d_pad][WB]
  int* s_lead = reinterpret_cast<int*>(s_depth4 + (size_t)d_pad * WB);           // [d_pad][WB]: see below
  __shared__ float s_cam[12];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int c4 = CH4 ? CH4 : (prm.c >> 2);
  const int C = 4 * c4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const int grp = lane / LG, sl = lane % LG;

  // 3-D grid (w-block, h-block, image): no integer divisions to find the block (they were 7 % of the kernel's
  // instructions with a flat index); dispatch order is still image-major
  const int bw = blockIdx.x, bh = blockIdx.y, bn = blockIdx.z;
  const int h0 = bh * kScH, w0 = bw * WB;
  const int hw = prm.h * prm.w;
  const int64_t img_base = (int64_t)bn * prm.d * hw;
  pdl_wait();
#ifdef BEVPOOL_TIMELINE
  unsigned long long tl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  SC_STAMP(0);
  // camera matrix -> shared memory; the barrier that publishes it sits BEHIND the first pass's loads (below), so its
  // memory round trip overlaps theirs
  float cam_val = 0.f;
  if (prm.from_geometry && threadIdx.x < 12)
    cam_val = threadIdx.x < 9 ? __ldg(rots + bn * 9 + threadIdx.x) : __ldg(trans + bn * 3 + threadIdx.x - 9);
  // ---- stage ranks / depths of the block: WB consecutive w per (d, h) (8: one 32-byte sector); a warp covers
  //      32 / (4 * WB) bins per pass
  for (int i = threadIdx.x; i < (d_pad - prm.d) * WB; i += kScThreads) {
    s_lead[prm.d * WB + i] = -1;
    s_depth4[prm.d * WB + i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {
    constexpr int BPI = 32 / (4 * WB);        // bins per warp and pass
    const int wl = lane % WB, hl = (lane / WB) & 3, bs = lane / (4 * WB);
    const unsigned colmask = (WB == 8 ? 0x01010101u : 0x00001111u) << (wl + 4 * WB * bs);
    const bool in = h0 + hl < prm.h && w0 + wl < prm.w;
    const int pix = (h0 + hl) * prm.w + w0 + wl;
    // 32-bit arithmetic throughout the staging loop: ranks fit int32 (checked by the launcher) and offsets inside one image
    // too; the 64-bit image base is folded into three pointers once
    const int frame_base = (bn / prm.n_cams) * (prm.nx * prm.ny * prm.nz);
    const T* depth_img = depth + img_base;
    int* rank_img = point_rank + img_base;
    for (int d0 = 0; d0 < prm.d; d0 += 4 * kScWarps * BPI) {
      int r[4];
      float dv[4];
      float fu[4], fv[4], fd[4];
      // all loads of the pass first (the geometry below branches — guard band, range tests — and loads do not move
      // across branches: issued per bin they cost one memory round trip each, r1 capture: half of the kernel's samples)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = d0 + (k * kScWarps + warp) * BPI + bs;
        r[k] = -1;
        dv[k] = fu[k] = fv[k] = fd[k] = 0.f;
        if (in && dd < prm.d) {
          const int o = dd * hw + pix;
          dv[k] = Vec4<T>::load1(depth_img, o);
          if (prm.from_geometry) {
            const float* fp = frustum + 3 * o;
            fu[k] = __ldg(fp + 0);
            fv[k] = __ldg(fp + 1);
            fd[k] = __ldg(fp + 2);
          } else {
            r[k] = ldg_stream_i32(rank_img + o);
          }
        }
      }
      SC_STAMP(1);   // loads of the pass issued
      if (prm.from_geometry) {
        if (d0 == 0) {   // CTA-uniform
          if (threadIdx.x < 12) s_cam[threadIdx.x] = cam_val;
          __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int dd = d0 + (k * kScWarps + warp) * BPI + bs;
          // computed for every lane (out-of-range lanes hold zeros): no reconvergence scope around the arithmetic
          float x, y, z;
          cam_point_of(fu[k], fv[k], fd[k], s_cam, x, y, z);
          int vx, vy, vz;
          const bool ok = voxel_index3(x, y, z, prm.lo, prm.dx, prm.inv, prm.nx, prm.ny, prm.nz, vx, vy, vz);
          const bool live = in && dd < prm.d;
          r[k] = (ok && live) ? frame_base + (vz * prm.ny + vy) * prm.nx + vx : -1;
          if (live) rank_img[dd * hw + pix] = r[k];
        }
      }
      SC_STAMP(2);   // geometry + ranks of the pass done
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int dd = d0 + (k * kScWarps + warp) * BPI + bs;
        // column summary for the walk: -1 nothing kept, rank >= 0 all kept rows share that voxel, -2 mixed
        int lead = max(r[k], __shfl_xor_sync(kFullMask, r[k], WB));
        lead = max(lead, __shfl_xor_sync(kFullMask, lead, 2 * WB));
        const unsigned agree = __ballot_sync(kFullMask, r[k] < 0 || r[k] == lead);
        const bool shared = (agree & colmask) == colmask;
        if (dd < prm.d) {
          reinterpret_cast<int*>(s_rank4 + dd * WB + wl)[hl] = r[k];
          reinterpret_cast<float*>(s_depth4 + dd * WB + wl)[hl] = r[k] >= 0 ? dv[k] : 0.f;   // dropped: weight 0
          if (hl == 0) s_lead[dd * WB + wl] = lead < 0 ? -1 : (shared ? lead : -2);
        }
      }
    }
  }
  SC_STAMP(3);
  __syncthreads();
  SC_STAMP(4);

  const int wcol = warp % WB, half = warp / WB;
  const int ww = w0 + wcol;   // this warp's image column
  if (ww >= prm.w) return;    // warp-uniform; no barrier below
  const int d_lo = half * (d_pad / HALVES), d_hi = d_lo + d_pad / HALVES;   // this warp's share of the depth range
  const int rl = X ? 16 : c4;                  // lanes a row occupies
  const bool act = sl < rl;
  const int sc = act ? sl : rl - 1;            // idle lanes alias the last slice; they never issue a RED
  Frag<X> fv[kScH], acc[kScH];
  int cur[kScH];
#pragma unroll
  for (int p = 0; p < kScH; ++p) {
    acc[p] = frag_zero<X>();
    cur[p] = -1;
    fv[p] = (h0 + p < prm.h) ? frag_load<T, X>(feat + ((int64_t)bn * hw + (h0 + p) * prm.w + ww) * C, sc) : frag_zero<X>();
  }
  const int4* rank_col = s_rank4 + wcol + grp * WB;       // this lane group's first bin
  const float4* depth_col = s_depth4 + wcol + grp * WB;
  const int* lead_col = s_lead + wcol + grp * WB;

  // point (row p, rank rp, depth dp): join the accumulator holding voxel rp, else evict row p's accumulator
#define BEVPOOL_SC_PUT(p, rp, dp)                                                        \
  if ((rp) >= 0) {                                                                       \
    if ((rp) == cur[0]) acc[0] = frag_fma<X>(fv[p], (dp), acc[0]);                       \
    else if ((rp) == cur[1]) acc[1] = frag_fma<X>(fv[p], (dp), acc[1]);                  \
    else if ((rp) == cur[2]) acc[2] = frag_fma<X>(fv[p], (dp), acc[2]);                  \
    else if ((rp) == cur[3]) acc[3] = frag_fma<X>(fv[p], (dp), acc[3]);                  \
    else {                                                                               \
      if (cur[p] >= 0 && act) frag_red<X>(acc_grid + (int64_t)cur[p] * C, sc, acc[p]);   \
      cur[p] = (rp);                                                                     \
      acc[p] = frag_mul<X>(fv[p], (dp));                                                 \
    }                                                                                    \
  }

  // The arrays are padded to a multiple of 4 bins (pad bins: lead = -1), so the loop needs no bounds tests and
  // every shared-memory access has an immediate offset. (Routing the summaries through a warp reduction to get
  // them into uniform registers removes the BSSY/BSYNC bookkeeping but CREDUX costs as much: measured equal.)
  int cur0 = -1;
#ifdef BEVPOOL_TIMELINE
  if (__any_sync(kFullMask, fv[0].v.x == 1.2345e33f)) tl[7] = 1;   // feature rows have landed
#endif
  SC_STAMP(5);
  for (int d0 = d_lo; d0 < d_hi; d0 += 4 * G) {
    int lead[4];
    float4 dp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lead[u] = lead_col[(d0 + G * u) * WB];     // broadcast LDS (per lane group)
      dp[u] = depth_col[(d0 + G * u) * WB];      // broadcast LDS.128
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int lu = lead[u];
      if (lu == -1) continue;   // nothing kept in this bin
      if (lu >= 0) {
        // fast path (every bin of a Z == 1 grid, most bins otherwise): all kept rows of the column share ONE
        // voxel. One comparison against the open accumulator, four unconditional FMAs (dropped rows weigh 0).
        if (lu != cur0) {
          if (cur0 >= 0 && act) frag_red<X>(acc_grid + (int64_t)cur0 * C, sc, acc[0]);
          cur0 = lu;
          acc[0] = frag_zero<X>();
        }
        acc[0] = frag_fma<X>(fv[0], dp[u].x, acc[0]);
        acc[0] = frag_fma<X>(fv[1], dp[u].y, acc[0]);
        acc[0] = frag_fma<X>(fv[2], dp[u].z, acc[0]);
        acc[0] = frag_fma<X>(fv[3], dp[u].w, acc[0]);
        continue;
      }
      cur[0] = cur0;
      const int4 r = rank_col[(d0 + G * u) * WB];
      BEVPOOL_SC_PUT(0, r.x, dp[u].x)
      BEVPOOL_SC_PUT(1, r.y, dp[u].y)
      BEVPOOL_SC_PUT(2, r.z, dp[u].z)
      BEVPOOL_SC_PUT(3, r.w, dp[u].w)
      cur0 = cur[0];
    }
  }
  cur[0] = cur0;
#undef BEVPOOL_SC_PUT
  if (act) {
#pragma unroll
    for (int p = 0; p < kScH; ++p)
      if (cur[p] >= 0) frag_red<X>(acc_grid + (int64_t)cur[p] * C, sc, acc[p]);
  }
#ifdef BEVPOOL_TIMELINE
  if (threadIdx.x == 0) {
    const int slot = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (slot < 4096) {
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      tl[6] = sc_now();
      tl[7] = smid;
      for (int k = 0; k < 8; ++k) g_sc_timeline[8 * slot + k] = tl[k];
    }
  }
#endif
}

// fp32 [F][V][C] -> T [F][C][V]; a CTA moves all channels of 64 consecutive voxels, 128-bit on both sides.
constexpr int kAlCols = 64;
template <typename T>
__global__ void __launch_bounds__(256)
acc_layout_kernel(const float* __restrict__ src_cl, T* __restrict__ dst, int c, int64_t vpf, int64_t tiles_per_frame) {
  extern __shared__ float t[];            // [c][kAlCols + 1]
  pdl_wait();
  const int64_t b = blockIdx.x / tiles_per_frame;
  const int64_t v0 = (blockIdx.x % tiles_per_frame) * kAlCols;
  const int ncol = (int)min((int64_t)kAlCols, vpf - v0);
  const int64_t rank0 = b * vpf + v0;
  const int c4 = c >> 2;
  for (int i = threadIdx.x; i < ncol * c4; i += 256) {
    const int col = i / c4, q = i - col * c4;
    const float4 v = Vec4<float>::load_stream(src_cl, (rank0 + col) * c + 4 * q);
    float* o = t + (4 * q) * (kAlCols + 1) + col;
    o[0] = v.x; o[kAlCols + 1] = v.y; o[2 * (kAlCols + 1)] = v.z; o[3 * (kAlCols + 1)] = v.w;
  }
  __syncthreads();
  T* d = dst + (b * c) * vpf + v0;
  const bool vec = ncol == kAlCols && (vpf & 3) == 0 && ((((uintptr_t)dst) & 15) == 0);
  if (vec) {
    for (int i = threadIdx.x; i < c * (kAlCols / 4); i += 256) {
      const int ch = i / (kAlCols / 4), q = i % (kAlCols / 4);
      const float* p = t + ch * (kAlCols + 1) + 4 * q;
      Vec4<T>::store(d, (int64_t)ch * vpf + 4 * q, make_float4(p[0], p[1], p[2], p[3]));
    }
  } else {
    for (int i = threadIdx.x; i < c * kAlCols; i += 256) {
      const int ch = i / kAlCols, q = i % kAlCols;
      if (q < ncol) Vec4<T>::store1s(d, (int64_t)ch * vpf + q, t[ch * (kAlCols + 1) + q]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
acc_convert_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n4) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256)
    Vec4<T>::store(dst, 4 * i, Vec4<float>::load_stream(src, 4 * i));
}

template <typename T, int CH4, int HALVES>
static void scatter_launch_h(const void* depth, const void* feat, const float* frustum, const float* rots, const float* trans,
                             const ScatterParams& prm, int32_t* point_rank, float* acc, unsigned blocks, size_t smem,
                             cudaStream_t st) {
  if (ensure_dynamic_smem(view_fwd_scatter_kernel<T, CH4, HALVES>, smem)) return;   // reported by launch_status()
  (void)blocks;
  launch_pdl(view_fwd_scatter_kernel<T, CH4, HALVES>, dim3((unsigned)prm.blocks_w, (unsigned)prm.blocks_h, (unsigned)prm.bn),
             dim3(kScThreads), smem, st, (const T*)depth, (const T*)feat, frustum, rots, trans, prm, point_rank, acc);
}

// Few waves of 8-column CTAs: use 4-column CTAs with the depth range split over two warps (shorter tail).
template <typename T, int CH4>
static int scatter_launch(const void* depth, const void* feat, const float* frustum, const float* rots, const float* trans,
                          ScatterParams prm, int32_t* point_rank, float* acc, cudaStream_t st) {
  prm.blocks_h = (prm.h + kScH - 1) / kScH;
  const int64_t blocks8 = (int64_t)prm.bn * ((prm.w + kScW - 1) / kScW) * prm.blocks_h;
  const char* env = getenv("BEVPOOL_FWD_HALVES");
  const bool split = env ? atoi(env) == 2 : blocks8 < (int64_t)kNumSMs * 3 * 4;   // under 4 waves (measured: no gain beyond)
  const int wb = split ? kScW / 2 : kScW;
  prm.blocks_w = (prm.w + wb - 1) / wb;
  const int64_t blocks = (int64_t)prm.bn * prm.blocks_w * prm.blocks_h;
  if (blocks > INT32_MAX || prm.blocks_h > 65535 || prm.bn > 65535) return BEVPOOL_ERR_OVERFLOW;
  const size_t smem = (size_t)((prm.d + 31) & ~31) * wb * (sizeof(int4) + sizeof(float4) + sizeof(int));
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_ARG;
  if (blocks == 0) return 0;
  if (split) scatter_launch_h<T, CH4, 2>(depth, feat, frustum, rots, trans, prm, point_rank, acc, (unsigned)blocks, smem, st);
  else scatter_launch_h<T, CH4, 1>(depth, feat, frustum, rots, trans, prm, point_rank, acc, (unsigned)blocks, smem, st);
  count_launch();
  return 0;
}

template <typename T>
static int view_forward_t(const void* depth, const void* feat, const float* frustum, const float* rots, const float* trans,
                          ScatterParams prm, int32_t* point_rank, void* out, int64_t n_frames, int64_t rows_per_frame,
                          int layout, void* scratch, cudaStream_t st) {
  const int64_t n_vox = (int64_t)(prm.bn / prm.n_cams) * prm.nx * prm.ny * prm.nz;
  const bool direct = layout == BEVPOOL_LAYOUT_BZYXC && sizeof(T) == 4;   // REDs land in the caller's tensor
  float* acc = direct ? (float*)out : (float*)scratch;
  cudaMemsetAsync(acc, 0, (size_t)n_vox * prm.c * sizeof(float), st);
  int rc;
  switch (prm.c) {
    case 32: rc = scatter_launch<T, 8>(depth, feat, frustum, rots, trans, prm, point_rank, acc, st); break;
    case 64: rc = scatter_launch<T, 16>(depth, feat, frustum, rots, trans, prm, point_rank, acc, st); break;
    case 80: rc = scatter_launch<T, 20>(depth, feat, frustum, rots, trans, prm, point_rank, acc, st); break;
    case 128: rc = scatter_launch<T, 32>(depth, feat, frustum, rots, trans, prm, point_rank, acc, st); break;
    default: rc = scatter_launch<T, 0>(depth, feat, frustum, rots, trans, prm, point_rank, acc, st); break;
  }
  if (rc) return rc;
  if (direct) return launch_status();
  if (layout == BEVPOOL_LAYOUT_BZYXC) {
    const int64_t n4 = n_vox * prm.c / 4;
    int64_t blocks_c = (n4 + 255) / 256;
    if (blocks_c > (int64_t)kNumSMs * 16) blocks_c = (int64_t)kNumSMs * 16;
    if (blocks_c > 0) {
      launch_pdl(acc_convert_kernel<T>, dim3((unsigned)blocks_c), dim3(256), 0, st, (const float*)acc, (T*)out, n4);
      count_launch();
    }
    return launch_status();
  }
  const int64_t vpf = rows_per_frame * prm.nx;
  const int64_t tiles_per_frame = (vpf + kAlCols - 1) / kAlCols;
  const int64_t total = tiles_per_frame * n_frames;
  if (total > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  if (total > 0) {
    const size_t sm = sizeof(float) * (size_t)prm.c * (kAlCols + 1);
    if (int rc = ensure_dynamic_smem(acc_layout_kernel<T>, sm)) return rc;
    launch_pdl(acc_layout_kernel<T>, dim3((unsigned)total), dim3(256), sm, st, (const float*)acc, (T*)out, prm.c, vpf,
               tiles_per_frame);
    count_launch();
  }
  return launch_status();
}

}  // namespace bevpool

using namespace bevpool;

extern "C" size_t bevpool_view_forward_scratch_bytes(int64_t n_voxels, int c, int layout, int dtype) {
  if (n_voxels < 0 || c <= 0) return 0;
  if (layout == BEVPOOL_LAYOUT_BZYXC && dtype == BEVPOOL_F32) return 0;
  return (size_t)n_voxels * c * sizeof(float);
}

extern "C" int bevpool_view_forward(const void* depth, const void* feat, const float* frustum, const float* rots,
                                    const float* trans, const bevpool_grid_t* g, int c, int32_t* point_rank,
                                    int from_geometry, void* out, int64_t n_frames, int64_t rows_per_frame, int layout,
                                    int dtype, void* scratch, size_t scratch_bytes, void* stream) {
  if (!g || g->b < 0 || g->n <= 0 || g->d <= 0 || g->h < 0 || g->w < 0) return BEVPOOL_ERR_BAD_ARG;
  if (g->nx[0] <= 0 || g->nx[1] <= 0 || g->nx[2] <= 0) return BEVPOOL_ERR_BAD_ARG;
  if (c <= 0 || c % 4 || c > 128) return BEVPOOL_ERR_BAD_CHANNELS;
  if (layout != BEVPOOL_LAYOUT_BZYXC && layout != BEVPOOL_LAYOUT_BCZYX) return BEVPOOL_ERR_BAD_ARG;
  if (dtype != BEVPOOL_F32 && dtype != BEVPOOL_BF16) return BEVPOOL_ERR_BAD_ARG;
  const int64_t n_vox = (int64_t)g->b * g->nx[0] * g->nx[1] * g->nx[2];
  const int64_t p0 = (int64_t)g->b * g->n * g->d * g->h * g->w;
  if (n_vox >= INT32_MAX || p0 >= INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  if (layout == BEVPOOL_LAYOUT_BCZYX && n_frames * rows_per_frame * g->nx[0] != n_vox) return BEVPOOL_ERR_BAD_ARG;
  if (n_vox == 0) return BEVPOOL_OK;
  if (!out || !point_rank) return BEVPOOL_ERR_BAD_ARG;
  if (p0 > 0 && (!depth || !feat)) return BEVPOOL_ERR_BAD_ARG;
  if (from_geometry && p0 > 0 && (!frustum || !rots || !trans)) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)feat % 16 || (uintptr_t)out % 16 || (uintptr_t)scratch % 16) return BEVPOOL_ERR_BAD_ARG;
  if (scratch_bytes < bevpool_view_forward_scratch_bytes(n_vox, c, layout, dtype)) return BEVPOOL_ERR_WORKSPACE;
  if (scratch_bytes && !scratch) return BEVPOOL_ERR_BAD_ARG;
  ScatterParams prm;
  prm.c = c;
  prm.d = g->d;
  prm.h = g->h;
  prm.w = g->w;
  prm.bn = g->b * g->n;
  prm.n_cams = g->n;
  prm.nx = g->nx[0];
  prm.ny = g->nx[1];
  prm.nz = g->nx[2];
  for (int a = 0; a < 3; ++a) {
    prm.lo[a] = g->lo[a];
    prm.dx[a] = g->dx[a];
    prm.inv[a] = exact_reciprocal_or_zero(g->dx[a]);
  }
  prm.from_geometry = from_geometry ? 1 : 0;
  prm.blocks_w = prm.blocks_h = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    return view_forward_t<float>(depth, feat, frustum, rots, trans, prm, point_rank, out, n_frames, rows_per_frame, layout,
                                 scratch, st);
  return view_forward_t<__nv_bfloat16>(depth, feat, frustum, rots, trans, prm, point_rank, out, n_frames, rows_per_frame,
                                       layout, scratch, st);
}

#ifdef BEVPOOL_TIMELINE
extern "C" int bevpool_debug_scatter_timeline(unsigned long long* host_out, int n) {
  cudaMemcpyFromSymbol(host_out, bevpool::g_sc_timeline, sizeof(unsigned long long) * 8 * (n < 4096 ? n : 4096));
  return (int)cudaGetLastError();
}
#endif
