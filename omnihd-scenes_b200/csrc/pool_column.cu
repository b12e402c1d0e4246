// Column-GEMM backward for column-coherent grids (Z == 1 BEV grids): same contract as pool_bwd_joint_kernel
// (pool_dense.cu) — sort-free, walks point_rank, writes every element of depth_grad and feat_grad — restructured so
// that the arithmetic is two small dense products per image column and the gathered out_grad rows are moved by the
// Blackwell bulk-copy engine instead of by the warps.
//
// Structure that is exploited. On a Z == 1 grid the 16 pixels (h0..h0+15, w) of an image column land, at one depth
// bin d, in the same voxel or outside the z-range (the ray of pixel (u, v) at depth d differs between rows only in
// height). So a column has at most D "items": item = (bin, voxel rank, 16-bit mask of the rows that sit in it). With
//     R[i][c]  = out_grad[rank_i][c]              the gathered row of item i                (n_items x C)
//     Wt[i][h] = depth[h][bin_i] if h in mask_i   the depth weights of the column           (n_items x 16)
//     F[h][c]  = feat[h][w][c]                    the column's 16 feature rows              (16 x C)
// the whole backward of the column is
//     feat_grad[h][c]      = sum_i Wt[i][h] * R[i][c]          (16 x C,       K = n_items)
//     depth_grad[bin_i][h] = sum_c R[i][c] * F[h][c]           (n_items x 16, K = C), kept for h in mask_i.
// One out_grad row is gathered per ITEM (not per point, and not per 4-pixel group as in the joint kernel): 16x fewer
// gathered bytes than the per-point formulation, and both products run as packed FP32 FMAs on register tiles.
//
// A warp owns one column. Lane (rg, cg) = (lane >> 3, lane & 7) holds a 4-row x C/8-channel tile of F and of the
// feat_grad accumulator (rows 4rg..4rg+3; channels 32k + 4cg.. and 64 + 2cg..): per item it reads C/8 values of R and 4
// weights from shared memory for 8 * C/8 FMAs — the register tile that minimises shared-memory traffic per FMA (the
// first versions, 1 row x C/4 channels per lane on two warps per column, were bound by the shared-memory pipe: 44
// wavefront cycles per item against 15 here; profiles/r2_ncu_bwd_column.md). The dot products are completed across the
// 8 channel lanes by a reduce-scatter (7 shuffles per item pair), after which every lane owns one (item, row) result.
// The rows of R for the next kChunk items are fetched with cp.async.bulk (one C*e-byte copy per row, issued by one lane
// each, completion counted on an mbarrier) into a two-stage ring per warp while the current chunk is consumed — the
// gather latency never sits in a register dependency chain.
// A CTA is 8 adjacent columns (8 consecutive w = one 32-byte sector of depth / point_rank per (d, h)); ranks are reduced
// to per-(column, bin) summaries (lead rank, row mask) while they are staged, so the rank array never lives in shared
// memory. Bins whose kept rows do NOT all share one voxel (camera roll / pitch; never on the synthetic ring) are
// finished on a slow path, one extra item per additional voxel, with the row loaded straight into registers: results
// are exact on any grid, only the speed differs. Z > 1 grids keep the block kernels (pool_dense.cu): there an item is
// barely longer than a point and the dense 16-row products would be wasted.
#include <stdlib.h>

#include "common.cuh"

namespace bevpool {

constexpr int kCgRows = 16;                 // image rows per column tile
// Columns per CTA (template parameter WL of the kernel; warp = column, WL * 32 threads): 8 consecutive w = one 32-byte
// sector of depth / point_rank per (d, h), two CTAs per SM; or 16 = 64 bytes per (d, h), one 16-warp CTA per SM. The
// staging phase and the write-outs are bound by L1 wavefronts (one per distinct 128-byte line a warp access touches,
// profiles/r2_ncu_bwd_column.md): 16-column tiles halve them.
constexpr int kCgChunk = 8;                 // items per stage of the out_grad row ring (a multiple of 2)
constexpr int kCgStages = 2;

#ifdef BEVPOOL_TIMELINE   // measurement builds only (profiles/timeline_col.py): per-CTA phase stamps in ns
__device__ unsigned long long g_col_timeline[8 * 4096];
__device__ __forceinline__ unsigned long long col_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define COL_STAMP(k) do { if (threadIdx.x == 0) tl[k] = col_now(); } while (0)
#else
#define COL_STAMP(k) do { } while (0)
#endif

struct ColParams {
  int d, h, w;
  int feat_grad_nchw;
  int d_pad, cs, hw;   // (d + 7) & ~7; column stride d_pad * 16 + 4 of the [w][d][16] arrays; h * w (set by the launcher)
};

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion reported to `bar` in bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- a lane's slice of a C-channel row: K4 float4 pieces at channels 32k + 4cg (k < K4) and, if K2, one float2 at
//      32*K4 + 2cg — the 8 lanes cg = 0..7 cover a row with conflict-free 128 / 64-bit accesses. C = 32*K4 + 16*K2.
template <int K4, bool K2>
struct Slice {
  float4 v[K4];
  float2 t;
};
template <typename T>
struct RowIO;
template <>
struct RowIO<float> {
  static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <>
struct RowIO<__nv_bfloat16> {
  using B = __nv_bfloat16;
  static __device__ __forceinline__ float2 un2(uint32_t r) {
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ float4 ld4(const B* p) { return Vec4<B>::unpack(*reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ float2 ld2(const B* p) { return un2(*reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ float4 ldg4(const B* p) { return Vec4<B>::unpack(__ldg(reinterpret_cast<const uint2*>(p))); }
  static __device__ __forceinline__ float2 ldg2(const B* p) { return un2(__ldg(reinterpret_cast<const uint32_t*>(p))); }
  static __device__ __forceinline__ void st2(B* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};
template <typename T, int K4, bool K2, bool GLOBAL>
__device__ __forceinline__ Slice<K4, K2> slice_load(const T* row, int cg) {
  Slice<K4, K2> s;
#pragma unroll
  for (int k = 0; k < K4; ++k)
    s.v[k] = GLOBAL ? RowIO<T>::ldg4(row + 32 * k + 4 * cg) : RowIO<T>::ld4(row + 32 * k + 4 * cg);
  s.t = make_float2(0.f, 0.f);
  if (K2) s.t = GLOBAL ? RowIO<T>::ldg2(row + 32 * K4 + 2 * cg) : RowIO<T>::ld2(row + 32 * K4 + 2 * cg);
  return s;
}
template <int K4, bool K2>
__device__ __forceinline__ Slice<K4, K2> slice_zero() {
  Slice<K4, K2> s;
#pragma unroll
  for (int k = 0; k < K4; ++k) s.v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  s.t = make_float2(0.f, 0.f);
  return s;
}

// One item against this lane's tile (4 rows x C/8 channels): feat_grad[p] += w[p] * R, partial dots <R, F[p]>.
template <int K4, bool K2>
__device__ __forceinline__ void item_fma(const Slice<K4, K2>& r, const float (&w)[4], const Slice<K4, K2> (&fv)[4],
                                         Slice<K4, K2> (&fg)[4], float (&dot)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K4; ++k) {
      fg[p].v[k] = fma4(r.v[k], w[p], fg[p].v[k]);
      a = __ffma2_rn(make_float2(r.v[k].x, r.v[k].y), make_float2(fv[p].v[k].x, fv[p].v[k].y), a);
      a = __ffma2_rn(make_float2(r.v[k].z, r.v[k].w), make_float2(fv[p].v[k].z, fv[p].v[k].w), a);
    }
    if (K2) {
      fg[p].t = __ffma2_rn(r.t, make_float2(w[p], w[p]), fg[p].t);
      a = __ffma2_rn(r.t, fv[p].t, a);
    }
    dot[p] = a.x + a.y;
  }
}

// Sum the partial dots of items A and B (4 rows each) over the 8 channel lanes of a row group: reduce-scatter, 7
// shuffles for 8 values. On return lane (rg, cg) holds the complete dot of item (cg & 4 ? B : A), row 4rg + (cg & 3).
__device__ __forceinline__ float reduce_pair(const float (&a)[4], const float (&b)[4], int cg) {
  const bool h4 = cg & 4, h2 = cg & 2, h1 = cg & 1;
  float k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) k[i] = (h4 ? b[i] : a[i]) + __shfl_xor_sync(kFullMask, h4 ? a[i] : b[i], 4);
  const float k0 = (h2 ? k[2] : k[0]) + __shfl_xor_sync(kFullMask, h2 ? k[0] : k[2], 2);
  const float k1 = (h2 ? k[3] : k[1]) + __shfl_xor_sync(kFullMask, h2 ? k[1] : k[3], 2);
  return (h1 ? k1 : k0) + __shfl_xor_sync(kFullMask, h1 ? k0 : k1, 1);
}

template <typename T, int K4, bool K2, int WL, bool WIDE>
__global__ void __launch_bounds__(WL * 32, 16 / WL)
pool_bwd_column_kernel(const T* __restrict__ og, const T* __restrict__ depth, const T* __restrict__ feat,
                       const int* __restrict__ point_rank, ColParams prm, T* __restrict__ depth_grad,
                       T* __restrict__ feat_grad) {
  constexpr int C = 32 * K4 + (K2 ? 16 : 0);
  constexpr uint32_t kRowBytes = C * sizeof(T);
  constexpr int kCgW = WL, kCgWarps = WL;
  constexpr int HL = 32 / WL;             // rows covered by one warp-wide staging access (4 or 2)
  constexpr int KS = kCgRows / HL;        // row slots per thread while staging (4 or 8)
  constexpr int kBins = 16 / KS;          // depth bins per warp in flight while staging: 32 loads per thread either way
  constexpr unsigned kColBits = WL == 8 ? 0x01010101u : 0x00010001u;   // bit of column 0 in every row of a ballot
  using Sl = Slice<K4, K2>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d_pad = prm.d_pad;
  const int CS = prm.cs;                    // column stride of the [w][d][16] arrays: (4 w + h) mod 32 distinct per warp access
  float* s_depth = reinterpret_cast<float*>(smem_raw);                                // [8][CS] lead-masked depth weights
  T* s_R = reinterpret_cast<T*>(s_depth + kCgW * CS);                                // [WL columns][stages][chunk][C]
  // (wide staging parks the raw ranks, [WL][CS] ints, in the ring's place: the region is the larger of the two)
  constexpr size_t kRingBytes = sizeof(T) * kCgW * kCgStages * kCgChunk * C;
  const size_t ring_region = WIDE ? (kRingBytes > sizeof(int) * kCgW * (size_t)CS ? kRingBytes : sizeof(int) * kCgW * (size_t)CS)
                                  : kRingBytes;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_R) + ring_region);   // [WL][stages]
  int* s_lead = reinterpret_cast<int*>(s_full + kCgW * kCgStages);                   // [WL][d_pad] lead rank, -1 = empty bin
  int* s_items = s_lead + kCgW * d_pad;   // [WL][d_pad]: per-bin summary (row mask | more << 16), later compacted in place to
                                          // the kept bins: bin | row mask << 8 | more << 24
  float* s_tile = s_depth;                // epilogue: [C][129] feat_grad transpose (over s_depth and the rings)
  // depth_grad goes straight to global memory: zeros for dropped points while staging, one 4-byte store per kept point from
  // the lane that holds its dot product (no shared-memory copy of the tile's gradients: half the footprint, D = 118 fits)

  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int hl = lane / WL, wl = lane % WL;
  const int h0 = blockIdx.y * kCgRows, w0 = blockIdx.x * kCgW, bn = blockIdx.z;
  const int hw = prm.hw;
  const int64_t img_base = (int64_t)bn * prm.d * hw;
  pdl_wait();
#ifdef BEVPOOL_TIMELINE
  unsigned long long tl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  COL_STAMP(0);
  if (threadIdx.x < kCgW * kCgStages) mbar_init(s_full + threadIdx.x, 1);

  // ---- stage the depth WEIGHTS of the tile and reduce the ranks to per-(column, bin) summaries. A warp takes bins
  //      warp, warp + 8, ...; one warp-wide load covers rows hl + 4k of 8 consecutive w; 4 bins (32 loads) are in flight
  //      per pass. Summary of a (column, bin): lead = rank of its first kept row, row mask of the rows in that voxel,
  //      "more" if kept rows sit in other voxels too. The staged weight of a row is its depth if it belongs to the lead
  //      voxel and 0 otherwise, so the main loop needs no masking.
  if constexpr (WIDE) {
    // ---- 16-column tiles, W % 4 == 0. A lane takes 4 consecutive columns of one row with 128-bit loads; a warp-wide access
    //      covers rows row8 (+8) of all 16 columns of one depth bin, so the warp that loaded a bin holds all of its 256
    //      values. They go to shared memory as they are (ranks into s_rank, which aliases the not yet used row ring; depth
    //      into its final place) and then ONE LANE per (column, bin) reduces the 16 ranks of its pair to the summary and
    //      masks the weights in place: ~150 instructions per pair and lane instead of ~100 ballot / shuffle instructions per
    //      pair and WARP (the ballot form spent 3.7 M of the kernel's 17.4 M warp instructions here, r2_col_e capture).
    static_assert(!WIDE || WL == 16, "wide staging is written for 16-column tiles");
    constexpr int kBinsW = 4;   // depth bins per warp in flight: 16 128-bit loads per thread
    int* s_rank = reinterpret_cast<int*>(s_R);   // [16][CS] raw ranks, dead before the first bulk copy is issued
    const int row8 = lane >> 2, g = lane & 3;
    const int wq = w0 + 4 * g;
    const bool q_in = wq < prm.w;
    for (int d0 = warp; d0 < d_pad; d0 += kBinsW * kCgWarps) {
      int4 r[kBinsW][2];
      float4 dv[kBinsW][2];
#pragma unroll
      for (int u = 0; u < kBinsW; ++u) {
        const int d = d0 + u * kCgWarps;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int h = h0 + 8 * sl + row8;
          r[u][sl] = make_int4(-1, -1, -1, -1);
          dv[u][sl] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q_in && h < prm.h && d < prm.d) {
            const int64_t o = img_base + (int64_t)d * hw + h * prm.w + wq;
            r[u][sl] = ldg_stream_i32x4(point_rank + o);
            dv[u][sl] = Vec4<T>::load_stream(depth, o);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBinsW; ++u) {
        const int d = d0 + u * kCgWarps;
        if (d >= d_pad) continue;   // warp-uniform
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int4 q = r[u][sl];
          const float4 v = dv[u][sl];
          // zeros for dropped points (nobody else writes them): one 16-byte store if all 4 are
          const int h = h0 + 8 * sl + row8;
          if (q_in && h < prm.h && d < prm.d) {
            const int64_t o = img_base + (int64_t)d * hw + h * prm.w + wq;
            if ((q.x & q.y & q.z & q.w) < 0) {
              Vec4<T>::store(depth_grad, o, make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
              if (q.x < 0) Vec4<T>::store1s(depth_grad, o + 0, 0.f);
              if (q.y < 0) Vec4<T>::store1s(depth_grad, o + 1, 0.f);
              if (q.z < 0) Vec4<T>::store1s(depth_grad, o + 2, 0.f);
              if (q.w < 0) Vec4<T>::store1s(depth_grad, o + 3, 0.f);
            }
          }
          const int e = (4 * g) * CS + d * kCgRows + 8 * sl + row8;   // column 4g, this row
          s_rank[e] = q.x, s_rank[e + CS] = q.y, s_rank[e + 2 * CS] = q.z, s_rank[e + 3 * CS] = q.w;
          s_depth[e] = v.x, s_depth[e + CS] = v.y, s_depth[e + 2 * CS] = v.z, s_depth[e + 3 * CS] = v.w;
        }
      }
      __syncwarp();
      // one lane per (column, bin): lanes 0-15 take the columns of bin 2 pass, lanes 16-31 those of bin 2 pass + 1
#pragma unroll
      for (int pass = 0; pass < kBinsW / 2; ++pass) {
        const int d = d0 + (2 * pass + (lane >> 4)) * kCgWarps;
        const int col = lane & 15;
        if (d < d_pad) {
          const int4* pr4 = reinterpret_cast<const int4*>(s_rank + col * CS + d * kCgRows);
          float4* pw4 = reinterpret_cast<float4*>(s_depth + col * CS + d * kCgRows);
          int rk[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int4 t = pr4[q4];
            rk[4 * q4] = t.x, rk[4 * q4 + 1] = t.y, rk[4 * q4 + 2] = t.z, rk[4 * q4 + 3] = t.w;
          }
          unsigned kept = 0, same = 0;
          int lead = -1;
#pragma unroll
          for (int i2 = 15; i2 >= 0; --i2) {
            kept |= (unsigned)(rk[i2] >= 0) << i2;
            lead = rk[i2] >= 0 ? rk[i2] : lead;   // ends as the rank of the first kept row
          }
#pragma unroll
          for (int i2 = 0; i2 < 16; ++i2) same |= (unsigned)(rk[i2] == lead) << i2;
          same &= kept;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 t = pw4[q4];
            t.x = (same >> (4 * q4 + 0)) & 1u ? t.x : 0.f;
            t.y = (same >> (4 * q4 + 1)) & 1u ? t.y : 0.f;
            t.z = (same >> (4 * q4 + 2)) & 1u ? t.z : 0.f;
            t.w = (same >> (4 * q4 + 3)) & 1u ? t.w : 0.f;
            pw4[q4] = t;
          }
          s_lead[col * d_pad + d] = kept ? lead : -1;
          s_items[col * d_pad + d] = (int)(same | (kept != same ? 1u << 16 : 0u));
        }
      }
      __syncwarp();
    }
    // the ring that aliases s_rank is written next by the async proxy (bulk copies)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  } else
  {
    const bool w_in = w0 + wl < prm.w;
    // rows HL*k .. HL*k + HL-1 of column wl out of a ballot: bit WL*hl' + wl per row hl'
    auto rows_of = [&](unsigned ballot) -> unsigned {
      const unsigned b = (ballot >> wl) & kColBits;
      return WL == 8 ? (b * 0x10204080u) >> 28 : (b | (b >> 15)) & 3u;
    };
    for (int d0 = warp; d0 < d_pad; d0 += kBins * kCgWarps) {
      int r[kBins][KS];
      float dv[kBins][KS];
#pragma unroll
      for (int u = 0; u < kBins; ++u) {
        const int d = d0 + u * kCgWarps;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int h = h0 + HL * k + hl;
          r[u][k] = -1;
          dv[u][k] = 0.f;
          if (w_in && h < prm.h && d < prm.d) {
            const int64_t o = img_base + (int64_t)d * hw + h * prm.w + w0 + wl;
            r[u][k] = ldg_stream_i32(point_rank + o);
            dv[u][k] = Vec4<T>::load1(depth, o);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBins; ++u) {
        const int d = d0 + u * kCgWarps;
        // kept rows of column wl as a 16-bit mask, and the rank of its first kept row (lower slots override)
        unsigned kept = 0;
        int lead = -1;
#pragma unroll
        for (int k = KS - 1; k >= 0; --k) {
          const unsigned kq = rows_of(__ballot_sync(kFullMask, r[u][k] >= 0));   // kept rows HL*k .. of column wl
          // row HL*k + hl' sits in lane WL*hl' + wl
          const int first = __shfl_sync(kFullMask, r[u][k], WL * (kq ? __ffs(kq) - 1 : 0) + wl);
          lead = kq ? first : lead;
          kept |= kq << (HL * k);
        }
        unsigned same = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k)
          same |= rows_of(__ballot_sync(kFullMask, r[u][k] >= 0 && r[u][k] == lead)) << (HL * k);
        if (d < d_pad) {
          float* pd = s_depth + wl * CS + d * kCgRows + hl;
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            pd[HL * k] = (r[u][k] >= 0 && r[u][k] == lead) ? dv[u][k] : 0.f;
            // depth_grad of a dropped point is 0 and nobody else writes it: stored here, WL consecutive w per (d, h);
            // every kept point is written exactly once by the item that owns it (main loop or slow path)
            const int h = h0 + HL * k + hl;
            if (r[u][k] < 0 && w_in && h < prm.h && d < prm.d)
              Vec4<T>::store1s(depth_grad, img_base + (int64_t)d * hw + h * prm.w + w0 + wl, 0.f);
          }
          if (hl == 0) {
            s_lead[wl * d_pad + d] = kept ? lead : -1;
            s_items[wl * d_pad + d] = (int)(same | (kept != same ? 1u << 16 : 0u));
          }
        }
      }
    }
  }
  // make the barrier initialisation visible to the async proxy before any bulk copy names it
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  COL_STAMP(1);

  const int ww = w0 + warp;                 // this warp's image column
  const int rg = lane >> 3, cg = lane & 7;  // row group (rows 4rg .. 4rg + 3), channel group
  Sl fv[4], fg[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    fg[p] = slice_zero<K4, K2>();
    fv[p] = (ww < prm.w && h0 + 4 * rg + p < prm.h)
                ? slice_load<T, K4, K2, true>(feat + ((int64_t)bn * hw + (h0 + 4 * rg + p) * prm.w + ww) * C, cg)
                : slice_zero<K4, K2>();
  }

  if (ww < prm.w) {   // warp-uniform
    const int* lead_col = s_lead + warp * d_pad;
    int* items = s_items + warp * d_pad;
    const float* depth_col = s_depth + warp * CS + 4 * rg;
    // depth_grad of (bin, row) is addressed with ONE 32-bit offset from the tensor base (the launcher checks that the
    // tensor has fewer than 2^31 elements): the address is rebuilt per store — no register survives the FMA block —
    // and this form costs 3 instructions instead of a 64-bit multiply chain
    T* dg_img = depth_grad;
    T* ring = s_R + (size_t)warp * kCgStages * kCgChunk * C;
    uint64_t* full = s_full + warp * kCgStages;

    // kept bins of the column, in bin order, compacted in place over the per-bin summaries (position <= bin, and
    // every lane has read its summary before any lane writes)
    int n_items = 0;
    bool any_more = false;
    for (int d0 = 0; d0 < d_pad; d0 += 32) {
      const int d = d0 + lane;
      const int meta = d < d_pad ? items[d] : 0;
      const bool kept = d < d_pad && lead_col[d] >= 0;
      const unsigned b = __ballot_sync(kFullMask, kept);
      any_more |= __any_sync(kFullMask, kept && (meta >> 16));
      __syncwarp();
      if (kept) items[n_items + __popc(b & ((1u << lane) - 1u))] = d | (meta << 8);   // mask: bits 8..23, more: bit 24
      n_items += __popc(b);
    }
    __syncwarp();

    auto issue = [&](int chunk, int stage) {
      const int j0 = chunk * kCgChunk;
      const int n = min(kCgChunk, n_items - j0);
      if (lane == 0) mbar_arrive_expect_tx(full + stage, (uint32_t)n * kRowBytes);
      __syncwarp();
      if (lane < n) {
        const int rank = lead_col[items[j0 + lane] & 255];
        bulk_g2s(ring + ((size_t)stage * kCgChunk + lane) * C, og + (int64_t)rank * C, kRowBytes, full + stage);
      }
    };
    // one lead item: the staged weights of this lane's 4 rows (0 outside the lead voxel), then the FMAs
    auto run_item = [&](const Sl& r, int bin, float (&dot)[4]) {
      const float4 dp = *reinterpret_cast<const float4*>(depth_col + bin * kCgRows);
      const float w[4] = {dp.x, dp.y, dp.z, dp.w};
      item_fma<K4, K2>(r, w, fv, fg, dot);
    };
    const int my_row = 4 * rg + (cg & 3);   // the row whose dot product this lane holds after reduce_pair
    const int dg_lane = (int)img_base + (h0 + my_row) * prm.w + ww;

    const int n_chunks = (n_items + kCgChunk - 1) / kCgChunk;
    if (n_chunks > 0) issue(0, 0);
    COL_STAMP(2);
#ifdef BEVPOOL_TIMELINE
    if (threadIdx.x == 0) tl[7] = ((unsigned long long)n_items << 16) | ((unsigned long long)any_more << 15);
#endif
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int stage = ch & 1;
      if (ch + 1 < n_chunks) issue(ch + 1, stage ^ 1);
      mbar_wait(full + stage, (uint32_t)(ch >> 1) & 1u);
#ifdef BEVPOOL_TIMELINE
      if (ch == 0) COL_STAMP(3);
#endif
      const int j0 = ch * kCgChunk;
      const int np = min(kCgChunk, n_items - j0) >> 1;   // full pairs of this chunk; an odd last item is done after the loop
      const T* rows = ring + (size_t)stage * kCgChunk * C;
#pragma unroll 1
      for (int pr = 0; pr < kCgChunk / 2; ++pr) {
        if (pr >= np) break;   // warp-uniform
        const int rec_a = items[j0 + 2 * pr], rec_b = items[j0 + 2 * pr + 1];
        float da[4], db[4];
        run_item(slice_load<T, K4, K2, false>(rows + (size_t)(2 * pr) * C, cg), rec_a & 255, da);
        run_item(slice_load<T, K4, K2, false>(rows + (size_t)(2 * pr + 1) * C, cg), rec_b & 255, db);
        const float dot = reduce_pair(da, db, cg);
        const int mine = (cg & 4) ? rec_b : rec_a;
        if ((mine >> (8 + my_row)) & 1) Vec4<T>::store1(dg_img, dg_lane + (mine & 255) * hw, dot);
      }
      __syncwarp();   // every lane is done with this stage before the next issue() refills it
    }
    // ---- an odd last item (it sits in the last chunk's stage, which is still valid)
    if (n_items & 1) {
      const int j = n_items - 1;
      const int rec = items[j];
      const T* rowp = ring + ((size_t)((n_chunks - 1) & 1) * kCgChunk + (j - (n_chunks - 1) * kCgChunk)) * C;
      float da[4];
      const float db[4] = {0.f, 0.f, 0.f, 0.f};
      run_item(slice_load<T, K4, K2, false>(rowp, cg), rec & 255, da);
      const float dot = reduce_pair(da, db, cg);
      if (!(cg & 4) && ((rec >> (8 + my_row)) & 1)) Vec4<T>::store1(dg_img, dg_lane + (rec & 255) * hw, dot);
    }

    // ---- slow path: bins whose kept rows sit in more than one voxel; one extra item per additional voxel, its
    //      out_grad row loaded straight into registers
    if (any_more) {
      for (int j = 0; j < n_items; ++j) {
        const int rec = items[j];
        if (!(rec >> 24)) continue;   // warp-uniform
        const int bin = rec & 255;
        int r = -1;
        if (lane < kCgRows && h0 + lane < prm.h)
          r = ldg_stream_i32(point_rank + img_base + (int64_t)bin * hw + (h0 + lane) * prm.w + ww);
        unsigned rest = __ballot_sync(kFullMask, r >= 0) & ~((unsigned)(rec >> 8) & 0xffffu);
        while (rest) {   // warp-uniform
          const int first = __ffs(rest) - 1;
          const int rank = __shfl_sync(kFullMask, r, first);
          const unsigned mask = __ballot_sync(kFullMask, r == rank) & rest;
          rest &= ~mask;
          float da[4], w[4];
          const float db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int p = 0; p < 4; ++p)   // the staged weights are those of the lead voxel: read this item's from memory
            w[p] = ((mask >> (4 * rg + p)) & 1)
                       ? Vec4<T>::load1(depth, img_base + (int64_t)bin * hw + (h0 + 4 * rg + p) * prm.w + ww)
                       : 0.f;
          item_fma<K4, K2>(slice_load<T, K4, K2, true>(og + (int64_t)rank * C, cg), w, fv, fg, da);
          const float dot = reduce_pair(da, db, cg);
          if (!(cg & 4) && ((mask >> my_row) & 1)) Vec4<T>::store1(dg_img, dg_lane + bin * hw, dot);
        }
      }
    }
  }
  COL_STAMP(4);
  __syncthreads();   // all warps are done with s_depth and the rings: the feat_grad tile may overwrite them
  COL_STAMP(5);

  // ---- feat_grad of the 16 x 8 pixels
  if (prm.feat_grad_nchw) {
    constexpr int TS = kCgRows * kCgW + 1;   // channel stride of the transpose tile [c][16 rows][WL w] (129 / 257)
    if (ww < prm.w) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        // column index XOR (row >> 2): the four row groups of a warp would otherwise hit the same banks (both lane
        // coordinates step by 4); the reader below undoes it
        float* t = s_tile + (4 * rg + p) * kCgW + (warp ^ rg);
#pragma unroll
        for (int k = 0; k < K4; ++k) {
          t[(32 * k + 4 * cg + 0) * TS] = fg[p].v[k].x;
          t[(32 * k + 4 * cg + 1) * TS] = fg[p].v[k].y;
          t[(32 * k + 4 * cg + 2) * TS] = fg[p].v[k].z;
          t[(32 * k + 4 * cg + 3) * TS] = fg[p].v[k].w;
        }
        if (K2) {
          t[(32 * K4 + 2 * cg + 0) * TS] = fg[p].t.x;
          t[(32 * K4 + 2 * cg + 1) * TS] = fg[p].t.y;
        }
      }
    }
    __syncthreads();
    if (w0 + wl < prm.w) {
      // runs of WL consecutive w; a warp-wide store covers HL rows of one channel
      for (int cr = warp; cr < C * KS; cr += kCgWarps) {
        const int c = cr / KS, trow = HL * (cr % KS) + hl;
        if (h0 + trow < prm.h)
          Vec4<T>::store1s(feat_grad, ((int64_t)bn * C + c) * hw + (h0 + trow) * prm.w + w0 + wl,
                           s_tile[c * TS + trow * kCgW + (wl ^ (trow >> 2))]);
      }
    }
  } else if (ww < prm.w) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (h0 + 4 * rg + p < prm.h) {
        T* o = feat_grad + ((int64_t)bn * hw + (h0 + 4 * rg + p) * prm.w + ww) * C;
#pragma unroll
        for (int k = 0; k < K4; ++k) Vec4<T>::store(o, 32 * k + 4 * cg, fg[p].v[k]);
        if (K2) RowIO<T>::st2(o + 32 * K4 + 2 * cg, fg[p].t);
      }
  }
#ifdef BEVPOOL_TIMELINE
  __syncthreads();
  if (threadIdx.x == 0) {
    const int slot = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (slot < 4096) {
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      tl[6] = col_now();
      tl[7] |= smid;
      for (int k = 0; k < 8; ++k) g_col_timeline[8 * slot + k] = tl[k];
    }
  }
#endif
}

template <typename T, int K4, bool K2, int WL, bool WIDE>
static int backward_column_launch_w(const void* og, void* dg, void* fg, const void* depth, const void* feat,
                                    const int* point_rank, int bn, const ColParams& prm, cudaStream_t st) {
  constexpr int C = 32 * K4 + (K2 ? 16 : 0);
  const size_t d_pad = (size_t)((prm.d + 7) & ~7);
  const size_t col_bytes = sizeof(float) * WL * (d_pad * kCgRows + 4);
  size_t ring_bytes = sizeof(T) * WL * kCgStages * kCgChunk * C;
  if (WIDE && ring_bytes < sizeof(int) * WL * (d_pad * kCgRows + 4)) ring_bytes = sizeof(int) * WL * (d_pad * kCgRows + 4);
  const size_t main_bytes = col_bytes + ring_bytes + sizeof(uint64_t) * WL * kCgStages + 2 * sizeof(int) * WL * d_pad;
  const size_t tile_bytes = prm.feat_grad_nchw ? sizeof(float) * C * (kCgRows * WL + 1) : 0;
  const size_t smem = main_bytes > tile_bytes ? main_bytes : tile_bytes;
  // 16 warps per SM (two 8-column CTAs or one 16-column CTA) are what keeps the FMA pipe fed
  if (smem > (size_t)(113 * 1024) * (WL / 8)) return BEVPOOL_ERR_BAD_ARG;
  const int blocks_w = (prm.w + WL - 1) / WL, blocks_h = (prm.h + kCgRows - 1) / kCgRows;
  if (blocks_h > 65535 || bn > 65535) return BEVPOOL_ERR_OVERFLOW;
  auto kern = pool_bwd_column_kernel<T, K4, K2, WL, WIDE>;
  if (int rc = ensure_dynamic_smem(kern, smem)) return rc;
  launch_pdl(kern, dim3((unsigned)blocks_w, (unsigned)blocks_h, (unsigned)bn), dim3(WL * 32), smem, st, (const T*)og,
             (const T*)depth, (const T*)feat, point_rank, prm, (T*)dg, (T*)fg);
  count_launch();
  return launch_status();
}

template <typename T, int K4, bool K2>
static int backward_column_launch(const void* og, void* dg, void* fg, const void* depth, const void* feat,
                                  const int* point_rank, int bn, const ColParams& prm, cudaStream_t st) {
  // Deep frusta (cfg 3: D = 118) take the 8-column kernel as long as two CTAs fit an SM (the shared-memory test in
  // backward_column_launch_w): 146 us against 152 us for the joint kernel at B = 4, 265 against 286 us in the step at
  // B = 8 (before the staging / loop clean-up of this round the joint kernel won there, 152 vs 169 us).
  static const int max_d = [] {
    const char* e = getenv("BEVPOOL_BWD_COLUMN_MAXD");   // measurement only
    return e ? atoi(e) : 248;
  }();
  if (prm.d > max_d) return BEVPOOL_ERR_BAD_ARG;
  // 16-column tiles (64-byte pieces of depth / point_rank / the gradients per (d, h): half the L1 wavefronts of staging
  // and write-out) when they cover the image width as tightly as 8-column tiles do. BEVPOOL_BWD_TILE_W=8|16 overrides
  // (measurement only).
  static const int forced = [] {
    const char* e = getenv("BEVPOOL_BWD_TILE_W");
    return e ? atoi(e) : 0;
  }();
  // the 16-column kernel stages with 128-bit loads: 4 consecutive columns per lane need W % 4 == 0 and 16-byte aligned arrays
  const bool can16 = prm.w % 4 == 0 && ((uintptr_t)point_rank % 16) == 0 && ((uintptr_t)depth % 16) == 0 &&
                     ((uintptr_t)dg % 16) == 0;
  const bool wide = can16 && (forced ? forced == 16 : (prm.w + 15) / 16 * 16 == (prm.w + 7) / 8 * 8);
  if (wide) {
    const int rc = backward_column_launch_w<T, K4, K2, 16, true>(og, dg, fg, depth, feat, point_rank, bn, prm, st);
    if (rc != BEVPOOL_ERR_BAD_ARG) return rc;
  }
  return backward_column_launch_w<T, K4, K2, 8, false>(og, dg, fg, depth, feat, point_rank, bn, prm, st);
}

// Column-GEMM backward when there is an instantiation for the channel count and the tile fits shared memory.
// *handled = false: the caller falls back to the joint / block kernels.
int backward_column(const void* og, void* dg, void* fg, const void* depth, const void* feat, const int* point_rank,
                    int bn, int d, int h, int w, int c, int feat_grad_nchw, int dtype, cudaStream_t st, bool* handled) {
  ColParams prm;
  prm.d = d;
  prm.h = h;
  prm.w = w;
  prm.feat_grad_nchw = feat_grad_nchw;
  prm.d_pad = (d + 7) & ~7;
  prm.cs = prm.d_pad * kCgRows + 4;
  prm.hw = h * w;
  *handled = true;
  if (d > 248 || (int64_t)bn * d * h * w >= INT32_MAX) {
    *handled = false;
    return 0;
  }
  int rc = BEVPOOL_ERR_BAD_ARG;
#define BEVPOOL_COL(T)                                                                                            \
  switch (c) {                                                                                                    \
    case 32: rc = backward_column_launch<T, 1, false>(og, dg, fg, depth, feat, point_rank, bn, prm, st); break;   \
    case 64: rc = backward_column_launch<T, 2, false>(og, dg, fg, depth, feat, point_rank, bn, prm, st); break;   \
    case 80: rc = backward_column_launch<T, 2, true>(og, dg, fg, depth, feat, point_rank, bn, prm, st); break;    \
    default: *handled = false; return 0;                                                                          \
  }
  if (dtype == BEVPOOL_F32) {
    BEVPOOL_COL(float)
  } else {
    BEVPOOL_COL(__nv_bfloat16)
  }
#undef BEVPOOL_COL
  if (rc == BEVPOOL_ERR_BAD_ARG) {   // tile does not fit shared memory (very deep frusta): older kernels
    *handled = false;
    return 0;
  }
  return rc;
}

}  // namespace bevpool

#ifdef BEVPOOL_TIMELINE
extern "C" int bevpool_debug_col_timeline(unsigned long long* host_out, int n) {
  cudaMemcpyFromSymbol(host_out, bevpool::g_col_timeline, sizeof(unsigned long long) * 8 * (n < 4096 ? n : 4096));
  return (int)cudaGetLastError();
}
#endif
