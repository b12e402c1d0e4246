// Shared device/host helpers for the bevpool_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/bevpool_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "bevpool_b200 is written for sm_100a (B200) only"
#endif

namespace bevpool {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr unsigned kFullMask = 0xffffffffu;

void count_launch(int n = 1);  // abi.cu

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The kernels of one step form a dependent chain of short launches; with the programmatic-stream-
// serialization attribute the next grid is set up and its CTAs are dispatched while the previous grid
// drains, and every kernel blocks at pdl_wait() before it touches global memory (a no-op for a kernel
// launched the ordinary way). Captured into CUDA graphs as programmatic dependency edges.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // let the NEXT grid be set up as soon as every CTA of this one has got this far: its CTAs take the slots
  // this grid frees and park at their own pdl_wait() until this grid has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();   // abi.cu (BEVPOOL_PDL=0 disables)

// Pinned landing slot + event for a small device->host hand-back inside one entry point (abi.cu)
struct CountSlot {
  int32_t* host;      // 2 x int32, page-locked
  cudaEvent_t ev;
  int dev;
};
CountSlot* acquire_count_slot();   // nullptr on failure
void release_count_slot(CountSlot* s);

// Opt a kernel in to `smem` bytes of dynamic shared memory (> 48 KB) on the CURRENT device; cached per
// (kernel, device). Returns 0 or the cudaError_t.
int ensure_dynamic_smem_impl(const void* kern, size_t smem);   // abi.cu
template <typename... KArgs>
inline int ensure_dynamic_smem(void (*kern)(KArgs...), size_t smem) {
  return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), smem);
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- cache-hinted memory ops ---------------------------------------------------------------
// Streaming (read-once) data: do not allocate in L1 so gathered feature rows keep it.
__device__ __forceinline__ int ldg_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ldg_stream_i32x4(const int* p) {   // p 16-byte aligned
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Output rows are written once and not re-read by this kernel: streaming store.
__device__ __forceinline__ void stg_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_f32(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_stream_u2(uint2* p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- element access: 4 consecutive channels as float4, fp32 or bf16 storage ------------------
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  // gathered (re-used) rows: default caching (L1 + L2)
  static __device__ __forceinline__ float4 load(const float* base, int64_t elem) {
    return __ldg(reinterpret_cast<const float4*>(base + elem));
  }
  static __device__ __forceinline__ float4 load_stream(const float* base, int64_t elem) {
    return ldg_stream_f4(reinterpret_cast<const float4*>(base + elem));
  }
  static __device__ __forceinline__ void store(float* base, int64_t elem, float4 v) {
    stg_stream_f4(reinterpret_cast<float4*>(base + elem), v);
  }
  // row that a following kernel gathers from: ordinary write-back store (stays in L2)
  static __device__ __forceinline__ void store_keep(float* base, int64_t elem, float4 v) {
    *reinterpret_cast<float4*>(base + elem) = v;
  }
  static __device__ __forceinline__ float load1(const float* base, int64_t elem) { return __ldg(base + elem); }
  static __device__ __forceinline__ void store1(float* base, int64_t elem, float v) { base[elem] = v; }
  static __device__ __forceinline__ void store1s(float* base, int64_t elem, float v) { stg_stream_f32(base + elem, v); }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 unpack(uint2 r) {
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
  }
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* base, int64_t elem) {
    return unpack(__ldg(reinterpret_cast<const uint2*>(base + elem)));
  }
  static __device__ __forceinline__ float4 load_stream(const __nv_bfloat16* base, int64_t elem) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(base + elem));
    return unpack(r);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* base, int64_t elem, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    stg_stream_u2(reinterpret_cast<uint2*>(base + elem), r);
  }
  static __device__ __forceinline__ void store_keep(__nv_bfloat16* base, int64_t elem, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(base + elem) = r;
  }
  static __device__ __forceinline__ float4 load_smem(const __nv_bfloat16* p) {
    return unpack(*reinterpret_cast<const uint2*>(p));
  }
  static __device__ __forceinline__ float load1(const __nv_bfloat16* base, int64_t elem) {
    return __bfloat162float(base[elem]);
  }
  static __device__ __forceinline__ void store1(__nv_bfloat16* base, int64_t elem, float v) {
    base[elem] = __float2bfloat16_rn(v);
  }
  static __device__ __forceinline__ void store1s(__nv_bfloat16* base, int64_t elem, float v) {
    base[elem] = __float2bfloat16_rn(v);
  }
};

// Blackwell packed FP32: one FFMA2 issues two IEEE fused multiply-adds (same rounding as fmaf), which
// halves the FMA instruction count of these issue-bound kernels.
__device__ __forceinline__ float4 fma4(float4 a, float s, float4 acc) {
  const float2 s2 = make_float2(s, s);
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), s2, make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), s2, make_float2(acc.z, acc.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}
// <a, b> with two packed FMAs and one add (summation order differs from dot4: tolerance-level only)
__device__ __forceinline__ float dot4_packed(float4 a, float4 b) {
  float2 p = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(0.f, 0.f));
  p = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), p);
  return p.x + p.y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------ row fragments
__device__ __forceinline__ float4 mul4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// A lane's slice of a C-channel row: channels 4*sl .. 4*sl+3 and, in the "4 + 1" mapping used for C = 80
// (X = true: 16 lanes x 5 channels, so two lane groups fit in a warp), also channel 64 + sl.
template <bool X>
struct Frag {
  float4 v;
  float s;
};
template <typename T, bool X>
__device__ __forceinline__ Frag<X> frag_load(const T* row, int sl) {
  Frag<X> f;
  f.v = Vec4<T>::load(row, 4 * sl);
  f.s = X ? Vec4<T>::load1(row, 64 + sl) : 0.f;
  return f;
}
template <bool X>
__device__ __forceinline__ Frag<X> frag_zero() {
  Frag<X> f;
  f.v = make_float4(0.f, 0.f, 0.f, 0.f);
  f.s = 0.f;
  return f;
}
template <bool X>
__device__ __forceinline__ Frag<X> frag_fma(const Frag<X>& a, float d, Frag<X> acc) {
  acc.v = fma4(a.v, d, acc.v);
  if (X) acc.s = fmaf(a.s, d, acc.s);
  return acc;
}
template <bool X>
__device__ __forceinline__ Frag<X> frag_mul(const Frag<X>& a, float d) {
  Frag<X> r;
  r.v = mul4(a.v, d);
  r.s = X ? a.s * d : 0.f;
  return r;
}
template <bool X>
__device__ __forceinline__ float frag_dot(const Frag<X>& a, const Frag<X>& b) {
  const float d = dot4_packed(a.v, b.v);
  return X ? fmaf(a.s, b.s, d) : d;
}

// ------------------------------------------------------------------------------------------ geometry
// Exactness (SURVEY.md §7 hard part 3): r0*x + r1*y + r2*z + t with every product and sum rounded separately,
// voxel index = trunc((coor - lo) / dx) with an IEEE fp32 subtract and divide (no reciprocal, no FMA).
__device__ __forceinline__ void cam_point(const float* __restrict__ frustum, const float* cam /*R[9] t[3]*/,
                                          int64_t dhw, float& x, float& y, float& z) {
  const float u = __ldg(frustum + 3 * dhw + 0), v = __ldg(frustum + 3 * dhw + 1), dd = __ldg(frustum + 3 * dhw + 2);
  const float px = __fmul_rn(u, dd), py = __fmul_rn(v, dd), pz = dd;
  x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[0], px), __fmul_rn(cam[1], py)), __fmul_rn(cam[2], pz)), cam[9]);
  y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[3], px), __fmul_rn(cam[4], py)), __fmul_rn(cam[5], pz)), cam[10]);
  z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[6], px), __fmul_rn(cam[7], py)), __fmul_rn(cam[8], pz)), cam[11]);
}

// Same arithmetic on frustum values that are already in registers (callers that issue all their loads first)
__device__ __forceinline__ void cam_point_of(float u, float v, float dd, const float* cam, float& x, float& y, float& z) {
  const float px = __fmul_rn(u, dd), py = __fmul_rn(v, dd), pz = dd;
  x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[0], px), __fmul_rn(cam[1], py)), __fmul_rn(cam[2], pz)), cam[9]);
  y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[3], px), __fmul_rn(cam[4], py)), __fmul_rn(cam[5], pz)), cam[10]);
  z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cam[6], px), __fmul_rn(cam[7], py)), __fmul_rn(cam[8], pz)), cam[11]);
}

// Voxel index of one coordinate: trunc((c - lo) / dx) with the reference's IEEE fp32 subtract and DIVIDE, bit for bit.
//   inv > 0 : dx is a power of two and inv = 1/dx exactly: the divide IS a multiply (both are the correctly rounded
//             value of the same real number);
//   inv < 0 : -inv = fl(1/dx). q' = fl(t * fl(1/dx)) is within 2^-23 |q| of the true quotient and so is the IEEE
//             quotient q = fl(t / dx): unless q' lies within a guard band of 2^-21 |q'| of an integer, no integer
//             separates q from q', so trunc(q) == trunc(q') and both compare alike with the integers -1 and n — the
//             divide (about 10 instructions, three per point) is skipped. Inside the band (about 1 point in 10^4),
//             or for NaN, the true divide decides. tests/test_gpu_parity.py::test_voxel_boundaries_bit_exact pins it;
//   inv == 0: always divide.
__device__ __forceinline__ bool voxel_index(float c, float lo, float dx, float inv, int n, int& v) {
  const float t = __fsub_rn(c, lo);
  // one multiply for every mode and ONE rarely taken branch (the three-way branch on the launch-uniform `inv` cost a
  // fifth of the staging instructions of the kernels that call this three times per point)
  float q = __fmul_rn(t, fabsf(inv));
  const bool settled = inv > 0.f || (inv < 0.f && fabsf(q - rintf(q)) > fabsf(q) * 4.76837158203125e-7f);   // 2^-21
  if (!settled) q = __fdiv_rn(t, dx);   // inside the guard band, NaN, or inv == 0
  // .long() truncates toward zero, so (-1, 0) lands in voxel 0 and is KEPT; trunc(q) in [0, n) <=> -1 < q < n
  // (n < 2^24 is exact in fp32). NaN / inf fail both comparisons (the CPU's INT64_MIN is dropped too).
  const bool in = q > -1.0f && q < (float)n;
  v = (int)(in ? q : 0.f);
  return in;
}

// The three axes of one point at once: ONE rarely taken branch for the guard-band / divide cases of all axes, everything
// else straight-line (three calls of voxel_index cost three reconvergence scopes per point). Same values as voxel_index.
__device__ __forceinline__ bool voxel_index3(float x, float y, float z, const float (&lo)[3], const float (&dx)[3],
                                             const float (&inv)[3], int nx, int ny, int nz, int& vx, int& vy, int& vz) {
  constexpr float kBand = 4.76837158203125e-7f;   // 2^-21
  const float tx = __fsub_rn(x, lo[0]), ty = __fsub_rn(y, lo[1]), tz = __fsub_rn(z, lo[2]);
  float qx = __fmul_rn(tx, fabsf(inv[0])), qy = __fmul_rn(ty, fabsf(inv[1])), qz = __fmul_rn(tz, fabsf(inv[2]));
  const bool sx = inv[0] > 0.f || (inv[0] < 0.f && fabsf(qx - rintf(qx)) > fabsf(qx) * kBand);
  const bool sy = inv[1] > 0.f || (inv[1] < 0.f && fabsf(qy - rintf(qy)) > fabsf(qy) * kBand);
  const bool sz = inv[2] > 0.f || (inv[2] < 0.f && fabsf(qz - rintf(qz)) > fabsf(qz) * kBand);
  if (!(sx && sy && sz)) {   // inside a guard band, NaN, or an axis without a usable reciprocal
    if (!sx) qx = __fdiv_rn(tx, dx[0]);
    if (!sy) qy = __fdiv_rn(ty, dx[1]);
    if (!sz) qz = __fdiv_rn(tz, dx[2]);
  }
  const bool in = qx > -1.0f && qx < (float)nx && qy > -1.0f && qy < (float)ny && qz > -1.0f && qz < (float)nz;
  vx = (int)(in ? qx : 0.f);
  vy = (int)(in ? qy : 0.f);
  vz = (int)(in ? qz : 0.f);
  return in;
}

// see voxel_index: 1/dx if dx is a positive power of two (exact), -fl(1/dx) for any other positive finite dx, else 0
inline float exact_reciprocal_or_zero(float dx) {
  int e = 0;
  if (!(dx > 0.f) || !isfinite(dx)) return 0.f;
  if (frexpf(dx, &e) == 0.5f) return 1.0f / dx;
  const float r = 1.0f / dx;
  return (isfinite(r) && r > 0.f && getenv("BEVPOOL_EXACT_DIVIDE") == nullptr) ? -r : 0.f;
}

}  // namespace bevpool
