// Cross-modal fusion glue: the bandwidth-bound pieces of the reference's
// Cross_Modal_Fusion.forward (rcfusion/detectors/BEVCross_modal_attention.py:31-43) around its three convolutions:
//
//   channel_avg_max   x [B,C,H,W] -> [B,2,H,W] = cat(mean over C, max over C)      (:32-34 and :36-38)
//   gate_concat       cat([img_bev * radar_att, radar_bev * img_att], dim=1)       (:40-42)
//
// each with its backward. One thread owns one pixel (consecutive threads = consecutive pixels, so every access is
// coalesced) and walks the channels; nothing is staged, each input is read once and each output written once.
#include <initializer_list>

#include "common.cuh"

namespace bevpool {

// CTA = 32 lanes x V consecutive pixels each x 8 warps; warp w takes channels w, w + 8, ... V = 4 (hw % 4 == 0 and
// 16-byte aligned tensors, i.e. every real BEV grid): 128-bit accesses, a CTA keeps 8 x 512-byte row segments in flight
// per step; V = 1 is the any-shape fallback. Per-pixel results that need all channels are combined through shared memory
// in channel order. (The first version, 4-byte accesses only, ran at 1.4 TB/s on the RCFusion size — torch parity.)
constexpr int kFuWarps = 8;

template <typename T, int V>
struct PixIO {   // V consecutive pixels of one channel plane
  static __device__ __forceinline__ void load(const T* p, float (&v)[V]) {
    if (V == 4) {
      const float4 f = Vec4<T>::load_stream(p, 0);
      v[0] = f.x; v[V > 1 ? 1 : 0] = f.y; v[V > 2 ? 2 : 0] = f.z; v[V > 3 ? 3 : 0] = f.w;
    } else {
      v[0] = Vec4<T>::load1(p, 0);
    }
  }
  static __device__ __forceinline__ void store(T* p, const float (&v)[V]) {
    if (V == 4) Vec4<T>::store(p, 0, make_float4(v[0], v[V > 1 ? 1 : 0], v[V > 2 ? 2 : 0], v[V > 3 ? 3 : 0]));
    else Vec4<T>::store1(p, 0, v[0]);
  }
};

template <typename T, int V>
__global__ void __launch_bounds__(256)
channel_avg_max_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int* __restrict__ argmax, int c, int64_t hw,
                           int64_t total) {
  __shared__ float s_sum[kFuWarps][32 * V], s_max[kFuWarps][32 * V];
  __shared__ int s_arg[kFuWarps][32 * V];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * V;
  const bool in = i < total;
  const int64_t b = in ? i / hw : 0, p = in ? i - b * hw : 0;
  const T* xp = x + b * c * hw + p;
  float sum[V], mx[V];
  int am[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sum[k] = 0.f;
    mx[k] = 0.f;
    am[k] = -1;
  }
  // eight channels per warp are requested before any is consumed: the max / argmax update is data-dependent control
  // flow, and loads do not move across it (one memory round trip per channel otherwise)
  constexpr int kBatch = 8;
  for (int ch0 = warp; ch0 < c; ch0 += kBatch * kFuWarps) {
    float v[kBatch][V];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int ch = ch0 + u * kFuWarps;
#pragma unroll
      for (int k = 0; k < V; ++k) v[u][k] = 0.f;
      if (in && ch < c) PixIO<T, V>::load(xp + (int64_t)ch * hw, v[u]);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int ch = ch0 + u * kFuWarps;
      if (ch >= c) break;   // warp-uniform
#pragma unroll
      for (int k = 0; k < V; ++k) {
        sum[k] += v[u][k];
        // the first maximum wins; the first NaN wins and sticks (torch.max semantics)
        if (am[k] < 0 || v[u][k] > mx[k] || (v[u][k] != v[u][k] && mx[k] == mx[k])) {
          mx[k] = v[u][k];
          am[k] = ch;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    s_sum[warp][lane * V + k] = sum[k];
    s_max[warp][lane * V + k] = mx[k];
    s_arg[warp][lane * V + k] = am[k];
  }
  __syncthreads();
  if (warp == 0 && in) {
    float mean[V], best[V];
    int arg[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float tot = 0.f;
      best[k] = 0.f;
      arg[k] = -1;
#pragma unroll
      for (int w = 0; w < kFuWarps; ++w) {
        tot += s_sum[w][lane * V + k];
        const float v = s_max[w][lane * V + k];
        const int a = s_arg[w][lane * V + k];
        if (a < 0) continue;
        const bool v_nan = v != v, b_nan = best[k] != best[k];
        // larger value wins; equal values: the smaller channel index; NaN: the first NaN channel
        if (arg[k] < 0 || (!b_nan && (v_nan || v > best[k])) || (((v_nan && b_nan) || v == best[k]) && a < arg[k])) {
          best[k] = v;
          arg[k] = a;
        }
      }
      mean[k] = tot / (float)c;
      argmax[i + k] = arg[k];
    }
    PixIO<T, V>::store(out + b * 2 * hw + p, mean);
    PixIO<T, V>::store(out + b * 2 * hw + hw + p, best);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
channel_avg_max_bwd_kernel(const T* __restrict__ g, const int* __restrict__ argmax, T* __restrict__ dx, int c, int64_t hw,
                           int64_t total) {
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * V;
  if (i >= total) return;
  const int64_t b = i / hw, p = i - b * hw;
  float ga[V], gm[V];
  int am[V];
  PixIO<T, V>::load(g + b * 2 * hw + p, ga);
  PixIO<T, V>::load(g + b * 2 * hw + hw + p, gm);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ga[k] /= (float)c;
    am[k] = argmax[i + k];
  }
  T* dp = dx + b * c * hw + p;
#pragma unroll 4
  for (int ch = warp; ch < c; ch += kFuWarps) {
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = ch == am[k] ? ga[k] + gm[k] : ga[k];
    PixIO<T, V>::store(dp + (int64_t)ch * hw, v);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
gate_concat_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ att_for_a,
                       const T* __restrict__ att_for_b, T* __restrict__ out, int ca, int cb, int64_t hw, int64_t total) {
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * V;
  if (i >= total) return;
  const int64_t n = i / hw, p = i - n * hw;
  float wa[V], wb[V];
  PixIO<T, V>::load(att_for_a + i, wa);
  PixIO<T, V>::load(att_for_b + i, wb);
  const T* ap = a + n * ca * hw + p;
  const T* bp = b + n * cb * hw + p;
  T* op = out + n * (ca + cb) * hw + p;
#pragma unroll 4
  for (int ch = warp; ch < ca + cb; ch += kFuWarps) {
    float v[V];
    const bool first = ch < ca;
    PixIO<T, V>::load(first ? ap + (int64_t)ch * hw : bp + (int64_t)(ch - ca) * hw, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] *= first ? wa[k] : wb[k];
    PixIO<T, V>::store(op + (int64_t)ch * hw, v);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
gate_concat_bwd_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ b,
                       const T* __restrict__ att_for_a, const T* __restrict__ att_for_b, T* __restrict__ da,
                       T* __restrict__ db, T* __restrict__ datt_a, T* __restrict__ datt_b, int ca, int cb, int64_t hw,
                       int64_t total) {
  __shared__ float s_a[kFuWarps][32 * V], s_b[kFuWarps][32 * V];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * V;
  const bool in = i < total;
  const int64_t n = in ? i / hw : 0, p = in ? i - n * hw : 0;
  float sa[V], sb[V];
#pragma unroll
  for (int k = 0; k < V; ++k) sa[k] = sb[k] = 0.f;
  if (in) {
    float wa[V], wb[V];
    PixIO<T, V>::load(att_for_a + i, wa);
    PixIO<T, V>::load(att_for_b + i, wb);
    const T* gp = g + n * (ca + cb) * hw + p;
    const T* ap = a + n * ca * hw + p;
    const T* bp = b + n * cb * hw + p;
#pragma unroll 4
    for (int ch = warp; ch < ca; ch += kFuWarps) {
      float gv[V], av[V];
      PixIO<T, V>::load(gp + (int64_t)ch * hw, gv);
      PixIO<T, V>::load(ap + (int64_t)ch * hw, av);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        sa[k] = fmaf(gv[k], av[k], sa[k]);
        gv[k] *= wa[k];
      }
      PixIO<T, V>::store(da + n * ca * hw + (int64_t)ch * hw + p, gv);
    }
#pragma unroll 4
    for (int ch = warp; ch < cb; ch += kFuWarps) {
      float gv[V], bv[V];
      PixIO<T, V>::load(gp + (int64_t)(ca + ch) * hw, gv);
      PixIO<T, V>::load(bp + (int64_t)ch * hw, bv);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        sb[k] = fmaf(gv[k], bv[k], sb[k]);
        gv[k] *= wb[k];
      }
      PixIO<T, V>::store(db + n * cb * hw + (int64_t)ch * hw + p, gv);
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    s_a[warp][lane * V + k] = sa[k];
    s_b[warp][lane * V + k] = sb[k];
  }
  __syncthreads();
  if (warp == 0 && in) {
    float ta[V], tb[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      ta[k] = tb[k] = 0.f;
#pragma unroll
      for (int w = 0; w < kFuWarps; ++w) {
        ta[k] += s_a[w][lane * V + k];
        tb[k] += s_b[w][lane * V + k];
      }
    }
    PixIO<T, V>::store(datt_a + i, ta);
    PixIO<T, V>::store(datt_b + i, tb);
  }
}

// V = 4 when every plane starts on a 16-byte boundary and holds a multiple of 4 pixels
static bool vec4_ok(int64_t hw, std::initializer_list<const void*> ptrs) {
  if (hw & 3) return false;
  for (const void* q : ptrs)
    if ((uintptr_t)q & 15) return false;
  return true;
}
static unsigned pixel_grid(int64_t total, int v) { return (unsigned)((total / v + 31) / 32); }

// launch kernel<T, V> for the runtime (dtype, vec) pair
#define BEVPOOL_FU_LAUNCH(KERN, VEC, TOTAL, ...)                                                                   \
  do {                                                                                                             \
    if (dtype == BEVPOOL_F32) {                                                                                    \
      using T = float;                                                                                             \
      if (VEC) launch_pdl(KERN<T, 4>, dim3(pixel_grid(TOTAL, 4)), dim3(256), 0, st, __VA_ARGS__);                  \
      else launch_pdl(KERN<T, 1>, dim3(pixel_grid(TOTAL, 1)), dim3(256), 0, st, __VA_ARGS__);                      \
    } else if (dtype == BEVPOOL_BF16) {                                                                            \
      using T = __nv_bfloat16;                                                                                     \
      if (VEC) launch_pdl(KERN<T, 4>, dim3(pixel_grid(TOTAL, 4)), dim3(256), 0, st, __VA_ARGS__);                  \
      else launch_pdl(KERN<T, 1>, dim3(pixel_grid(TOTAL, 1)), dim3(256), 0, st, __VA_ARGS__);                      \
    } else {                                                                                                       \
      return BEVPOOL_ERR_BAD_ARG;                                                                                  \
    }                                                                                                              \
  } while (0)

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_channel_avg_max_forward(const void* x, void* out, int32_t* argmax, int b, int c, int64_t hw, int dtype,
                                               void* stream) {
  if (b < 0 || c <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t total = (int64_t)b * hw;
  if (total == 0) return BEVPOOL_OK;
  if (!x || !out || !argmax) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec4_ok(hw, {x, out});
  BEVPOOL_FU_LAUNCH(channel_avg_max_fwd_kernel, vec, total, (const T*)x, (T*)out, (int*)argmax, c, hw, total);
  count_launch();
  return launch_status();
}

extern "C" int bevpool_channel_avg_max_backward(const void* out_grad, const int32_t* argmax, void* x_grad, int b, int c,
                                                int64_t hw, int dtype, void* stream) {
  if (b < 0 || c <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t total = (int64_t)b * hw;
  if (total == 0) return BEVPOOL_OK;
  if (!out_grad || !x_grad || !argmax) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec4_ok(hw, {out_grad, x_grad});
  BEVPOOL_FU_LAUNCH(channel_avg_max_bwd_kernel, vec, total, (const T*)out_grad, (const int*)argmax, (T*)x_grad, c, hw, total);
  count_launch();
  return launch_status();
}

extern "C" int bevpool_gate_concat_forward(const void* a, const void* b, const void* att_for_a, const void* att_for_b, void* out,
                                           int n, int ca, int cb, int64_t hw, int dtype, void* stream) {
  if (n < 0 || ca <= 0 || cb <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t total = (int64_t)n * hw;
  if (total == 0) return BEVPOOL_OK;
  if (!a || !b || !att_for_a || !att_for_b || !out) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec4_ok(hw, {a, b, att_for_a, att_for_b, out});
  BEVPOOL_FU_LAUNCH(gate_concat_fwd_kernel, vec, total, (const T*)a, (const T*)b, (const T*)att_for_a, (const T*)att_for_b,
                    (T*)out, ca, cb, hw, total);
  count_launch();
  return launch_status();
}

extern "C" int bevpool_gate_concat_backward(const void* out_grad, const void* a, const void* b, const void* att_for_a,
                                            const void* att_for_b, void* a_grad, void* b_grad, void* att_for_a_grad,
                                            void* att_for_b_grad, int n, int ca, int cb, int64_t hw, int dtype,
                                            void* stream) {
  if (n < 0 || ca <= 0 || cb <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t total = (int64_t)n * hw;
  if (total == 0) return BEVPOOL_OK;
  if (!out_grad || !a || !b || !att_for_a || !att_for_b || !a_grad || !b_grad || !att_for_a_grad || !att_for_b_grad)
    return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec4_ok(hw, {out_grad, a, b, att_for_a, att_for_b, a_grad, b_grad, att_for_a_grad, att_for_b_grad});
  BEVPOOL_FU_LAUNCH(gate_concat_bwd_kernel, vec, total, (const T*)out_grad, (const T*)a, (const T*)b, (const T*)att_for_a,
                    (const T*)att_for_b, (T*)a_grad, (T*)b_grad, (T*)att_for_a_grad, (T*)att_for_b_grad, ca, cb, hw, total);
  count_launch();
  return launch_status();
}
