// Cross-modal fusion glue: the bandwidth-bound pieces of the reference's
// Cross_Modal_Fusion.forward (rcfusion/detectors/BEVCross_modal_attention.py:31-43) around its three convolutions:
//
//   channel_avg_max   x [B,C,H,W] -> [B,2,H,W] = cat(mean over C, max over C)      (:32-34 and :36-38)
//   gate_concat       cat([img_bev * radar_att, radar_bev * img_att], dim=1)       (:40-42)
//
// each with its backward. One thread owns one pixel (consecutive threads = consecutive pixels, so every access is
// coalesced) and walks the channels; nothing is staged, each input is read once and each output written once.
#include "common.cuh"

namespace bevpool {

// CTA = 32 consecutive pixels (lanes) x 8 warps; warp w takes channels w, w + 8, ... so a CTA keeps 8 x 128-byte row
// segments in flight per step (one thread per pixel walking all channels left 3/4 of the machine idle at 2 x 160 x 240
// pixels). Per-pixel results that need all channels are combined through shared memory in channel order.
constexpr int kFuWarps = 8;

template <typename T>
__global__ void __launch_bounds__(256)
channel_avg_max_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int* __restrict__ argmax, int c, int64_t hw,
                           int64_t total) {
  __shared__ float s_sum[kFuWarps][32], s_max[kFuWarps][32];
  __shared__ int s_arg[kFuWarps][32];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool in = i < total;
  const int64_t b = in ? i / hw : 0, p = in ? i - b * hw : 0;
  const T* xp = x + b * c * hw + p;
  float sum = 0.f, mx = 0.f;
  int am = -1;
#pragma unroll 8
  for (int ch = warp; ch < c; ch += kFuWarps) {
    const float v = in ? Vec4<T>::load1(xp, (int64_t)ch * hw) : 0.f;
    sum += v;
    // the first maximum wins; the first NaN wins and sticks (torch.max semantics)
    if (am < 0 || v > mx || (v != v && mx == mx)) {
      mx = v;
      am = ch;
    }
  }
  s_sum[warp][lane] = sum;
  s_max[warp][lane] = mx;
  s_arg[warp][lane] = am;
  __syncthreads();
  if (warp == 0 && in) {
    float tot = 0.f, best = 0.f;
    int arg = -1;
#pragma unroll
    for (int w = 0; w < kFuWarps; ++w) {
      tot += s_sum[w][lane];
      const float v = s_max[w][lane];
      const int a = s_arg[w][lane];
      if (a < 0) continue;
      const bool v_nan = v != v, b_nan = best != best;
      // larger value wins; equal values: the smaller channel index; NaN: the first NaN channel
      if (arg < 0 || (!b_nan && (v_nan || v > best)) || (((v_nan && b_nan) || v == best) && a < arg)) {
        best = v;
        arg = a;
      }
    }
    Vec4<T>::store1(out, b * 2 * hw + p, tot / (float)c);
    Vec4<T>::store1(out, b * 2 * hw + hw + p, best);
    argmax[i] = arg;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
channel_avg_max_bwd_kernel(const T* __restrict__ g, const int* __restrict__ argmax, T* __restrict__ dx, int c, int64_t hw,
                           int64_t total) {
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  if (i >= total) return;
  const int64_t b = i / hw, p = i - b * hw;
  const float ga = Vec4<T>::load1(g, b * 2 * hw + p) / (float)c;
  const float gm = Vec4<T>::load1(g, b * 2 * hw + hw + p);
  const int am = argmax[i];
  T* dp = dx + b * c * hw + p;
#pragma unroll 8
  for (int ch = warp; ch < c; ch += kFuWarps) Vec4<T>::store1(dp, (int64_t)ch * hw, ch == am ? ga + gm : ga);
}

template <typename T>
__global__ void __launch_bounds__(256)
gate_concat_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ att_for_a,
                       const T* __restrict__ att_for_b, T* __restrict__ out, int ca, int cb, int64_t hw, int64_t total) {
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  if (i >= total) return;
  const int64_t n = i / hw, p = i - n * hw;
  const float wa = Vec4<T>::load1(att_for_a, i), wb = Vec4<T>::load1(att_for_b, i);
  const T* ap = a + n * ca * hw + p;
  const T* bp = b + n * cb * hw + p;
  T* op = out + n * (ca + cb) * hw + p;
#pragma unroll 8
  for (int ch = warp; ch < ca; ch += kFuWarps) Vec4<T>::store1(op, (int64_t)ch * hw, Vec4<T>::load1(ap, (int64_t)ch * hw) * wa);
#pragma unroll 8
  for (int ch = warp; ch < cb; ch += kFuWarps)
    Vec4<T>::store1(op, (int64_t)(ca + ch) * hw, Vec4<T>::load1(bp, (int64_t)ch * hw) * wb);
}

template <typename T>
__global__ void __launch_bounds__(256)
gate_concat_bwd_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ b,
                       const T* __restrict__ att_for_a, const T* __restrict__ att_for_b, T* __restrict__ da,
                       T* __restrict__ db, T* __restrict__ datt_a, T* __restrict__ datt_b, int ca, int cb, int64_t hw,
                       int64_t total) {
  __shared__ float s_a[kFuWarps][32], s_b[kFuWarps][32];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool in = i < total;
  const int64_t n = in ? i / hw : 0, p = in ? i - n * hw : 0;
  float sa = 0.f, sb = 0.f;
  if (in) {
    const float wa = Vec4<T>::load1(att_for_a, i), wb = Vec4<T>::load1(att_for_b, i);
    const T* gp = g + n * (ca + cb) * hw + p;
    const T* ap = a + n * ca * hw + p;
    const T* bp = b + n * cb * hw + p;
#pragma unroll 8
    for (int ch = warp; ch < ca; ch += kFuWarps) {
      const float gv = Vec4<T>::load1(gp, (int64_t)ch * hw);
      sa = fmaf(gv, Vec4<T>::load1(ap, (int64_t)ch * hw), sa);
      Vec4<T>::store1(da, n * ca * hw + (int64_t)ch * hw + p, gv * wa);
    }
#pragma unroll 8
    for (int ch = warp; ch < cb; ch += kFuWarps) {
      const float gv = Vec4<T>::load1(gp, (int64_t)(ca + ch) * hw);
      sb = fmaf(gv, Vec4<T>::load1(bp, (int64_t)ch * hw), sb);
      Vec4<T>::store1(db, n * cb * hw + (int64_t)ch * hw + p, gv * wb);
    }
  }
  s_a[warp][lane] = sa;
  s_b[warp][lane] = sb;
  __syncthreads();
  if (warp == 0 && in) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < kFuWarps; ++w) {
      ta += s_a[w][lane];
      tb += s_b[w][lane];
    }
    Vec4<T>::store1(datt_a, i, ta);
    Vec4<T>::store1(datt_b, i, tb);
  }
}

static unsigned pixel_grid(int64_t total) { return (unsigned)((total + 31) / 32); }

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_channel_avg_max_forward(const void* x, void* out, int32_t* argmax, int b, int c, int64_t hw, int dtype,
                                               void* stream) {
  if (b < 0 || c <= 0 || hw < 0) return BEVPOOL_ERR_BAD_ARG;
  const int64_t total = (int64_t)b * hw;
  if (total == 0) return BEVPOOL_OK;
  if (!x || !out || !argmax) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    launch_pdl(channel_avg_max_fwd_kernel<float>, dim3(pixel_grid(total)), dim3(256), 0, st, (const float*)x, (float*)out,
               (int*)argmax, c, hw, total);
  else if (dtype == BEVPOOL_BF16)
    launch_pdl(channel_avg_max_fwd_kernel<__nv_bfloat16>, dim3(pixel_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)x,
               (__nv_bfloat16*)out, (int*)argmax, c, hw, total);
  else
    return BEVPOOL_ERR_BAD_ARG;
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
  if (dtype == BEVPOOL_F32)
    launch_pdl(channel_avg_max_bwd_kernel<float>, dim3(pixel_grid(total)), dim3(256), 0, st, (const float*)out_grad,
               (const int*)argmax, (float*)x_grad, c, hw, total);
  else if (dtype == BEVPOOL_BF16)
    launch_pdl(channel_avg_max_bwd_kernel<__nv_bfloat16>, dim3(pixel_grid(total)), dim3(256), 0, st,
               (const __nv_bfloat16*)out_grad, (const int*)argmax, (__nv_bfloat16*)x_grad, c, hw, total);
  else
    return BEVPOOL_ERR_BAD_ARG;
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
  if (dtype == BEVPOOL_F32)
    launch_pdl(gate_concat_fwd_kernel<float>, dim3(pixel_grid(total)), dim3(256), 0, st, (const float*)a, (const float*)b,
               (const float*)att_for_a, (const float*)att_for_b, (float*)out, ca, cb, hw, total);
  else if (dtype == BEVPOOL_BF16)
    launch_pdl(gate_concat_fwd_kernel<__nv_bfloat16>, dim3(pixel_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)a,
               (const __nv_bfloat16*)b, (const __nv_bfloat16*)att_for_a, (const __nv_bfloat16*)att_for_b, (__nv_bfloat16*)out,
               ca, cb, hw, total);
  else
    return BEVPOOL_ERR_BAD_ARG;
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
  if (dtype == BEVPOOL_F32)
    launch_pdl(gate_concat_bwd_kernel<float>, dim3(pixel_grid(total)), dim3(256), 0, st, (const float*)out_grad, (const float*)a,
               (const float*)b, (const float*)att_for_a, (const float*)att_for_b, (float*)a_grad, (float*)b_grad,
               (float*)att_for_a_grad, (float*)att_for_b_grad, ca, cb, hw, total);
  else if (dtype == BEVPOOL_BF16)
    launch_pdl(gate_concat_bwd_kernel<__nv_bfloat16>, dim3(pixel_grid(total)), dim3(256), 0, st, (const __nv_bfloat16*)out_grad,
               (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)att_for_a,
               (const __nv_bfloat16*)att_for_b, (__nv_bfloat16*)a_grad, (__nv_bfloat16*)b_grad, (__nv_bfloat16*)att_for_a_grad,
               (__nv_bfloat16*)att_for_b_grad, ca, cb, hw, total);
  else
    return BEVPOOL_ERR_BAD_ARG;
  count_launch();
  return launch_status();
}
