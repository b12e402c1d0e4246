// "Lift" head of the view transform: the depth softmax + context split that feed
// bev_pool_v2 (CamEncode.get_depth_feat, cam_stream_lss_bevpoolv2.py:134-141), fused with the NCHW->NHWC
// transpose of the context features the pool wants (:282). One pass over the depthnet output:
//
//   x [BN, D+C, H, W]  ->  depth [BN, D, H, W] = softmax over the first D channels
//                          feat  [BN, H, W, C] = channels D..D+C-1, channels-last (or [BN, C, H, W] untouched layout)
//
// and its backward (softmax Jacobian + inverse transpose) in one pass as well. Pure streaming kernels: a CTA owns
// 32 consecutive pixels of one image; every global access is a 128-byte row (depth side) or a contiguous C-row
// (feature side). Bound: HBM, 2*(D+C)*e bytes per pixel.
#include "common.cuh"

namespace bevpool {

constexpr int kLiftPix = 32;
constexpr int kLiftWarps = 8;
constexpr int kLiftThreads = kLiftWarps * 32;
constexpr int kLiftMaxPerThread = 32;   // D <= 8 * 32 = 256 depth bins
constexpr int kLiftPre = 8;             // context values per thread prefetched ahead of the softmax barriers

// NV = depth bins per thread (D <= 8*NV). All global loads of a CTA (depth logits and the first 64 context channels)
// are issued before the first barrier so a single resident wave keeps enough bytes in flight.
template <typename T, int NV>
__global__ void __launch_bounds__(kLiftThreads)
lift_fwd_kernel(const T* __restrict__ x, T* __restrict__ depth, T* __restrict__ feat, int d, int c, int hw,
                int feat_channels_last, int64_t blocks_per_img) {
  extern __shared__ float s_tile[];             // [c][33] feature transpose tile
  __shared__ float s_red[kLiftWarps][kLiftPix];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t img = blockIdx.x / blocks_per_img;
  const int p0 = (int)(blockIdx.x % blocks_per_img) * kLiftPix;
  const int pix = p0 + lane;
  const bool in = pix < hw;
  const T* xi = x + img * (int64_t)(d + c) * hw;
  const T* xf = xi + (int64_t)d * hw;
  // ---- softmax over D: warp w owns bins w, w+8, ...; values stay in registers
  float v[NV], f[kLiftPre];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int dd = warp + i * kLiftWarps;
    v[i] = (in && dd < d) ? Vec4<T>::load1(xi, (int64_t)dd * hw + pix) : -INFINITY;
  }
#pragma unroll
  for (int i = 0; i < kLiftPre; ++i) {
    const int cc = warp + i * kLiftWarps;
    f[i] = (in && cc < c) ? Vec4<T>::load1(xf, (int64_t)cc * hw + pix) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) m = fmaxf(m, v[i]);
  s_red[warp][lane] = m;
  __syncthreads();
  float gm = s_red[0][lane];
#pragma unroll
  for (int w = 1; w < kLiftWarps; ++w) gm = fmaxf(gm, s_red[w][lane]);
  if (!in) gm = 0.f;
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = expf(v[i] - gm);                     // padding bins: exp(-inf) = 0
    sum += v[i];
  }
  s_red[warp][lane] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kLiftWarps; ++w) tot += s_red[w][lane];
  const float inv = 1.f / tot;
  T* di = depth + img * (int64_t)d * hw;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int dd = warp + i * kLiftWarps;
    if (in && dd < d) Vec4<T>::store1(di, (int64_t)dd * hw + pix, v[i] * inv);
  }
  // ---- context features
  if (!feat_channels_last) {
    T* fo = feat + img * (int64_t)c * hw;
#pragma unroll
    for (int i = 0; i < kLiftPre; ++i) {
      const int cc = warp + i * kLiftWarps;
      if (in && cc < c) Vec4<T>::store1(fo, (int64_t)cc * hw + pix, f[i]);
    }
    for (int cc = warp + kLiftPre * kLiftWarps; cc < c; cc += kLiftWarps)
      if (in) Vec4<T>::store1(fo, (int64_t)cc * hw + pix, Vec4<T>::load1(xf, (int64_t)cc * hw + pix));
    return;
  }
#pragma unroll
  for (int i = 0; i < kLiftPre; ++i) {
    const int cc = warp + i * kLiftWarps;
    if (cc < c) s_tile[cc * (kLiftPix + 1) + lane] = f[i];
  }
  for (int cc = warp + kLiftPre * kLiftWarps; cc < c; cc += kLiftWarps)
    s_tile[cc * (kLiftPix + 1) + lane] = in ? Vec4<T>::load1(xf, (int64_t)cc * hw + pix) : 0.f;
  __syncthreads();
  T* fo = feat + (img * hw + p0) * (int64_t)c;
  const int npix = min(kLiftPix, hw - p0);
  int px = threadIdx.x / c, cc = threadIdx.x - px * c;       // i = threadIdx.x + k*256 walked without divisions
  const int dpx = kLiftThreads / c, dcc = kLiftThreads - dpx * c;
  for (int i = threadIdx.x; i < npix * c; i += kLiftThreads) {
    Vec4<T>::store1(fo, i, s_tile[cc * (kLiftPix + 1) + px]);
    px += dpx;
    cc += dcc;
    if (cc >= c) {
      cc -= c;
      ++px;
    }
  }
}

// dx[:, :D] = depth * (g - sum_d g*depth);  dx[:, D:] = feat_grad (transposed back when channels-last)
template <typename T, int NV>
__global__ void __launch_bounds__(kLiftThreads)
lift_bwd_kernel(const T* __restrict__ depth, const T* __restrict__ depth_grad, const T* __restrict__ feat_grad,
                T* __restrict__ dx, int d, int c, int hw, int feat_channels_last, int64_t blocks_per_img) {
  extern __shared__ float s_tile[];
  __shared__ float s_red[kLiftWarps][kLiftPix];
  pdl_wait();
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t img = blockIdx.x / blocks_per_img;
  const int p0 = (int)(blockIdx.x % blocks_per_img) * kLiftPix;
  const int pix = p0 + lane;
  const bool in = pix < hw;
  const int npix = min(kLiftPix, hw - p0);
  const T* di = depth + img * (int64_t)d * hw;
  const T* gi = depth_grad + img * (int64_t)d * hw;
  T* xo = dx + img * (int64_t)(d + c) * hw;
  const T* fi = feat_channels_last ? feat_grad + (img * hw + p0) * (int64_t)c : feat_grad + img * (int64_t)c * hw;
  float y[NV], g[NV], f[kLiftPre];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int dd = warp + i * kLiftWarps;
    const bool ok = in && dd < d;
    y[i] = ok ? Vec4<T>::load1(di, (int64_t)dd * hw + pix) : 0.f;
    g[i] = ok ? Vec4<T>::load1(gi, (int64_t)dd * hw + pix) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < kLiftPre; ++i) {
    if (feat_channels_last) {
      const int e = threadIdx.x + i * kLiftThreads;
      f[i] = e < npix * c ? Vec4<T>::load1(fi, e) : 0.f;
    } else {
      const int cc = warp + i * kLiftWarps;
      f[i] = (in && cc < c) ? Vec4<T>::load1(fi, (int64_t)cc * hw + pix) : 0.f;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) dot = fmaf(y[i], g[i], dot);
  s_red[warp][lane] = dot;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kLiftWarps; ++w) tot += s_red[w][lane];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int dd = warp + i * kLiftWarps;
    if (in && dd < d) Vec4<T>::store1(xo, (int64_t)dd * hw + pix, y[i] * (g[i] - tot));
  }
  T* xf = xo + (int64_t)d * hw;
  if (!feat_channels_last) {
#pragma unroll
    for (int i = 0; i < kLiftPre; ++i) {
      const int cc = warp + i * kLiftWarps;
      if (in && cc < c) Vec4<T>::store1(xf, (int64_t)cc * hw + pix, f[i]);
    }
    for (int cc = warp + kLiftPre * kLiftWarps; cc < c; cc += kLiftWarps)
      if (in) Vec4<T>::store1(xf, (int64_t)cc * hw + pix, Vec4<T>::load1(fi, (int64_t)cc * hw + pix));
    return;
  }
  int px = threadIdx.x / c, cc = threadIdx.x - px * c;
  const int dpx = kLiftThreads / c, dcc = kLiftThreads - dpx * c;
  for (int i = threadIdx.x, k = 0; i < npix * c; i += kLiftThreads, ++k) {
    float val = 0.f;
    if (k < kLiftPre) {
#pragma unroll
      for (int j = 0; j < kLiftPre; ++j)
        if (j == k) val = f[j];
    } else {
      val = Vec4<T>::load1(fi, i);
    }
    s_tile[cc * (kLiftPix + 1) + px] = val;
    px += dpx;
    cc += dcc;
    if (cc >= c) {
      cc -= c;
      ++px;
    }
  }
  __syncthreads();
  for (int cc2 = warp; cc2 < c; cc2 += kLiftWarps)
    if (in) Vec4<T>::store1(xf, (int64_t)cc2 * hw + pix, s_tile[cc2 * (kLiftPix + 1) + lane]);
}

template <typename T, int NV>
static void lift_launch_nv(bool fwd, const void* a, const void* b, const void* cc_, void* o1, void* o2, int d, int c, int hw,
                           int cl, int64_t bpi, int64_t blocks, size_t smem, cudaStream_t st) {
  if (fwd) {
    if (ensure_dynamic_smem(lift_fwd_kernel<T, NV>, smem)) return;   // the failed launch is reported by launch_status()
    launch_pdl(lift_fwd_kernel<T, NV>, dim3((unsigned)blocks), dim3(kLiftThreads), smem, st, (const T*)a, (T*)o1, (T*)o2, d,
               c, hw, cl, bpi);
  } else {
    if (ensure_dynamic_smem(lift_bwd_kernel<T, NV>, smem)) return;
    launch_pdl(lift_bwd_kernel<T, NV>, dim3((unsigned)blocks), dim3(kLiftThreads), smem, st, (const T*)a, (const T*)b,
               (const T*)cc_, (T*)o1, d, c, hw, cl, bpi);
  }
}

template <typename T>
static int lift_launch(bool fwd, const void* a, const void* b, const void* cc_, void* o1, void* o2, int bn, int d, int c,
                       int hw, int cl, cudaStream_t st) {
  const int64_t bpi = (hw + kLiftPix - 1) / kLiftPix;
  const int64_t blocks = bpi * bn;
  if (blocks == 0) return 0;
  if (blocks > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const size_t smem = cl ? sizeof(float) * (size_t)c * (kLiftPix + 1) : 0;
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_CHANNELS;
  if (d <= 8 * kLiftWarps) lift_launch_nv<T, 8>(fwd, a, b, cc_, o1, o2, d, c, hw, cl, bpi, blocks, smem, st);
  else if (d <= 16 * kLiftWarps) lift_launch_nv<T, 16>(fwd, a, b, cc_, o1, o2, d, c, hw, cl, bpi, blocks, smem, st);
  else lift_launch_nv<T, 32>(fwd, a, b, cc_, o1, o2, d, c, hw, cl, bpi, blocks, smem, st);
  count_launch();
  return launch_status();
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_lift_forward(const void* x, void* depth, void* feat, int bn, int d, int c, int hw,
                                    int feat_channels_last, int dtype, void* stream) {
  if (bn < 0 || d <= 0 || c < 0 || hw < 0 || d > kLiftWarps * kLiftMaxPerThread) return BEVPOOL_ERR_BAD_ARG;
  if ((int64_t)bn * hw == 0) return BEVPOOL_OK;
  if (!x || !depth || (c > 0 && !feat)) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32) return lift_launch<float>(true, x, nullptr, nullptr, depth, feat, bn, d, c, hw, feat_channels_last, st);
  if (dtype == BEVPOOL_BF16)
    return lift_launch<__nv_bfloat16>(true, x, nullptr, nullptr, depth, feat, bn, d, c, hw, feat_channels_last, st);
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" int bevpool_lift_backward(const void* depth, const void* depth_grad, const void* feat_grad, void* x_grad,
                                     int bn, int d, int c, int hw, int feat_channels_last, int dtype, void* stream) {
  if (bn < 0 || d <= 0 || c < 0 || hw < 0 || d > kLiftWarps * kLiftMaxPerThread) return BEVPOOL_ERR_BAD_ARG;
  if ((int64_t)bn * hw == 0) return BEVPOOL_OK;
  if (!depth || !depth_grad || (c > 0 && !feat_grad) || !x_grad) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32)
    return lift_launch<float>(false, depth, depth_grad, feat_grad, x_grad, nullptr, bn, d, c, hw, feat_channels_last, st);
  if (dtype == BEVPOOL_BF16)
    return lift_launch<__nv_bfloat16>(false, depth, depth_grad, feat_grad, x_grad, nullptr, bn, d, c, hw, feat_channels_last,
                                      st);
  return BEVPOOL_ERR_BAD_ARG;
}
