// Radar / lidar pillar scatter: the reference's `pts_middle_encoder`
// (rcfusion/detectors/rcfusion_faster_rcnn.py:100, config RCFusion_NewScenes/rcfusion_lss.py:63-64) is mmdet3d's
// PointPillarsScatter (mmdet3d v0.17.1, README.md:153-156; not vendored in the reference tree):
//   canvas[b, :, y*nx + x] = voxel_features[i, :]   for every pillar i with coors[i] = (b, z, y, x)
// on a zero canvas [B, C, ny, nx]. Duplicated cells: the LAST pillar wins (what the sequential CPU indexing does).
//
//   memset(pillar_index, -1)        int32 [B][ny*nx]
//   pillar_index_kernel             atomicMax(pillar_index[b][y*nx+x], i)   (max index = last writer, deterministic)
//   pillar_canvas_kernel            ONE dense pass over the canvas: 64 cells x all channels per CTA, pillar rows
//                                   gathered as contiguous C-rows, zeros for empty cells, NCHW written coalesced
//                                   (no memset of the 39 MB/sample canvas, no scattered 4-byte stores)
//   pillar_grad_kernel              backward: voxel_grad[i, :] = canvas_grad[b, :, y, x]   (every duplicate gets the
//                                   cell's gradient, as autograd of index_put_ does)
// Pillars whose (b, y, x) is outside the canvas are ignored (PyTorch would raise).
#include "common.cuh"

namespace bevpool {

__global__ void __launch_bounds__(256)
pillar_index_kernel(const int* __restrict__ coors, int n_pillars, int b, int ny, int nx, int* __restrict__ pillar_index) {
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_pillars) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coors) + i);   // (batch, z, y, x)
  if (c.x < 0 || c.x >= b || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  atomicMax(pillar_index + ((int64_t)c.x * ny + c.z) * nx + c.w, i);
}

constexpr int kPcCols = 64;
template <typename T>
__global__ void __launch_bounds__(256)
pillar_canvas_kernel(const T* __restrict__ feats, const int* __restrict__ pillar_index, T* __restrict__ canvas, int c,
                     int64_t cells, int64_t tiles_per_sample) {
  extern __shared__ float t[];            // [c][kPcCols + 1]
  __shared__ int s_idx[kPcCols];
  pdl_wait();
  const int64_t bi = blockIdx.x / tiles_per_sample;
  const int64_t c0 = (blockIdx.x % tiles_per_sample) * kPcCols;
  const int ncol = (int)min((int64_t)kPcCols, cells - c0);
  if (threadIdx.x < kPcCols) s_idx[threadIdx.x] = threadIdx.x < ncol ? __ldg(pillar_index + bi * cells + c0 + threadIdx.x) : -1;
  __syncthreads();
  const bool vec_in = (c & 3) == 0 && (((uintptr_t)feats) & 15) == 0;
  if (vec_in) {   // pillar rows as 128-bit pieces (a row is c*e contiguous bytes)
    const int c4 = c >> 2;
    // four row pieces per thread are requested before any is stored (a load and its dependent shared-memory stores per
    // iteration would cost one memory round trip each: 64 cells x 64 channels are 4 iterations per thread)
    for (int i0 = threadIdx.x; i0 < ncol * c4; i0 += 4 * 256) {
      float4 v[4];
      int col[4], q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        col[u] = i / c4;
        q[u] = i - col[u] * c4;
        const int p = i < ncol * c4 ? s_idx[col[u]] : -1;
        v[u] = p >= 0 ? Vec4<T>::load_stream(feats, (int64_t)p * c + 4 * q[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * 256 < ncol * c4) {
          float* o = t + (4 * q[u]) * (kPcCols + 1) + col[u];
          o[0] = v[u].x; o[kPcCols + 1] = v[u].y; o[2 * (kPcCols + 1)] = v[u].z; o[3 * (kPcCols + 1)] = v[u].w;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < ncol * c; i += 256) {
      const int col = i / c, ch = i - col * c;
      const int p = s_idx[col];
      t[ch * (kPcCols + 1) + col] = p >= 0 ? Vec4<T>::load1(feats, (int64_t)p * c + ch) : 0.f;
    }
  }
  __syncthreads();
  T* d = canvas + (bi * c) * cells + c0;
  const bool vec_out = ncol == kPcCols && (cells & 3) == 0 && (((uintptr_t)canvas) & 15) == 0;
  if (vec_out) {   // 4 consecutive cells of one channel per store: 128-bit (fp32) / 64-bit (bf16)
    for (int i = threadIdx.x; i < c * (kPcCols / 4); i += 256) {
      const int ch = i / (kPcCols / 4), q = i % (kPcCols / 4);
      const float* pp = t + ch * (kPcCols + 1) + 4 * q;
      Vec4<T>::store(d, (int64_t)ch * cells + 4 * q, make_float4(pp[0], pp[1], pp[2], pp[3]));
    }
  } else {
    for (int i = threadIdx.x; i < c * kPcCols; i += 256) {
      const int ch = i / kPcCols, q = i % kPcCols;
      if (q < ncol) Vec4<T>::store1s(d, (int64_t)ch * cells + q, t[ch * (kPcCols + 1) + q]);
    }
  }
}

// one warp per pillar: lanes stride the channels (each a 4-byte gather at stride ny*nx), the row is written coalesced
template <typename T>
__global__ void __launch_bounds__(256)
pillar_grad_kernel(const T* __restrict__ canvas_grad, const int* __restrict__ coors, T* __restrict__ voxel_grad,
                   int n_pillars, int c, int b, int ny, int nx) {
  pdl_wait();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_pillars) return;
  const int lane = lane_id();
  const int4 co = __ldg(reinterpret_cast<const int4*>(coors) + i);
  const bool in = !(co.x < 0 || co.x >= b || co.z < 0 || co.z >= ny || co.w < 0 || co.w >= nx);
  const int64_t cells = (int64_t)ny * nx;
  const T* g = canvas_grad + (int64_t)co.x * c * cells + (int64_t)co.z * nx + co.w;
  for (int ch = lane; ch < c; ch += 32)
    Vec4<T>::store1(voxel_grad, (int64_t)i * c + ch, in ? Vec4<T>::load1(g, (int64_t)ch * cells) : 0.f);
}

template <typename T>
static int pillar_forward_t(const void* feats, const int* coors, void* canvas, int* pillar_index, int n_pillars, int c, int b,
                            int ny, int nx, cudaStream_t st) {
  const int64_t cells = (int64_t)ny * nx;
  cudaMemsetAsync(pillar_index, 0xff, sizeof(int) * (size_t)b * cells, st);
  if (n_pillars > 0) {
    launch_pdl(pillar_index_kernel, dim3((n_pillars + 255) / 256), dim3(256), 0, st, coors, n_pillars, b, ny, nx, pillar_index);
    count_launch();
  }
  const int64_t tiles = (cells + kPcCols - 1) / kPcCols;
  const int64_t total = tiles * b;
  if (total > INT32_MAX) return BEVPOOL_ERR_OVERFLOW;
  const size_t smem = sizeof(float) * (size_t)c * (kPcCols + 1);
  if (smem > 200 * 1024) return BEVPOOL_ERR_BAD_CHANNELS;
  if (int rc = ensure_dynamic_smem(pillar_canvas_kernel<T>, smem)) return rc;
  launch_pdl(pillar_canvas_kernel<T>, dim3((unsigned)total), dim3(256), smem, st, (const T*)feats, (const int*)pillar_index,
             (T*)canvas, c, cells, tiles);
  count_launch();
  return launch_status();
}

}  // namespace bevpool

using namespace bevpool;

extern "C" int bevpool_pillar_scatter_forward(const void* voxel_features, const int32_t* coors, void* canvas,
                                              int32_t* pillar_index, int n_pillars, int c, int b, int ny, int nx, int dtype,
                                              void* stream) {
  if (n_pillars < 0 || c <= 0 || b < 0 || ny < 0 || nx < 0) return BEVPOOL_ERR_BAD_ARG;
  if ((int64_t)b * ny * nx == 0) return BEVPOOL_OK;
  if (!canvas || !pillar_index || (n_pillars > 0 && (!voxel_features || !coors))) return BEVPOOL_ERR_BAD_ARG;
  if ((uintptr_t)coors % 16) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BEVPOOL_F32) return pillar_forward_t<float>(voxel_features, coors, canvas, pillar_index, n_pillars, c, b, ny, nx, st);
  if (dtype == BEVPOOL_BF16)
    return pillar_forward_t<__nv_bfloat16>(voxel_features, coors, canvas, pillar_index, n_pillars, c, b, ny, nx, st);
  return BEVPOOL_ERR_BAD_ARG;
}

extern "C" int bevpool_pillar_scatter_backward(const void* canvas_grad, const int32_t* coors, void* voxel_grad, int n_pillars,
                                               int c, int b, int ny, int nx, int dtype, void* stream) {
  if (n_pillars < 0 || c <= 0 || b < 0 || ny < 0 || nx < 0) return BEVPOOL_ERR_BAD_ARG;
  if (n_pillars == 0) return BEVPOOL_OK;
  if (!canvas_grad || !coors || !voxel_grad || (uintptr_t)coors % 16) return BEVPOOL_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((n_pillars + 7) / 8);
  if (dtype == BEVPOOL_F32)
    launch_pdl(pillar_grad_kernel<float>, dim3(blocks), dim3(256), 0, st, (const float*)canvas_grad, (const int*)coors,
               (float*)voxel_grad, n_pillars, c, b, ny, nx);
  else if (dtype == BEVPOOL_BF16)
    launch_pdl(pillar_grad_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)canvas_grad,
               (const int*)coors, (__nv_bfloat16*)voxel_grad, n_pillars, c, b, ny, nx);
  else
    return BEVPOOL_ERR_BAD_ARG;
  count_launch();
  return launch_status();
}
