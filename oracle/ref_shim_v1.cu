// ORACLE — TEST INFRASTRUCTURE ONLY. extern "C" doorway onto the reference's UNMODIFIED v1 translation unit
// (-DREF_CU_V1='"/root/reference/projects/mmdet3d_plugin/ops/bev_pool/src/bev_pool_cuda.cu"'), compiled from
// where it lies. Launchers bev_pool() / bev_pool_grad() (bev_pool_cuda.cu:86-98) use the legacy default stream.
#include <cuda_runtime.h>
#include <math.h>
#include REF_CU_V1

extern "C" int ref_bev_pool_v1_fwd(int b, int d, int h, int w, int n, int c, int n_intervals, const float* x,
                                   const int* geom, const int* starts, const int* lengths, float* out) {
  bev_pool(b, d, h, w, n, c, n_intervals, x, geom, starts, lengths, out);
  return (int)cudaGetLastError();
}
extern "C" int ref_bev_pool_v1_bwd(int b, int d, int h, int w, int n, int c, int n_intervals, const float* out_grad,
                                   const int* geom, const int* starts, const int* lengths, float* x_grad) {
  bev_pool_grad(b, d, h, w, n, c, n_intervals, out_grad, geom, starts, lengths, x_grad);
  return (int)cudaGetLastError();
}
