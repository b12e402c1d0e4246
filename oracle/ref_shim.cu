// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/bevpool_oracle.c header).
//
// extern "C" doorway onto the reference's UNMODIFIED CUDA translation unit,
// which is compiled from where it lies under /root/reference (never copied):
//   -DREF_CU='"/root/reference/projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu"'
// The reference launchers bev_pool_v2() / bev_pool_v2_grad() (bev_pool_cuda.cu:125-140)
// take plain device pointers and launch on the legacy default stream.
// Output: oracle/_ref/libref_bevpool_v2.so (git-ignored, travels to the GPU box).
#include <cuda_runtime.h>
#include <math.h>
#include REF_CU

extern "C" int ref_bev_pool_v2_fwd(int c, int n_intervals, const float* depth, const float* feat,
                                   const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                                   const int* interval_starts, const int* interval_lengths, float* out) {
  bev_pool_v2(c, n_intervals, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
              interval_lengths, out);
  return (int)cudaGetLastError();
}

extern "C" int ref_bev_pool_v2_bwd(int c, int n_intervals, const float* out_grad, const float* depth,
                                   const float* feat, const int* ranks_depth, const int* ranks_feat,
                                   const int* ranks_bev, const int* interval_starts,
                                   const int* interval_lengths, float* depth_grad, float* feat_grad) {
  bev_pool_v2_grad(c, n_intervals, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                   interval_starts, interval_lengths, depth_grad, feat_grad);
  return (int)cudaGetLastError();
}
