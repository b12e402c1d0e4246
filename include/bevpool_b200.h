/*
 * bevpool_b200 — C ABI of the B200-native (sm_100a) camera->BEV view transform.
 *
 * This is the drop-in boundary. Every entry point takes plain device pointers and
 * sizes (no torch types), launches hand-written CUDA kernels asynchronously on the
 * stream the caller passes (a cudaStream_t cast to void*; NULL = legacy default
 * stream) and returns 0 or a negative bevpool_status / positive cudaError_t.
 * Nothing here synchronises the device (bevpool_prepare_v2_counts alone waits, for two integers) and nothing
 * falls back to the CPU.
 *
 * Reference interfaces replaced (paths relative to
 * /root/reference/projects/mmdet3d_plugin):
 *
 *   bevpool_v2_forward            ops/bev_pool_v2/src/bev_pool.cpp:30-57   bev_pool_v2_forward(...)
 *                                 -> src/bev_pool_cuda.cu:21-48,125-131    bev_pool_v2_kernel
 *   bevpool_v2_backward           ops/bev_pool_v2/src/bev_pool.cpp:74-104  bev_pool_v2_backward(...)
 *                                 -> src/bev_pool_cuda.cu:67-121,133-140   bev_pool_grad_kernel
 *   bevpool_v2_backward_regroup   ops/bev_pool_v2/bev_pool.py:47-57        argsort by ranks_feat + run-length
 *   bevpool_geometry              bevfusion/detectors/cam_stream_lss_bevpoolv2.py:244-251  get_geometry
 *   bevpool_prepare_v2            same file :294-351                        voxel_pooling_prepare_v2
 *   bevpool_prepare_v2_counts     same file :324-351                        ... with its exact-length outputs
 *   bevpool_voxel_table +         the fused forms of the above used by the view-transform shim:
 *   bevpool_v2_forward_dense /    bev_pool.py:27 (zeros) + :29 (kernel) + :91 (permute) in one pass;
 *   bevpool_v2_backward_dense     bev_pool.py:47-57,67-70 (argsort, zeros, kernel) in one pass
 *
 *   bevpool_lift_forward / _backward  bevfusion/detectors/cam_stream_lss_bevpoolv2.py:134-141 (+ :282)
 *   bevpool_channel_avg_max_*, bevpool_gate_concat_*  rcfusion/detectors/BEVCross_modal_attention.py:31-43
 *   bevpool_pillar_scatter_*          rcfusion/detectors/rcfusion_faster_rcnn.py:100 (mmdet3d PointPillarsScatter)
 *   bevpool_v1_forward / _backward ops/bev_pool/src/bev_pool.cpp:27-94, src/bev_pool_cuda.cu:20-84 (v1 op)
 *
 * Argument order note: like the reference's native entry points, interval_lengths
 * precedes interval_starts here (bev_pool.cpp:37-38), the opposite of the Python API.
 */
#ifndef BEVPOOL_B200_H_
#define BEVPOOL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEVPOOL_B200_ABI_VERSION 1

/* element type of depth / feat / out / grads */
enum { BEVPOOL_F32 = 0, BEVPOOL_BF16 = 1 };
/* memory layout of the BEV grid tensor */
enum {
  BEVPOOL_LAYOUT_BZYXC = 0, /* channels last: what QuickCumsumCuda.forward returns (bev_pool.py:27) */
  BEVPOOL_LAYOUT_BCZYX = 1  /* what bev_pool_v2() returns after its permute (bev_pool.py:91)      */
};
/* negative status codes (positive values are cudaError_t) */
enum {
  BEVPOOL_OK = 0,
  BEVPOOL_ERR_BAD_ARG = -1,      /* null pointer, negative size, unsupported dtype/layout      */
  BEVPOOL_ERR_BAD_CHANNELS = -2, /* c <= 0                                                     */
  BEVPOOL_ERR_WORKSPACE = -3,    /* workspace too small (see *_workspace_bytes)                */
  BEVPOOL_ERR_OVERFLOW = -4      /* a rank/index would not fit the int32 the API mandates      */
};

int bevpool_b200_abi_version(void);
/* human-readable text for a code returned by any function below (static storage) */
const char* bevpool_b200_strerror(int code);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t bevpool_b200_launch_count(void);

/* ------------------------------------------------------------------ pooling, reference contract
 * out[ranks_bev[s_k]*c + ch] = sum_i depth[ranks_depth[s_k+i]] * feat[ranks_feat[s_k+i]*c + ch]
 * `out` is [B,Z,Y,X,C] and PRE-ZEROED by the caller; only touched voxels are written.
 * The per-channel sum runs in sorted point order with fused multiply-add, i.e. it is
 * bit-identical to the reference kernel for fp32. */
int bevpool_v2_forward(const void* depth, const void* feat, void* out,
                       const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                       const int32_t* interval_lengths, const int32_t* interval_starts,
                       int64_t n_points, int64_t n_intervals, int c, int dtype, void* stream);

/* Rank arrays here are the ones REGROUPED by ranks_feat and the intervals are the
 * backward intervals (one per feature pixel), exactly what bev_pool.py:47-57 builds.
 * depth_grad / feat_grad are PRE-ZEROED by the caller. out_grad is [B,Z,Y,X,C]. */
int bevpool_v2_backward(const void* out_grad, void* depth_grad, void* feat_grad,
                        const void* depth, const void* feat,
                        const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                        const int32_t* interval_lengths, const int32_t* interval_starts,
                        int64_t n_points, int64_t n_intervals, int c, int dtype, void* stream);

/* Stable regrouping of the three rank arrays by ranks_feat and run-length segmentation.
 * Outputs have n_points (ranks) / n_points (intervals, worst case) capacity;
 * *n_intervals_bp_dev (device int32) receives the interval count. */
size_t bevpool_v2_backward_regroup_workspace_bytes(int64_t n_points);
int bevpool_v2_backward_regroup(const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                                int64_t n_points, int32_t max_ranks_feat,
                                int32_t* ranks_depth_bp, int32_t* ranks_feat_bp, int32_t* ranks_bev_bp,
                                int32_t* interval_starts_bp, int32_t* interval_lengths_bp,
                                int32_t* n_intervals_bp_dev,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ geometry
 * coor[b,n,d,h,w,:] = rots[b,n] . (u*d, v*d, d) + trans[b,n] with every multiply/add rounded
 * separately (no FMA contraction), matching the reference's CPU result bit for bit.
 * frustum is [D,H,W,3] fp32 (the reference's nn.Parameter), rots [BN,3,3], trans [BN,3]. */
int bevpool_geometry(const float* frustum, const float* rots, const float* trans, float* coor,
                     int bn, int d, int hw, void* stream);

/* ------------------------------------------------------------------ prepare
 * Grid description shared by the prepare / dense entry points. */
typedef struct {
  int32_t b, n, d, h, w;  /* frustum batch shape: B frames x N cameras x D bins x H x W            */
  int32_t nx[3];          /* X, Y, Z voxel counts                                                   */
  float lo[3];            /* fp32 (bx - dx/2), computed by the caller in fp32 as the reference does */
  float dx[3];            /* voxel size                                                             */
} bevpool_grid_t;

size_t bevpool_prepare_v2_workspace_bytes(const bevpool_grid_t* g);

/* voxel_pooling_prepare_v2. Exactly one of (coor) or (frustum, rots, trans) is given:
 *   coor != NULL  : reads the materialised [B,N,D,H,W,3] fp32 coordinates (the unchanged API);
 *   coor == NULL  : geometry is fused, coordinates are never written to memory.
 * Outputs (capacity P0 = b*n*d*h*w each for ranks_*, min(P0, B*V) for intervals):
 *   ranks_bev/ranks_depth/ranks_feat (sorted by ranks_bev, ties ascending ranks_depth),
 *   interval_starts/interval_lengths, counts_dev[0] = P (kept points), counts_dev[1] = I.
 *   point_rank (optional, may be NULL): int32[P0] voxel rank of every frustum point, -1 if
 *   dropped — the inverse table the sort-free backward uses.
 * No host synchronisation: the caller reads counts_dev when it needs exact lengths. */
int bevpool_prepare_v2(const float* coor, const float* frustum, const float* rots, const float* trans,
                       const bevpool_grid_t* g,
                       int32_t* ranks_bev, int32_t* ranks_depth, int32_t* ranks_feat,
                       int32_t* interval_starts, int32_t* interval_lengths,
                       int32_t* counts_dev, int32_t* point_rank,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same work, plus the two counts handed back to the HOST (host_counts[0] = P, host_counts[1] = I) — what the
 * reference API needs to return exact-length tensors (cam_stream_lss_bevpoolv2.py:324-351 reads them through
 * boolean-mask indexing and torch.where). P is final after the rank kernel and I equals the number of occupied
 * voxels (a byte map filled by the same kernel with plain stores), so a small kernel stores both into page-locked host memory BEFORE the sort
 * and the segmentation: the call returns as soon as that kernel is done, with the remaining kernels still running on
 * `stream`. This is the only entry point that waits for the device; it refuses a capturing stream (BEVPOOL_ERR_BAD_ARG). */
int bevpool_prepare_v2_counts(const float* coor, const float* frustum, const float* rots, const float* trans,
                              const bevpool_grid_t* g,
                              int32_t* ranks_bev, int32_t* ranks_depth, int32_t* ranks_feat,
                              int32_t* interval_starts, int32_t* interval_lengths,
                              int32_t* counts_dev, int32_t* point_rank,
                              void* workspace, size_t workspace_bytes, void* stream, int32_t* host_counts);

/* ------------------------------------------------------------------ fused ("dense") pooling
 * Lower-bound table over the sorted voxel ranks: vox_pt[v] = number of sorted points with rank < v,
 * for v in [0, n_voxels_total] (n_voxels_total + 1 entries); voxel v owns sorted points
 * [vox_pt[v], vox_pt[v+1]). It stands in for interval_starts / interval_lengths on the device.
 * n_points is the point count, or an upper bound with the true count read from counts_dev[0]. */
int bevpool_voxel_table(const int32_t* ranks_bev_sorted, int64_t n_points, const int32_t* counts_dev,
                        int64_t n_voxels_total, int32_t* vox_pt, void* stream);

/* Forward that also zero-fills empty voxels and writes either layout directly (`out` need NOT be
 * zeroed): new_zeros + kernel + permute of bev_pool.py:27,29,91 in one pass.
 * ranks_depth / ranks_bev are the sorted per-point lists; ranks_bev must be non-decreasing. ranks_feat may be NULL when it is derivable from ranks_depth, as for every output of
 * voxel_pooling_prepare_v2: rf = (rd / dhw) * hw + rd % hw  (dhw = D*H*W, hw = H*W).
 * The grid is n_frames x rows_per_frame (= Z*Y) x x voxels.
 * n_points: number of sorted points, or an upper bound with the true count in counts_dev[0].
 * scratch (bevpool_v2_forward_dense_scratch_bytes, 256-byte aligned): the pooled rows are produced
 * channels-last by a barrier-free streaming kernel (one warp per fixed-size chunk of the sorted point list,
 * voxels cut by a chunk border finished in chunk order by a fix-up kernel) and, for layout BCZYX, transposed
 * + zero-filled by a second, bandwidth-bound pass; with no / too little scratch, or c > 128, the single-pass
 * shared-memory tile kernel is used instead. */
size_t bevpool_v2_forward_dense_scratch_bytes(int64_t n_points, int64_t n_voxels, int c, int layout, int dtype);
int bevpool_v2_forward_dense(const void* depth, const void* feat, void* out,
                             const int32_t* ranks_depth, const int32_t* ranks_feat, const int32_t* ranks_bev,
                             const int32_t* vox_pt, int64_t n_points, const int32_t* counts_dev,
                             int c, int64_t n_frames, int64_t rows_per_frame, int x, int dhw, int hw,
                             int layout, int dtype, void* scratch, size_t scratch_bytes, void* stream);

/* Sort-free backward for rank arrays that came from bevpool_prepare_v2: walks the D depth
 * bins of every feature pixel through point_rank, writes EVERY element of depth_grad
 * ([BN,D,H,W], zeros for dropped points) and feat_grad — no memset, no argsort.
 * out_grad is [B,Z,Y,X,C] (use bevpool_grid_transpose for a [B,C,Z,Y,X] gradient).
 * feat is [BN,H,W,C]; feat_grad is [BN,H,W,C] (feat_grad_nchw == 0) or [BN,C,H,W] (!= 0).
 * column_hint != 0 selects the kernel that walks the 4 pixels of an image column jointly and loads an
 * out_grad row once for all of them that share the voxel — the common case on Z == 1 BEV grids; results
 * are the same either way, only the speed differs (pass nx[2] == 1). */
int bevpool_v2_backward_dense(const void* out_grad, void* depth_grad, void* feat_grad,
                              const void* depth, const void* feat, const int32_t* point_rank,
                              int bn, int d, int h, int w, int c, int feat_grad_nchw, int column_hint,
                              int dtype, void* stream);

/* [B,C,Z,Y,X] <-> [B,Z,Y,X,C] tile transpose (bev_pool.py:69 / :91 as one coalesced pass).
 * to_channels_last != 0: src is BCZYX, dst is BZYXC; else the reverse. */
int bevpool_grid_transpose(const void* src, void* dst, int b, int c, int64_t zyx,
                           int to_channels_last, int dtype, void* stream);

/* ------------------------------------------------------------------ v1 op (ops/bev_pool)
 * Replaces bev_pool_forward / bev_pool_backward of ops/bev_pool/src/bev_pool.cpp:27-94 and the kernels of
 * ops/bev_pool/src/bev_pool_cuda.cu:20-84. x is [n, c], already multiplied by depth and sorted by voxel rank;
 * geom_feats is int32 [n, 4] = (h, w, d, b) indices as the reference kernel reads them (cur_geom_feats[0..3]);
 * out / out_grad are [b, d, h, w, c]. `out` is PRE-ZEROED by the caller (the reference allocates torch::zeros);
 * x_grad is fully written for every row that belongs to an interval. */
int bevpool_v1_forward(const void* x, const int32_t* geom_feats, const int32_t* interval_lengths,
                       const int32_t* interval_starts, void* out, int b, int d, int h, int w, int64_t n,
                       int64_t n_intervals, int c, int dtype, void* stream);
int bevpool_v1_backward(const void* out_grad, const int32_t* geom_feats, const int32_t* interval_lengths,
                        const int32_t* interval_starts, void* x_grad, int b, int d, int h, int w, int64_t n,
                        int64_t n_intervals, int c, int dtype, void* stream);

/* ------------------------------------------------------------------ sort-free fused view transform (forward)
 * get_geometry + voxel_pooling_prepare_v2 + bev_pool_v2 forward (cam_stream_lss_bevpoolv2.py:229-258, 294-351,
 * 260-292) as ONE pixel-major pass: ranks are computed per camera pixel block (and written to point_rank
 * [B*N*D*H*W] for bevpool_v2_backward_dense), depth-weighted feature rows are accumulated per run of points
 * sharing a voxel and pushed into an fp32 grid with vector REDs; no sort, no ranks_* arrays. The summation order
 * across image columns is therefore not fixed (see csrc/pool_scatter.cu); the sorted entry points above are the
 * deterministic alternative. feat is channels-last [B*N, H, W, C], C % 4 == 0, C <= 128.
 * from_geometry == 0: point_rank is an INPUT (e.g. from bevpool_prepare_v2) and frustum/rots/trans are unused.
 * Output layouts as bevpool_v2_forward_dense (n_frames x rows_per_frame x X voxels; s2c = (B*Z, Y)).
 * scratch: bevpool_view_forward_scratch_bytes() bytes, 16-byte aligned (0 for fp32 channels-last output). */
size_t bevpool_view_forward_scratch_bytes(int64_t n_voxels, int c, int layout, int dtype);
int bevpool_view_forward(const void* depth, const void* feat, const float* frustum, const float* rots,
                         const float* trans, const bevpool_grid_t* g, int c, int32_t* point_rank, int from_geometry,
                         void* out, int64_t n_frames, int64_t rows_per_frame, int layout, int dtype, void* scratch,
                         size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------ lift head
 * CamEncode.get_depth_feat (cam_stream_lss_bevpoolv2.py:134-141) fused with the NCHW->NHWC transpose of the
 * context features (:282): x [BN, D+C, H, W] -> depth [BN, D, H, W] = softmax over the first D channels and
 * feat = channels D..D+C-1 as [BN, H, W, C] (feat_channels_last != 0) or [BN, C, H, W]. One streaming pass each
 * way; the backward applies the softmax Jacobian (dx = y * (g - sum_d g*y)) and the inverse transpose. D <= 256. */
int bevpool_lift_forward(const void* x, void* depth, void* feat, int bn, int d, int c, int hw,
                         int feat_channels_last, int dtype, void* stream);
int bevpool_lift_backward(const void* depth, const void* depth_grad, const void* feat_grad, void* x_grad,
                          int bn, int d, int c, int hw, int feat_channels_last, int dtype, void* stream);

/* ------------------------------------------------------------------ pillar scatter
 * `pts_middle_encoder` of the RCFusion detector (rcfusion/detectors/rcfusion_faster_rcnn.py:100; config
 * RCFusion_NewScenes/rcfusion_lss.py:63-64) = mmdet3d v0.17.1 PointPillarsScatter (dependency, not in the tree):
 * canvas [B, C, ny, nx] = zeros, canvas[b, :, y, x] = voxel_features[i, :] for coors[i] = (b, z, y, x) int32
 * (16-byte aligned [P, 4]); the last pillar wins on duplicated cells. pillar_index: int32 [B, ny*nx] output
 * (winning pillar per cell, -1 = empty). Every canvas element is written (no memset needed).
 * Backward: voxel_grad [P, C] = canvas_grad[b, :, y, x]. */
int bevpool_pillar_scatter_forward(const void* voxel_features, const int32_t* coors, void* canvas,
                                   int32_t* pillar_index, int n_pillars, int c, int b, int ny, int nx, int dtype,
                                   void* stream);
int bevpool_pillar_scatter_backward(const void* canvas_grad, const int32_t* coors, void* voxel_grad, int n_pillars,
                                    int c, int b, int ny, int nx, int dtype, void* stream);

/* ------------------------------------------------------------------ cross-modal fusion glue
 * The bandwidth-bound pieces of Cross_Modal_Fusion.forward (rcfusion/detectors/BEVCross_modal_attention.py:31-43)
 * around its convolutions, NCHW contiguous tensors:
 *   channel_avg_max: out [B,2,H,W] = cat(mean over C, max over C) of x [B,C,H,W] (:32-34, :36-38); argmax [B,H,W]
 *                    int32 is written for the backward (first maximum wins, as torch.max).
 *   gate_concat:     out [N,Ca+Cb,H,W] = cat(a * att_for_a, b * att_for_b) with att_* [N,1,H,W] (:40-42; the
 *                    reference passes att_for_a = radar_att for a = img_bev and att_for_b = img_att for b = radar_bev).
 * hw = H*W. The backward entry points write every element of their outputs. */
int bevpool_channel_avg_max_forward(const void* x, void* out, int32_t* argmax, int b, int c, int64_t hw, int dtype,
                                    void* stream);
int bevpool_channel_avg_max_backward(const void* out_grad, const int32_t* argmax, void* x_grad, int b, int c,
                                     int64_t hw, int dtype, void* stream);
int bevpool_gate_concat_forward(const void* a, const void* b, const void* att_for_a, const void* att_for_b, void* out,
                                int n, int ca, int cb, int64_t hw, int dtype, void* stream);
int bevpool_gate_concat_backward(const void* out_grad, const void* a, const void* b, const void* att_for_a,
                                 const void* att_for_b, void* a_grad, void* b_grad, void* att_for_a_grad,
                                 void* att_for_b_grad, int n, int ca, int cb, int64_t hw, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVPOOL_B200_H_ */
