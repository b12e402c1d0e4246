/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not shipped, not on the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this. The product (omnihd-scenes_b200) must never
 * import, link or execute anything under oracle/.
 *
 * Plain-C CPU restatement of the reference's two CUDA kernels, one output
 * element at a time, in the reference's own summation order:
 *
 *   oracle_bev_pool_v2_fwd   follows  projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu:30-47
 *   oracle_bev_pool_v2_bwd   follows  projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu:78-120
 *                            (intervals already regrouped by ranks_feat, as
 *                             ops/bev_pool_v2/bev_pool.py:47-57 does on the host)
 *
 * Parity pin: the reference's known-answer test (bev_pool.py:145-176, stored as
 * tests/golden/kat_bev_pool_v2.npz), the reference's own CPU cumsum pooling run
 * on seeded inputs (tests/golden/tiny_*.npz, field cumsum_pooled), and — on the
 * GPU box — the reference's unmodified .cu compiled into oracle/_ref.
 *
 * acc_mode 0: float accumulator updated with fmaf() — what nvcc's default
 *             -fmad=true makes of `psum += *cur_feat * *cur_depth` (setup_bevpool2.py
 *             passes no -fmad flag).
 * acc_mode 1: double accumulator (the "exact" value both GPU kernels are judged against).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

void oracle_bev_pool_v2_fwd(int c, int n_intervals,
                            const float *depth, const float *feat,
                            const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                            const int *interval_starts, const int *interval_lengths,
                            float *out, int acc_mode)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = interval_starts[k], len = interval_lengths[k];
        float *o = out + (int64_t)ranks_bev[s] * c;
        for (int ch = 0; ch < c; ++ch) {
            if (acc_mode == 0) {
                float psum = 0.f;
                for (int i = 0; i < len; ++i)
                    psum = fmaf(feat[(int64_t)ranks_feat[s + i] * c + ch], depth[ranks_depth[s + i]], psum);
                o[ch] = psum;
            } else {
                double psum = 0.0;
                for (int i = 0; i < len; ++i)
                    psum += (double)feat[(int64_t)ranks_feat[s + i] * c + ch] * (double)depth[ranks_depth[s + i]];
                o[ch] = (float)psum;
            }
        }
    }
}

void oracle_bev_pool_v2_bwd(int c, int n_intervals,
                            const float *out_grad, const float *depth, const float *feat,
                            const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                            const int *interval_starts, const int *interval_lengths,
                            float *depth_grad, float *feat_grad, int acc_mode)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = interval_starts[k], len = interval_lengths[k];
        /* d(depth): one dot product over channels per point (.cu:91-105) */
        for (int i = 0; i < len; ++i) {
            const float *og = out_grad + (int64_t)ranks_bev[s + i] * c;
            const float *f = feat + (int64_t)ranks_feat[s + i] * c;
            if (acc_mode == 0) {
                float g = 0.f;
                for (int ch = 0; ch < c; ++ch) g = fmaf(og[ch], f[ch], g);
                depth_grad[ranks_depth[s + i]] = g;
            } else {
                double g = 0.0;
                for (int ch = 0; ch < c; ++ch) g += (double)og[ch] * (double)f[ch];
                depth_grad[ranks_depth[s + i]] = (float)g;
            }
        }
        /* d(feat): per channel, sum over the interval's points (.cu:107-120) */
        float *fg = feat_grad + (int64_t)ranks_feat[s] * c;
        for (int ch = 0; ch < c; ++ch) {
            if (acc_mode == 0) {
                float g = 0.f;
                for (int i = 0; i < len; ++i)
                    g = fmaf(out_grad[(int64_t)ranks_bev[s + i] * c + ch], depth[ranks_depth[s + i]], g);
                fg[ch] = g;
            } else {
                double g = 0.0;
                for (int i = 0; i < len; ++i)
                    g += (double)out_grad[(int64_t)ranks_bev[s + i] * c + ch] * (double)depth[ranks_depth[s + i]];
                fg[ch] = (float)g;
            }
        }
    }
}

/*
 * Voxel rank of every frustum point, -1 when outside the grid. Follows
 * cam_stream_lss_bevpoolv2.py:317-335: ((coor - (bx - dx/2)) / dx) in fp32 with a
 * true IEEE divide, .long() (truncation toward zero — values in (-1,0) become 0
 * and are KEPT), range mask, rank = b*Z*Y*X + z*Y*X + y*X + x.
 * `lo[a]` must be the fp32 value of (bx[a] - dx[a]/2) computed by the caller in fp32.
 * Compile with -ffp-contract=off (see Makefile) so nothing here is fused.
 */
void oracle_voxel_rank(const float *coor, int64_t n_points, int64_t points_per_frame,
                       const float *lo, const float *dx, const int64_t *nx, int64_t *rank_out)
{
    for (int64_t i = 0; i < n_points; ++i) {
        int64_t v[3];
        int ok = 1;
        for (int a = 0; a < 3; ++a) {
            volatile float t = coor[3 * i + a] - lo[a];
            volatile float q = t / dx[a];
            v[a] = (int64_t)q;
            ok &= (v[a] >= 0) & (v[a] < nx[a]);
        }
        const int64_t b = i / points_per_frame;
        rank_out[i] = ok ? b * (nx[2] * nx[1] * nx[0]) + v[2] * (nx[1] * nx[0]) + v[1] * nx[0] + v[0] : -1;
    }
}

/*
 * v1 op (ops/bev_pool): follows bev_pool_cuda.cu:20-42 (forward) and :61-84 (backward).
 * x is [n, c] sorted by rank; geom is int[n, 4]; out / out_grad are [b, d, h, w, c].
 */
void oracle_bev_pool_v1_fwd(int d, int h, int w, int c, int n_intervals, const float *x, const int *geom,
                            const int *starts, const int *lengths, float *out)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = starts[k], len = lengths[k];
        const int *g = geom + 4 * (int64_t)s;
        float *o = out + ((((int64_t)g[3] * d + g[2]) * h + g[0]) * w + g[1]) * c;
        for (int ch = 0; ch < c; ++ch) {
            float psum = 0.f;
            for (int i = 0; i < len; ++i) psum += x[(int64_t)(s + i) * c + ch];
            o[ch] = psum;
        }
    }
}

void oracle_bev_pool_v1_bwd(int d, int h, int w, int c, int n_intervals, const float *out_grad, const int *geom,
                            const int *starts, const int *lengths, float *x_grad)
{
    for (int k = 0; k < n_intervals; ++k) {
        const int s = starts[k], len = lengths[k];
        const int *g = geom + 4 * (int64_t)s;
        const float *o = out_grad + ((((int64_t)g[3] * d + g[2]) * h + g[0]) * w + g[1]) * c;
        for (int i = 0; i < len; ++i)
            for (int ch = 0; ch < c; ++ch) x_grad[(int64_t)(s + i) * c + ch] = o[ch];
    }
}
