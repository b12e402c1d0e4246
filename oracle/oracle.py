"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's camera->BEV hot path. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import this module; the product package never does.

Each function cites the reference lines it follows (paths relative to
/root/reference/projects/mmdet3d_plugin):

  gen_dx_bx            bevfusion/detectors/cam_stream_lss_bevpoolv2.py:77-82
  create_frustum       same :216-227
  get_geometry         same :229-258   (the live `else` branch :244-251 only)
  prepare_v2           same :294-351
  bev_pool_v2_forward  ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48  (+ zeros, ops/bev_pool_v2/bev_pool.py:27)
  bev_pool_v2_backward ops/bev_pool_v2/bev_pool.py:44-83 + src/bev_pool_cuda.cu:67-121
  bev_pool_v1          ops/bev_pool/bev_pool.py:83-97 + ops/bev_pool/src/bev_pool_cuda.cu:20-84 (v1 op)
  cumsum_voxel_pooling cam_stream_lss_bevpoolv2.py:85-122 (QuickCumsum) inside upstream-LSS glue
                       — the CPU BASELINE, not a parity oracle (global cumsum loses precision)

Third-party arithmetic on the path: PyTorch (`torch==1.9.1+cu111` pinned by the
reference README:143) — trunc-toward-zero `.long()`, IEEE fp32 divide, `argsort`
tie order. None is pinned by a reference test; this oracle is pinned instead by
golden vectors produced by running the reference's own Python on CPU under
torch 2.11 (tests/golden/make_golden.py) where argsort is stable, and by the
reference's KAT (bev_pool.py:145-176).

Integer work is numpy (vectorised) or C (oracle_voxel_rank); fp32 pooling is the C
library liboracle.so built by oracle/Makefile.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_f, _i, _l = ctypes.c_float, ctypes.c_int, ctypes.c_int64


# --------------------------------------------------------------------------- grid / frustum
def gen_dx_bx(xbound, ybound, zbound):
    rows = [xbound, ybound, zbound]
    dx = np.array([r[2] for r in rows], dtype=np.float32)
    bx = np.array([r[0] + r[2] / 2.0 for r in rows], dtype=np.float32)
    nx = np.array([int((r[1] - r[0]) / r[2]) for r in rows], dtype=np.int64)   # LongTensor(float) truncates
    return dx, bx, nx


def _linspace_f32(lo, hi, n):
    """torch.linspace(lo, hi, n, dtype=float32) CPU semantics: step rounded to fp32, first half
    counted up from lo, second half counted down from hi, each element a single-rounding
    multiply-add (FMA) — reproduced here in float64 (exact product) then rounded once."""
    lo, hi = np.float32(lo), np.float32(hi)
    if n == 1:
        return np.array([lo], dtype=np.float32)
    step = np.float64(np.float32((hi - lo) / np.float32(n - 1)))
    i = np.arange(n)
    up = np.float64(lo) + step * i
    down = np.float64(hi) - step * (n - 1 - i)
    return np.where(i < n // 2, up, down).astype(np.float32)


def create_frustum(final_dim, downsample, dbound):
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    lo, hi, st = dbound
    n = int(np.ceil((hi - lo) / st))
    ds = (np.float32(lo) + np.arange(n, dtype=np.float32) * np.float32(st)).astype(np.float32)
    xs = _linspace_f32(0, ogfW - 1, fW)
    ys = _linspace_f32(0, ogfH - 1, fH)
    fr = np.empty((n, fH, fW, 3), dtype=np.float32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = ds[:, None, None]
    return fr


def get_geometry(frustum, rots, trans):
    """coor[b,n,d,h,w,:] = rots[b,n] @ (u*d, v*d, d) + trans[b,n], every multiply and
    add rounded separately in fp32, products summed left to right (no FMA): this is what
    the CPU batched matmul at :250 produces bit for bit (checked against the goldens)."""
    fr = frustum.astype(np.float32)
    px = (fr[..., 0] * fr[..., 2]).astype(np.float32)[None, None]
    py = (fr[..., 1] * fr[..., 2]).astype(np.float32)[None, None]
    pz = fr[..., 2][None, None]
    r = rots.astype(np.float32)[:, :, None, None, None]
    t = trans.astype(np.float32)[:, :, None, None, None]
    out = np.empty(rots.shape[:2] + fr.shape[:3] + (3,), dtype=np.float32)
    for a in range(3):
        acc = (r[..., a, 0] * px).astype(np.float32)
        acc = (acc + (r[..., a, 1] * py).astype(np.float32)).astype(np.float32)
        acc = (acc + (r[..., a, 2] * pz).astype(np.float32)).astype(np.float32)
        out[..., a] = (acc + t[..., a]).astype(np.float32)
    return out


# --------------------------------------------------------------------------- prepare
def voxel_rank(coor, dx, bx, nx):
    """int64 rank per frustum point, -1 if outside (C loop, exact fp32 semantics)."""
    coor = np.ascontiguousarray(coor, dtype=np.float32)
    B = coor.shape[0]
    n = coor.size // 3
    dx = np.asarray(dx, dtype=np.float32)
    bx = np.asarray(bx, dtype=np.float32)
    lo = (bx - (dx / np.float32(2.0)).astype(np.float32)).astype(np.float32)
    nx = np.ascontiguousarray(nx, dtype=np.int64)
    out = np.empty(n, dtype=np.int64)
    _lib().oracle_voxel_rank(_p(coor, _f), _l(n), _l(n // max(B, 1)), _p(lo, _f), _p(dx, _f), _p(nx, _l), _p(out, _l))
    return out


def prepare_v2(coor, dx, bx, nx):
    """-> (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths) int32,
    or five Nones when no point is kept (:344-345)."""
    B, N, D, H, W, _ = coor.shape
    rank = voxel_rank(coor, dx, bx, nx)
    kept = np.nonzero(rank >= 0)[0]
    if kept.size == 0:
        return None, None, None, None, None
    ranks_depth = kept.astype(np.int64)
    ranks_feat = (kept // (D * H * W)) * (H * W) + kept % (H * W)
    ranks_bev = rank[kept]
    order = np.argsort(ranks_bev, kind="stable")
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    head = np.ones(ranks_bev.size, dtype=bool)
    head[1:] = ranks_bev[1:] != ranks_bev[:-1]
    starts = np.nonzero(head)[0]
    lengths = np.empty_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.size - starts[-1]
    return tuple(np.ascontiguousarray(a.astype(np.int32)) for a in (ranks_bev, ranks_depth, ranks_feat, starts, lengths))


def intervals_from_sorted(ranks):
    head = np.ones(ranks.size, dtype=bool)
    head[1:] = ranks[1:] != ranks[:-1]
    starts = np.nonzero(head)[0].astype(np.int32)
    lengths = np.diff(np.append(starts, ranks.size)).astype(np.int32)
    return starts, lengths


# --------------------------------------------------------------------------- pooling
def bev_pool_v2_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                        interval_starts, interval_lengths, exact=False):
    """out[B,Z,Y,X,C] fp32; untouched voxels stay 0 (bev_pool.py:27)."""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    c = feat.shape[-1]
    out = np.zeros(tuple(int(v) for v in bev_feat_shape), dtype=np.float32)
    args = [np.ascontiguousarray(a, dtype=np.int32) for a in
            (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)]
    _lib().oracle_bev_pool_v2_fwd(_i(c), _i(args[3].size), _p(depth, _f), _p(feat, _f),
                                  *[_p(a, _i) for a in args], _p(out, _f), _i(1 if exact else 0))
    return out


def bev_pool_v2_backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, exact=False):
    """(depth_grad like depth, feat_grad like feat); out_grad is [B,Z,Y,X,C].
    Regroups by ranks_feat first, as QuickCumsumCuda.backward does (bev_pool.py:47-57)."""
    out_grad = np.ascontiguousarray(out_grad, dtype=np.float32)
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    c = feat.shape[-1]
    order = np.argsort(np.asarray(ranks_feat), kind="stable")
    rf = np.ascontiguousarray(np.asarray(ranks_feat)[order], dtype=np.int32)
    rd = np.ascontiguousarray(np.asarray(ranks_depth)[order], dtype=np.int32)
    rb = np.ascontiguousarray(np.asarray(ranks_bev)[order], dtype=np.int32)
    starts, lengths = intervals_from_sorted(rf)
    dg = np.zeros_like(depth)
    fg = np.zeros_like(feat)
    _lib().oracle_bev_pool_v2_bwd(_i(c), _i(starts.size), _p(out_grad, _f), _p(depth, _f), _p(feat, _f),
                                  _p(rd, _i), _p(rf, _i), _p(rb, _i), _p(starts, _i), _p(lengths, _i),
                                  _p(dg, _f), _p(fg, _f), _i(1 if exact else 0))
    return dg, fg


# --------------------------------------------------------------------------- lift head
def lift_head(x, D, C):
    """CamEncode.get_depth_feat (cam_stream_lss_bevpoolv2.py:134-141) after the depthnet: softmax over the first D
    channels (float64 inside, rounded once) and the C context channels as a slice. -> (depth, feat) NCHW."""
    x = np.asarray(x, dtype=np.float32)
    z = x[:, :D].astype(np.float64)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32), x[:, D:D + C].copy()


def lift_head_backward(depth, depth_grad, feat_grad):
    """Softmax Jacobian (dx = y * (g - sum_d g*y)) and the pass-through of the context gradient. -> x_grad NCHW."""
    y, g = depth.astype(np.float64), depth_grad.astype(np.float64)
    dx = y * (g - (g * y).sum(axis=1, keepdims=True))
    return np.concatenate([dx.astype(np.float32), np.asarray(feat_grad, dtype=np.float32)], axis=1)


# --------------------------------------------------------------------------- cross-modal fusion glue
def channel_avg_max(x):
    """BEVCross_modal_attention.py:32-34 / :36-38: cat([mean over C, max over C], 1), float64 mean rounded once."""
    x = np.asarray(x, dtype=np.float32)
    return np.stack([x.astype(np.float64).mean(axis=1).astype(np.float32), x.max(axis=1)], axis=1)


def channel_avg_max_backward(x, g):
    """mean: g/C to every channel; max: g to the FIRST maximal channel (torch.max's gradient)."""
    x = np.asarray(x, dtype=np.float32)
    B, C, H, W = x.shape
    dx = np.repeat((g[:, 0:1].astype(np.float64) / C), C, axis=1)
    am = x.argmax(axis=1)
    bi, hi, wi = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
    dx[bi, am, hi, wi] += g[:, 1].astype(np.float64)
    return dx.astype(np.float32)


def gate_concat(a, b, att_for_a, att_for_b):
    """:40-42: cat([a * att_for_a, b * att_for_b], 1)."""
    return np.concatenate([a * att_for_a, b * att_for_b], axis=1).astype(np.float32)


def gate_concat_backward(g, a, b, att_for_a, att_for_b):
    ca = a.shape[1]
    g64 = g.astype(np.float64)
    return ((g[:, :ca] * att_for_a).astype(np.float32), (g[:, ca:] * att_for_b).astype(np.float32),
            (g64[:, :ca] * a).sum(axis=1, keepdims=True).astype(np.float32),
            (g64[:, ca:] * b).sum(axis=1, keepdims=True).astype(np.float32))


# --------------------------------------------------------------------------- pillar scatter
def pillar_scatter(voxel_features, coors, batch_size, ny, nx):
    """mmdet3d v0.17.1 PointPillarsScatter.forward_batch (the reference's pts_middle_encoder,
    rcfusion_faster_rcnn.py:100; mmdet3d is a dependency outside the reference tree — PARITY UNPINNED for this
    function: restated from the published algorithm, checked against torch index assignment in the tests):
    per sample a zero canvas [C, ny*nx], canvas[:, y*nx + x] = features.T in pillar order (last wins)."""
    f = np.asarray(voxel_features, dtype=np.float32)
    co = np.asarray(coors).astype(np.int64)
    out = np.zeros((batch_size, f.shape[1], ny * nx), dtype=np.float32)
    for i in range(f.shape[0]):
        b, _, y, x = co[i]
        out[b, :, y * nx + x] = f[i]
    return out.reshape(batch_size, f.shape[1], ny, nx)


def pillar_scatter_backward(canvas_grad, coors):
    g = np.asarray(canvas_grad, dtype=np.float32)
    co = np.asarray(coors).astype(np.int64)
    return np.stack([g[b, :, y, x] for b, _, y, x in co]) if len(co) else np.zeros((0, g.shape[1]), np.float32)


# --------------------------------------------------------------------------- v1 op
def bev_pool_v1(feats, coords, B, D, H, W):
    """ops/bev_pool/bev_pool.py:83-97 + src/bev_pool_cuda.cu:20-42: -> (out [B,C,D,H,W], order, starts, lengths)."""
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    coords = np.asarray(coords).astype(np.int64)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    order = np.argsort(ranks, kind="stable")
    x = np.ascontiguousarray(feats[order])
    geom = np.ascontiguousarray(coords[order].astype(np.int32))
    starts, lengths = intervals_from_sorted(ranks[order])
    c = x.shape[1]
    out = np.zeros((B, D, H, W, c), dtype=np.float32)
    _lib().oracle_bev_pool_v1_fwd(_i(D), _i(H), _i(W), _i(c), _i(starts.size), _p(x, _f), _p(geom, _i), _p(starts, _i),
                                  _p(lengths, _i), _p(out, _f))
    return out.transpose(0, 4, 1, 2, 3).copy(), order, geom, starts, lengths


def bev_pool_v1_backward(out_grad_bcdhw, order, geom, starts, lengths, D, H, W):
    """Gradient w.r.t. the ORIGINAL (unsorted) feats rows."""
    og = np.ascontiguousarray(out_grad_bcdhw.transpose(0, 2, 3, 4, 1), dtype=np.float32)
    c = og.shape[-1]
    xg = np.zeros((order.size, c), dtype=np.float32)
    _lib().oracle_bev_pool_v1_bwd(_i(D), _i(H), _i(W), _i(c), _i(starts.size), _p(og, _f), _p(geom, _i), _p(starts, _i),
                                  _p(lengths, _i), _p(xg, _f))
    out = np.zeros_like(xg)
    out[order] = xg
    return out


# --------------------------------------------------------------------------- CPU baseline (torch)
def cumsum_voxel_pooling(coor, depth, feat, dx, bx, nx):
    """The reference's PyTorch cumsum voxel-pooling path on CPU: outer product depth x feat,
    voxelise with the reference's expression, mask, rank, argsort, then the cumsum trick
    (x.cumsum(0); keep the last row of each rank run; first-difference) and a dense scatter
    into [B,C,Z,Y,X]. Port of QuickCumsum.forward (:96-113) + upstream LSS voxel_pooling glue.
    torch tensors in, torch tensor out; used only as the timed CPU baseline."""
    import torch
    B, N, D, H, W, _ = coor.shape
    C = feat.shape[2]
    x = (depth.unsqueeze(-1) * feat.permute(0, 1, 3, 4, 2).unsqueeze(2)).reshape(-1, C)
    g = ((coor - (bx - dx / 2.)) / dx).long().view(-1, 3)
    bidx = torch.arange(B).view(B, 1).expand(B, N * D * H * W).reshape(-1, 1)
    g = torch.cat((g, bidx), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x, g = x[kept], g[kept]
    ranks = g[:, 3] * (nx[2] * nx[1] * nx[0]) + g[:, 2] * (nx[1] * nx[0]) + g[:, 1] * nx[0] + g[:, 0]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x = x.cumsum(0)
    last = torch.ones(x.shape[0], dtype=torch.bool)
    last[:-1] = ranks[1:] != ranks[:-1]
    x, g = x[last], g[last]
    x = torch.cat((x[:1], x[1:] - x[:-1]))
    final = torch.zeros((B, C, int(nx[2]), int(nx[1]), int(nx[0])), dtype=x.dtype)
    final[g[:, 3], :, g[:, 2], g[:, 1], g[:, 0]] = x
    return final


# --------------------------------------------------------------------------- CPU reference step (torch)
def _quick_cumsum_function():
    import torch

    class QuickCumsumPort(torch.autograd.Function):
        """Port of QuickCumsum (cam_stream_lss_bevpoolv2.py:96-122): cumsum over all kept points,
        keep the last row of every rank run, first-difference; backward gathers gradx by run id."""

        @staticmethod
        def forward(ctx, x, ranks):
            x = x.cumsum(0)
            last = torch.ones(x.shape[0], dtype=torch.bool)
            last[:-1] = ranks[1:] != ranks[:-1]
            x = x[last]
            x = torch.cat((x[:1], x[1:] - x[:-1]))
            ctx.save_for_backward(last)
            return x, last

        @staticmethod
        def backward(ctx, gradx, _gl):
            (last,) = ctx.saved_tensors
            run = torch.cumsum(last, 0)
            run[last] -= 1
            return gradx[run], None

    return QuickCumsumPort


def cpu_view_transform_step(frustum, rots, trans, depth, feat, out_grad, dx, bx, nx):
    """One training-step pass of the whole path on CPU with the reference's PyTorch code shape:
    get_geometry (:244-251) -> voxelise/mask/rank/argsort (:317-338) -> cumsum pooling (QuickCumsum,
    :96-122) -> dense [B,C,Z,Y,X] -> backward through it (autograd) for d(depth), d(feat).
    torch CPU tensors; returns (bev, depth_grad, feat_grad). This is what bench.py times as the
    reference arm and the cpu_baseline — never a parity oracle."""
    import torch
    QC = _quick_cumsum_function()
    B, N = trans.shape[:2]
    D, H, W, _ = frustum.shape
    C = feat.shape[2]
    depth = depth.detach().requires_grad_()
    feat = feat.detach().requires_grad_()
    pts = frustum.repeat(B, N, 1, 1, 1, 1).unsqueeze(-1)
    pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
    coor = rots.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1) + trans.view(B, N, 1, 1, 1, 3)
    g = ((coor - (bx - dx / 2.)) / dx).long().view(-1, 3)
    bidx = torch.arange(B).view(B, 1).expand(B, N * D * H * W).reshape(-1, 1)
    g = torch.cat((g, bidx), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x = (depth.unsqueeze(-1) * feat.permute(0, 1, 3, 4, 2).unsqueeze(2)).reshape(-1, C)
    x, g = x[kept], g[kept]
    ranks = g[:, 3] * (nx[2] * nx[1] * nx[0]) + g[:, 2] * (nx[1] * nx[0]) + g[:, 1] * nx[0] + g[:, 0]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x, last = QC.apply(x, ranks)
    g = g[last]
    final = _scatter_dense((B, C, int(nx[2]), int(nx[1]), int(nx[0])), g, x)
    final.backward(out_grad)
    return final.detach(), depth.grad, feat.grad


def _scatter_dense(shape, g, x):
    import torch
    B, C, Z, Y, X = shape
    flat = ((g[:, 3] * Z + g[:, 2]) * Y + g[:, 1]) * X + g[:, 0]              # voxel id without channel
    out = torch.zeros((B * Z * Y * X, C), dtype=x.dtype).index_copy(0, flat, x)
    return out.view(B, Z, Y, X, C).permute(0, 4, 1, 2, 3).contiguous()


def cpu_reference_class_step(ref, lss, rots, trans, depth, feat, out_grad):
    """The same CPU step with the REFERENCE'S OWN code wherever it still exists (bench.py `kind: "reference"`):
    `lss` is an unmodified reference LiftSplatShoot (oracle/refimport.py) — its get_geometry (:229-258) and its
    QuickCumsum autograd Function (:96-122) run as they are; the voxelise / mask / rank / argsort glue and the dense
    scatter are restated from upstream LSS `voxel_pooling` (the reference deleted that method and kept only the cumsum
    core). torch CPU tensors; returns (bev, depth_grad, feat_grad)."""
    import torch
    B, N = trans.shape[:2]
    D, H, W, _ = lss.frustum.shape
    C = feat.shape[2]
    dx, bx, nx = lss.dx, lss.bx, lss.nx
    depth = depth.detach().requires_grad_()
    feat = feat.detach().requires_grad_()
    with torch.no_grad():
        coor = lss.get_geometry(rots, trans)
    g = ((coor - (bx - dx / 2.)) / dx).long().view(-1, 3)
    bidx = torch.arange(B).view(B, 1).expand(B, N * D * H * W).reshape(-1, 1)
    g = torch.cat((g, bidx), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x = (depth.unsqueeze(-1) * feat.permute(0, 1, 3, 4, 2).unsqueeze(2)).reshape(-1, C)
    x, g = x[kept], g[kept]
    ranks = g[:, 3] * (nx[2] * nx[1] * nx[0]) + g[:, 2] * (nx[1] * nx[0]) + g[:, 1] * nx[0] + g[:, 0]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x, g = ref.QuickCumsum.apply(x, g, ranks)
    final = _scatter_dense((B, C, int(nx[2]), int(nx[1]), int(nx[0])), g, x)
    final.backward(out_grad)
    return final.detach(), depth.grad, feat.grad
