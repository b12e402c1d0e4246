"""Host-side mirror of the reference operator module
projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py (same names, argument order,
return layout and None conventions) on top of the sm_100a C-ABI library.

    bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths) -> [B, C, Z, Y, X]
    QuickCumsumCuda            reference-contract autograd Function -> [B, Z, Y, X, C]
    TRTBEVPoolv2               inference wrapper + ONNX symbolic (mmdeploy::bev_pool_v2)

What differs from the reference is only HOW: `bev_pool_v2` runs one fused kernel that
zero-fills, pools and writes [B,C,Z,Y,X] directly (reference: new_zeros + kernel + permute,
bev_pool.py:27,29,91); its backward is sort-free when the rank tensors are the ones returned
by this package's `voxel_pooling_prepare_v2` (reference: argsort + where + two new_zeros per
step, bev_pool.py:47-68) and otherwise regroups on the device with a radix sort.
Errors are loud: wrong device/dtype/shape raise ValueError, a failing kernel raises
BevPoolError; nothing falls back to PyTorch ops or the CPU.
"""
import os
import weakref

import torch

from . import _lib

__all__ = ['bev_pool_v2', 'TRTBEVPoolv2']


# ----------------------------------------------------------------------------- helpers
def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object and
    # costs ~4 us a call; this path is taken 6+ times per step)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise ValueError(f"unsupported dtype {t.dtype}")


def _as_int(v):
    # bev_feat_shape entries may be 0-dim (even CUDA) tensors: cam_stream_lss_bevpoolv2.py:283-285
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


def _require_cuda(name, t):
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name} must be a tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (bevpool_b200 has no CPU path)")


def _canon_floats(depth, feat):
    """Casts of bev_pool.py:19-20. fp32 unless BOTH depth and feat are bf16 (explicit extension)."""
    _require_cuda("depth", depth)
    _require_cuda("feat", feat)
    if feat.device != depth.device:
        raise ValueError(f"feat is on {feat.device}, depth on {depth.device}")
    if depth.dtype == torch.bfloat16 and feat.dtype == torch.bfloat16:
        depth, feat = depth.contiguous(), feat.contiguous()
    else:
        depth, feat = depth.contiguous().float(), feat.contiguous().float()
    if feat.dim() < 1 or feat.shape[-1] <= 0:
        raise ValueError("feat must be [..., C] with C > 0 (channels last)")
    return depth, feat


def _canon_ints(device, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths):
    """Casts of bev_pool.py:21-25 + the consistency checks the reference leaves to its kernel."""
    for n, t in (("ranks_depth", ranks_depth), ("ranks_feat", ranks_feat), ("ranks_bev", ranks_bev),
                 ("interval_starts", interval_starts), ("interval_lengths", interval_lengths)):
        _require_cuda(n, t)
        if t.device != device:
            raise ValueError(f"{n} is on {t.device}, depth on {device}")
    ints = [t.contiguous().int() for t in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)]
    n_points = ints[0].numel()
    if ints[1].numel() != n_points or ints[2].numel() != n_points:
        raise ValueError("ranks_depth, ranks_feat and ranks_bev must have the same length")
    if ints[3].numel() != ints[4].numel():
        raise ValueError("interval_starts and interval_lengths must have the same length")
    return tuple(ints)


def _canon_inputs(depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths):
    depth, feat = _canon_floats(depth, feat)
    return (depth, feat) + _canon_ints(depth.device, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)


def _shape5(bev_feat_shape, feat):
    if len(bev_feat_shape) != 5:
        raise ValueError("bev_feat_shape must be (B, Z, Y, X, C)")
    B, Z, Y, X, C = (_as_int(v) for v in bev_feat_shape)
    if C != feat.shape[-1]:
        raise ValueError(f"bev_feat_shape C={C} does not match feat channels {feat.shape[-1]}")
    if min(B, Z, Y, X) < 0:
        raise ValueError("negative bev_feat_shape")
    if B * Z * Y * X >= 2 ** 31:
        raise ValueError("B*Z*Y*X must fit int32 ranks")
    return B, Z, Y, X, C


# ----------------------------------------------------------------------------- plans
class PreparePlan:
    """Side information attached to the tensors `voxel_pooling_prepare_v2` returns.

    point_rank[P0] (voxel rank of every frustum point, -1 = dropped) is the inverse table
    that lets the backward walk a pixel's depth bins without sorting by ranks_feat.
    The plan rides on the `ranks_bev` tensor object itself (a Python attribute: it lives and dies with that
    tensor, no registry) and is honoured only if all five tensors are the very objects prepare returned, unmodified.
    """
    __slots__ = ("others", "versions", "point_rank", "bn", "d", "h", "w", "hw")

    def __init__(self, tensors, point_rank, bn, d, h, w):
        self.others = tuple(tensors[1:])
        self.versions = tuple(t._version for t in tensors)
        self.point_rank, self.bn, self.d, self.h, self.w, self.hw = point_rank, bn, d, h, w, h * w

    def matches(self, tensors):
        o, v = self.others, self.versions
        return (tensors[1] is o[0] and tensors[2] is o[1] and tensors[3] is o[2] and tensors[4] is o[3] and
                all(t._version == ver for t, ver in zip(tensors, v)))


_PLAN_ATTR = "_bevpool_b200_plan"


def register_plan(ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths, point_rank, bn, d, h, w):
    setattr(ranks_bev, _PLAN_ATTR, PreparePlan((ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths),
                                                 point_rank, bn, d, h, w))


def _find_plan(ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths, depth, feat):
    plan = getattr(ranks_bev, _PLAN_ATTR, None)
    if plan is None or not plan.matches((ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths)):
        return None
    if depth.numel() != plan.bn * plan.d * plan.hw or feat.numel() != plan.bn * plan.hw * feat.shape[-1]:
        return None
    return plan


# Rank tensors that did not come from this package's prepare: the fused forward walks the sorted point list by voxel,
# which equals "pool exactly the given intervals" only if the intervals are the run-length segmentation of a sorted,
# in-range ranks_bev. That is verified ON THE DEVICE once per tensor set (one boolean read-back, cached by identity
# and version like the plans); anything else (a subset of intervals, split intervals, unsorted or out-of-range
# ranks) takes the reference-contract kernel, which pools precisely the intervals it is given.
_CANONICAL = {}


def _intervals_are_canonical(plan_key, rb, starts, lengths, n_vox):
    key = id(plan_key[0])
    hit = _CANONICAL.get(key)
    if hit is not None and hit[2] == n_vox and all(r() is t and t._version == v for r, t, v in zip(hit[0], plan_key, hit[1])):
        return hit[3]
    n, i = rb.numel(), starts.numel()
    if n == 0 or i == 0:
        ok = False
    else:
        st, ln = starts.long(), lengths.long()
        heads = rb[st.clamp(0, n - 1)]
        ok = bool(((rb[1:] >= rb[:-1]).all() & (rb[0] >= 0) & (rb[-1] < n_vox) & (st[0] == 0) &
                   (st[1:] == st[:-1] + ln[:-1]).all() & (st[-1] + ln[-1] == n) & (ln > 0).all() &
                   (heads[1:] > heads[:-1]).all()).item())
    _CANONICAL[key] = ([weakref.ref(t) for t in plan_key], [t._version for t in plan_key], n_vox, ok)
    weakref.finalize(plan_key[0], _CANONICAL.pop, key, None)
    return ok


# ----------------------------------------------------------------------------- raw launches
def _launch_forward(depth, feat, out, rd, rf, rb, starts, lengths):
    lib = _lib.load()
    _lib.check(lib.bevpool_v2_forward(_ptr(depth), _ptr(feat), _ptr(out), _ptr(rd), _ptr(rf), _ptr(rb),
                                      _ptr(lengths), _ptr(starts), rd.numel(), starts.numel(), feat.shape[-1],
                                      _dtype_code(feat), _stream()), "bevpool_v2_forward")


def _launch_voxel_table(rb_sorted, n_points, counts_dev, n_vox_total):
    """vox_pt[v] = #sorted points with rank < v; stands in for the interval arrays on the device."""
    lib = _lib.load()
    vox_pt = torch.empty(n_vox_total + 1, dtype=torch.int32, device=rb_sorted.device)
    _lib.check(lib.bevpool_voxel_table(_ptr(rb_sorted), n_points, _ptr(counts_dev), n_vox_total, _ptr(vox_pt),
                                       _stream()), "bevpool_voxel_table")
    return vox_pt


def _launch_forward_dense(depth, feat, out, rd, rf, rb, vox_pt, frames, rows, x, layout, dhw=0, hw=0,
                          n_points=None, counts_dev=None):
    """rf=None: ranks_feat is derived from ranks_depth on the fly (dhw = D*H*W, hw = H*W).
    n_points: sorted-point count (default: len(rd)); with counts_dev it is an upper bound and the true count
    is read on the device. A scratch buffer is handed to the library for the streaming forward + layout pass."""
    lib = _lib.load()
    n_points = rd.numel() if n_points is None else n_points
    c = feat.shape[-1]
    code = _dtype_code(feat)
    nbytes = lib.bevpool_v2_forward_dense_scratch_bytes(n_points, frames * rows * x, c, layout, code) if c <= 128 else 0
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=out.device) if nbytes else None
    _lib.check(lib.bevpool_v2_forward_dense(_ptr(depth), _ptr(feat), _ptr(out), _ptr(rd), _ptr(rf), _ptr(rb), _ptr(vox_pt),
                                            n_points, _ptr(counts_dev), c, frames, rows, x, dhw, hw, layout, code,
                                            _ptr(scratch), nbytes, _stream()), "bevpool_v2_forward_dense")


def _launch_transpose(src, dst, b, c, zyx, to_channels_last):
    lib = _lib.load()
    _lib.check(lib.bevpool_grid_transpose(_ptr(src), _ptr(dst), b, c, zyx, 1 if to_channels_last else 0,
                                          _dtype_code(src), _stream()), "bevpool_grid_transpose")


def _backward_general(out_grad_cl, depth, feat, rd, rf, rb):
    """Regroup by ranks_feat on the device (bev_pool.py:47-57), then the pixel-major kernel."""
    lib = _lib.load()
    n = rd.numel()
    dev = feat.device
    depth_grad = torch.zeros_like(depth)
    feat_grad = torch.zeros_like(feat)
    if n == 0:
        return depth_grad, feat_grad
    n_pixels = feat.numel() // feat.shape[-1]
    bp = torch.empty((5, (n + 63) // 64 * 64), dtype=torch.int32, device=dev)[:, :n]   # rd, rf, rb, starts, lengths; rows 256 B aligned
    n_bp = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.bevpool_v2_backward_regroup_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.bevpool_v2_backward_regroup(_ptr(rd), _ptr(rf), _ptr(rb), n, max(n_pixels - 1, 0),
                                               _ptr(bp[0]), _ptr(bp[1]), _ptr(bp[2]), _ptr(bp[3]), _ptr(bp[4]),
                                               _ptr(n_bp), _ptr(ws), ws.numel(), _stream()),
               "bevpool_v2_backward_regroup")
    n_intervals = int(n_bp.item())      # the reference syncs here too (torch.where, bev_pool.py:53)
    _lib.check(lib.bevpool_v2_backward(_ptr(out_grad_cl), _ptr(depth_grad), _ptr(feat_grad), _ptr(depth), _ptr(feat),
                                       _ptr(bp[0]), _ptr(bp[1]), _ptr(bp[2]), _ptr(bp[4]), _ptr(bp[3]),
                                       n, n_intervals, feat.shape[-1], _dtype_code(feat), _stream()),
               "bevpool_v2_backward")
    return depth_grad, feat_grad


def _column_hint(Z):
    """Which sort-free backward kernel to prefer: the column kernels pay when the pixels of an image column share
    voxels (always for Z == 1). BEVPOOL_BWD_KERNEL=joint|block overrides (measurement / tests only)."""
    env = os.environ.get("BEVPOOL_BWD_KERNEL")
    if env in ("joint", "block"):
        return 1 if env == "joint" else 0
    return 1 if Z == 1 else 0


def _backward_dense(out_grad_cl, depth, feat, plan, column_hint, feat_nchw=False):
    """feat is channels-last [..., H, W, C]; feat_nchw: write feat_grad as [..., C, H, W] instead."""
    lib = _lib.load()
    depth_grad = torch.empty_like(depth)
    if feat_nchw:
        feat_grad = feat.new_empty(feat.shape[:-3] + (feat.shape[-1], feat.shape[-3], feat.shape[-2]))
    else:
        feat_grad = torch.empty_like(feat)
    _lib.check(lib.bevpool_v2_backward_dense(_ptr(out_grad_cl), _ptr(depth_grad), _ptr(feat_grad), _ptr(depth),
                                             _ptr(feat), _ptr(plan.point_rank), plan.bn, plan.d, plan.h, plan.w,
                                             feat.shape[-1], 1 if feat_nchw else 0, 1 if column_hint else 0,
                                             _dtype_code(feat), _stream()),
               "bevpool_v2_backward_dense")
    return depth_grad, feat_grad


# ----------------------------------------------------------------------------- autograd functions
class QuickCumsumCuda(torch.autograd.Function):
    """Reference-contract op (bev_pool.py:11-83): returns channels-last [B, Z, Y, X, C]."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths):
        depth, feat, rd, rf, rb, starts, lengths = _canon_inputs(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                                                 interval_starts, interval_lengths)
        shape = _shape5(bev_feat_shape, feat)
        out = feat.new_zeros(shape)
        _launch_forward(depth, feat, out, rd, rf, rb, starts, lengths)
        ctx.save_for_backward(rb, depth, feat, rf, rd)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        rb, depth, feat, rf, rd = ctx.saved_tensors
        out_grad = out_grad.contiguous().to(feat.dtype)
        depth_grad, feat_grad = _backward_general(out_grad, depth, feat, rd, rf, rb)
        return depth_grad, feat_grad, None, None, None, None, None, None


def _nchw_view_of(feat, depth):
    """The reference hands bev_pool_v2 `feat.permute(0, 1, 3, 4, 2)` of the neck's [B,N,C,H,W] tensor
    (cam_stream_lss_bevpoolv2.py:282) and lets `.contiguous()` copy it (bev_pool.py:20). If `feat` is such a view,
    return the contiguous [B,N,C,H,W] tensor behind it: our transpose kernel then makes the channels-last copy, and
    the sort-free backward writes the gradient straight back in [B,N,C,H,W] (returned as the same permuted view)."""
    if feat.dim() != 5 or feat.is_contiguous():
        return None
    if not (feat.dtype == torch.float32 or (feat.dtype == torch.bfloat16 and depth.dtype == torch.bfloat16)):
        return None
    v = feat.permute(0, 1, 4, 2, 3)
    return v if v.is_contiguous() else None


class _BevPoolV2Fused(torch.autograd.Function):
    """bev_pool_v2 as one fused pass: returns [B, C, Z, Y, X] directly."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths):
        plan_key = (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths)
        in_depth_dtype, in_feat_dtype = depth.dtype, feat.dtype
        _require_cuda("feat", feat)
        nchw = _nchw_view_of(feat, depth)
        if nchw is not None:
            Bf, Nf, Cf, Hf, Wf = nchw.shape
            feat = nchw.new_empty((Bf, Nf, Hf, Wf, Cf))
            if feat.numel():
                _launch_transpose(nchw, feat, Bf * Nf, Cf, Hf * Wf, True)       # [BN,C,HW] -> [BN,HW,C]
        depth, feat = _canon_floats(depth, feat)
        ctx.plan = _find_plan(*plan_key, depth, feat) if isinstance(ranks_bev, torch.Tensor) else None
        if ctx.plan is not None and ranks_bev.device == depth.device:
            rd, rf, rb, starts, lengths = ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths
        else:       # foreign rank tensors: casts and checks
            ctx.plan = None
            rd, rf, rb, starts, lengths = _canon_ints(depth.device, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                                                      interval_lengths)
        B, Z, Y, X, C = _shape5(bev_feat_shape, feat)
        ctx.shape = (B, Z, Y, X, C)
        ctx.in_dtypes = (in_depth_dtype, in_feat_dtype)
        ctx.nchw = nchw is not None
        # the fused kernel walks the sorted point list by voxel. Tensors that came from our prepare are canonical by
        # construction; anything else is verified on the device (see _intervals_are_canonical).
        fused_ok = C % 4 == 0 and rb.numel() > 0 and feat.data_ptr() % 16 == 0 and \
            (ctx.plan is not None or _intervals_are_canonical(plan_key, rb, starts, lengths, B * Z * Y * X))
        if not fused_ok:
            # odd channel counts (the reference's KAT has C=2), or intervals that are not the plain run-length
            # segmentation of a sorted ranks_bev: reference-contract kernel + transpose kernel
            out_cl = feat.new_zeros((B, Z, Y, X, C))
            _launch_forward(depth, feat, out_cl, rd, rf, rb, starts, lengths)
            out = feat.new_empty((B, C, Z, Y, X))
            _launch_transpose(out_cl, out, B, C, Z * Y * X, to_channels_last=False)
        else:
            out = feat.new_empty((B, C, Z, Y, X))
            vox_pt = _launch_voxel_table(rb, rb.numel(), None, B * Z * Y * X)
            _launch_forward_dense(depth, feat, out, rd, rf, rb, vox_pt, B, Z * Y, X, _lib.LAYOUT_BCZYX)
        ctx.save_for_backward(rb, depth, feat, rf, rd)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        rb, depth, feat, rf, rd = ctx.saved_tensors
        B, Z, Y, X, C = ctx.shape
        out_grad = out_grad.contiguous().to(feat.dtype)
        og_cl = out_grad.new_empty((B, Z, Y, X, C))
        _launch_transpose(out_grad, og_cl, B, C, Z * Y * X, to_channels_last=True)
        if ctx.plan is not None and C % 4 == 0:
            depth_grad, feat_grad = _backward_dense(og_cl, depth, feat, ctx.plan, column_hint=_column_hint(Z),
                                                    feat_nchw=ctx.nchw)
            if ctx.nchw:
                feat_grad = feat_grad.permute(0, 1, 3, 4, 2)
        else:
            depth_grad, feat_grad = _backward_general(og_cl, depth, feat, rd, rf, rb)
        return depth_grad.to(ctx.in_dtypes[0]), feat_grad.to(ctx.in_dtypes[1]), None, None, None, None, None, None


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                bev_feat_shape, interval_starts, interval_lengths):
    """Drop-in for the reference `bev_pool_v2` (bev_pool.py:86-92).

    depth [B,N,D,H,W]; feat [B,N,H,W,C] (channels last); ranks_* int [P]; interval_* int [I];
    bev_feat_shape (B,Z,Y,X,C) of ints or 0-dim tensors. Returns a new contiguous [B,C,Z,Y,X]
    tensor, differentiable w.r.t. depth and feat.
    """
    return _BevPoolV2Fused.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                 interval_starts, interval_lengths)


class TRTBEVPoolv2(torch.autograd.Function):
    """Inference wrapper (bev_pool.py:95-142): depth [N,D,H,W], feat [N,H,W,C] -> [1,out_h,out_w,C].
    The channels-last result is produced directly by the fused kernel (Z == 1, so
    [B,Z,Y,X,C] IS [1,out_h,out_w,C]) instead of pool + two permutes."""

    @staticmethod
    def symbolic(g, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths,
                 out_height=128, out_width=128):
        return g.op('mmdeploy::bev_pool_v2', depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts,
                    interval_lengths, out_height_i=out_height, out_width_i=out_width)

    @staticmethod
    def forward(g, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths,
                out_height=128, out_width=128):
        depth, feat, rd, rf, rb, starts, lengths = _canon_inputs(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                                                 interval_starts, interval_lengths)
        C = feat.shape[-1]
        key = (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths)
        if C % 4 != 0 or rb.numel() == 0 or feat.data_ptr() % 16 != 0 or \
                not _intervals_are_canonical(key, rb, starts, lengths, out_height * out_width):
            out = feat.new_zeros((1, out_height, out_width, C))
            _launch_forward(depth, feat, out, rd, rf, rb, starts, lengths)
            return out
        out = feat.new_empty((1, out_height, out_width, C))
        vox_pt = _launch_voxel_table(rb, rb.numel(), None, out_height * out_width)
        _launch_forward_dense(depth, feat, out, rd, rf, rb, vox_pt, 1, out_height, out_width, _lib.LAYOUT_BZYXC)
        return out
