"""Host-side mirror of the reference's v1 operator module projects/mmdet3d_plugin/ops/bev_pool/bev_pool.py
(MIT-BEVFusion style pooling): same names, argument order and return layout.

    bev_pool(feats [N,C], coords [N,4], B, D, H, W) -> [B, C, D, H, W]

`feats` are already multiplied by depth; `coords[:, 0..3]` index (H, W, D, B) of the output, exactly as the
reference kernel reads them. Everything after the rank arithmetic runs on the sm_100a library: the stable argsort and
the run-length segmentation (bev_pool.py:40-46, 84-92 do them with torch.argsort / torch.where) are the device radix
sort + head-flag scan of csrc/prepare.cu (`bevpool_v2_backward_regroup`: sort by one int32 key, carry the permutation,
emit interval starts / lengths and their count), the interval sums and their backward are `bevpool_v1_forward/_backward`.
One 4-byte device->host read per call (the interval count: the API returns nothing sized by it, but the kernels' grid
is); the reference synchronises at the same place (torch.where).
"""
import torch

from . import _lib
from .bev_pool import _dtype_code, _ptr, _require_cuda, _stream

__all__ = ["bev_pool"]


def _sort_and_segment(ranks, max_rank):
    """Stable argsort of int32 `ranks` + run-length segmentation on the device.
    Returns (order int32[n], interval_starts int32[I], interval_lengths int32[I])."""
    n = ranks.numel()
    dev = ranks.device
    if n == 0:
        e = torch.empty(0, dtype=torch.int32, device=dev)
        return e, e, e
    if n >= 2 ** 30 or max_rank >= 2 ** 31 - 1:
        raise ValueError("bev_pool v1: too many points / voxels for int32 ranks; shard the batch")
    lib = _lib.load()
    ident = torch.arange(n, dtype=torch.int32, device=dev)
    cols = (n + 63) // 64 * 64                                   # rows stay 256-byte aligned
    buf = torch.empty((5, cols), dtype=torch.int32, device=dev)   # order, sorted ranks, (unused), starts, lengths
    count = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.bevpool_v2_backward_regroup_workspace_bytes(n), dtype=torch.uint8, device=dev)
    _lib.check(lib.bevpool_v2_backward_regroup(_ptr(ident), _ptr(ranks), _ptr(ident), n, int(max_rank), _ptr(buf[0]),
                                               _ptr(buf[1]), _ptr(buf[2]), _ptr(buf[3]), _ptr(buf[4]), _ptr(count), _ptr(ws),
                                               ws.numel(), _stream()), "bevpool_v2_backward_regroup")
    n_int = int(count.item())
    return buf[0, :n], buf[3, :n_int], buf[4, :n_int]


class _PoolSorted(torch.autograd.Function):
    """Interval sums over rows of x that are already grouped: out[b, d, h, w, :] = sum of the interval's rows."""

    @staticmethod
    def forward(ctx, x, geom_feats, interval_starts, interval_lengths, B, D, H, W):
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        geom_feats = geom_feats.int().contiguous()
        out = x.new_zeros((B, D, H, W, x.shape[1]))
        _lib.check(_lib.load().bevpool_v1_forward(_ptr(x), _ptr(geom_feats), _ptr(interval_lengths), _ptr(interval_starts),
                                                  _ptr(out), B, D, H, W, x.shape[0], interval_starts.numel(), x.shape[1],
                                                  _dtype_code(x), _stream()), "bevpool_v1_forward")
        ctx.save_for_backward(interval_starts, interval_lengths, geom_feats)
        ctx.saved_shapes = B, D, H, W
        return out

    @staticmethod
    def backward(ctx, out_grad):
        interval_starts, interval_lengths, geom_feats = ctx.saved_tensors
        B, D, H, W = ctx.saved_shapes
        out_grad = out_grad.contiguous()
        n, c = geom_feats.shape[0], out_grad.shape[4]
        x_grad = out_grad.new_zeros((n, c))
        _lib.check(_lib.load().bevpool_v1_backward(_ptr(out_grad), _ptr(geom_feats), _ptr(interval_lengths),
                                                   _ptr(interval_starts), _ptr(x_grad), B, D, H, W, n, interval_starts.numel(),
                                                   c, _dtype_code(out_grad), _stream()), "bevpool_v1_backward")
        return x_grad, None, None, None, None, None, None, None


def _check_inputs(x, geom_feats):
    _require_cuda("x", x)
    _require_cuda("geom_feats", geom_feats)
    if x.dim() != 2 or geom_feats.shape != (x.shape[0], 4):
        raise ValueError("x must be [N, C] and geom_feats [N, 4]")


class QuickCumsumCuda:
    """v1 contract (ops/bev_pool/bev_pool.py:37-80): x and geom_feats are already sorted by `ranks`; returns
    [B, D, H, W, C]. The run-length segmentation of `ranks` happens on the device."""

    @staticmethod
    def apply(x, geom_feats, ranks, B, D, H, W):
        _check_inputs(x, geom_feats)
        _, starts, lengths = _sort_and_segment(ranks.int().contiguous(), B * D * H * W)   # sorted input: order = identity
        return _PoolSorted.apply(x, geom_feats, starts, lengths, B, D, H, W)


def bev_pool(feats, coords, B, D, H, W):
    """Drop-in for the reference `bev_pool` (ops/bev_pool/bev_pool.py:83-97): feats [N, C], coords [N, 4] -> [B, C, D, H, W]."""
    assert feats.shape[0] == coords.shape[0]
    _check_inputs(feats, coords)
    ranks = (coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]).int()
    order, starts, lengths = _sort_and_segment(ranks, B * D * H * W)
    order = order.long()
    x = _PoolSorted.apply(feats[order], coords[order], starts, lengths, B, D, H, W)
    return x.permute(0, 4, 1, 2, 3).contiguous()
