"""Host-side mirror of the reference's v1 operator module projects/mmdet3d_plugin/ops/bev_pool/bev_pool.py
(MIT-BEVFusion style pooling): same names, argument order and return layout.

    bev_pool(feats [N,C], coords [N,4], B, D, H, W) -> [B, C, D, H, W]

`feats` are already multiplied by depth; `coords[:, 0..3]` index (H, W, D, B) of the output, exactly as the
reference kernel reads them. Ranking and the argsort stay torch ops on the device, as in the reference
(bev_pool.py:84-92); the interval sums and their backward run on the sm_100a library.
"""
import torch

from . import _lib
from .bev_pool import _dtype_code, _ptr, _require_cuda, _stream

__all__ = ["bev_pool"]


class QuickCumsumCuda(torch.autograd.Function):
    """v1 contract (ops/bev_pool/bev_pool.py:37-80): x and geom_feats are sorted by `ranks`; returns [B,D,H,W,C]."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks, B, D, H, W):
        _require_cuda("x", x)
        _require_cuda("geom_feats", geom_feats)
        if x.dim() != 2 or geom_feats.shape != (x.shape[0], 4):
            raise ValueError("x must be [N, C] and geom_feats [N, 4]")
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        kept = torch.ones(x.shape[0], device=x.device, dtype=torch.bool)
        kept[1:] = ranks[1:] != ranks[:-1]
        interval_starts = torch.where(kept)[0].int()
        interval_lengths = torch.zeros_like(interval_starts)
        if interval_starts.numel():
            interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
            interval_lengths[-1] = x.shape[0] - interval_starts[-1]
        geom_feats = geom_feats.int().contiguous()
        out = x.new_zeros((B, D, H, W, x.shape[1]))
        lib = _lib.load()
        _lib.check(lib.bevpool_v1_forward(_ptr(x), _ptr(geom_feats), _ptr(interval_lengths), _ptr(interval_starts),
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
        lib = _lib.load()
        _lib.check(lib.bevpool_v1_backward(_ptr(out_grad), _ptr(geom_feats), _ptr(interval_lengths),
                                           _ptr(interval_starts), _ptr(x_grad), B, D, H, W, n, interval_starts.numel(), c,
                                           _dtype_code(out_grad), _stream()), "bevpool_v1_backward")
        return x_grad, None, None, None, None, None, None


def bev_pool(feats, coords, B, D, H, W):
    """Drop-in for the reference `bev_pool` (ops/bev_pool/bev_pool.py:83-97)."""
    assert feats.shape[0] == coords.shape[0]
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    indices = ranks.argsort(stable=True)
    feats, coords, ranks = feats[indices], coords[indices], ranks[indices]
    x = QuickCumsumCuda.apply(feats, coords, ranks, B, D, H, W)
    return x.permute(0, 4, 1, 2, 3).contiguous()
