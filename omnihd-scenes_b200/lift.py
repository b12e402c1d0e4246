"""Lift head: depth softmax + context split.

Mirrors `CamEncode.get_depth_feat` / `get_depth_dist` of the reference
(bevfusion/detectors/cam_stream_lss_bevpoolv2.py:131-141): the depthnet output x [BN, D+C, H, W] is cut into
D depth logits (softmax over dim 1) and C context channels. The reference returns the context as an NCHW slice
and `voxel_pooling_v2` permutes it to channels-last afterwards (:282); here one CUDA kernel produces the softmax
and, on request, the channels-last features in the same pass, and one kernel does the whole backward.
"""
import torch

from . import _lib
from .bev_pool import _dtype_code, _ptr, _stream


class _LiftSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, D, C, channels_last):
        if not x.is_cuda:
            raise ValueError("get_depth_feat: CUDA tensors only (this library has no CPU path)")
        if x.dim() != 4 or x.shape[1] < D + C:
            raise ValueError(f"x must be [BN, >= D+C = {D + C}, H, W], got {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        n_in = x.shape[1]
        if n_in != D + C:
            x = x[:, :D + C]
        x = x.contiguous()
        BN, _, H, W = x.shape
        depth = x.new_empty((BN, D, H, W))
        feat = x.new_empty((BN, H, W, C) if channels_last else (BN, C, H, W))
        _lib.check(_lib.load().bevpool_lift_forward(_ptr(x), _ptr(depth), _ptr(feat), BN, D, C, H * W,
                                                    1 if channels_last else 0, _dtype_code(x), _stream()),
                   "bevpool_lift_forward")
        ctx.save_for_backward(depth)
        ctx.dims = (BN, D, C, H, W, channels_last, n_in)
        return depth, feat

    @staticmethod
    def backward(ctx, depth_grad, feat_grad):
        (depth,) = ctx.saved_tensors
        BN, D, C, H, W, cl, n_in = ctx.dims
        if feat_grad is None:
            feat_grad = depth.new_zeros((BN, H, W, C) if cl else (BN, C, H, W))
        if depth_grad is None:
            depth_grad = torch.zeros_like(depth)
        depth_grad = depth_grad.contiguous().to(depth.dtype)
        feat_grad = feat_grad.contiguous().to(depth.dtype)
        x_grad = depth.new_empty((BN, D + C, H, W))
        _lib.check(_lib.load().bevpool_lift_backward(_ptr(depth), _ptr(depth_grad), _ptr(feat_grad), _ptr(x_grad), BN, D,
                                                     C, H * W, 1 if cl else 0, _dtype_code(depth), _stream()),
                   "bevpool_lift_backward")
        if n_in != D + C:                       # channels past D+C were never read
            x_grad = torch.cat([x_grad, x_grad.new_zeros((BN, n_in - D - C, H, W))], 1)
        return x_grad, None, None, None


def get_depth_feat(x, D, C, channels_last=False):
    """x [BN, D+C, H, W] -> (depth [BN, D, H, W] = softmax(x[:, :D], dim=1), feat = x[:, D:D+C]).
    feat is [BN, C, H, W] (reference layout) or, with channels_last, [BN, H, W, C] as the pool consumes it."""
    return _LiftSplit.apply(x, int(D), int(C), bool(channels_last))


def get_depth_dist(x):
    """Reference name (:131-132): softmax over dim 1."""
    depth, _ = _LiftSplit.apply(x, x.shape[1], 0, False)
    return depth
