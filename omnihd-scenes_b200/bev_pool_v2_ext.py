"""`bev_pool_v2_ext` on the sm_100a C-ABI library — the binding a maintainer adds inside the plugin.

The reference binds its native code with pybind as the module `bev_pool_v2_ext`
(ops/bev_pool_v2/src/bev_pool.cpp:106-111), imported by ops/bev_pool_v2/bev_pool.py:6 and called at
:29-38 (forward) and :70-81 (backward). This module exposes the SAME two entry points with the same
positional arguments (note: interval_lengths BEFORE interval_starts, bev_pool.cpp:37-38 / :83-84), so the
reference's own, unmodified `QuickCumsumCuda` / `bev_pool_v2` / `TRTBEVPoolv2` run on the new kernels:

    plugin.install_ext()        # sys.modules['projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool_v2_ext'] = this

Contract (as the reference's): tensors are CUDA, contiguous, float32 / int32 (the reference's Python side
casts them, bev_pool.py:19-25), `out` / `depth_grad` / `feat_grad` are pre-zeroed by the caller and only
touched elements are written; `c` is read from feat.size(4) / out_grad.size(4) (bev_pool.cpp:40,86), so
feat and out_grad are 5-D. Launches go to the CURRENT stream (the reference uses the legacy stream).
Unlike the reference (no checks at all) a wrong device / dtype / layout raises ValueError.
"""
import torch

from . import _lib

__all__ = ["bev_pool_v2_forward", "bev_pool_v2_backward"]


def _check(name, t, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (bevpool_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _check_all(floats, ints):
    for n, t in floats:
        _check(n, t, torch.float32)
    for n, t in ints:
        _check(n, t, torch.int32)
    dev = floats[0][1].device
    for n, t in floats + ints:
        if t.device != dev:
            raise ValueError(f"{n} is on {t.device}, expected {dev}")


def bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts):
    """bev_pool.cpp:30-57. out[ranks_bev[start_k], :] = sum_i depth[ranks_depth[i]] * feat[ranks_feat[i], :]."""
    _check_all([("depth", depth), ("feat", feat), ("out", out)],
               [("ranks_depth", ranks_depth), ("ranks_feat", ranks_feat), ("ranks_bev", ranks_bev),
                ("interval_lengths", interval_lengths), ("interval_starts", interval_starts)])
    if feat.dim() != 5:
        raise ValueError("feat must be 5-D [B, N, H, W, C] at this boundary (c = feat.size(4))")
    with torch.cuda.device(depth.device):
        rc = _lib.load().bevpool_v2_forward(
            depth.data_ptr(), feat.data_ptr(), out.data_ptr(), ranks_depth.data_ptr(), ranks_feat.data_ptr(),
            ranks_bev.data_ptr(), interval_lengths.data_ptr(), interval_starts.data_ptr(), ranks_depth.numel(),
            interval_lengths.numel(), feat.size(4), _lib.F32, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bevpool_v2_forward")


def bev_pool_v2_backward(out_grad, depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                         interval_lengths, interval_starts):
    """bev_pool.cpp:74-104. Rank arrays regrouped by ranks_feat, one interval per feature pixel (bev_pool.py:47-57)."""
    _check_all([("out_grad", out_grad), ("depth_grad", depth_grad), ("feat_grad", feat_grad), ("depth", depth),
                ("feat", feat)],
               [("ranks_depth", ranks_depth), ("ranks_feat", ranks_feat), ("ranks_bev", ranks_bev),
                ("interval_lengths", interval_lengths), ("interval_starts", interval_starts)])
    if out_grad.dim() != 5:
        raise ValueError("out_grad must be 5-D [B, Z, Y, X, C] at this boundary (c = out_grad.size(4))")
    with torch.cuda.device(depth.device):
        rc = _lib.load().bevpool_v2_backward(
            out_grad.data_ptr(), depth_grad.data_ptr(), feat_grad.data_ptr(), depth.data_ptr(), feat.data_ptr(),
            ranks_depth.data_ptr(), ranks_feat.data_ptr(), ranks_bev.data_ptr(), interval_lengths.data_ptr(),
            interval_starts.data_ptr(), ranks_depth.numel(), interval_lengths.numel(), out_grad.size(4), _lib.F32,
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bevpool_v2_backward")
