"""Pillar scatter: drop-in for the `pts_middle_encoder` the RCFusion / BEVFusion detectors
call (`rcfusion/detectors/rcfusion_faster_rcnn.py:100`, `bevfusion/detectors/bevf_faster_rcnn.py:99`; configured as
`dict(type='PointPillarsScatter', in_channels=64, output_shape=[320, 480])` in
`configs/RCFusion_NewScenes/rcfusion_lss.py:63-64`). The class itself lives in mmdet3d v0.17.1 (a dependency that is
not part of the reference tree); this mirrors its constructor, `forward(voxel_features, coors, batch_size=None)` and
result: a dense [B, C, ny, nx] canvas, zeros where no pillar lands.
"""
import torch
from torch import nn

from . import _lib
from .bev_pool import _dtype_code, _ptr, _stream


class _PillarScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, voxel_features, coors, batch_size, ny, nx):
        if not voxel_features.is_cuda:
            raise ValueError("PointPillarsScatter: CUDA tensors only (this library has no CPU path)")
        if voxel_features.dim() != 2 or coors.dim() != 2 or coors.shape[1] != 4 or coors.shape[0] != voxel_features.shape[0]:
            raise ValueError("expected voxel_features [P, C] and coors [P, 4] = (batch, z, y, x)")
        if voxel_features.dtype not in (torch.float32, torch.bfloat16):
            voxel_features = voxel_features.float()
        voxel_features = voxel_features.contiguous()
        coors = coors.to(torch.int32).contiguous()
        P, C = voxel_features.shape
        canvas = voxel_features.new_empty((batch_size, C, ny, nx))
        index = torch.empty((batch_size, ny * nx), dtype=torch.int32, device=canvas.device)
        _lib.check(_lib.load().bevpool_pillar_scatter_forward(_ptr(voxel_features), _ptr(coors), _ptr(canvas), _ptr(index), P, C,
                                                              batch_size, ny, nx, _dtype_code(voxel_features), _stream()),
                   "bevpool_pillar_scatter_forward")
        ctx.save_for_backward(coors)
        ctx.dims = (P, C, batch_size, ny, nx)
        return canvas

    @staticmethod
    def backward(ctx, canvas_grad):
        (coors,) = ctx.saved_tensors
        P, C, B, ny, nx = ctx.dims
        canvas_grad = canvas_grad.contiguous()
        if canvas_grad.dtype not in (torch.float32, torch.bfloat16):
            canvas_grad = canvas_grad.float()
        grad = canvas_grad.new_empty((P, C))
        _lib.check(_lib.load().bevpool_pillar_scatter_backward(_ptr(canvas_grad), _ptr(coors), _ptr(grad), P, C, B, ny, nx,
                                                               _dtype_code(canvas_grad), _stream()),
                   "bevpool_pillar_scatter_backward")
        return grad, None, None, None, None


class PointPillarsScatter(nn.Module):
    """Same constructor and call as mmdet3d's module: `PointPillarsScatter(in_channels, output_shape=[ny, nx])`,
    `forward(voxel_features [P, C], coors [P, 4] = (batch, z, y, x), batch_size=None) -> [B, C, ny, nx]`."""

    def __init__(self, in_channels, output_shape):
        super().__init__()
        self.output_shape = output_shape
        self.ny, self.nx = int(output_shape[0]), int(output_shape[1])
        self.in_channels = in_channels

    def forward(self, voxel_features, coors, batch_size=None):
        if voxel_features.shape[1] != self.in_channels:
            raise ValueError(f"voxel_features must have {self.in_channels} channels, got {voxel_features.shape[1]}")
        if batch_size is None:          # mmdet3d's forward_single
            batch_size = 1
            coors = torch.cat([torch.zeros_like(coors[:, :1]), coors[:, -3:]], 1) if coors.shape[1] == 4 else coors
        return _PillarScatter.apply(voxel_features, coors, int(batch_size), self.ny, self.nx)
