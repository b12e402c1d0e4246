"""Seeded synthetic 6-camera OmniHD-shaped inputs.

Host-side only (torch CPU): builds the `rots` / `trans` the reference's
detectors build from `img_metas[..]['lidar2img']`
(projects/mmdet3d_plugin/bevfusion/detectors/bevf_faster_rcnn.py:114-128)
for a ring of N cameras, plus the grid constants of each BASELINE.json config.
No dataset, no network: everything is a function of the seed.
"""
import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass(frozen=True)
class ViewConfig:
    """One camera->BEV workload. Bounds are (lo, hi, step) triples as in the
    reference's `grid_conf` (cam_stream_lss_bevpoolv2.py:163-168)."""
    name: str
    final_dim: Tuple[int, int]          # network input H, W
    downsample: int
    dbound: Tuple[float, float, float]
    xbound: Tuple[float, float, float]
    ybound: Tuple[float, float, float]
    zbound: Tuple[float, float, float]
    channels: int
    batch: int
    n_cams: int = 6
    dtype: str = "f32"

    @property
    def fH(self):
        return self.final_dim[0] // self.downsample

    @property
    def fW(self):
        return self.final_dim[1] // self.downsample

    @property
    def D(self):
        # torch.arange(*dbound) length (create_frustum, cam_stream_lss_bevpoolv2.py:220)
        return int(torch.arange(*self.dbound, dtype=torch.float).numel())


# BASELINE.json configs[0..4], concretised as in SURVEY.md §8(d).
CONFIGS = {
    "bevdet_r50_cpu": ViewConfig("bevdet_r50_cpu", (256, 704), 16, (1.0, 60.0, 1.0),
                                 (-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8), (-5.0, 3.0, 8.0), 80, 1),
    "bevdet_r50_b8": ViewConfig("bevdet_r50_b8", (256, 704), 16, (1.0, 60.0, 1.0),
                                (-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8), (-5.0, 3.0, 8.0), 80, 8),
    "bevdepth_hires_b16": ViewConfig("bevdepth_hires_b16", (512, 1408), 16, (1.0, 60.0, 0.5),
                                     (-51.2, 51.2, 0.512), (-51.2, 51.2, 0.512), (-5.0, 3.0, 8.0),
                                     80, 16, dtype="bf16"),
    "rcfusion_omnihd_b32": ViewConfig("rcfusion_omnihd_b32", (544, 960), 4, (1.0, 60.0, 1.0),
                                      (-60.0, 60.0, 0.5), (-40.0, 40.0, 0.5), (-3.0, 5.0, 0.5), 64, 32),
    "occ_200x200x16_b64": ViewConfig("occ_200x200x16_b64", (256, 704), 16, (1.0, 45.0, 0.5),
                                     (-40.0, 40.0, 0.4), (-40.0, 40.0, 0.4), (-1.0, 5.4, 0.4), 32, 64),
}


def camera_ring(batch: int, n_cams: int, final_dim, seed: int = 0, roll: float = 0.0, pitch: float = 0.0):
    """rots [B,N,3,3], trans [B,N,3] fp32 = inverse(lidar2img)[:3,:3], [:3,3].

    Ring geometry per SURVEY.md §8(d): yaw_i = 2*pi*i/N + (u-0.5)*0.05 with u drawn
    in (b, i) order; cam->lidar columns [right | down | fwd]; pinhole K with
    f = 0.6*W. The inverse is taken per matrix in fp32 exactly as
    bevf_faster_rcnn.py:119-121 does (`torch.Tensor(mat).inverse()`).
    roll / pitch (radians, default 0 = the SURVEY ring, bit-identical to before): every camera is additionally
    rotated about its optical axis / its right axis, so the pixels of one image column no longer share a voxel
    at a given depth (real calibrations; exercises the mixed-bin paths of the column kernels).
    """
    H, W = final_dim
    g = torch.Generator().manual_seed(seed)
    rots = torch.empty(batch, n_cams, 3, 3, dtype=torch.float32)
    trans = torch.empty(batch, n_cams, 3, dtype=torch.float32)
    K = torch.eye(4, dtype=torch.float64)
    K[0, 0] = 0.6 * W
    K[1, 1] = 0.6 * W
    K[0, 2] = W / 2.0
    K[1, 2] = H / 2.0
    for b in range(batch):
        for i in range(n_cams):
            u = torch.rand((), generator=g).item()
            yaw = 2.0 * math.pi * i / n_cams + (u - 0.5) * 0.05
            c, s = math.cos(yaw), math.sin(yaw)
            cam2lidar = torch.eye(4, dtype=torch.float64)
            cam2lidar[:3, 0] = torch.tensor([s, -c, 0.0], dtype=torch.float64)   # right
            cam2lidar[:3, 1] = torch.tensor([0.0, 0.0, -1.0], dtype=torch.float64)  # down
            cam2lidar[:3, 2] = torch.tensor([c, s, 0.0], dtype=torch.float64)    # fwd
            cam2lidar[:3, 3] = torch.tensor([0.5 * c, 0.5 * s, 1.5], dtype=torch.float64)
            if roll != 0.0 or pitch != 0.0:
                cr, sr, cp, sp = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch)
                rz = torch.tensor([[cr, -sr, 0.0], [sr, cr, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
                rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]], dtype=torch.float64)
                cam2lidar[:3, :3] = cam2lidar[:3, :3] @ rz @ rx
            lidar2img = (K @ torch.linalg.inv(cam2lidar)).numpy()
            mat = torch.Tensor(lidar2img)          # fp32, as the reference does
            inv = mat.inverse()
            rots[b, i] = inv[:3, :3]
            trans[b, i] = inv[:3, 3]
    return rots, trans


def pool_inputs(cfg: ViewConfig, batch: int = None, seed: int = 0, with_grad_out: bool = True):
    """depth (softmax over D) [B,N,D,fH,fW], feat [B,N,C,fH,fW], out_grad [B,C,Z,Y,X] fp32 (CPU)."""
    from .view_transform import gen_dx_bx
    B = cfg.batch if batch is None else batch
    g = torch.Generator().manual_seed(seed)
    depth = torch.randn(B, cfg.n_cams, cfg.D, cfg.fH, cfg.fW, generator=g).softmax(dim=2)
    feat = torch.randn(B, cfg.n_cams, cfg.channels, cfg.fH, cfg.fW, generator=g)
    out = [depth, feat]
    if with_grad_out:
        _, _, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        X, Y, Z = (int(v) for v in nx)
        out.append(torch.randn(B, cfg.channels, Z, Y, X, generator=g))
    return tuple(out)
