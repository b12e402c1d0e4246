"""omnihd-scenes_b200 — B200-native (sm_100a) camera->BEV view transform.

Drop-in for the hot path of TJRadarLab/OmniHD-Scenes' mmdet3d plugin:
`voxel_pooling_prepare_v2(coor)` and `bev_pool_v2(...)` (+ autograd backward),
frustum/geometry and the LSS view-transform shim around them. Host code is
Python/PyTorch; all device work is hand-written CUDA reached through the C-ABI
library `csrc/libbevpool_b200.so` (include/bevpool_b200.h). There is no CPU
fallback: a missing library raises at first use.
"""
from . import build, _lib, synthetic            # noqa: F401
from . import view_transform, bev_pool, bev_pool_v1, lift, pillar_scatter, cross_modal, plugin, sharding   # noqa: F401
from .lift import get_depth_feat, get_depth_dist   # noqa: F401
from .bev_pool import bev_pool_v2, TRTBEVPoolv2, QuickCumsumCuda   # noqa: F401
from .view_transform import (gen_dx_bx, create_frustum, get_geometry,       # noqa: F401
                             voxel_pooling_prepare_v2, LSSViewTransform)

__all__ = ["bev_pool_v2", "TRTBEVPoolv2", "QuickCumsumCuda", "gen_dx_bx", "create_frustum",
           "get_geometry", "voxel_pooling_prepare_v2", "LSSViewTransform"]
