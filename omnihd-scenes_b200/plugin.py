"""Patch-in for the reference plugin tree.

The reference's necks do `from projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool import bev_pool_v2`
(cam_stream_lss_bevpoolv2.py:17, ..._depthnet.py:18, rcfusion/...:18) and define
`voxel_pooling_prepare_v2` as a method that reads `self.dx / self.bx / self.nx`.
`install()` makes both resolve to this package without touching the reference sources:

  * registers a module object under `projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool`
    (and a stub `bev_pool_v2_ext`) in sys.modules exporting `bev_pool_v2`, `TRTBEVPoolv2`,
    `QuickCumsumCuda` — import it BEFORE the plugin package is imported;
  * `patch_lss_class(cls)` replaces `voxel_pooling_prepare_v2` / `get_geometry` on an already
    imported LiftSplatShoot-like class with shims that call the sm_100a kernels.
"""
import sys
import types

from . import bev_pool as _bev_pool
from . import view_transform as _vt

REF_MODULE = "projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool"
REF_MODULE_V1 = "projects.mmdet3d_plugin.ops.bev_pool.bev_pool"   # the plugin __init__ imports this one (:20)


def install(force=False):
    """Make `import projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool` resolve to this package."""
    if REF_MODULE in sys.modules and not force:
        mod = sys.modules[REF_MODULE]
    else:
        mod = types.ModuleType(REF_MODULE)
        sys.modules[REF_MODULE] = mod
    mod.bev_pool_v2 = _bev_pool.bev_pool_v2
    mod.TRTBEVPoolv2 = _bev_pool.TRTBEVPoolv2
    mod.QuickCumsumCuda = _bev_pool.QuickCumsumCuda
    mod.__all__ = ['bev_pool_v2', 'TRTBEVPoolv2']
    parent = sys.modules.get("projects.mmdet3d_plugin.ops.bev_pool_v2")
    if parent is not None:
        parent.bev_pool = mod
    # v1 op: `from .ops.bev_pool import *` in mmdet3d_plugin/__init__.py:20 needs bev_pool_ext built otherwise
    from . import bev_pool_v1 as _v1
    if REF_MODULE_V1 not in sys.modules or force:
        m1 = types.ModuleType(REF_MODULE_V1)
        m1.bev_pool, m1.QuickCumsumCuda, m1.__all__ = _v1.bev_pool, _v1.QuickCumsumCuda, ["bev_pool"]
        sys.modules[REF_MODULE_V1] = m1
    return mod


def patch_lss_class(cls):
    """Swap the two geometry/prepare methods of a reference LSS class for the CUDA-backed ones."""

    def voxel_pooling_prepare_v2(self, coor):
        return _vt.voxel_pooling_prepare_v2(coor, self.dx, self.bx, self.nx)

    def get_geometry(self, rots, trans, post_rots=None, post_trans=None, extra_rots=None, extra_trans=None):
        if any(v is not None for v in (post_rots, post_trans, extra_rots, extra_trans)):
            return cls._bevpool_b200_orig_get_geometry(self, rots, trans, post_rots, post_trans, extra_rots,
                                                       extra_trans)
        return _vt.get_geometry(self.frustum, rots, trans)

    if not hasattr(cls, "_bevpool_b200_orig_get_geometry"):
        cls._bevpool_b200_orig_get_geometry = cls.get_geometry
        cls._bevpool_b200_orig_prepare = cls.voxel_pooling_prepare_v2
    cls.voxel_pooling_prepare_v2 = voxel_pooling_prepare_v2
    cls.get_geometry = get_geometry
    return cls
