"""Patch-in for the reference plugin tree — nothing in the reference's sources is edited.

The reference's necks do `from projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool import bev_pool_v2`
(cam_stream_lss_bevpoolv2.py:17, ..._depthnet.py:18, rcfusion/...:18) and define `get_geometry`,
`voxel_pooling_prepare_v2`, `voxel_pooling_v2` and `get_voxels` as methods that read `self.dx / self.bx /
self.nx / self.frustum`. Three levels, each a drop-in for the one before:

  * `install_ext()`   registers `bev_pool_v2_ext` (this package's ctypes binding with the pybind module's two
                      entry points) so the reference's OWN `bev_pool.py` runs unmodified on the sm_100a kernels;
  * `install()`       registers this package's operator module under
                      `projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool` (and the v1 op under `...ops.bev_pool.bev_pool`)
                      — import it BEFORE the plugin package is imported;
  * `patch_lss_class(cls)`  swaps the methods of an (already imported, unmodified) LiftSplatShoot-like class:
                      `get_geometry` + `voxel_pooling_prepare_v2` -> CUDA kernels (the reference's `voxel_pooling_v2`
                      then runs as it is), the module-level name `bev_pool_v2` of the class's module -> ours, and
                      `get_voxels` -> the fully fused view transform (no `coor`, no rank arrays, no host sync).
"""
import sys
import types

from . import bev_pool as _bev_pool
from . import view_transform as _vt

REF_PKG = "projects.mmdet3d_plugin.ops.bev_pool_v2"
REF_MODULE = REF_PKG + ".bev_pool"
REF_EXT = REF_PKG + ".bev_pool_v2_ext"
REF_MODULE_V1 = "projects.mmdet3d_plugin.ops.bev_pool.bev_pool"   # the plugin __init__ imports this one (:20)


def install_ext():
    """Make `from . import bev_pool_v2_ext` (ops/bev_pool_v2/bev_pool.py:6) resolve to this package's binding."""
    from . import bev_pool_v2_ext as ext
    sys.modules[REF_EXT] = ext
    parent = sys.modules.get(REF_PKG)
    if parent is not None:
        parent.bev_pool_v2_ext = ext
    return ext


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
    parent = sys.modules.get(REF_PKG)
    if parent is not None:
        parent.bev_pool = mod
    # v1 op: `from .ops.bev_pool import *` in mmdet3d_plugin/__init__.py:20 needs bev_pool_ext built otherwise
    from . import bev_pool_v1 as _v1
    if REF_MODULE_V1 not in sys.modules or force:
        m1 = types.ModuleType(REF_MODULE_V1)
        m1.bev_pool, m1.QuickCumsumCuda, m1.__all__ = _v1.bev_pool, _v1.QuickCumsumCuda, ["bev_pool"]
        sys.modules[REF_MODULE_V1] = m1
    return mod


def fused_view_of(lss):
    """The LSSViewTransform that shares `lss`'s own grid constants and frustum Parameter. It is kept in the instance
    __dict__ (not registered as a submodule: the state_dict of the reference module must not change) and rebuilt
    when the module moved to another device or its dx / bx / nx / frustum were re-assigned."""
    cached = lss.__dict__.get("_bevpool_b200_view")
    key = (id(lss.frustum), lss.frustum.device, id(lss.dx), id(lss.bx), id(lss.nx))
    if cached is not None and cached[0] == key:
        return cached[1]
    view = _vt.LSSViewTransform.adopt(lss.frustum, lss.dx, lss.bx, lss.nx)
    lss.__dict__["_bevpool_b200_view"] = (key, view)
    return view


def _graphed_view_of(lss, depth, feat, rots, trans):
    """Per-signature cache of CUDA-graphed forms of `fused_view_of(lss)` (LSSViewTransform.graphed)."""
    view = fused_view_of(lss)
    import torch
    key = (tuple(depth.shape), tuple(feat.shape), depth.dtype, feat.dtype, depth.requires_grad, feat.requires_grad,
           torch.is_grad_enabled(), id(view))
    cache = lss.__dict__.setdefault("_bevpool_b200_graphs", {})
    fn = cache.get(key)
    if fn is None:
        if len(cache) >= 8:
            cache.clear()
        fn = cache[key] = view.graphed(depth, feat, rots, trans)
    return fn


def patch_lss_class(cls, fused=True, cuda_graph=False):
    """Swap the view-transform methods of a reference LSS class for the CUDA-backed ones (see module docstring).
    fused=False keeps the reference's own `get_voxels` (geometry -> prepare -> bev_pool_v2 call sequence).
    cuda_graph=True (opt-in, fixed shapes): `get_voxels` replays a CUDA graph of the fused view transform, forward and
    backward (one pair per input signature) — the step's launches and Python glue disappear; the returned BEV grid
    lives in a static buffer that the next call overwrites."""

    def voxel_pooling_prepare_v2(self, coor):
        return _vt.voxel_pooling_prepare_v2(coor, self.dx, self.bx, self.nx)

    def get_geometry(self, rots, trans, post_rots=None, post_trans=None, extra_rots=None, extra_trans=None):
        if any(v is not None for v in (post_rots, post_trans, extra_rots, extra_trans)):
            return cls._bevpool_b200_orig_get_geometry(self, rots, trans, post_rots, post_trans, extra_rots,
                                                       extra_trans)
        return _vt.get_geometry(self.frustum, rots, trans)

    def get_voxels(self, x, rots=None, trans=None, post_rots=None, post_trans=None, extra_rots=None, extra_trans=None):
        """cam_stream_lss_bevpoolv2.py:354-361 (…_depthnet.py:344-350 in both variants) as one fused pass: same
        inputs, same `(bev [B,C,Z,Y,X], depth [B,N,D,H,W])` result. One deliberate difference: when no frustum point
        falls inside the grid the reference prints a warning and returns `None` (its forward then fails in `s2c`);
        this path never synchronises with the host, so it returns an all-zero grid instead."""
        if any(v is not None for v in (post_rots, post_trans, extra_rots, extra_trans)):
            return cls._bevpool_b200_orig_get_voxels(self, x, rots, trans, post_rots, post_trans, extra_rots, extra_trans)
        feat, depth = self.get_cam_feats(x)
        if feat.shape[2] % 4:      # channel counts the fused kernels do not take: the reference's own sequence
            return self.voxel_pooling_v2(self.get_geometry(rots, trans), depth, feat), depth
        rots, trans = rots.float(), trans.float()
        if cls._bevpool_b200_cuda_graph and depth.is_cuda:
            return _graphed_view_of(self, depth, feat, rots, trans)(depth, feat, rots, trans), depth
        return fused_view_of(self)(depth, feat, rots, trans), depth

    if not hasattr(cls, "_bevpool_b200_orig_get_geometry"):
        cls._bevpool_b200_orig_get_geometry = cls.get_geometry
        cls._bevpool_b200_orig_prepare = cls.voxel_pooling_prepare_v2
        cls._bevpool_b200_orig_get_voxels = getattr(cls, "get_voxels", None)
    cls._bevpool_b200_cuda_graph = bool(cuda_graph)
    cls.voxel_pooling_prepare_v2 = voxel_pooling_prepare_v2
    cls.get_geometry = get_geometry
    if fused and cls._bevpool_b200_orig_get_voxels is not None:
        cls.get_voxels = get_voxels
    elif cls._bevpool_b200_orig_get_voxels is not None:
        cls.get_voxels = cls._bevpool_b200_orig_get_voxels
    # the class's module bound the reference `bev_pool_v2` at import time: point it at ours, so the unmodified
    # `voxel_pooling_v2` reaches the new kernels even if install() came after the import
    mod = sys.modules.get(cls.__module__)
    if mod is not None and hasattr(mod, "bev_pool_v2"):
        if not hasattr(mod, "_bevpool_b200_orig_bev_pool_v2"):
            mod._bevpool_b200_orig_bev_pool_v2 = mod.bev_pool_v2
        mod.bev_pool_v2 = _bev_pool.bev_pool_v2
    return cls


def unpatch_lss_class(cls):
    """Undo patch_lss_class (tests; A/B runs against the reference's own torch-op methods)."""
    if hasattr(cls, "_bevpool_b200_orig_get_geometry"):
        cls.get_geometry = cls._bevpool_b200_orig_get_geometry
        cls.voxel_pooling_prepare_v2 = cls._bevpool_b200_orig_prepare
        if cls._bevpool_b200_orig_get_voxels is not None:
            cls.get_voxels = cls._bevpool_b200_orig_get_voxels
        del cls._bevpool_b200_orig_get_geometry, cls._bevpool_b200_orig_prepare, cls._bevpool_b200_orig_get_voxels
        if hasattr(cls, "_bevpool_b200_cuda_graph"):
            del cls._bevpool_b200_cuda_graph
    mod = sys.modules.get(cls.__module__)
    if mod is not None and hasattr(mod, "_bevpool_b200_orig_bev_pool_v2"):
        mod.bev_pool_v2 = mod._bevpool_b200_orig_bev_pool_v2
        del mod._bevpool_b200_orig_bev_pool_v2
    return cls
