"""ORACLE — TEST INFRASTRUCTURE ONLY. Imports the reference's UNMODIFIED Python files of the path.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's reference / comparator legs may use this module; the
product package never does (tests/test_host_cpu.py::test_product_never_imports_oracle).

Source of the files, in this order:
  1. oracle/_ref/py/   — staged by `make -C oracle refpy` from /root/reference (git-ignored; it travels to the GPU
                         box with the snapshot, where /root/reference does not exist);
  2. /root/reference   — the read-only reference tree (build container only).

mmcv / mmdet / mmdet3d / matplotlib are not installed in this image: they are stubbed in sys.modules with permissive
modules (any attribute is a pass-through decorator factory) and the package chain `projects.mmdet3d_plugin...` is
registered as bare namespace packages so the heavy plugin `__init__` (which needs mmdet3d) never runs
(SURVEY.md Appendix A.2). The reference code that then executes is its own, byte for byte.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "py")
REFERENCE = "/root/reference"

LSS_MODULES = {
    "bevfusion": "projects.mmdet3d_plugin.bevfusion.detectors.cam_stream_lss_bevpoolv2",
    "bevfusion_depth": "projects.mmdet3d_plugin.bevfusion.detectors.cam_stream_lss_bevpoolv2_depthnet",
    "rcfusion_depth": "projects.mmdet3d_plugin.rcfusion.detectors.cam_stream_lss_bevpoolv2_depthnet",
}
OP_MODULE = "projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool"
OP_EXT = "projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool_v2_ext"


class _Permissive(types.ModuleType):
    """Any attribute is a pass-through decorator factory / dummy callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def factory(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f
        return factory


def ref_root():
    """Directory holding `projects/mmdet3d_plugin/...`, or None when neither source exists."""
    for root in (STAGED, REFERENCE):
        if os.path.isfile(os.path.join(root, "projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py")):
            return root
    return None


def _register_chain(root):
    for name in ["mmcv", "mmcv.runner", "mmcv.cnn", "mmdet", "mmdet.models", "mmdet.models.backbones",
                 "mmdet.models.backbones.resnet", "mmdet3d", "mmdet3d.models", "mmdet3d.models.fusion_layers",
                 "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "projects.mmdet3d_plugin.utils", "projects.mmdet3d_plugin.utils.gaussian"]:
        if name not in sys.modules:
            sys.modules[name] = _Permissive(name)
    for mod in ["projects", "projects.mmdet3d_plugin", "projects.mmdet3d_plugin.ops",
                "projects.mmdet3d_plugin.ops.bev_pool_v2", "projects.mmdet3d_plugin.ops.bev_pool",
                "projects.mmdet3d_plugin.bevfusion", "projects.mmdet3d_plugin.bevfusion.detectors",
                "projects.mmdet3d_plugin.rcfusion", "projects.mmdet3d_plugin.rcfusion.detectors"]:
        m = sys.modules.get(mod)
        if m is None or not hasattr(m, "__path__"):
            m = types.ModuleType(mod)
            sys.modules[mod] = m
        m.__path__ = [os.path.join(root, *mod.split("."))]


def import_reference_op(ext=None):
    """The reference's own ops/bev_pool_v2/bev_pool.py (QuickCumsumCuda, bev_pool_v2, TRTBEVPoolv2), executed
    unmodified. `ext` is the object its `from . import bev_pool_v2_ext` resolves to (a module exposing
    bev_pool_v2_forward / bev_pool_v2_backward); None = a permissive stub (the functions are never called)."""
    root = ref_root()
    if root is None:
        raise ImportError("reference Python files are neither staged (make -C oracle refpy) nor under /root/reference")
    _register_chain(root)
    ext = ext if ext is not None else _Permissive("bev_pool_v2_ext")
    sys.modules[OP_EXT] = ext
    sys.modules["projects.mmdet3d_plugin.ops.bev_pool_v2"].bev_pool_v2_ext = ext
    sys.modules.pop(OP_MODULE, None)
    return importlib.import_module(OP_MODULE)


def import_reference_lss(variant="bevfusion", fresh=True):
    """The reference's LSS neck module (`LiftSplatShoot` / `LiftSplatShoot_Depth`, gen_dx_bx, QuickCumsum, CamEncode).
    Whatever is registered under the operator module path at that moment is what its
    `from ...ops.bev_pool_v2.bev_pool import bev_pool_v2` binds (register the package's module with plugin.install()
    first, or the reference's own with import_reference_op())."""
    root = ref_root()
    if root is None:
        raise ImportError("reference Python files are neither staged (make -C oracle refpy) nor under /root/reference")
    _register_chain(root)
    if OP_MODULE not in sys.modules:
        import_reference_op()
    name = LSS_MODULES[variant]
    if fresh:
        sys.modules.pop(name, None)
    return importlib.import_module(name)


def make_reference_lss(ref, final_dim, downsample, dbound, xb, yb, zb, inputC=8, camC=8):
    """A `LiftSplatShoot` (light variant) with per-axis grid constants: the class forces one scalar step for x, y, z
    (cam_stream_lss_bevpoolv2.py:163-168), so dx / bx / nx are overwritten with its own gen_dx_bx output."""
    lss = ref.LiftSplatShoot(lss=False, final_dim=final_dim, camera_depth_range=list(dbound),
                             pc_range=[xb[0], yb[0], zb[0], xb[1], yb[1], zb[1]],
                             downsample=downsample, grid=xb[2], inputC=inputC, camC=camC)
    lss.dx, lss.bx, lss.nx = ref.gen_dx_bx(list(xb), list(yb), list(zb))
    return lss
