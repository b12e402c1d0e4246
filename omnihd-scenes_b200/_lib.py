"""ctypes binding of csrc/libbevpool_b200.so (include/bevpool_b200.h).

No torch types cross this boundary: tensors are passed as `data_ptr()` integers
and the current CUDA stream as a raw handle. If the library is missing or lacks
a symbol the import of the op fails loudly — there is no CPU or eager fallback.
"""
import ctypes
import os

from . import build as _build

_LIB = None

c_void_p, c_int, c_i64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

F32, BF16 = 0, 1
LAYOUT_BZYXC, LAYOUT_BCZYX = 0, 1


class GridT(ctypes.Structure):
    """bevpool_grid_t"""
    _fields_ = [("b", ctypes.c_int32), ("n", ctypes.c_int32), ("d", ctypes.c_int32),
                ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("nx", ctypes.c_int32 * 3), ("lo", ctypes.c_float * 3), ("dx", ctypes.c_float * 3)]


# name -> (restype, argtypes); must list every symbol include/bevpool_b200.h declares
SIGNATURES = {
    "bevpool_b200_abi_version": (c_int, []),
    "bevpool_b200_strerror": (ctypes.c_char_p, [c_int]),
    "bevpool_b200_launch_count": (c_i64, []),
    "bevpool_v2_forward": (c_int, [c_void_p] * 8 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "bevpool_v2_backward": (c_int, [c_void_p] * 10 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "bevpool_v2_backward_regroup_workspace_bytes": (c_size_t, [c_i64]),
    "bevpool_v2_backward_regroup": (c_int, [c_void_p] * 3 + [c_i64, ctypes.c_int32] + [c_void_p] * 6 +
                                    [c_void_p, c_size_t, c_void_p]),
    "bevpool_geometry": (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_void_p]),
    "bevpool_prepare_v2_workspace_bytes": (c_size_t, [ctypes.POINTER(GridT)]),
    "bevpool_prepare_v2": (c_int, [c_void_p] * 4 + [ctypes.POINTER(GridT)] + [c_void_p] * 7 +
                           [c_void_p, c_size_t, c_void_p]),
    "bevpool_prepare_v2_counts": (c_int, [c_void_p] * 4 + [ctypes.POINTER(GridT)] + [c_void_p] * 7 +
                                  [c_void_p, c_size_t, c_void_p, ctypes.POINTER(ctypes.c_int32 * 2)]),
    "bevpool_voxel_table": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p]),
    "bevpool_v2_forward_dense_scratch_bytes": (c_size_t, [c_i64, c_i64, c_int, c_int, c_int]),
    "bevpool_v2_forward_dense": (c_int, [c_void_p] * 7 + [c_i64, c_void_p, c_int, c_i64, c_i64, c_int, c_int, c_int,
                                                          c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "bevpool_v2_backward_dense": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "bevpool_v1_forward": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "bevpool_v1_backward": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "bevpool_view_forward_scratch_bytes": (ctypes.c_size_t, [c_i64, c_int, c_int, c_int]),
    "bevpool_view_forward": (c_int, [c_void_p] * 5 + [ctypes.POINTER(GridT), c_int, c_void_p, c_int, c_void_p, c_i64, c_i64,
                                     c_int, c_int, c_void_p, ctypes.c_size_t, c_void_p]),
    "bevpool_pillar_scatter_forward": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "bevpool_pillar_scatter_backward": (c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    "bevpool_channel_avg_max_forward": (c_int, [c_void_p] * 3 + [c_int, c_int, c_i64, c_int, c_void_p]),
    "bevpool_channel_avg_max_backward": (c_int, [c_void_p] * 3 + [c_int, c_int, c_i64, c_int, c_void_p]),
    "bevpool_gate_concat_forward": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_i64, c_int, c_void_p]),
    "bevpool_gate_concat_backward": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_i64, c_int, c_void_p]),
    "bevpool_lift_forward": (c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    "bevpool_lift_backward": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "bevpool_grid_transpose": (c_int, [c_void_p, c_void_p, c_int, c_int, c_i64, c_int, c_int, c_void_p]),
}


class BevPoolError(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load():
    """Load (building first if nvcc is present and sources are newer) and type every symbol."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path) or (os.path.exists(_build.NVCC) and _build._stale()):
        if not os.path.exists(_build.NVCC):
            raise BevPoolError(
                f"{path} is missing and nvcc is not available: the sm_100a extension must be built "
                "(python __graft_entry__.py). There is no CPU fallback.")
        _build.build_native()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise BevPoolError(f"{path} does not export {name}") from e
        fn.restype, fn.argtypes = res, args
    if lib.bevpool_b200_abi_version() != 1:
        raise BevPoolError("libbevpool_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().bevpool_b200_strerror(code).decode()
        raise BevPoolError(f"{what} failed: {msg} (code {code})")


def launch_count():
    return int(load().bevpool_b200_launch_count())
