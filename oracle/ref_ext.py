"""ORACLE — TEST / BENCH-COMPARATOR INFRASTRUCTURE ONLY.

`bev_pool_v2_ext` on the REFERENCE'S OWN CUDA kernels: oracle/_ref/libref_bevpool_v2.so is the unmodified
ops/bev_pool_v2/src/bev_pool_cuda.cu compiled for sm_100a (oracle/Makefile, oracle/ref_shim.cu). This object exposes
the two entry points of the reference's pybind module (src/bev_pool.cpp:30-57, :74-104: same argument order, c read
from feat.size(4) / out_grad.size(4), launches on the legacy default stream as bev_pool_cuda.cu:125-140 does), so the
reference's unmodified `bev_pool.py` runs on the reference's unmodified kernels on this GPU — "the existing kernel to
beat" (SURVEY.md 8(d)). Used by tests and by bench.py's `variants.reference_cuda_ext`; never by the product.
"""
import ctypes
import os
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_bevpool_v2.so")


def available():
    return os.path.exists(LIB)


def load():
    lib = ctypes.CDLL(LIB)
    p = ctypes.c_void_p
    lib.ref_bev_pool_v2_fwd.restype = ctypes.c_int
    lib.ref_bev_pool_v2_fwd.argtypes = [ctypes.c_int, ctypes.c_int] + [p] * 8
    lib.ref_bev_pool_v2_bwd.restype = ctypes.c_int
    lib.ref_bev_pool_v2_bwd.argtypes = [ctypes.c_int, ctypes.c_int] + [p] * 10

    def bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts):
        rc = lib.ref_bev_pool_v2_fwd(feat.size(4), interval_lengths.size(0), depth.data_ptr(), feat.data_ptr(),
                                     ranks_depth.data_ptr(), ranks_feat.data_ptr(), ranks_bev.data_ptr(),
                                     interval_starts.data_ptr(), interval_lengths.data_ptr(), out.data_ptr())
        if rc:
            raise RuntimeError(f"reference bev_pool_v2 kernel failed: cudaError {rc}")

    def bev_pool_v2_backward(out_grad, depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                             interval_lengths, interval_starts):
        rc = lib.ref_bev_pool_v2_bwd(out_grad.size(4), interval_lengths.size(0), out_grad.data_ptr(), depth.data_ptr(),
                                     feat.data_ptr(), ranks_depth.data_ptr(), ranks_feat.data_ptr(), ranks_bev.data_ptr(),
                                     interval_starts.data_ptr(), interval_lengths.data_ptr(), depth_grad.data_ptr(),
                                     feat_grad.data_ptr())
        if rc:
            raise RuntimeError(f"reference bev_pool_v2_grad kernel failed: cudaError {rc}")

    ext = types.ModuleType("bev_pool_v2_ext")
    ext.bev_pool_v2_forward, ext.bev_pool_v2_backward, ext._lib = bev_pool_v2_forward, bev_pool_v2_backward, lib
    return ext
