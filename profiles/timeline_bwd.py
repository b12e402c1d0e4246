"""Per-CTA timeline of the joint backward kernel (debug hook)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WHICH", "bwd")
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(__file__), "time_kernels.py")).read().split("which = os.environ")[0])
lib2 = ctypes.CDLL(pkg._lib.library_path())
bwd(0); torch.cuda.synchronize()
lib2.bevpool_debug_fwd_timeline(1, None, 0)
bwd(1); torch.cuda.synchronize()
n = 1152
buf = np.zeros(8 * n, dtype=np.uint64)
lib2.bevpool_debug_fwd_timeline(0, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), n)
t = buf.reshape(n, 8).astype(np.int64)
t0 = t[:, 2].min()
dur = (t[:, 3] - t[:, 2]) / 1e3
print("kernel span us", (t[:, 3].max() - t0) / 1e3, "ctas", n)
print("cta dur us: mean %.1f p50 %.1f p90 %.1f max %.1f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), dur.max()))
ph = np.stack([t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 3] - t[:, 5]], 1) / 1e3
print("phase us (stage, process, writeout): mean", ph.mean(0).round(2), "p90", np.percentile(ph, 90, axis=0).round(2))
sm = t[:, 1]
last = np.array([((t[sm == s, 3].max() - t0) / 1e3) if (sm == s).any() else 0 for s in range(148)])
cnt = np.array([(sm == s).sum() for s in range(148)])
print("per-SM ctas min/max", cnt.min(), cnt.max(), "last-end min/mean/max", last.min(), last.mean(), last.max())
starts = np.sort((t[:, 2] - t0) / 1e3)
print("cta start times us: p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(starts, [10, 50, 90, 100])))
