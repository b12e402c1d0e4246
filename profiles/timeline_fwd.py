"""Per-CTA timeline of the forward tile kernel (debug hook): who is on the critical path?"""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WHICH", "fwd")
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(__file__), "time_kernels.py")).read().split("which = os.environ")[0])
lib2 = ctypes.CDLL(pkg._lib.library_path())
fwd(0); torch.cuda.synchronize()
lib2.bevpool_debug_fwd_timeline(1, None, 0)
fwd(1); torch.cuda.synchronize()
n = 1024
buf = np.zeros(8 * n, dtype=np.uint64)
lib2.bevpool_debug_fwd_timeline(0, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), n)
t = buf.reshape(n, 8).astype(np.int64)
t0 = t[:, 2].min()
dur = (t[:, 3] - t[:, 2]) / 1e3
print("kernel span us", (t[:, 3].max() - t0) / 1e3, "tiles", n)
print("tile dur us: mean %.1f p50 %.1f p90 %.1f max %.1f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), dur.max()))
pts = t[:, 0]
print("corr pts-dur", np.corrcoef(pts, dur)[0, 1], "ns per point (sum dur/sum pts)", dur.sum() * 1e3 / pts.sum())
for lo, hi in [(0, 100), (100, 500), (500, 1500), (1500, 3000), (3000, 10000)]:
    m = (pts >= lo) & (pts < hi)
    if m.any(): print(f"pts[{lo},{hi}): n={m.sum()} mean dur {dur[m].mean():.1f} us")
# per-SM busy time
sm = t[:, 1]
busy = np.array([dur[sm == s].sum() for s in range(148)])
cnt = np.array([(sm == s).sum() for s in range(148)])
last = np.array([((t[sm == s, 3].max() - t0) / 1e3) if (sm == s).any() else 0 for s in range(148)])
print("per-SM: ctas min/max", cnt.min(), cnt.max(), "sum-dur min/mean/max", busy.min(), busy.mean(), busy.max(), "last-end min/mean/max", last.min(), last.mean(), last.max())
late = np.argsort(-t[:, 3])[:8]
for i in late: print("late tile", i, "pts", pts[i], "start", (t[i, 2] - t0) / 1e3, "end", (t[i, 3] - t0) / 1e3, "sm", sm[i])

ph = np.stack([t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5], t[:, 7] - t[:, 6], t[:, 3] - t[:, 7]], 1) / 1e3
print("phase us (rowranges, zero, gather, fixup, writeout): mean", ph.mean(0).round(2), "light tiles (<300 pts):", ph[pts < 300].mean(0).round(2),
      "heavy (>3000):", ph[pts > 3000].mean(0).round(2))
