"""Per-CTA phase timeline of the column backward (needs a library built with -DBEVPOOL_TIMELINE, see README)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WHICH", "bwd")
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(__file__), "time_kernels.py")).read().split("which = os.environ")[0])
lib2 = ctypes.CDLL(pkg._lib.library_path())
for i in range(6):
    bwd(i)
torch.cuda.synchronize()
n = int(os.environ.get("NCTA", 288))
buf = np.zeros(8 * n, dtype=np.uint64)
lib2.bevpool_debug_col_timeline(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), n)
t = buf.reshape(n, 8).astype(np.int64)
t0 = t[:, 0].min()
print("kernel span us (first CTA start -> last CTA end)", (t[:, 6].max() - t0) / 1e3, "ctas", n)
names = ["staging(0->1)", "fv loads+compaction+issue(1->2)", "first chunk wait(2->3)", "item loop warp0(3->4)", "barrier wait(4->5)", "write-out(5->6)"]
ph = np.stack([t[:, k + 1] - t[:, k] for k in range(6)], 1) / 1e3
for k, nm in enumerate(names):
    print(f"  {nm:36s} mean {ph[:, k].mean():6.2f}  p10 {np.percentile(ph[:, k], 10):6.2f}  p90 {np.percentile(ph[:, k], 90):6.2f}  max {ph[:, k].max():6.2f}")
dur = (t[:, 6] - t[:, 0]) / 1e3
print("cta duration us: mean %.1f p10 %.1f p90 %.1f max %.1f" % (dur.mean(), np.percentile(dur, 10), np.percentile(dur, 90), dur.max()))
st = (t[:, 0] - t0) / 1e3
print("cta start us: p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(st, [50, 90, 100])))
items = t[:, 7] >> 16
sm = t[:, 7] & 0x7fff
loop = ph[:, 3]
print("items of warp 0: min %d mean %.1f max %d; corr(loop time, items) = %.2f; loop us per item: mean %.3f p10 %.3f p90 %.3f" % (
    items.min(), items.mean(), items.max(), np.corrcoef(loop, items)[0, 1], (loop / np.maximum(items, 1)).mean(),
    np.percentile(loop / np.maximum(items, 1), 10), np.percentile(loop / np.maximum(items, 1), 90)))
for lo, hi in ((0, 30), (30, 40), (40, 50), (50, 60), (60, 70)):
    m = (items >= lo) & (items < hi)
    if m.any():
        print(f"  items {lo}-{hi}: n={m.sum():3d} loop {loop[m].mean():5.1f} us, cta {((t[m, 6] - t[m, 0]) / 1e3).mean():5.1f} us")
cnt = np.bincount(sm, minlength=148)
print("CTAs per SM: min %d max %d; SMs with 2: %d" % (cnt.min(), cnt.max(), (cnt == 2).sum()))
one = np.isin(sm, np.where(cnt == 1)[0])
print("duration on SMs with 1 CTA: %.1f us (n=%d); with 2 CTAs: %.1f us" % (dur[one].mean() if one.any() else 0, one.sum(), dur[~one].mean()))
