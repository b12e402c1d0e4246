"""Per-CTA phase timeline of the scatter forward (needs csrc/pool_scatter.cu compiled with -DBEVPOOL_TIMELINE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
sets = []
for s in range(4):
    rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=s)
    depth, feat, _ = pkg.synthetic.pool_inputs(cfg, batch=B, seed=s)
    sets.append((depth.to(dev), feat.to(dev), rots.to(dev), trans.to(dev)))
with torch.no_grad():
    for i in range(6):
        view(*sets[i % 4])
torch.cuda.synchronize()
lib2 = ctypes.CDLL(pkg._lib.library_path())
n = 4096
buf = np.zeros(8 * n, dtype=np.uint64)
lib2.bevpool_debug_scatter_timeline(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), n)
t = buf.reshape(n, 8).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print("CTAs", len(t), "kernel span us", (t[:, 6].max() - t0) / 1e3)
names = ["issue the pass's loads (0->1)", "loads land + geometry + ranks (1->2)", "summaries + smem stores (2->3)", "barrier (3->4)",
         "feature rows land (4->5)", "walk + REDs (5->6)"]
ph = np.stack([t[:, k + 1] - t[:, k] for k in range(6)], 1) / 1e3
for k, nm in enumerate(names):
    print(f"  {nm:40s} mean {ph[:, k].mean():6.2f}  p10 {np.percentile(ph[:, k], 10):6.2f}  p90 {np.percentile(ph[:, k], 90):6.2f}  max {ph[:, k].max():6.2f}")
dur = (t[:, 6] - t[:, 0]) / 1e3
print("cta duration us: mean %.2f p10 %.2f p90 %.2f max %.2f" % (dur.mean(), np.percentile(dur, 10), np.percentile(dur, 90), dur.max()))
st = np.sort((t[:, 0] - t0) / 1e3)
print("cta start us: p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(st, [10, 50, 90, 100])))
sm = t[:, 7]
busy = np.array([dur[sm == s].sum() for s in range(148)])
print("sum of CTA durations per SM: mean %.1f max %.1f (3 slots: /3 = %.1f us of the span)" % (busy.mean(), busy.max(), busy.mean() / 3))
