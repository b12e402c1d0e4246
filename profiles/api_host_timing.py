"""Wall time per stage of the eager reference API sequence (no profiler): host time spent inside each call, with the
device kept busy by the previous step as in a training loop. BEVPOOL_LATE_COUNTS=1 gives the round-1 read-back."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
B = cfg.batch
NS = 4
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, seed=s)
    sets.append((rots.to(dev), trans.to(dev), gout.to(dev), depth.to(dev).requires_grad_(), feat.to(dev).requires_grad_()))
T = [0.0] * 5
pc = time.perf_counter


def step(i, rec):
    rots, trans, gout, d, f = sets[i % NS]
    d.grad = f.grad = None
    t0 = pc()
    coor = view.get_geometry(rots, trans)
    t1 = pc()
    ranks = view.voxel_pooling_prepare_v2(coor)
    t2 = pc()
    rb, rd, rf, st, ln = ranks
    X, Y, Z = (int(v) for v in view.nx)
    bev = pkg.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, (B, Z, Y, X, f.shape[2]), st, ln)
    t3 = pc()
    bev.backward(gout)
    t4 = pc()
    if rec:
        for k, (a, b) in enumerate(((t0, t1), (t1, t2), (t2, t3), (t3, t4))):
            T[k] += b - a


for mode in ("early", "late"):
    if mode == "late":
        os.environ["BEVPOOL_LATE_COUNTS"] = "1"
    for i in range(12):
        step(i, False)
    torch.cuda.synchronize()
    T = [0.0] * 5
    n = 200
    t0 = pc()
    for i in range(n):
        step(i, True)
    torch.cuda.synchronize()
    tot = (pc() - t0) / n * 1e6
    print(f"{mode}: {tot:.1f} us/step wall | host us: get_geometry {T[0]/n*1e6:.1f}, prepare (incl. wait) {T[1]/n*1e6:.1f}, "
          f"bev_pool_v2 fwd {T[2]/n*1e6:.1f}, backward {T[3]/n*1e6:.1f}")
