"""Where the eager reference call sequence (get_geometry -> voxel_pooling_prepare_v2 -> bev_pool_v2 -> backward) spends
its time: host wall clock per stage with a synchronize after each (so launch overhead and the D2H read are included)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda:0")
cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
B = cfg.batch
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=0)
depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=0)
rots, trans, gout = rots.to(dev), trans.to(dev), gout.to(dev)
depth, feat = depth.to(dev).requires_grad_(), feat.to(dev).requires_grad_()
X, Y, Z = (int(v) for v in view.nx)
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)


R = 30
for it in range(R + 5):
    if it == 5:
        acc.clear()
    depth.grad = feat.grad = None
    t = time.perf_counter(); coor = view.get_geometry(rots, trans); tick("get_geometry", t)
    t = time.perf_counter(); ranks = view.voxel_pooling_prepare_v2(coor); tick("voxel_pooling_prepare_v2", t)
    t = time.perf_counter(); f = feat.permute(0, 1, 3, 4, 2).contiguous(); tick("feat.permute.contiguous (torch)", t)
    t = time.perf_counter()
    bev = pkg.bev_pool_v2(depth, f, ranks[1], ranks[2], ranks[0], (B, Z, Y, X, cfg.channels), ranks[3], ranks[4]); tick("bev_pool_v2 forward", t)
    t = time.perf_counter(); bev.backward(gout); tick("backward", t)
print(json.dumps({k: round(v / R * 1e6, 1) for k, v in acc.items()}))
