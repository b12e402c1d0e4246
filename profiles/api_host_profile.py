"""cProfile of the reference API sequence (get_geometry -> voxel_pooling_prepare_v2 -> bev_pool_v2 -> backward), eager:
where the HOST time of that path goes (it is launch / Python bound, not kernel bound)."""
import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
B = cfg.batch
rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=0)
depth, feat, gout = pkg.synthetic.pool_inputs(cfg, seed=0)
rots, trans, gout = rots.to(dev), trans.to(dev), gout.to(dev)
d, f = depth.to(dev).requires_grad_(), feat.to(dev).requires_grad_()


def step():
    d.grad = f.grad = None
    bev = view.voxel_pooling_v2(view.get_geometry(rots, trans), d, f)
    bev.backward(gout)


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
torch.cuda.synchronize()
print("us per step", (time.perf_counter() - t0) / 200 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:45]))
