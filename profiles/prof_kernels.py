"""Run the dense forward / backward / transpose / prepare kernels a few times on cfg 2 (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=0)
depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=0)
dt = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
d = depth.to(dev, dt).requires_grad_()
f = feat.to(dev, dt).requires_grad_()
g = gout.to(dev, dt)
for _ in range(3):
    d.grad = f.grad = None
    bev = view(d, f, rots.to(dev), trans.to(dev))
    bev.backward(g)
torch.cuda.synchronize()
print("ok", bev.shape)
