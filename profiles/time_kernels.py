"""Time the dense forward / backward kernels alone (CUDA events, rotating buffer sets)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
lib = pkg._lib.load()
cfg = pkg.synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
NS = int(os.environ.get("NS", 4))
dev = torch.device("cuda:0")
bp = pkg.bev_pool
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
X, Y, Z = (int(v) for v in view.nx)
C, D, H, W, N = cfg.channels, view.D, view.fH, view.fW, cfg.n_cams
V = B * X * Y * Z
dt = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=s)
    pr = pkg.view_transform._prepare_device(None, view.frustum, rots.to(dev), trans.to(dev), B, N, D, H, W, view.dx, view.bx, view.nx, dev, want_intervals=False)
    f = feat.to(dev, dt)
    fcl = f.new_empty((B * N, H, W, C))
    bp._launch_transpose(f, fcl, B * N, C, H * W, True)
    sets.append(dict(pr=pr, depth=depth.to(dev, dt), fcl=fcl, f=f, out=torch.empty((B, C, Z, Y, X), dtype=dt, device=dev),
                     tab=bp._launch_voxel_table(pr.rb, pr.p0, pr.counts, V), og=torch.randn((B, Z, Y, X, C), device=dev).to(dt),
                     dg=torch.empty((B, N, D, H, W), dtype=dt, device=dev), fg=torch.empty_like(f)))
code = bp._dtype_code(sets[0]["fcl"])
st = torch.cuda.current_stream().cuda_stream
def fwd(i):
    s = sets[i % NS]; p = s["pr"]
    bp._launch_forward_dense(s["depth"], s["fcl"], s["out"], p.rd, None, p.rb, s["tab"], B, Z * Y, X, pkg._lib.LAYOUT_BCZYX, dhw=D * H * W, hw=H * W, n_points=p.p0, counts_dev=p.counts)
def bwd(i):
    s = sets[i % NS]; p = s["pr"]
    lib.bevpool_v2_backward_dense(s["og"].data_ptr(), s["dg"].data_ptr(), s["fg"].data_ptr(), s["depth"].data_ptr(), s["fcl"].data_ptr(),
                                  p.point_rank.data_ptr(), p.bn, p.d, p.h, p.w, C, 1, int(os.environ.get('HINT', 1 if Z == 1 else 0)), code, st)
def timeit(fn, reps=40):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(5 + i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
which = os.environ.get("WHICH", "fwd,bwd").split(",")
res = {k: round(timeit({"fwd": fwd, "bwd": bwd}[k]), 1) for k in which}
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("BEVPOOL_")}, "us": res, "P": int(sets[0]["pr"].counts[0])}))
