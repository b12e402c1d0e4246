"""Device time of bevpool_view_forward with in-kernel geometry vs precomputed point_rank (what the exact geometry costs)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
lib = pkg._lib.load()
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"
cfg = pkg.synthetic.CONFIGS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
vt, bp = pkg.view_transform, pkg.bev_pool
X, Y, Z = (int(v) for v in view.nx)
C, D, H, W, N = cfg.channels, view.D, view.fH, view.fW, cfg.n_cams
NS = 4
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=s)
    depth, feat, _ = pkg.synthetic.pool_inputs(cfg, batch=B, seed=s)
    fcl = feat.to(dev).permute(0, 1, 3, 4, 2).contiguous()
    sets.append(dict(rots=rots.to(dev), trans=trans.to(dev), depth=depth.to(dev), fcl=fcl,
                     prk=torch.empty(B * N * D * H * W, dtype=torch.int32, device=dev),
                     acc=torch.empty((B, Z, Y, X, C), device=dev)))
g = vt._grid_struct(B, N, D, H, W, view.dx, view.bx, view.nx)
st = torch.cuda.current_stream().cuda_stream


def call(s, geom):
    lib.bevpool_view_forward(s["depth"].data_ptr(), s["fcl"].data_ptr(), view.frustum.data_ptr(), s["rots"].data_ptr(),
                             s["trans"].data_ptr(), ctypes.byref(g), C, s["prk"].data_ptr(), geom, s["acc"].data_ptr(), B, Z * Y, 0, 0,
                             None, 0, st)


def timeit(geom, reps=40):
    for s in sets: call(s, geom)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): call(sets[i % NS], geom)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


t1 = timeit(1)
t0 = timeit(0)
tm = 0.0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(40): sets[i % NS]["acc"].zero_()
b.record(); torch.cuda.synchronize()
tm = a.elapsed_time(b) / 40 * 1e3
print(json.dumps({"cfg": name, "B": B, "memset+scatter_with_geometry_us": round(t1, 1), "memset+scatter_ranks_given_us": round(t0, 1),
                  "memset_alone_us": round(tm, 1)}))
