"""In-context time of every stage of one training step: CUDA graphs of growing prefixes of the step are
replayed over rotating buffer sets and differenced (so each stage sees the cache state the real step gives it)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
lib = pkg._lib.load()
cfg = pkg.synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
NS = 4
dev = torch.device("cuda:0")
bp, vt = pkg.bev_pool, pkg.view_transform
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
X, Y, Z = (int(v) for v in view.nx)
C, D, H, W, N = cfg.channels, view.D, view.fH, view.fW, cfg.n_cams
V = B * X * Y * Z
dt = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=s)
    sets.append(dict(rots=rots.to(dev), trans=trans.to(dev), depth=depth.to(dev, dt), feat=feat.to(dev, dt), gout=gout.to(dev, dt)))
SCATTER = os.environ.get("PATH_KIND", "scatter") == "scatter"      # PATH_KIND=sorted: radix sort + streaming forward
stages = (["feat_transpose", "scatter_fwd(memset+kernel)", "acc_layout", "og_transpose", "pool_bwd"] if SCATTER else
          ["prepare", "feat_transpose", "voxel_table", "pool_fwd", "og_transpose", "pool_bwd"])
import ctypes


def run_scatter(s, upto):
    st = torch.cuda.current_stream().cuda_stream
    fcl = s["feat"].new_empty((B * N, H, W, C)); bp._launch_transpose(s["feat"], fcl, B * N, C, H * W, True)
    if upto == 0: return fcl
    g = vt._grid_struct(B, N, D, H, W, view.dx, view.bx, view.nx)
    prk = torch.empty(B * N * D * H * W, dtype=torch.int32, device=dev)
    acc = torch.empty((B, Z, Y, X, C), dtype=torch.float32, device=dev)
    code = bp._dtype_code(fcl)
    if upto == 1 and dt == torch.float32:      # kernel + memset only: channels-last fp32 output, no layout pass
        lib.bevpool_view_forward(s["depth"].data_ptr(), fcl.data_ptr(), view.frustum.data_ptr(), s["rots"].data_ptr(), s["trans"].data_ptr(),
                                 ctypes.byref(g), C, prk.data_ptr(), 1, acc.data_ptr(), B, Z * Y, 0, code, None, 0, st)
        return acc, prk
    out = s["feat"].new_empty((B, C, Z, Y, X))
    lib.bevpool_view_forward(s["depth"].data_ptr(), fcl.data_ptr(), view.frustum.data_ptr(), s["rots"].data_ptr(), s["trans"].data_ptr(),
                             ctypes.byref(g), C, prk.data_ptr(), 1, out.data_ptr(), B, Z * Y, 1, code, acc.data_ptr(), acc.numel() * 4, st)
    if upto <= 2: return out, prk, acc
    og = s["gout"].new_empty((B, Z, Y, X, C)); bp._launch_transpose(s["gout"], og, B, C, Z * Y * X, True)
    if upto == 3: return og
    dg = torch.empty_like(s["depth"]); fg = torch.empty_like(s["feat"])
    lib.bevpool_v2_backward_dense(og.data_ptr(), dg.data_ptr(), fg.data_ptr(), s["depth"].data_ptr(), fcl.data_ptr(), prk.data_ptr(),
                                  B * N, D, H, W, C, 1, int(os.environ.get('HINT', 1 if Z == 1 else 0)), code, st)
    return (out, dg, fg, acc, prk)


def run(s, upto):
    if SCATTER:
        return run_scatter(s, upto)
    st = torch.cuda.current_stream().cuda_stream
    pr = vt._prepare_device(None, view.frustum, s["rots"], s["trans"], B, N, D, H, W, view.dx, view.bx, view.nx, dev, want_intervals=False)
    if upto == 0: return pr
    fcl = s["feat"].new_empty((B * N, H, W, C)); bp._launch_transpose(s["feat"], fcl, B * N, C, H * W, True)
    if upto == 1: return fcl
    tab = bp._launch_voxel_table(pr.rb, pr.p0, pr.counts, V)
    if upto == 2: return tab
    out = s["feat"].new_empty((B, C, Z, Y, X))
    bp._launch_forward_dense(s["depth"], fcl, out, pr.rd, None, pr.rb, tab, B, Z * Y, X, pkg._lib.LAYOUT_BCZYX, dhw=D * H * W, hw=H * W, n_points=pr.p0, counts_dev=pr.counts)
    if upto == 3: return out
    og = s["gout"].new_empty((B, Z, Y, X, C)); bp._launch_transpose(s["gout"], og, B, C, Z * Y * X, True)
    if upto == 4: return og
    dg = torch.empty_like(s["depth"]); fg = torch.empty_like(s["feat"])
    lib.bevpool_v2_backward_dense(og.data_ptr(), dg.data_ptr(), fg.data_ptr(), s["depth"].data_ptr(), fcl.data_ptr(), pr.point_rank.data_ptr(),
                                  pr.bn, pr.d, pr.h, pr.w, C, 1, int(os.environ.get('HINT', 1 if Z == 1 else 0)), bp._dtype_code(fcl), st)
    return (out, dg, fg)

def time_prefix(upto, reps=60):
    graphs, keep = [], []
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets: run(s, upto)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(run(s, upto))
        graphs.append(g)
    for i in range(8): graphs[i % NS].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): graphs[i % NS].replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

prev, res = 0.0, {}
for k, name in enumerate(stages):
    t = time_prefix(k)
    res[name] = round(t - prev, 1); prev = t
res["total_us"] = round(prev, 1)
res["frames_per_s"] = round(B / (prev * 1e-6))
print(json.dumps({"cfg": cfg.name, "B": B, "stages_us": res}))
