"""Device time of the §8(f) rank-4 side ops (pillar scatter, cross-modal glue) vs their torch formulations."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


# ---- pillar scatter, RCFusion size: 4 samples, 40 000 pillars each, 64 ch, 320 x 480
B, C, ny, nx, P = 4, 64, 320, 480, 40000
g = torch.Generator().manual_seed(0)
cells = torch.cat([torch.randperm(ny * nx, generator=g)[:P] + b * ny * nx for b in range(B)])
coors = torch.stack([cells // (ny * nx), torch.zeros_like(cells), (cells % (ny * nx)) // nx, cells % nx], 1).to(dev, torch.int32)
feats = torch.randn(B * P, C, device=dev)
m = pkg.pillar_scatter.PointPillarsScatter(C, [ny, nx])


def torch_scatter():
    out = []
    for b in range(B):
        canvas = torch.zeros(C, ny * nx, device=dev)
        mask = coors[:, 0] == b
        idx = (coors[mask, 2] * nx + coors[mask, 3]).long()
        canvas[:, idx] = feats[mask].t()
        out.append(canvas)
    return torch.stack(out, 0).view(B, C, ny, nx)


t_ps, t_ps_ref = timeit(lambda: m(feats, coors, B)), timeit(torch_scatter)
canvas_bytes = B * C * ny * nx * 4
res = {"pillar_scatter_us": round(t_ps, 1), "torch_index_assign_us": round(t_ps_ref, 1),
       "pillar_scatter_GBps": round((canvas_bytes + B * P * C * 4) / t_ps / 1e3, 1)}

# ---- cross-modal glue, RCFusion size
N, Ca, Cb, H, W = 2, 256, 384, 160, 240
a, b = torch.randn(N, Ca, H, W, device=dev), torch.randn(N, Cb, H, W, device=dev)
wa, wb = torch.rand(N, 1, H, W, device=dev), torch.rand(N, 1, H, W, device=dev)
cm = pkg.cross_modal
t_am = timeit(lambda: cm.channel_avg_max(b))
t_am_ref = timeit(lambda: torch.cat([b.mean(1, keepdim=True), b.max(1, keepdim=True)[0]], 1))
t_gc = timeit(lambda: cm.gate_concat(a, b, wa, wb))
t_gc_ref = timeit(lambda: torch.cat([a * wa, b * wb], 1))
res.update({"channel_avg_max_us": round(t_am, 1), "torch_mean_max_cat_us": round(t_am_ref, 1),
            "channel_avg_max_GBps": round(b.numel() * 4 / t_am / 1e3, 1),
            "gate_concat_us": round(t_gc, 1), "torch_mul_mul_cat_us": round(t_gc_ref, 1),
            "gate_concat_GBps": round(2 * (a.numel() + b.numel()) * 4 / t_gc / 1e3, 1)})
print(json.dumps(res))
