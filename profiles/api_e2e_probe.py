"""Where the API sequence's end-to-end step goes when bulk H2D / D2H copies run on other streams: wall time of the
prepare call (it waits for two integers from the device) and of the whole step, early vs late count read-back."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
B, NS = cfg.batch, 4
sets, host, out_host = [], [], []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, seed=s)
    h = [t.pin_memory() for t in (rots, trans, depth, feat, gout)]
    host.append(h)
    r, t, d, f, g = (x.to(dev) for x in h)
    sets.append(dict(rots=r, trans=t, depth=d.requires_grad_(), feat=f.requires_grad_(), gout=g))
    out_host.append([torch.empty_like(x).pin_memory() for x in (gout, depth, feat)])
st_h2d, st_cmp, st_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ev_h2d = [torch.cuda.Event() for _ in range(NS)]; ev_cmp = [torch.cuda.Event() for _ in range(NS)]; ev_d2h = [torch.cuda.Event() for _ in range(NS)]
for evs, stream in ((ev_h2d, st_h2d), (ev_cmp, st_cmp), (ev_d2h, st_d2h)):
    for ev in evs: ev.record(stream)
pc = time.perf_counter
T = {"h2d": 0.0, "geom": 0.0, "prepare": 0.0, "pool_fwd_bwd": 0.0, "d2h": 0.0}


def h2d_of(i):
    k = i % NS; s, h = sets[k], host[k]
    with torch.cuda.stream(st_h2d), torch.no_grad():
        st_h2d.wait_event(ev_cmp[k])
        for dst, src in zip((s["rots"], s["trans"], s["depth"], s["feat"], s["gout"]), h):
            dst.copy_(src, non_blocking=True)
        ev_h2d[k].record(st_h2d)


def step(i, copies):
    k = i % NS; s = sets[k]
    t0 = pc()
    if copies: h2d_of(i + 1)
    t1 = pc()
    with torch.cuda.stream(st_cmp):
        st_cmp.wait_event(ev_h2d[k]); st_cmp.wait_event(ev_d2h[k])
        d, f = (s["depth"].detach().requires_grad_(), s["feat"].detach().requires_grad_()) if FRESH else (s["depth"], s["feat"])
        d.grad = f.grad = None
        coor = view.get_geometry(s["rots"], s["trans"])
        t2 = pc()
        ranks = view.voxel_pooling_prepare_v2(coor)
        t3 = pc()
        rb, rd, rf, stt, ln = ranks
        X, Y, Z = (int(v) for v in view.nx)
        bev = pkg.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, (B, Z, Y, X, f.shape[2]), stt, ln)
        bev.backward(s["gout"])
        s["bev"], s["dg"], s["fg"] = bev, d.grad, f.grad
        ev_cmp[k].record(st_cmp)
    t4 = pc()
    if copies:
        with torch.cuda.stream(st_d2h):
            st_d2h.wait_event(ev_cmp[k])
            for dst, src in zip(out_host[k], (s["bev"], s["dg"], s["fg"])):
                dst.copy_(src, non_blocking=True)
            ev_d2h[k].record(st_d2h)
    t5 = pc()
    for key, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): T[key] += v


FRESH = False
AHEAD = 0
for mode, FRESH, AHEAD in (("early", False, 0), ("early", True, 0), ("early", True, 2), ("early", False, 2), ("late", False, 2)):
    os.environ["BEVPOOL_LATE_COUNTS"] = "1" if mode == "late" else "0"
    for copies in (False, True):
        h2d_of(0)
        for i in range(12): step(i, copies)
        torch.cuda.synchronize()
        for key in T: T[key] = 0.0
        n = 60; t0 = pc()
        for i in range(12, 12 + n):
            step(i, copies)
            if AHEAD and i >= 12 + AHEAD: ev_cmp[(i - AHEAD) % NS].synchronize()   # host at most AHEAD steps in front
        torch.cuda.synchronize()
        tot = (pc() - t0) / n * 1e6
        print(f"{mode:5s} fresh_leaves={FRESH!s:5s} ahead<={AHEAD} copies={copies!s:5s}: {tot:8.1f} us/step | host us: " + ", ".join(f"{k} {v / n * 1e6:.0f}" for k, v in T.items()))
