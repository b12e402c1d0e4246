"""Does splitting the frame batch over S concurrent streams (parallel graph branches) help? Frames are independent."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
cfg = pkg.synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bevdet_r50_b8"]
B = cfg.batch
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
NS = 4
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=s)
    sets.append(dict(rots=rots.to(dev), trans=trans.to(dev), depth=depth.to(dev).requires_grad_(), feat=feat.to(dev).requires_grad_(), gout=gout.to(dev)))

def step(s, S, streams):
    cur = torch.cuda.current_stream()
    n = B // S
    outs = []
    for i in range(S):
        st = streams[i] if S > 1 else cur
        if S > 1:
            st.wait_stream(cur)
        with torch.cuda.stream(st):
            sl = slice(i * n, (i + 1) * n)
            d = s["depth"][sl].detach().requires_grad_(); f = s["feat"][sl].detach().requires_grad_()
            bev = view(d, f, s["rots"][sl], s["trans"][sl])
            bev.backward(s["gout"][sl])
            outs.append((bev, d.grad, f.grad))
    if S > 1:
        for i in range(S):
            cur.wait_stream(streams[i])
    return outs

res = {}
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets: step(s, S, streams)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    graphs, keep = [], []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(step(s, S, streams))
        graphs.append(g)
    for i in range(8): graphs[i % NS].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(100): graphs[i % NS].replay()
    b.record(); torch.cuda.synchronize()
    res[S] = round(a.elapsed_time(b) / 100 * 1e3, 1)
print(json.dumps({"cfg": cfg.name, "us_per_step_by_streams": res, "frames_per_s": {k: round(B / (v * 1e-6)) for k, v in res.items()}}))
