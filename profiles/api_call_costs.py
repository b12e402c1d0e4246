"""Where the host time of the eager API sequence goes: every C-ABI call is wrapped with a wall-clock timer (time spent
inside the library = CUDA driver calls), the rest of each stage is Python / torch glue. Also: cost of one tiny launch."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
lib = pkg._lib.load()
pc = time.perf_counter
acc = collections.defaultdict(lambda: [0, 0.0])


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t = pc()
        r = self.fn(*a)
        e = acc[self.name]
        e[0] += 1
        e[1] += pc() - t
        return r


class LibProxy:
    def __init__(self, lib):
        self._lib = lib
        self._cache = {}

    def __getattr__(self, name):
        f = self._cache.get(name)
        if f is None:
            f = self._cache[name] = Timed(name, getattr(self._lib, name))
        return f


pkg._lib._LIB = LibProxy(lib)
cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
dev = torch.device("cuda:0")
view = pkg.LSSViewTransform.from_config(cfg).to(dev)
B, NS = cfg.batch, 4
sets = []
for s in range(NS):
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=s)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, seed=s)
    sets.append((rots.to(dev), trans.to(dev), gout.to(dev), depth.to(dev).requires_grad_(), feat.to(dev).requires_grad_()))


def step(i):
    rots, trans, gout, d, f = sets[i % NS]
    d.grad = f.grad = None
    bev = view.voxel_pooling_v2(view.get_geometry(rots, trans), d, f)
    bev.backward(gout)


for mode in ("early", "late"):
    if mode == "late":
        os.environ["BEVPOOL_LATE_COUNTS"] = "1"
    for i in range(12):
        step(i)
    torch.cuda.synchronize()
    acc.clear()
    n = 200
    t0 = pc()
    for i in range(n):
        step(i)
    torch.cuda.synchronize()
    tot = (pc() - t0) / n * 1e6
    inside = sum(v[1] for v in acc.values()) / n * 1e6
    print(f"{mode}: {tot:.1f} us/step wall, {inside:.1f} us inside the library")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:45s} {v[0] / n:4.1f} calls/step  {v[1] / n * 1e6:7.1f} us/step")

# one tiny launch, GPU idle vs GPU busy
a = torch.zeros(1, 4, 8, device=dev)
b = torch.empty(1, 8, 4, device=dev)
raw = lib.bevpool_grid_transpose
st = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
t0 = pc()
for _ in range(2000):
    raw(a.data_ptr(), b.data_ptr(), 1, 4, 8, 1, 0, st)
t1 = pc()
torch.cuda.synchronize()
print(f"tiny transpose launch (PDL attr): {(t1 - t0) / 2000 * 1e6:.2f} us per call (queue filling)")
t0 = pc()
for _ in range(2000):
    a.zero_()
t1 = pc()
torch.cuda.synchronize()
print(f"torch zero_ (memset/fill kernel): {(t1 - t0) / 2000 * 1e6:.2f} us per call")
t0 = pc()
for _ in range(2000):
    x = torch.empty(1 << 20, device=dev)
t1 = pc()
print(f"torch.empty(4 MB): {(t1 - t0) / 2000 * 1e6:.2f} us per call")
os.environ["BEVPOOL_PDL"] = "0"
