#!/usr/bin/env python
"""bench.py — BEV-pool frames/s and achieved HBM GB/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm, rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic 6-camera frames (BASELINE.json configs[1]:
BEVDet-R50 shapes, B = 8 frames per GPU, fp32, forward + backward): frustum geometry + voxel ranking + depth-weighted
pooling into the [B,C,Z,Y,X] grid + both gradients, through the public module call `LSSViewTransform.forward /
.backward` — the call an UNMODIFIED reference `LiftSplatShoot.get_voxels` reaches after `plugin.patch_lss_class`
(variants.reference_class_zero_edit runs exactly that). Its default forward is the sort-free pixel-major kernel (fp32
vector REDs into the grid); the sorted, atomics-free forward the north star describes is `variants.deterministic`, the
reference's own three-call sequence (get_geometry -> voxel_pooling_prepare_v2 -> bev_pool_v2) is
`variants.reference_api_sequence`, and the reference's own CUDA extension + torch-op prepare on the same GPU is
`variants.reference_cuda_ext`.

value : whole-job frames/s, inputs resident in HBM, CUDA graph replay timed with CUDA events, max over ranks.
e2e   : the same step fed from pinned HOST buffers (H2D of rots/trans/depth/feat/out_grad every step, D2H of
        bev/depth_grad/feat_grad every step) inside the timed region.
roofline / cpu_baseline / clocks / gpu_launches / configs: see DESIGN.md §6.
Under torchrun (WORLD_SIZE > 1) the line additionally carries `variants.occ_allgather` (BASELINE configs[4]: B = 64
frames sharded over the ranks + NCCL all-gather of the BEV-grid shards, off / synchronous / overlapped, fp32 / bf16
payload) and `variants.rcfusion_strong` (configs[3]: B = 32 frames sharded, strong scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bev_pool_fwd_bwd_frames_per_s"
UNIT = "frames/s"
WORKLOAD = "bevdet_r50_b8"          # BASELINE.json configs[1]
N_BUFFER_SETS = 4                   # rotated so no step finds its inputs in the 126 MB L2
# eager (non-graph) variants allocate their outputs per call and keep one result per buffer set alive: every set must
# have been through the caching allocator twice before timing, or cudaMalloc calls land in the timed region
EAGER_WARMUP = 2 * N_BUFFER_SETS + 1
NVLINK_PEAK_GBS = 900.0             # per direction per GPU, nominal (B200_PROFILING.md; measured peer copy 770)
NVLINK_MEASURED_GBS = 770.0


def algorithmic_bytes(P0, P, I, F, V, C, e=4):
    """every input read once, every output written once at the function-level contract."""
    idx = 4 * (2 * P) + 4 * (3 * I)
    fwd = e * P + e * C * F + idx + e * C * V
    bwd = e * C * V + e * P + e * C * F + idx + e * P0 + e * C * F
    prep = 4 * (3 * P) + 4 * (2 * I)
    return dict(fwd=fwd, bwd=bwd, prep=prep)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, reasons, mx = [], set(), None
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step_factory(cfg_name, frames, cams=None):
    """The reference's CPU implementation of the path, all host threads (torch CPU ops).
    kind "reference": the reference's own LiftSplatShoot.get_geometry and QuickCumsum run unmodified from the staged
    files (oracle/_ref/py) or the reference tree; kind "port": the oracle's restatement (bit-identical results), used
    only when neither is present. cams: use only the first `cams` cameras of each frame (a smaller bounded sample)."""
    import torch
    from __graft_entry__ import load_package
    from oracle import oracle as orc
    from oracle import refimport
    pkg = load_package()
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    torch.set_num_threads(os.cpu_count() or 1)
    rots, trans = pkg.synthetic.camera_ring(frames, cfg.n_cams, cfg.final_dim, seed=0)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=frames, seed=0)
    if cams is not None:
        rots, trans, depth, feat = (t[:, :cams].contiguous() for t in (rots, trans, depth, feat))
    if refimport.ref_root() is not None:
        ref = refimport.import_reference_lss("bevfusion")
        lss = refimport.make_reference_lss(ref, cfg.final_dim, cfg.downsample, cfg.dbound, cfg.xbound, cfg.ybound, cfg.zbound)

        def step():
            return orc.cpu_reference_class_step(ref, lss, rots, trans, depth, feat, gout)
        return step, torch.get_num_threads(), "reference"
    view = pkg.LSSViewTransform.from_config(cfg)
    fr = view.frustum.data

    def step():
        return orc.cpu_view_transform_step(fr, rots, trans, depth, feat, gout, view.dx, view.bx, view.nx)
    return step, torch.get_num_threads(), "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 1.0   # bounded sample: 1 of the workload's 8 frames per step
    step, threads, kind = cpu_reference_step_factory(WORKLOAD, 1)
    step()         # first call pays torch's one-off thread-pool / allocator start-up
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    cams = 6
    if t1 * (args.steps + args.warmup) > 300.0:      # keep the whole run within a few minutes
        cams = max(1, int(6 * 300.0 / (t1 * (args.steps + args.warmup))))
        step, threads, kind = cpu_reference_step_factory(WORKLOAD, 1, cams=cams)
        frames = cams / 6.0
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = frames * args.steps / dt
    sample = (f"{cams} of 6 cameras of 1 of the workload's 8 frames per step (frames are independent), "
              "fwd+bwd incl. geometry+prepare, the reference's get_geometry + QuickCumsum on torch CPU ops")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames, "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ our arm: helpers
def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads (and so its first-touched pinned buffers) to the CPUs next to its GPU, as any
    multi-GPU launcher does: with 8 ranks the e2e copies otherwise cross the socket interconnect. Best effort."""
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"pci": dev, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError):
        pass
    return None


class Ctx:
    """Rank / device / timing plumbing shared by every measurement."""

    def __init__(self, torch, dist, world, rank, dev):
        self.torch, self.dist, self.world, self.rank, self.dev = torch, dist, world, rank, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, finalize=None):
        """ms for `steps` calls of fn(i) on the current stream: barrier + synchronize on both sides, CUDA events,
        max over ranks. finalize(): joins side streams into the current one before the closing event."""
        torch = self.torch
        for i in range(warmup):
            fn(i)
        if finalize is not None:
            finalize()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if finalize is not None:
            finalize()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def make_sets(pkg, torch, cfg, B, dev, dt, n_sets, seed0, on_device=False):
    """n_sets independent input sets. on_device: the big tensors are drawn on the GPU (same distributions; used for
    the large secondary configs where drawing hundreds of MB on the host would dominate the run)."""
    sets, host = [], []
    X, Y, Z = (int(v) for v in pkg.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)[2])
    for s in range(n_sets):
        rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=seed0 + s)
        if on_device:
            g = torch.Generator(device=dev).manual_seed(seed0 + s)
            depth = torch.randn(B, cfg.n_cams, cfg.D, cfg.fH, cfg.fW, device=dev, generator=g).softmax(dim=2).to(dt)
            feat = torch.randn(B, cfg.n_cams, cfg.channels, cfg.fH, cfg.fW, device=dev, generator=g).to(dt)
            gout = torch.randn(B, cfg.channels, Z, Y, X, device=dev, generator=g).to(dt)
            sets.append(dict(rots=rots.to(dev), trans=trans.to(dev), depth=depth.requires_grad_(),
                             feat=feat.requires_grad_(), gout=gout))
        else:
            depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=seed0 + s)
            h = [t.pin_memory() for t in (rots, trans, depth.to(dt), feat.to(dt), gout.to(dt))]
            host.append(h)
            r, t, d, f, g = (x.to(dev) for x in h)
            sets.append(dict(rots=r, trans=t, depth=d.requires_grad_(), feat=f.requires_grad_(), gout=g))
    return sets, host


def capture_graphs(torch, view, sets, out_key="bev", gout_key="gout", **fwd_kwargs):
    """One CUDA graph of the whole step per buffer set (after eager warm-up on a side stream)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            for _ in range(2):
                s["depth"].grad = s["feat"].grad = None
                view(s["depth"], s["feat"], s["rots"], s["trans"], **fwd_kwargs).backward(s[gout_key])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = []
    for s in sets:
        s["depth"].grad = s["feat"].grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            s[out_key] = view(s["depth"], s["feat"], s["rots"], s["trans"], **fwd_kwargs)
            s[out_key].backward(s[gout_key])
        graphs.append(g)
    torch.cuda.synchronize()
    return graphs


def measured_counts(pkg, view, s, B, cfg, dev):
    pr = pkg.view_transform._prepare_device(None, view.frustum, s["rots"], s["trans"], B, cfg.n_cams, view.D, view.fH,
                                            view.fW, view.dx, view.bx, view.nx, dev, want_intervals=True)
    P, I = (int(v) for v in pr.counts.tolist())
    return P, I


def config_record(ctx, pkg, name, B, steps, n_sets=2, groups=2):
    """Step time + step roofline of one more BASELINE config at a bounded per-GPU batch (graph replay, rotating sets)."""
    torch = ctx.torch
    cfg = pkg.synthetic.CONFIGS[name]
    dt = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
    view = pkg.LSSViewTransform.from_config(cfg, frame_groups=groups if B % groups == 0 else 1).to(ctx.dev)
    sets, _ = make_sets(pkg, torch, cfg, B, ctx.dev, dt, n_sets, 7000 + 100 * ctx.rank, on_device=True)
    graphs = capture_graphs(torch, view, sets)
    ms = ctx.timed(lambda i: graphs[i % n_sets].replay(), steps, 3)
    X, Y, Z = (int(v) for v in view.nx)
    P, I = measured_counts(pkg, view, sets[0], B, cfg, ctx.dev)
    e = 2 if dt == torch.bfloat16 else 4
    ab = algorithmic_bytes(B * cfg.n_cams * view.D * view.fH * view.fW, P, I, B * cfg.n_cams * view.fH * view.fW,
                           B * X * Y * Z, cfg.channels, e)
    ms_step = ms / steps
    gbs = (ab["fwd"] + ab["bwd"]) / (ms_step * 1e-3) / 1e9
    del graphs, sets
    torch.cuda.empty_cache()
    return {"workload": name, "frames_per_gpu": B, "dtype": "bf16" if e == 2 else "f32", "ms_per_step": ms_step,
            "value": ctx.world * B / (ms_step * 1e-3), "unit": UNIT, "P": P, "I": I,
            "algorithmic_bytes_fwd_plus_bwd": ab["fwd"] + ab["bwd"], "step_GBps": gbs}


def timed_local(torch, fn, reps):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(5 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps      # ms per launch


# ------------------------------------------------------------------------------------------ reference-on-GPU legs
def reference_gpu_variants(ctx, pkg, cfg, sets, steps):
    """(a) the UNMODIFIED reference LiftSplatShoot with the zero-edit patch (plugin.install + patch_lss_class):
    its own get_voxels -> our fused path; (b) the reference's own Python + its own CUDA extension compiled for sm_100a
    (oracle/_ref/libref_bevpool_v2.so) + its torch-op get_geometry / voxel_pooling_prepare_v2, eager, as the plugin
    runs it today: the existing kernel to beat. Both on the headline config's device-resident inputs."""
    torch = ctx.torch
    from oracle import refimport, ref_ext
    out = {}
    if refimport.ref_root() is None:
        return out
    B = sets[0]["depth"].shape[0]

    def purge():
        for name in list(sys.modules):
            if name.startswith("projects.mmdet3d_plugin"):
                del sys.modules[name]

    def build_lss(ref):
        lss = refimport.make_reference_lss(ref, cfg.final_dim, cfg.downsample, cfg.dbound, cfg.xbound, cfg.ybound,
                                           cfg.zbound, inputC=8, camC=cfg.channels).to(ctx.dev)
        return lss

    # ---- (a) zero-edit drop-in on the unmodified class
    purge()
    pkg.plugin.install(force=True)
    ref = refimport.import_reference_lss("bevfusion")
    lss = build_lss(ref)
    pkg.plugin.patch_lss_class(ref.LiftSplatShoot)
    cur = {}
    lss.__dict__["get_cam_feats"] = lambda x: (cur["feat"], cur["depth"])      # the conv nets are out of scope

    def zero_edit(i):
        s = sets[i % len(sets)]
        s["depth"].grad = s["feat"].grad = None
        cur["feat"], cur["depth"] = s["feat"], s["depth"]
        bev, _ = lss.get_voxels(None, s["rots"], s["trans"])
        bev.backward(s["gout"])
    n = max(3, min(steps, 50))
    ms = ctx.timed(zero_edit, n, EAGER_WARMUP)
    out["reference_class_zero_edit"] = {
        "value": ctx.world * B * n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n,
        "note": "unmodified reference LiftSplatShoot.get_voxels after plugin.install() + patch_lss_class(): same kernels "
                "as the headline, eager launches (no CUDA graph)"}
    pkg.plugin.unpatch_lss_class(ref.LiftSplatShoot)
    # ---- (a') the same class with patch_lss_class(cuda_graph=True): get_voxels replays graphs of the fused transform
    #      (inputs are copied into the graphs' static buffers every call: rotating buffer sets never match them)
    try:
        pkg.plugin.patch_lss_class(ref.LiftSplatShoot, cuda_graph=True)
        ms = ctx.timed(zero_edit, n, EAGER_WARMUP)
        out["reference_class_zero_edit_graphed"] = {
            "value": ctx.world * B * n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n,
            "note": "as reference_class_zero_edit with patch_lss_class(cuda_graph=True): torch.cuda.make_graphed_callables "
                    "over the fused view transform, autograd intact; includes the copies into / out of static buffers"}
    except Exception as ex:
        out["reference_class_zero_edit_graphed"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    pkg.plugin.unpatch_lss_class(ref.LiftSplatShoot)
    lss.__dict__.pop("_bevpool_b200_graphs", None)

    # ---- (b) the reference's own extension + torch-op prepare
    if ref_ext.available():
        purge()
        ext = ref_ext.load()
        refimport.import_reference_op(ext=ext)                       # the reference's unmodified bev_pool.py
        ref = refimport.import_reference_lss("bevfusion")            # binds the reference's own bev_pool_v2
        lss = build_lss(ref)

        def ref_step(i):
            s = sets[i % len(sets)]
            s["depth"].grad = s["feat"].grad = None
            bev = lss.voxel_pooling_v2(lss.get_geometry(s["rots"], s["trans"]), s["depth"], s["feat"])
            bev.backward(s["gout"])
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                          # torch.range deprecation in the reference
            n = max(3, min(steps, 20))
            ms = ctx.timed(ref_step, n, EAGER_WARMUP)
            # the two reference kernels alone on the same rank arrays (CUDA events on the legacy stream they use)
            s = sets[0]
            with torch.no_grad():
                rb, rd, rf, st, ln = lss.voxel_pooling_prepare_v2(lss.get_geometry(s["rots"], s["trans"]))
                feat_cl = s["feat"].detach().permute(0, 1, 3, 4, 2).contiguous()
                depth = s["depth"].detach()
                X, Y, Z = (int(v) for v in lss.nx)
                C = feat_cl.shape[-1]
                outz = torch.zeros((B, Z, Y, X, C), device=ctx.dev)
                order = rf.argsort()
                rf_s, rd_s, rb_s = rf[order].contiguous(), rd[order].contiguous(), rb[order].contiguous()
                kept = torch.ones(rf_s.numel(), dtype=torch.bool, device=ctx.dev)
                kept[1:] = rf_s[1:] != rf_s[:-1]
                st_bp = torch.where(kept)[0].int()
                ln_bp = torch.diff(torch.cat([st_bp, torch.tensor([rf_s.numel()], device=ctx.dev, dtype=torch.int32)])).int()
                og = s["gout"].permute(0, 2, 3, 4, 1).contiguous()
                dg, fg = torch.zeros_like(depth), torch.zeros_like(feat_cl)
                t_fwd = timed_local(torch, lambda i: ext.bev_pool_v2_forward(depth, feat_cl, outz, rd, rf, rb, ln, st), 20)
                t_bwd = timed_local(torch, lambda i: ext.bev_pool_v2_backward(og, dg, fg, depth, feat_cl, rd_s, rf_s, rb_s,
                                                                              ln_bp, st_bp), 20)
        out["reference_cuda_ext"] = {
            "value": ctx.world * B * n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n,
            "kernel_us": {"bev_pool_v2_kernel": t_fwd * 1e3, "bev_pool_grad_kernel": t_bwd * 1e3},
            "note": "the reference's unmodified Python (get_geometry, voxel_pooling_prepare_v2 torch ops, bev_pool.py "
                    "QuickCumsumCuda incl. zeros / argsort / permute) on its unmodified bev_pool_cuda.cu compiled for "
                    "sm_100a; kernel_us = its two kernels alone (L2-warm, same arrays)"}
    purge()
    return out


# ------------------------------------------------------------------------------------------ multi-GPU legs
def occ_allgather_variant(ctx, pkg, steps):
    """BASELINE configs[4]: semantic-occupancy grid 200 x 200 x 16, C = 32, B = 64 frames sharded over the ranks, with the
    optional NCCL all-gather of the BEV-grid shards: off, synchronous (gather after the step on the same stream) and
    overlapped (gather of step i on a side stream while step i + 1 runs). Payload fp32, or bf16 (one elementwise cast of
    the shard before the gather, which then moves half the bytes)."""
    torch, dist = ctx.torch, ctx.dist
    cfg = pkg.synthetic.CONFIGS["occ_200x200x16_b64"]
    Btot = cfg.batch
    if Btot % ctx.world:
        return None
    B = Btot // ctx.world
    view = pkg.LSSViewTransform.from_config(cfg, frame_groups=2 if B % 2 == 0 else 1).to(ctx.dev)
    n_sets = 2
    sets, _ = make_sets(pkg, torch, cfg, B, ctx.dev, torch.float32, n_sets, 9000 + 100 * ctx.rank, on_device=True)
    graphs = capture_graphs(torch, view, sets)
    steps = max(3, min(steps, 20))
    res = {"workload": cfg.name, "frames_total": Btot, "frames_per_gpu": B}
    ms_off = ctx.timed(lambda i: graphs[i % n_sets].replay(), steps, 3)
    res["off"] = {"ms_per_step": ms_off / steps, "value": Btot * steps / (ms_off * 1e-3), "unit": UNIT}
    shard_bytes = sets[0]["bev"].numel() * 4
    for payload in ("f32", "bf16"):
        e = 4 if payload == "f32" else 2
        full = [torch.empty((Btot,) + tuple(sets[0]["bev"].shape[1:]), dtype=torch.float32 if e == 4 else torch.bfloat16,
                            device=ctx.dev) for _ in range(n_sets)]
        stage = [torch.empty_like(sets[0]["bev"], dtype=torch.bfloat16) for _ in range(n_sets)] if e == 2 else None
        recv_bytes = shard_bytes // 4 * e * (ctx.world - 1)

        def gather(k):
            src = sets[k]["bev"]
            if e == 2:
                stage[k].copy_(src)
                src = stage[k]
            dist.all_gather_into_tensor(full[k], src)

        def sync_step(i):
            k = i % n_sets
            graphs[k].replay()
            gather(k)
        ms_sync = ctx.timed(sync_step, steps, 3)
        # overlapped: the gather of step i runs on a side stream while the graph of step i + 1 (other buffer set) runs
        side = torch.cuda.Stream()
        done = [torch.cuda.Event() for _ in range(n_sets)]
        for ev in done:
            ev.record()

        def ovl_step(i):
            k = i % n_sets
            cur = torch.cuda.current_stream()
            cur.wait_event(done[k])                 # the previous gather of this set has read its shard
            graphs[k].replay()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gather(k)
                done[k].record(side)
        ms_ovl = ctx.timed(ovl_step, steps, 3, finalize=lambda: torch.cuda.current_stream().wait_stream(side))
        torch.cuda.synchronize()
        # the collective alone
        ms_g = ctx.timed(lambda i: gather(i % n_sets), steps, 3)
        res[payload] = {
            "gather_ms": ms_g / steps, "recv_bytes_per_gpu": recv_bytes,
            "gather_GBps_per_gpu": recv_bytes / (ms_g / steps * 1e-3) / 1e9,
            "frac_of_nvlink5_nominal": recv_bytes / (ms_g / steps * 1e-3) / 1e9 / NVLINK_PEAK_GBS,
            "frac_of_measured_peer_copy": recv_bytes / (ms_g / steps * 1e-3) / 1e9 / NVLINK_MEASURED_GBS,
            "synchronous": {"ms_per_step": ms_sync / steps, "value": Btot * steps / (ms_sync * 1e-3),
                            "exposed_ms": (ms_sync - ms_off) / steps},
            "overlapped": {"ms_per_step": ms_ovl / steps, "value": Btot * steps / (ms_ovl * 1e-3),
                           "exposed_ms": (ms_ovl - ms_off) / steps}}
        del full, stage
    del graphs, sets
    torch.cuda.empty_cache()
    return res


def rcfusion_strong_variant(ctx, pkg, steps):
    """BASELINE configs[3]: RCFusion camera BEV (240 x 160 x 16 grid, 136 x 240 features, C = 64), B = 32 frames in
    total sharded over the ranks — strong scaling; the radar pillar BEV of the same grid is a dense tensor added /
    concatenated downstream and adds no work to the pooled path."""
    cfg = pkg.synthetic.CONFIGS["rcfusion_omnihd_b32"]
    if cfg.batch % ctx.world:
        return None
    B = cfg.batch // ctx.world
    rec = config_record(ctx, pkg, cfg.name, B, max(3, min(steps, 10)), n_sets=2, groups=2)
    rec.update(frames_total=cfg.batch, scaling="strong")
    return rec


# ------------------------------------------------------------------------------------------ our arm
def run_native_arm(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    lib = pkg._lib.load()          # fails loudly if the sm_100a library is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 and not args.no_numa_bind else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(torch, dist, world, rank, dev)

    cfg = pkg.synthetic.CONFIGS[args.config]
    B = cfg.batch if args.frames is None else args.frames     # frames per GPU (weak scaling)
    dt_t = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
    groups = args.frame_groups if B % max(args.frame_groups, 1) == 0 else 1
    view = pkg.LSSViewTransform.from_config(cfg, frame_groups=groups, deterministic=args.deterministic).to(dev)
    X, Y, Z = (int(v) for v in view.nx)
    C, D, H, W, N = cfg.channels, view.D, view.fH, view.fW, cfg.n_cams
    V, F, P0 = B * X * Y * Z, B * N * H * W, B * N * D * H * W

    # ---- buffer sets (device-resident inputs, pinned host copies for e2e), rotated every step
    sets, host = make_sets(pkg, torch, cfg, B, dev, dt_t, N_BUFFER_SETS, 100 * rank)
    graphs = capture_graphs(torch, view, sets)
    l0 = pkg._lib.launch_count()
    sets[0]["depth"].grad = sets[0]["feat"].grad = None
    view(sets[0]["depth"], sets[0]["feat"], sets[0]["rots"], sets[0]["trans"]).backward(sets[0]["gout"])
    torch.cuda.synchronize()
    launches_per_step = pkg._lib.launch_count() - l0

    # measured P, I of set 0 (reported with the result; they depend on the camera ring)
    P, I = measured_counts(pkg, view, sets[0], B, cfg, dev)
    e = 2 if dt_t == torch.bfloat16 else 4
    ab = algorithmic_bytes(P0, P, I, F, V, C, e)

    # ---- (1) device-resident throughput: graph replay, rotating buffer sets
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = ctx.timed(lambda i: graphs[i % N_BUFFER_SETS].replay(), args.steps, max(args.warmup, 3))
    if rank == 0 and ms_total < 400.0:
        # the timed region is shorter than the sampler's 100 ms period: keep the same load running (untimed) so the
        # clocks line is sampled under load at least three times
        t_end = time.perf_counter() + 0.4
        i = 0
        while time.perf_counter() < t_end:
            graphs[i % N_BUFFER_SETS].replay()
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- (2) e2e: pinned host buffers -> H2D -> graph -> D2H, every step, same public API. Three streams: H2D of step
    #          i+1, the graph of step i and D2H of step i-1 overlap (separate copy engines); events order buffer re-use.
    def make_e2e(step_on_stream, out_of):
        out_host = [dict(bev=torch.empty((B, C, Z, Y, X), dtype=dt_t).pin_memory(),
                         dg=torch.empty((B, N, D, H, W), dtype=dt_t).pin_memory(),
                         fg=torch.empty((B, N, C, H, W), dtype=dt_t).pin_memory()) for _ in range(N_BUFFER_SETS)]
        st_h2d, st_cmp, st_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev_h2d = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
        ev_cmp = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
        ev_d2h = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
        for evs, stream in ((ev_h2d, st_h2d), (ev_cmp, st_cmp), (ev_d2h, st_d2h)):
            for ev in evs:
                ev.record(stream)

        def h2d_of(i):
            k = i % N_BUFFER_SETS
            s, h = sets[k], host[k]
            with torch.cuda.stream(st_h2d), torch.no_grad():
                st_h2d.wait_event(ev_cmp[k])          # the previous step on this set has consumed its inputs
                for dst, src in zip((s["rots"], s["trans"], s["depth"], s["feat"], s["gout"]), h):
                    dst.copy_(src, non_blocking=True)
                ev_h2d[k].record(st_h2d)

        def compute_and_d2h_of(i):
            k = i % N_BUFFER_SETS
            o = out_host[k]
            with torch.cuda.stream(st_cmp):
                st_cmp.wait_event(ev_h2d[k])
                st_cmp.wait_event(ev_d2h[k])          # the previous results of this set have left the device
                step_on_stream(k)
                ev_cmp[k].record(st_cmp)
            with torch.cuda.stream(st_d2h):
                st_d2h.wait_event(ev_cmp[k])
                bev, dg, fg = out_of(k)
                o["bev"].copy_(bev, non_blocking=True)
                o["dg"].copy_(dg, non_blocking=True)
                o["fg"].copy_(fg, non_blocking=True)
                ev_d2h[k].record(st_d2h)

        def pipeline(first, count):
            """`count` steps, every one with its H2D, compute and D2H issued here; the inputs of step i+1 are put on
            the copy stream BEFORE step i is issued (a data loader's prefetch of depth 1), so a step whose host code
            has to wait for the device (the API sequence reads two counts back) does not hold up the next upload."""
            h2d_of(first)
            for i in range(first, first + count):
                if i + 1 < first + count:
                    h2d_of(i + 1)
                compute_and_d2h_of(i)

        def run(steps, warmup):
            pipeline(0, warmup)
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream()
            e0.record(cur)
            for stream in (st_h2d, st_cmp, st_d2h):
                stream.wait_event(e0)
            pipeline(warmup, steps)
            for stream in (st_h2d, st_cmp, st_d2h):
                cur.wait_stream(stream)
            e1.record(cur)
            ctx.barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())
        d2h = sum(t.numel() * t.element_size() for t in out_host[0].values())
        return run, d2h

    h2d = sum(t.numel() * t.element_size() for t in host[0])
    e2e_run, d2h = make_e2e(lambda k: graphs[k].replay(), lambda k: (sets[k]["bev"], sets[k]["depth"].grad, sets[k]["feat"].grad))
    e2e_steps = max(3, min(args.steps, 50))
    ms_e2e = e2e_run(e2e_steps, 4)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)

    # ---- (3) variants, reported beside the headline (never instead of it)
    variants = {}
    if not args.no_variants:
        # (3a) the sorted, fixed-summation-order, atomics-free forward (radix sort + voxel table + streaming pool)
        det_view = pkg.LSSViewTransform.from_config(cfg, frame_groups=groups, deterministic=True).to(dev)
        det_graphs = capture_graphs(torch, det_view, sets, out_key="bev_det")
        ms_d = ctx.timed(lambda i: det_graphs[i % N_BUFFER_SETS].replay(), args.steps, 3)
        variants["deterministic"] = {
            "value": world * B * args.steps / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d / args.steps,
            "step_GBps": (ab["fwd"] + ab["bwd"]) / (ms_d / args.steps * 1e-3) / 1e9,
            "note": "LSSViewTransform(deterministic=True): radix sort + interval table + sorted warp-per-chunk forward, "
                    "no atomics, bit-reproducible; same backward"}
        del det_graphs
        # (3b) the same step with the BEV grid and its gradient in torch.channels_last_3d — same values, both layout
        #      passes gone. The headline `value` stays on the reference's contiguous [B,C,Z,Y,X] contract.
        for s in sets:
            s["gout_cl"] = s["gout"].contiguous(memory_format=torch.channels_last_3d)
        vg = capture_graphs(torch, view, sets, out_key="bev_cl", gout_key="gout_cl", memory_format=torch.channels_last_3d)
        ms_v = ctx.timed(lambda i: vg[i % N_BUFFER_SETS].replay(), args.steps, 3)
        variants["channels_last_3d_grid"] = {"value": world * B * args.steps / (ms_v * 1e-3), "unit": UNIT,
                                             "ms_per_step": ms_v / args.steps}
        del vg
        # (3c) the reference's own call sequence through our drop-in functions (get_geometry -> voxel_pooling_prepare_v2
        #      -> bev_pool_v2, as LiftSplatShoot.get_voxels does): coor is materialised, prepare returns exact-length
        #      sorted rank tensors (one 8-byte device->host read per step), so it runs eagerly, not as a graph
        def api_step(i):
            s = sets[i % N_BUFFER_SETS]
            s["depth"].grad = s["feat"].grad = None
            s["bev_api"] = view.voxel_pooling_v2(view.get_geometry(s["rots"], s["trans"]), s["depth"], s["feat"])
            s["bev_api"].backward(s["gout"])
        api_steps = max(3, min(args.steps, 50))
        ms_a = ctx.timed(api_step, api_steps, EAGER_WARMUP)
        # end to end the step runs on a side stream next to the copy streams: fresh leaves per step, so that their
        # AccumulateGrad nodes live on that stream (nodes created earlier on the legacy default stream would make
        # every backward synchronise with the blocking copy streams: measured 2-16 ms per step instead of ~2)
        def api_step_e2e(k):
            s = sets[k]
            d, f = s["depth"].detach().requires_grad_(), s["feat"].detach().requires_grad_()
            bev = view.voxel_pooling_v2(view.get_geometry(s["rots"], s["trans"]), d, f)
            bev.backward(s["gout"])
            s["api_e2e"] = (bev, d.grad, f.grad)
        api_e2e_run, _ = make_e2e(api_step_e2e, lambda k: sets[k]["api_e2e"])
        ms_ae = api_e2e_run(api_steps, EAGER_WARMUP)
        variants["reference_api_sequence"] = {
            "value": world * B * api_steps / (ms_a * 1e-3), "unit": UNIT, "ms_per_step": ms_a / api_steps,
            "e2e": {"value": world * B * api_steps / (ms_ae * 1e-3), "unit": UNIT, "ms_per_step": ms_ae / api_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "step_GBps": (ab["fwd"] + ab["bwd"] + ab["prep"]) / (ms_a / api_steps * 1e-3) / 1e9,
            "note": "get_geometry + voxel_pooling_prepare_v2 + bev_pool_v2 + backward, eager; bytes = fwd + bwd + prepare outputs"}
        # (3d) the reference itself on this GPU
        try:
            variants.update(reference_gpu_variants(ctx, pkg, cfg, sets, args.steps))
        except Exception as ex:      # comparator leg only: never lose the headline line over it
            variants["reference_cuda_ext"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- (4) per-kernel timing for the roofline (rank 0; CUDA events on the launching stream,
    #          L2-cold: inputs rotate over the buffer sets)
    kernels = {}
    if rank == 0:
        bp = pkg.bev_pool
        prs = [pkg.view_transform._prepare_device(None, view.frustum, s["rots"], s["trans"], B, N, D, H, W,
                                                  view.dx, view.bx, view.nx, dev, want_intervals=False) for s in sets]
        feats = [s["feat"].detach() for s in sets]                                    # [B,N,C,H,W]
        feat_cl = [f.new_empty((B * N, H, W, C)) for f in feats]
        for f, fc in zip(feats, feat_cl):
            bp._launch_transpose(f, fc, B * N, C, H * W, True)
        outs = [torch.empty((B, C, Z, Y, X), dtype=dt_t, device=dev) for _ in sets]
        tables = [bp._launch_voxel_table(p.rb, p.p0, p.counts, V) for p in prs]
        og_cl = [torch.empty((B, Z, Y, X, C), dtype=dt_t, device=dev) for _ in sets]
        dgs = [torch.empty_like(s["depth"]) for s in sets]
        fgs = [torch.empty_like(f) for f in feats]
        code = bp._dtype_code(feat_cl[0])
        st = torch.cuda.current_stream().cuda_stream

        def k_fwd(i):
            k = i % N_BUFFER_SETS
            p = prs[k]
            bp._launch_forward_dense(sets[k]["depth"].detach(), feat_cl[k], outs[k], p.rd, None, p.rb, tables[k], B, Z * Y, X,
                                     pkg._lib.LAYOUT_BCZYX, dhw=D * H * W, hw=H * W, n_points=p.p0, counts_dev=p.counts)

        def k_tr(i):
            k = i % N_BUFFER_SETS
            bp._launch_transpose(sets[k]["gout"], og_cl[k], B, C, X * Y * Z, True)

        def k_trf(i):
            k = i % N_BUFFER_SETS
            bp._launch_transpose(feats[k], feat_cl[k], B * N, C, H * W, True)

        def k_tbl(i):
            k = i % N_BUFFER_SETS
            lib.bevpool_voxel_table(prs[k].rb.data_ptr(), prs[k].p0, prs[k].counts.data_ptr(), V, tables[k].data_ptr(), st)

        def k_bwd(i):
            k = i % N_BUFFER_SETS
            p = prs[k]
            lib.bevpool_v2_backward_dense(og_cl[k].data_ptr(), dgs[k].data_ptr(), fgs[k].data_ptr(),
                                          sets[k]["depth"].data_ptr(), feat_cl[k].data_ptr(), p.point_rank.data_ptr(),
                                          p.bn, p.d, p.h, p.w, C, 1, 1 if Z == 1 else 0, code, st)

        def k_view(i):      # default forward: view_fwd_scatter (geometry, ranks, pooling) + glue launches
            k = i % N_BUFFER_SETS
            pkg.view_transform._view_forward_scatter(sets[k]["depth"].detach(), feat_cl[k], outs[k], view, sets[k]["rots"],
                                                     sets[k]["trans"], B, N, D, H, W, C, B, Z * Y, pkg._lib.LAYOUT_BCZYX)

        def k_prep(i):
            k = i % N_BUFFER_SETS
            pkg.view_transform._prepare_device(None, view.frustum, sets[k]["rots"], sets[k]["trans"], B, N, D, H, W,
                                               view.dx, view.bx, view.nx, dev, want_intervals=False)

        def k_prep_api(i):      # prepare from a materialised coor with interval arrays (the API form), no read-back
            k = i % N_BUFFER_SETS
            pkg.view_transform._prepare_device(coors[k], None, None, None, B, N, D, H, W, view.dx, view.bx, view.nx, dev,
                                               want_intervals=True)
        coors = [view.get_geometry(s["rots"], s["trans"]) for s in sets]
        reps = 40
        todo = [("grid_transpose", k_tr), ("pool_bwd_dense", k_bwd), ("feat_transpose", k_trf), ("view_forward", k_view),
                ("pool_fwd_dense", k_fwd), ("prepare_fused_geometry", k_prep), ("prepare_from_coor", k_prep_api),
                ("voxel_table", k_tbl)]
        for name, fn in todo:
            kernels[name] = timed_local(torch, fn, reps) * 1e-3      # seconds per launch
        del coors, prs, outs, og_cl, dgs, fgs, tables

    # ---- (5) the other BASELINE configs at a bounded per-GPU batch (every rank runs them: the timings are max over ranks)
    configs = {}
    if not args.no_configs:
        for name, b in (("bevdepth_hires_b16", 8), ("rcfusion_omnihd_b32", 4), ("occ_200x200x16_b64", 8)):
            try:
                configs[name] = config_record(ctx, pkg, name, b, max(3, min(args.steps, 20)))
            except Exception as ex:
                configs[name] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    if world > 1 and not args.no_variants:
        for key, fn in (("occ_allgather", occ_allgather_variant), ("rcfusion_strong", rcfusion_strong_variant)):
            try:
                variants[key] = fn(ctx, pkg, args.steps)
            except Exception as ex:
                variants[key] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}
    elif not args.no_variants and not args.no_configs:
        try:      # the 1-GPU anchor of the strong-scaling series (B = 32 on one GPU)
            variants["rcfusion_strong"] = rcfusion_strong_variant(ctx, pkg, args.steps)
        except Exception as ex:
            variants["rcfusion_strong"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    # The step's default path is: feat transpose, view_forward (memset + view_fwd_scatter + acc_layout), out_grad
    # transpose, pool_bwd_dense. The roofline is reported for the largest SINGLE kernel launch of the step, which the ncu
    # launch list (profiles/) shows to be the backward kernel or view_fwd_scatter; view_forward here is the composite of
    # three launches, so the backward kernel is the one kernel timed alone.
    dom = "pool_bwd_dense"
    dom_bytes = ab["bwd"]
    achieved = dom_bytes / kernels[dom] / 1e9
    traffic, traffic_src = None, None
    try:     # dram bytes per launch from the committed `ncu --set full` capture (cannot be measured live)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if args.config == WORKLOAD and B == cfg.batch:
            traffic = tj.get(dom)
            traffic_src = {"file": "profiles/ncu_traffic.json", "capture": tj.get("_capture"), "commit": tj.get("_commit")}
    except OSError:
        pass

    def kern(name, nbytes=None, **extra):
        if name not in kernels:
            return None
        d = {"us": kernels[name] * 1e6}
        if nbytes is not None:
            d.update(bytes=nbytes, GBps=nbytes / kernels[name] / 1e9, frac=nbytes / kernels[name] / 1e9 / peak)
        d.update(extra)
        return d
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_8TBps_nominal": achieved / 8000.0, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "us_per_launch": kernels[dom] * 1e6,
                "all_kernels": {
                    "view_forward(memset+scatter+layout; geometry, ranks and pooling, default path)": kern("view_forward", ab["fwd"]),
                    "pool_fwd_dense(chunk+fixup+layout, 3 launches; deterministic path)": kern("pool_fwd_dense", ab["fwd"]),
                    "pool_bwd_dense": kern("pool_bwd_dense", ab["bwd"]),
                    "grid_transpose(out_grad)": kern("grid_transpose", 2 * e * C * V),
                    "feat_transpose": kern("feat_transpose", 2 * e * C * F),
                    "voxel_table(deterministic path)": kern("voxel_table"),
                    "prepare(fused geometry, sorted lists only; deterministic path)": kern("prepare_fused_geometry", ab["prep"]),
                    "prepare(from coor, with intervals; API path, all launches, no read-back)":
                        kern("prepare_from_coor", ab["prep"] + 12 * P0)},
                "step_GBps_fwd_plus_bwd": (ab["fwd"] + ab["bwd"]) / (ms_step * 1e-3) / 1e9,
                "step_frac": (ab["fwd"] + ab["bwd"]) / (ms_step * 1e-3) / 1e9 / peak}

    # ---- (6) CPU baseline (N=1 only): the reference's PyTorch cumsum path on this box's host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        step, threads, kind = cpu_reference_step_factory(args.config, 1)
        step()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 12.0 and reps < 400):
            step()
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": reps / dt, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"{reps} steps of 1 frame (of the workload's {B}), fwd+bwd incl. geometry+prepare, "
                                  f"the reference's get_geometry + QuickCumsum on torch CPU ops, {dt:.1f} s"}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if dt_t == torch.bfloat16 else "f32", "data": "synthetic",
        "config": {"workload": args.config, "frames_per_gpu": B, "cams": N, "feat": [H, W], "D": D, "C": C,
                   "grid": [X, Y, Z], "P0": P0, "P": P, "I": I,
                   "step": "geometry+ranks+fwd+bwd (LSSViewTransform.forward + backward = what the unmodified reference "
                           "LiftSplatShoot.get_voxels runs after plugin.patch_lss_class; one CUDA graph per buffer set)",
                   "forward_path": "deterministic(sorted, no atomics)" if view.deterministic else "scatter(sort-free, fp32 REDs)",
                   "frame_groups": groups,
                   "l2": f"inputs rotated over {N_BUFFER_SETS} buffer sets (> 126 MB L2 in total), no explicit flush",
                   "parallelism": f"frame-sharded x{world}, no collective on the path", "numa_bind": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "variants": variants, "configs": configs,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default=WORKLOAD)
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU (default: the config's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin ranks to the CPUs local to their GPU")
    ap.add_argument("--no-variants", action="store_true", help="skip the variant measurements (deterministic, API sequence, "
                    "reference CUDA ext, all-gather, strong scaling)")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--deterministic", action="store_true",
                    help="headline through the sorted, fixed-summation-order forward instead of the sort-free scatter")
    ap.add_argument("--frame-groups", type=int, default=2,
                    help="independent frame groups run on concurrent streams inside one step (1 = single stream)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
