#!/usr/bin/env python
"""bench.py — BEV-pool frames/s and achieved HBM GB/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm, rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic 6-camera frames
(BASELINE.json configs[1]: BEVDet-R50 shapes, B=8 frames per GPU, fp32, fwd+bwd):
frustum geometry + voxel ranking + radix sort + interval segmentation + fused
zero-fill/pool forward + sort-free backward, through the public API
(LSSViewTransform.forward / .backward), captured in a CUDA graph.

value : whole-job frames/s, inputs resident in HBM, timed with CUDA events, max over ranks.
e2e   : the same step fed from pinned HOST buffers (H2D of rots/trans/depth/feat/out_grad every
        step, D2H of bev/depth_grad/feat_grad every step) inside the timed region.
roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md §Measurement.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bev_pool_fwd_bwd_frames_per_s"
UNIT = "frames/s"
WORKLOAD = "bevdet_r50_b8"          # BASELINE.json configs[1]
N_BUFFER_SETS = 4                   # rotated so no step finds its inputs in the 126 MB L2


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
    """The reference's CPU implementation of the path (oracle port; torch CPU ops, all host threads).
    cams: use only the first `cams` cameras of each frame (a smaller bounded sample)."""
    import torch
    from __graft_entry__ import load_package
    from oracle import oracle as orc
    pkg = load_package()
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    torch.set_num_threads(os.cpu_count() or 1)
    view = pkg.LSSViewTransform.from_config(cfg)
    rots, trans = pkg.synthetic.camera_ring(frames, cfg.n_cams, cfg.final_dim, seed=0)
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=frames, seed=0)
    if cams is not None:
        rots, trans, depth, feat = (t[:, :cams].contiguous() for t in (rots, trans, depth, feat))
    fr = view.frustum.data

    def step():
        return orc.cpu_view_transform_step(fr, rots, trans, depth, feat, gout, view.dx, view.bx, view.nx)
    return step, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 1.0   # bounded sample: 1 of the workload's 8 frames per step
    step, threads = cpu_reference_step_factory(WORKLOAD, 1)
    step()         # first call pays torch's one-off thread-pool / allocator start-up
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    cams = 6
    if t1 * (args.steps + args.warmup) > 300.0:      # keep the whole run within a few minutes
        cams = max(1, int(6 * 300.0 / (t1 * (args.steps + args.warmup))))
        step, threads = cpu_reference_step_factory(WORKLOAD, 1, cams=cams)
        frames = cams / 6.0
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = frames * args.steps / dt
    sample = (f"{cams} of 6 cameras of 1 of the workload's 8 frames per step (frames are independent), "
              "fwd+bwd incl. geometry+prepare, torch CPU ops")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames, "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ our arm
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

    cfg = pkg.synthetic.CONFIGS[args.config]
    B = cfg.batch if args.frames is None else args.frames     # frames per GPU (weak scaling)
    dt_t = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
    groups = args.frame_groups if B % max(args.frame_groups, 1) == 0 else 1
    view = pkg.LSSViewTransform.from_config(cfg, frame_groups=groups, deterministic=args.deterministic).to(dev)
    X, Y, Z = (int(v) for v in view.nx)
    C, D, H, W, N = cfg.channels, view.D, view.fH, view.fW, cfg.n_cams
    V, F, P0 = B * X * Y * Z, B * N * H * W, B * N * D * H * W

    # ---- buffer sets (device-resident inputs), rotated every step
    host = []
    for s in range(N_BUFFER_SETS):
        rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=100 * rank + s)
        depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=100 * rank + s)
        host.append([t.pin_memory() for t in (rots, trans, depth.to(dt_t), feat.to(dt_t), gout.to(dt_t))])
    sets = []
    for h in host:
        rots, trans, depth, feat, gout = (t.to(dev) for t in h)
        sets.append(dict(rots=rots, trans=trans, depth=depth.requires_grad_(), feat=feat.requires_grad_(), gout=gout))

    def eager_step(s):
        s["depth"].grad = None
        s["feat"].grad = None
        bev = view(s["depth"], s["feat"], s["rots"], s["trans"])
        bev.backward(s["gout"])
        return bev

    # ---- capture one CUDA graph per buffer set (whole step: ~12 kernels + 3 memsets)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            for _ in range(2):
                eager_step(s)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    l0 = pkg._lib.launch_count()
    eager_step(sets[0])
    launches_per_step = pkg._lib.launch_count() - l0
    graphs = []
    for s in sets:
        s["depth"].grad = None
        s["feat"].grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            s["bev"] = view(s["depth"], s["feat"], s["rots"], s["trans"])
            s["bev"].backward(s["gout"])
        graphs.append(g)
    torch.cuda.synchronize()

    # measured P, I of set 0 (reported with the result; they depend on the camera ring)
    pr = pkg.view_transform._prepare_device(None, view.frustum, sets[0]["rots"], sets[0]["trans"], B, N, D, H, W,
                                            view.dx, view.bx, view.nx, dev, want_intervals=True)
    P, I = (int(v) for v in pr.counts.tolist())
    e = 2 if dt_t == torch.bfloat16 else 4
    ab = algorithmic_bytes(P0, P, I, F, V, C, e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- (1) device-resident throughput: graph replay, rotating buffer sets
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda i: graphs[i % N_BUFFER_SETS].replay(), args.steps, max(args.warmup, 3))
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

    # ---- (1b) variant, reported beside the headline (never instead of it): the same step with the BEV grid and
    # its gradient in torch.channels_last_3d (what a channels-last conv stack produces/consumes) — same values,
    # both layout passes gone. The headline `value` stays on the reference's contiguous [B,C,Z,Y,X] contract.
    variants = {}
    if not args.no_variants:
        vg = []
        for s in sets:
            s["gout_cl"] = s["gout"].contiguous(memory_format=torch.channels_last_3d)
            s["depth"].grad = None
            s["feat"].grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s["bev_cl"] = view(s["depth"], s["feat"], s["rots"], s["trans"], memory_format=torch.channels_last_3d)
                s["bev_cl"].backward(s["gout_cl"])
            vg.append(g)
        torch.cuda.synchronize()
        ms_v = timed(lambda i: vg[i % N_BUFFER_SETS].replay(), args.steps, max(args.warmup, 3))
        variants["channels_last_3d_grid"] = {"value": world * B * args.steps / (ms_v * 1e-3), "unit": UNIT,
                                             "ms_per_step": ms_v / args.steps}

        # the reference's own call sequence (get_geometry -> voxel_pooling_prepare_v2 -> bev_pool_v2, as
        # LiftSplatShoot.get_voxels does): coor is materialised, prepare returns exact-length sorted rank tensors
        # (one 8-byte device->host read per step), so it runs eagerly, not as a graph
        def api_step(i):
            s = sets[i % N_BUFFER_SETS]
            s["depth"].grad = None
            s["feat"].grad = None
            bev = view.voxel_pooling_v2(view.get_geometry(s["rots"], s["trans"]), s["depth"], s["feat"])
            bev.backward(s["gout"])
        api_steps = max(3, min(args.steps, 50))
        ms_a = timed(api_step, api_steps, 3)
        variants["reference_api_sequence"] = {"value": world * B * api_steps / (ms_a * 1e-3), "unit": UNIT,
                                              "ms_per_step": ms_a / api_steps,
                                              "note": "get_geometry + voxel_pooling_prepare_v2 + bev_pool_v2 + backward, eager"}

    # ---- (2) e2e: pinned host buffers -> H2D -> graph -> D2H, every step, same public API
    out_host = [dict(bev=torch.empty((B, C, Z, Y, X), dtype=dt_t).pin_memory(),
                     dg=torch.empty((B, N, D, H, W), dtype=dt_t).pin_memory(),
                     fg=torch.empty((B, N, C, H, W), dtype=dt_t).pin_memory()) for _ in range(N_BUFFER_SETS)]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = sum(t.numel() * t.element_size() for t in out_host[0].values())

    # three streams: H2D of step i+1, the graph of step i and D2H of step i-1 overlap (separate copy engines);
    # events order the re-use of each buffer set.
    st_h2d, st_cmp, st_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_h2d = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
    ev_cmp = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
    ev_d2h = [torch.cuda.Event() for _ in range(N_BUFFER_SETS)]
    for evs, stream in ((ev_h2d, st_h2d), (ev_cmp, st_cmp), (ev_d2h, st_d2h)):
        for ev in evs:
            ev.record(stream)

    def e2e_step(i):
        k = i % N_BUFFER_SETS
        s, h, o = sets[k], host[k], out_host[k]
        with torch.cuda.stream(st_h2d), torch.no_grad():
            st_h2d.wait_event(ev_cmp[k])          # the previous step on this set has consumed its inputs
            s["rots"].copy_(h[0], non_blocking=True)
            s["trans"].copy_(h[1], non_blocking=True)
            s["depth"].copy_(h[2], non_blocking=True)
            s["feat"].copy_(h[3], non_blocking=True)
            s["gout"].copy_(h[4], non_blocking=True)
            ev_h2d[k].record(st_h2d)
        with torch.cuda.stream(st_cmp):
            st_cmp.wait_event(ev_h2d[k])
            st_cmp.wait_event(ev_d2h[k])          # the previous results of this set have left the device
            graphs[k].replay()
            ev_cmp[k].record(st_cmp)
        with torch.cuda.stream(st_d2h):
            st_d2h.wait_event(ev_cmp[k])
            o["bev"].copy_(s["bev"], non_blocking=True)
            o["dg"].copy_(s["depth"].grad, non_blocking=True)
            o["fg"].copy_(s["feat"].grad, non_blocking=True)
            ev_d2h[k].record(st_d2h)

    def timed_streams(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for stream in (st_h2d, st_cmp, st_d2h):
            stream.wait_event(e0)
        for i in range(steps):
            fn(warmup + i)
        for stream in (st_h2d, st_cmp, st_d2h):
            cur.wait_stream(stream)
        e1.record(cur)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    e2e_steps = max(3, min(args.steps, 50))
    ms_e2e = timed_streams(e2e_step, e2e_steps, 4)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)

    # ---- (3) per-kernel timing for the roofline (rank 0; CUDA events on the launching stream,
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

        def k_view(i):      # default forward: memset + view_fwd_scatter (geometry, ranks, pooling) + layout pass
            k = i % N_BUFFER_SETS
            pkg.view_transform._view_forward_scatter(sets[k]["depth"].detach(), feat_cl[k], outs[k], view, sets[k]["rots"],
                                                     sets[k]["trans"], B, N, D, H, W, C, B, Z * Y, pkg._lib.LAYOUT_BCZYX)

        def k_prep(i):
            k = i % N_BUFFER_SETS
            pkg.view_transform._prepare_device(None, view.frustum, sets[k]["rots"], sets[k]["trans"], B, N, D, H, W,
                                               view.dx, view.bx, view.nx, dev, want_intervals=False)

        reps = 40
        todo = [("grid_transpose", k_tr), ("pool_bwd_dense", k_bwd), ("feat_transpose", k_trf), ("view_forward", k_view)]
        if args.deterministic or args.time_sorted_path:      # the sorted alternative of view_forward
            todo += [("pool_fwd_dense", k_fwd), ("prepare_all", k_prep), ("voxel_table", k_tbl)]
        for name, fn in todo:
            kernels[name] = timed_local(torch, fn, reps) * 1e-3      # seconds per launch

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
    # transpose, pool_bwd_dense. pool_bwd_dense is ONE kernel and the largest single launch of the step, so it is
    # the kernel the roofline is reported for. prepare / voxel_table / pool_fwd_dense are the deterministic
    # (sorted) alternative of view_forward and are timed for comparison.
    dom = "pool_bwd_dense"
    dom_bytes = ab["bwd"]
    achieved = dom_bytes / kernels[dom] / 1e9
    traffic = None
    try:     # dram bytes per launch from the committed `ncu --set full` capture (cannot be measured live)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom) \
            if args.config == WORKLOAD and B == cfg.batch else None
    except OSError:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_8TBps_nominal": achieved / 8000.0, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "us_per_launch": kernels[dom] * 1e6,
                "all_kernels": {
                    "view_forward(memset+scatter+layout; geometry, ranks and pooling, default path)": {
                        "us": kernels["view_forward"] * 1e6, "GBps": ab["fwd"] / kernels["view_forward"] / 1e9,
                        "bytes": ab["fwd"]},
                    "pool_fwd_dense(chunk+fixup+layout, 3 launches; deterministic path)": {
                        "us": kernels["pool_fwd_dense"] * 1e6, "GBps": ab["fwd"] / kernels["pool_fwd_dense"] / 1e9,
                        "bytes": ab["fwd"]} if "pool_fwd_dense" in kernels else None,
                    "pool_bwd_dense": {"us": kernels["pool_bwd_dense"] * 1e6, "GBps": ab["bwd"] / kernels["pool_bwd_dense"] / 1e9,
                                       "bytes": ab["bwd"]},
                    "grid_transpose(out_grad)": {"us": kernels["grid_transpose"] * 1e6,
                                                 "GBps": 2 * e * C * V / kernels["grid_transpose"] / 1e9},
                    "feat_transpose": {"us": kernels["feat_transpose"] * 1e6,
                                       "GBps": 2 * e * C * F / kernels["feat_transpose"] / 1e9},
                    "voxel_table(deterministic path)": {"us": kernels["voxel_table"] * 1e6} if "voxel_table" in kernels else None,
                    "prepare(all kernels, eager launches; deterministic path)":
                        {"us": kernels["prepare_all"] * 1e6, "bytes_out": ab["prep"]} if "prepare_all" in kernels else None},
                "step_GBps_fwd_plus_bwd": (ab["fwd"] + ab["bwd"]) / (ms_step * 1e-3) / 1e9}

    # ---- (4) CPU baseline (N=1 only): the reference's PyTorch cumsum path on this box's host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        step, threads = cpu_reference_step_factory(args.config, 1)
        step()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 12.0 and reps < 400):
            step()
            reps += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{reps} steps of 1 frame (of the workload's {B}), fwd+bwd incl. geometry+prepare, "
                                  f"torch CPU ops, {dt:.1f} s"}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if dt_t == torch.bfloat16 else "f32", "data": "synthetic",
        "config": {"workload": args.config, "frames_per_gpu": B, "cams": N, "feat": [H, W], "D": D, "C": C,
                   "grid": [X, Y, Z], "P0": P0, "P": P, "I": I,
                   "step": "geometry+ranks+fwd+bwd (LSSViewTransform.forward + backward, one CUDA graph per buffer set)",
                   "forward_path": "deterministic(sorted)" if view.deterministic else "scatter(sort-free, fp32 REDs)",
                   "frame_groups": groups,
                   "l2": f"inputs rotated over {N_BUFFER_SETS} buffer sets (> 126 MB L2 in total), no explicit flush",
                   "parallelism": f"frame-sharded x{world}, no collective on the path", "numa_bind": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "variants": variants,
    }))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-variants", action="store_true", help="skip the channels_last_3d variant measurement")
    ap.add_argument("--time-sorted-path", action="store_true",
                    help="also time the kernels of the sorted (deterministic) forward for comparison")
    ap.add_argument("--deterministic", action="store_true",
                    help="fused forward through the sorted, fixed-summation-order path instead of the sort-free scatter")
    ap.add_argument("--frame-groups", type=int, default=2,
                    help="independent frame groups run on concurrent streams inside one step (1 = single stream)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
