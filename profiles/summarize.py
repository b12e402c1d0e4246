"""Turn gpurun_out/{launches_r1.csv, pool_r1.ncu-rep, bench_r1.json} into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0][:70]; v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = [f"# {tag} — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (B200, cfg bevdet_r50_b8)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 700` — per-launch times are cold-cache and serialised: compare SHARES.", "",
       "| kernel | launches | avg us | share of captured GPU time |", "|---|---|---|---|"]
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{n}` | {a[0]} | {a[1]/a[0]/1e3:.2f} | {a[1]/tot:.3f} |")
out += ["", "The window covers bench.py's eager warm-up, the graph-captured steps, the channels-last variant and the per-kernel timing loops",
        "(which also time the sorted/deterministic alternative: point_rank, tile_scan, radix_scatter, voxel_table, pool_fwd_chunk, chunk_fixup,",
        "cl_to_bczyx_zero_fill), so launch COUNTS are not per step. A step of the default path issues 5 launches of our kernels per frame group",
        "(transpose(feat), view_fwd_scatter, acc_layout, transpose(out_grad), pool_bwd_joint) plus one memset; bench.py runs 2 groups = 10 launches.", ""]
extra = os.path.join(P, f"{tag}_step_breakdown.md")
if os.path.exists(extra):
    out += open(extra).read().splitlines()
open(os.path.join(P, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
open(os.path.join(P, f"{tag}_launches.csv"), "w").write(open(os.path.join(G, f"launches_{tag}.csv")).read())

raw = subprocess.run(["ncu", "-i", os.path.join(G, f"pool_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
lines = [f"# {tag} — `ncu --set full --clock-control none --import-source on -k regex:pool_bwd_joint|view_fwd_scatter` on bench.py", "",
         "Selected raw metrics (the .ncu-rep itself stays in gpurun_out/, ~8 MB).", ""]
traffic = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    lines += ["## " + name, "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in idx: lines.append(f"| {w} | {r[idx[w]]} | {rows[1][idx[w]]} |")
    lines.append("")
    def mb(k):
        v = float(r[idx[k]].replace(",", "")); u = rows[1][idx[k]]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[u]
    key = "pool_bwd_dense" if "bwd" in name else "view_fwd_scatter"
    traffic[key] = int(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
open(os.path.join(P, f"{tag}_ncu_pool_kernels.md"), "w").write("\n".join(lines))
traffic["_source"] = (f"profiles/{tag}_ncu_pool_kernels.md: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture per kernel "
                      "(cfg bevdet_r50_b8, B=8, frame_groups=1). Writes still sitting in the 126 MB L2 at kernel end are not counted by ncu.")
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for f in ("bench_r1.json", "bench_ref_r1.json"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        open(os.path.join(P, f.replace("bench_", f"{tag}_bench_").replace("_r1", "")), "w").write(open(src).read())
print("\n".join(out[:22])); print(traffic)
