"""What the box's PCIe link gives for the e2e step's copies: H2D only, D2H only, both directions at once."""
import torch, json
dev = torch.device("cuda:0")
n = 60733696 // 4
h_in = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
d_in = [torch.empty(n, device=dev) for _ in range(2)]
d_out = [torch.empty(n, device=dev) for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for i in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[i % 2].copy_(h_in[i % 2], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i % 2].copy_(d_out[i % 2], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for _ in range(2):
    res = {"h2d_ms": run(True, False), "d2h_ms": run(False, True), "both_ms": run(True, True)}
res["GBps_h2d"] = n * 4 / res["h2d_ms"] / 1e6
res["GBps_d2h"] = n * 4 / res["d2h_ms"] / 1e6
res["GBps_each_when_both"] = n * 4 / res["both_ms"] / 1e6
print(json.dumps(res))
