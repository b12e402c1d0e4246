"""Device time of the lift head (softmax + split + transpose) vs torch softmax + slice + our transpose, BASELINE shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda:0")


def timeit(fn, n=50):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


for name, B, dt in (("bevdet_r50_b8", 8, torch.float32), ("bevdepth_hires_b16", 4, torch.bfloat16),
                    ("rcfusion_omnihd_b32", 2, torch.float32)):
    cfg = pkg.synthetic.CONFIGS[name]
    v = pkg.LSSViewTransform.from_config(cfg)
    BN, D, C, H, W = B * cfg.n_cams, v.D, cfg.channels, v.fH, v.fW
    x = torch.randn(BN, D + C, H, W, device=dev).to(dt)
    e = x.element_size()
    byts = 2 * BN * (D + C) * H * W * e
    depth, feat = pkg.get_depth_feat(x, D, C, channels_last=True)
    gd, gf = torch.randn_like(depth), torch.randn_like(feat)
    lib = pkg._lib.load()
    code = 0 if dt == torch.float32 else 1
    t_f = timeit(lambda: lib.bevpool_lift_forward(x.data_ptr(), depth.data_ptr(), feat.data_ptr(), BN, D, C, H * W, 1, code,
                                                  torch.cuda.current_stream().cuda_stream))
    xg = torch.empty_like(x)
    t_b = timeit(lambda: lib.bevpool_lift_backward(depth.data_ptr(), gd.data_ptr(), gf.data_ptr(), xg.data_ptr(), BN, D, C,
                                                   H * W, 1, 0 if dt == torch.float32 else 1,
                                                   torch.cuda.current_stream().cuda_stream))
    t_t = timeit(lambda: (x[:, :D].softmax(dim=1), x[:, D:].permute(0, 2, 3, 1).contiguous()))
    print(json.dumps({"cfg": name, "B": B, "dtype": str(dt), "lift_fwd_us": round(t_f, 1), "lift_bwd_us": round(t_b, 1),
                      "torch_softmax_slice_permute_us": round(t_t, 1), "fwd_GBps": round(byts / t_f / 1e3, 1),
                      "bwd_GBps": round((3 * BN * D * H * W * e + 2 * BN * C * H * W * e) / t_b / 1e3, 1)}))
