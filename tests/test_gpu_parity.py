"""GPU (-m gpu): parity of the sm_100a path against the oracle, the committed reference goldens
and the reference's own CUDA kernels (oracle/_ref), all through the C-ABI library.

Bars: bit-exact (np.array_equal) for ranks, sort order, intervals and geometry; fp32 forward and
backward within 1e-5 of max|ref| (BASELINE.json north_star) — the reference-contract forward is in
fact bit-identical; bf16 within 2^-8 of max (one output rounding, fp32 accumulation).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, rel_to_max

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5           # fp32, relative to max|ref|  (north_star)
TOL_BF16 = 2.0 ** -8  # bf16 outputs


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def canon_ties(rb, rd, rf):
    order = np.lexsort((rd, rb))
    return rb[order], rd[order], rf[order]


def synth_pool_case(pkg, orc, cfg, B, seed=0, roll=0.0, pitch=0.0):
    """Full pipeline inputs on the host + oracle prepare outputs."""
    view = pkg.LSSViewTransform.from_config(cfg)
    rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=seed, roll=roll, pitch=pitch)
    coor = orc.get_geometry(view.frustum.numpy(), rots.numpy(), trans.numpy())
    ranks = orc.prepare_v2(coor, view.dx.numpy(), view.bx.numpy(), view.nx.numpy())
    depth, feat, gout = pkg.synthetic.pool_inputs(cfg, batch=B, seed=seed)
    return view, rots, trans, coor, ranks, depth, feat, gout


# --------------------------------------------------------------------------------------- KAT
def test_reference_kat_through_public_api(pkg):
    """ops/bev_pool_v2/bev_pool.py:145-176, same inputs, same asserts."""
    k = load("kat_bev_pool_v2")
    depth = cu(k["depth"]).requires_grad_()
    feat = cu(k["feat"]).requires_grad_()
    rd, rf, rb = cu(k["ranks_depth"]), cu(k["ranks_feat"]), cu(k["ranks_bev"])
    kept = torch.ones(rb.shape[0], device=DEV, dtype=torch.bool)
    kept[1:] = rb[1:] != rb[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb.shape[0] - starts[-1]
    bev = pkg.bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), starts, lengths)
    assert bev.shape == (1, 2, 1, 2, 2) and bev.is_contiguous()
    loss = torch.sum(bev)
    loss.backward()
    assert loss == 4.4
    assert depth.grad.allclose(cu(k["grad_depth"]))
    assert feat.grad.allclose(cu(k["grad_feat"]))
    # same through the reference-contract Function (channels-last result)
    d2, f2 = cu(k["depth"]).requires_grad_(), cu(k["feat"]).requires_grad_()
    out = pkg.QuickCumsumCuda.apply(d2, f2, rd, rf, rb, (1, 1, 2, 2, 2), starts, lengths)
    assert out.shape == (1, 1, 2, 2, 2)
    out.sum().backward()
    assert d2.grad.allclose(cu(k["grad_depth"])) and f2.grad.allclose(cu(k["grad_feat"]))


# --------------------------------------------------------------------------------------- geometry / prepare
@pytest.mark.parametrize("name", ["tiny_bev_z1", "tiny_occ_z16", "tiny_omnihd", "tiny_hires"])
def test_geometry_bit_exact_vs_reference_golden(pkg, name):
    g = load(name)
    coor = pkg.get_geometry(cu(g["frustum"]), cu(g["rots"]), cu(g["trans"]))
    assert np.array_equal(coor.cpu().numpy(), g["coor"])


@pytest.mark.parametrize("name", ["tiny_bev_z1", "tiny_occ_z16", "tiny_omnihd", "tiny_hires"])
def test_prepare_bit_exact_vs_reference_golden(pkg, name):
    g = load(name)
    out = pkg.voxel_pooling_prepare_v2(cu(g["coor"]), torch.from_numpy(g["dx"]), torch.from_numpy(g["bx"]),
                                       torch.from_numpy(g["nx"]))
    want = canon_ties(g["ranks_bev"], g["ranks_depth"], g["ranks_feat"]) + (g["interval_starts"], g["interval_lengths"])
    for got, ref, key in zip(out, want, ("ranks_bev", "ranks_depth", "ranks_feat", "starts", "lengths")):
        assert got.dtype == torch.int32 and got.is_contiguous()
        assert np.array_equal(got.cpu().numpy(), ref), key


@pytest.mark.parametrize("name", ["mid_bev_z1", "mid_occ_z16", "mid_omnihd"])
def test_prepare_tie_order_vs_reference_golden(pkg, orc, name):
    """P >= 5e4: the unmodified reference's output, tie order included. Geometry comes from our own
    kernel and must hash to the reference's coor."""
    import hashlib
    g = load(name)
    fr = pkg.create_frustum(tuple(int(v) for v in g["final_dim"]), int(g["downsample"]), tuple(float(v) for v in g["dbound"]))
    coor = pkg.get_geometry(fr.to(DEV), cu(g["rots"]), cu(g["trans"]))
    assert hashlib.sha256(coor.cpu().numpy().tobytes()).digest() == g["coor_sha256"].tobytes()
    out = pkg.voxel_pooling_prepare_v2(coor, torch.from_numpy(g["dx"]), torch.from_numpy(g["bx"]), torch.from_numpy(g["nx"]))
    for got, key in zip(out, ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")):
        assert np.array_equal(got.cpu().numpy(), g[key]), key


def test_prepare_none_when_nothing_in_range(pkg):
    g = load("tiny_bev_z1")
    out = pkg.voxel_pooling_prepare_v2(cu(g["coor"] + np.float32(1e4)), torch.from_numpy(g["dx"]),
                                       torch.from_numpy(g["bx"]), torch.from_numpy(g["nx"]))
    assert out == (None,) * 5
    empty = torch.zeros((0, 6, 59, 4, 11, 3), device=DEV)
    assert pkg.voxel_pooling_prepare_v2(empty, torch.from_numpy(g["dx"]), torch.from_numpy(g["bx"]),
                                        torch.from_numpy(g["nx"])) == (None,) * 5


@pytest.mark.parametrize("step", [0.8, 0.512, 0.4, 0.3, 0.5, 1.7])
def test_voxel_boundaries_bit_exact(pkg, orc, step):
    """The voxel index is trunc((c - lo) / dx) with an IEEE divide (cam_stream_lss_bevpoolv2.py:317-318). The kernels
    replace the divide by a multiply with fl(1/dx) outside a guard band around the integers (csrc/common.cuh
    voxel_index): coordinates ON every voxel boundary and 1..3 ulps either side of it — where a reciprocal multiply
    alone flips indices — must still give the reference's ranks bit for bit, from materialised coordinates and through
    the fused geometry (identity camera, so the frustum values ARE the coordinates)."""
    n = 200
    lo_edge = np.float32(-0.5 * n * step)
    dx = torch.tensor([step, step, 8.0])
    bx = torch.tensor([float(lo_edge) + step / 2.0, float(lo_edge) + step / 2.0, -1.0])
    nx = torch.tensor([n, n, 1])
    lo32 = (bx - dx / 2.0).numpy()[0]
    ks = np.arange(-2, n + 3, dtype=np.float64)
    base = (np.float64(lo32) + ks * np.float64(np.float32(step))).astype(np.float32)
    xs = [base]
    for _ in range(3):
        xs.append(np.nextafter(xs[-1], np.float32(np.inf)))
    lo_side = base
    for _ in range(3):
        lo_side = np.nextafter(lo_side, np.float32(-np.inf))
        xs.append(lo_side)
    xs = np.concatenate(xs + [base + np.float32(step * 0.37)]).astype(np.float32)     # + generic interior points
    W = xs.size
    coor = np.zeros((1, 1, 2, 1, W, 3), np.float32)
    coor[0, 0, 0, 0, :, 0] = xs                      # adversarial x, fixed y
    coor[0, 0, 0, 0, :, 1] = np.float32(0.1)
    coor[0, 0, 1, 0, :, 0] = np.float32(0.1)         # fixed x, adversarial y
    coor[0, 0, 1, 0, :, 1] = xs
    want = orc.prepare_v2(coor, dx.numpy(), bx.numpy(), nx.numpy())
    got = pkg.voxel_pooling_prepare_v2(cu(coor), dx, bx, nx)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)
    # a bare reciprocal multiply is NOT equivalent on these inputs (the test would be vacuous otherwise)
    t = coor[..., 0] - lo32
    assert np.any((t * np.float32(1.0 / np.float32(step))).astype(np.int64) != (t / np.float32(step)).astype(np.int64)) \
        or step in (0.5,)
    # fused geometry: rots = I, trans = 0, frustum = (x, y, 1) -> coor == frustum values
    vt = pkg.view_transform

    class V:
        pass
    v = V()
    v.dx, v.bx, v.nx = dx, bx, nx
    fr = np.zeros((2, 1, W, 3), np.float32)
    fr[..., 2] = 1.0
    fr[0, 0, :, 0], fr[0, 0, :, 1] = xs, np.float32(0.1)
    fr[1, 0, :, 0], fr[1, 0, :, 1] = np.float32(0.1), xs
    v.frustum = cu(fr)
    C = 8
    out = torch.empty((1, 1, n, n, C), device=DEV)
    pr = vt._view_forward_scatter(torch.ones((1, 1, 2, 1, W), device=DEV), torch.ones((1, 1, 1, W, C), device=DEV), out, v,
                                  torch.eye(3, device=DEV).view(1, 1, 3, 3), torch.zeros((1, 1, 3), device=DEV), 1, 1, 2, 1, W, C,
                                  1, n, pkg._lib.LAYOUT_BZYXC)
    rank = orc.voxel_rank(coor, dx.numpy(), bx.numpy(), nx.numpy()).reshape(-1)
    assert np.array_equal(pr.point_rank.cpu().numpy().astype(np.int64), rank)


def test_prepare_truncation_nan_inf(pkg, orc):
    dx = torch.tensor([1., 1., 1.])
    bx = torch.tensor([.5, .5, .5])
    nx = torch.tensor([4, 4, 1])
    pts = np.array([[-0.5, 0.2, 0.0], [-1.0, 0.2, 0.0], [3.999, 3.5, 0.5], [4.0, 0, 0], [np.nan, 0, 0],
                    [np.inf, 0, 0], [1e30, 1, 0], [2.5, 2.5, 0.99]], np.float32).reshape(1, 1, 8, 1, 1, 3)
    rb, rd, rf, st, ln = pkg.voxel_pooling_prepare_v2(cu(pts), dx, bx, nx)
    assert rd.tolist() == [0, 7, 2] and rb.tolist() == [0, 10, 15] and rf.tolist() == [0, 0, 0]
    assert st.tolist() == [0, 1, 2] and ln.tolist() == [1, 1, 1]
    want = orc.prepare_v2(pts[:, :, [0, 1, 2, 3, 7]], dx.numpy(), bx.numpy(), nx.numpy())     # finite subset on the CPU
    assert want[0].tolist() == [0, 10, 15]


@pytest.mark.parametrize("cfg_name,B", [("bevdet_r50_b8", 8), ("occ_200x200x16_b64", 3), ("bevdepth_hires_b16", 2)])
def test_prepare_full_size_vs_oracle(pkg, orc, cfg_name, B):
    """BASELINE.json sizes (frame count bounded so the numpy oracle stays in seconds): bit-exact against
    the oracle, both from materialised coor and with the geometry fused."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    view, rots, trans, coor, want, *_ = synth_pool_case(pkg, orc, cfg, B)
    view = view.to(DEV)
    got_coor = view.get_geometry(rots.to(DEV), trans.to(DEV))
    assert np.array_equal(got_coor.cpu().numpy(), coor)
    got = view.voxel_pooling_prepare_v2(got_coor)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)
    # fused geometry (coor never materialised)
    from importlib import import_module
    vt = pkg.view_transform
    pr = vt._prepare_device(None, view.frustum, rots.to(DEV), trans.to(DEV), B, cfg.n_cams, view.D, view.fH, view.fW,
                            view.dx, view.bx, view.nx, torch.device(DEV))
    P, I = pr.counts.tolist()
    assert (P, I) == (want[0].size, want[3].size)
    for a, b in zip((pr.rb[:P], pr.rd[:P], pr.rf[:P], pr.starts[:I], pr.lengths[:I]), want):
        assert np.array_equal(a.cpu().numpy(), b)
    # inverse table: voxel rank of every frustum point, -1 when dropped
    rank = orc.voxel_rank(coor, view.dx.numpy(), view.bx.numpy(), view.nx.numpy())
    assert np.array_equal(pr.point_rank.cpu().numpy().astype(np.int64), rank)


def test_prepare_properties_omnihd_shape(pkg):
    """OmniHD shape (cfg 4), 2 frames, 23 M frustum points: size-independent properties."""
    cfg = pkg.synthetic.CONFIGS["rcfusion_omnihd_b32"]
    B = 2
    view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=5)
    coor = view.get_geometry(rots.to(DEV), trans.to(DEV))
    rb, rd, rf, st, ln = view.voxel_pooling_prepare_v2(coor)
    P, I = rb.numel(), st.numel()
    V = int(view.nx.prod())
    assert 0 < I <= P < coor.numel() // 3
    assert bool((rb[1:] >= rb[:-1]).all()) and int(rb.min()) >= 0 and int(rb.max()) < B * V       # sortedness
    assert bool(((rd[1:] > rd[:-1]) | (rb[1:] != rb[:-1])).all())                                 # stable ties
    assert torch.unique(rd).numel() == P                                                          # a permutation of kept points
    assert int(ln.sum()) == P and int(st[0]) == 0 and bool((st[1:] == st[:-1] + ln[:-1]).all())     # intervals partition
    assert torch.unique(rb).numel() == I and bool((rb[st.long()][1:] > rb[st.long()][:-1]).all())
    dhw, hw = view.D * view.fH * view.fW, view.fH * view.fW
    assert torch.equal(rf, (rd // dhw) * hw + rd % hw)
    # recompute the voxel of every kept point with torch ops on the device (reference expression)
    lo = (view.bx - view.dx / 2.).to(DEV)
    vox = ((coor.view(-1, 3)[rd.long()] - lo) / view.dx.to(DEV)).long()
    frame = rd.long() // (coor.numel() // 3 // B)
    nx = view.nx.to(DEV)
    assert torch.equal(rb.long(), frame * V + vox[:, 2] * nx[1] * nx[0] + vox[:, 1] * nx[0] + vox[:, 0])
    # idempotence
    again = view.voxel_pooling_prepare_v2(coor)
    assert all(torch.equal(a, b) for a, b in zip((rb, rd, rf, st, ln), again))


@pytest.mark.parametrize("cfg_name,B", [("bevdet_r50_b8", 3), ("occ_200x200x16_b64", 1), ("tiny", 1)])
def test_prepare_early_host_counts_equal_device_counts(pkg, cfg_name, B, monkeypatch):
    """bevpool_prepare_v2_counts hands (P, I) to the host from the rank kernel's kept count and voxel byte map, before
    the sort has run: they must equal what the sort + segmentation leave in counts_dev, and the late read-back
    (BEVPOOL_LATE_COUNTS=1) must return the same tensors."""
    vt = pkg.view_transform
    dev = torch.device(DEV)
    if cfg_name == "tiny":      # a handful of points incl. NaN / Inf / out of range, and a case with nothing kept
        dx, bx, nx = torch.tensor([1., 1., 1.]), torch.tensor([.5, .5, .5]), torch.tensor([4, 4, 1])
        pts = np.array([[-0.5, 0.2, 0.0], [3.999, 3.5, 0.5], [4.0, 0, 0], [np.nan, 0, 0], [np.inf, 0, 0], [2.5, 2.5, 0.99],
                        [2.6, 2.4, 0.5], [0.1, 0.1, 0.1]], np.float32).reshape(1, 1, 8, 1, 1, 3)
        cases = [(cu(pts), dx, bx, nx, (1, 1, 8, 1, 1)), (cu(pts * 0 - 7), dx, bx, nx, (1, 1, 8, 1, 1))]
    else:
        cfg = pkg.synthetic.CONFIGS[cfg_name]
        view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
        rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=11, roll=0.05)
        coor = view.get_geometry(rots.to(DEV), trans.to(DEV))
        cases = [(coor, view.dx, view.bx, view.nx, tuple(coor.shape[:5]))]
    for coor, dx, bx, nx, shp in cases:
        pr = vt._prepare_device(coor, None, None, None, *shp, dx, bx, nx, dev, host_counts=True)
        assert tuple(pr.host_counts) == tuple(pr.counts.tolist())
        early = pkg.voxel_pooling_prepare_v2(coor, dx, bx, nx)
        monkeypatch.setenv("BEVPOOL_LATE_COUNTS", "1")
        late = pkg.voxel_pooling_prepare_v2(coor, dx, bx, nx)
        monkeypatch.delenv("BEVPOOL_LATE_COUNTS")
        assert all((a is None and b is None) or torch.equal(a, b) for a, b in zip(early, late))
        if pr.host_counts[0] == 0:
            assert early[0] is None
    # a stream under capture cannot hand anything to the host: refused, not deadlocked
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        coor, dx, bx, nx, shp = cases[0]
        with pytest.raises(pkg._lib.BevPoolError):
            with torch.cuda.graph(g, stream=side):
                vt._prepare_device(coor, None, None, None, *shp, dx, bx, nx, dev, host_counts=True)
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------- forward
def _forward_cases(pkg, orc, name_or_cfg, B=None, C=None, seed=0):
    if isinstance(name_or_cfg, str) and name_or_cfg.startswith(("tiny", "mid")):
        g = load(name_or_cfg)
        feat_cl = np.ascontiguousarray(g["feat"].transpose(0, 1, 3, 4, 2))
        X, Y, Z = (int(v) for v in g["nx"])
        ranks = canon_ties(g["ranks_bev"], g["ranks_depth"], g["ranks_feat"]) + (g["interval_starts"], g["interval_lengths"])
        return g["depth"], feat_cl, ranks, (g["depth"].shape[0], Z, Y, X, feat_cl.shape[-1])
    cfg = pkg.synthetic.CONFIGS[name_or_cfg]
    view, rots, trans, coor, ranks, depth, feat, gout = synth_pool_case(pkg, orc, cfg, B, seed)
    if C is not None:
        feat = torch.randn(B, cfg.n_cams, C, cfg.fH, cfg.fW, generator=torch.Generator().manual_seed(seed))
    X, Y, Z = (int(v) for v in view.nx)
    feat_cl = feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
    return depth.numpy(), feat_cl, ranks, (B, Z, Y, X, feat_cl.shape[-1])


@pytest.mark.parametrize("case", ["tiny_bev_z1", "tiny_occ_z16", "tiny_omnihd", "tiny_hires",
                                  ("bevdet_r50_b8", 2, None), ("bevdet_r50_b8", 1, 256), ("bevdet_r50_b8", 1, 4),
                                  ("bevdet_r50_b8", 1, 6), ("bevdepth_hires_b16", 1, None)])
def test_forward_vs_oracle_all_paths(pkg, orc, case):
    args = (case,) if isinstance(case, str) else case
    depth, feat_cl, (rb, rd, rf, st, ln), shape = _forward_cases(pkg, orc, *args)
    ref = orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, shape, st, ln)
    exact = orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, shape, st, ln, exact=True)
    t = [cu(a) for a in (depth, feat_cl, rd, rf, rb)] + [shape, cu(st), cu(ln)]
    # reference-contract kernel: same summation order and FMA as the reference -> bit-identical
    out = pkg.QuickCumsumCuda.apply(*t).cpu().numpy()
    assert out.shape == tuple(shape)
    assert np.array_equal(out, ref)
    assert rel_to_max(out, exact) <= TOL
    # fused dense kernel writing [B,C,Z,Y,X] (zero fill + pool + layout in one pass)
    bev = pkg.bev_pool_v2(*t)
    assert bev.shape == (shape[0], shape[4], shape[1], shape[2], shape[3]) and bev.is_contiguous()
    assert rel_to_max(bev.permute(0, 2, 3, 4, 1).cpu().numpy(), exact) <= TOL
    if shape[1] == 1:
        trt = pkg.TRTBEVPoolv2.apply(t[0][0], t[1][0], t[2], t[3], t[4], t[6], t[7], shape[2], shape[3]) \
            if shape[0] == 1 else None
        if trt is not None:
            assert trt.shape == (1, shape[2], shape[3], shape[4])
            assert rel_to_max(trt.cpu().numpy()[0], exact[0, 0]) <= TOL


def test_forward_long_intervals_and_garbage_output_buffer(pkg, orc):
    """Hand-built ranks: one interval of 5000 points, one of 97, many of 1; arbitrary (non-prepare)
    rank arrays; the dense kernel must overwrite a poisoned output completely."""
    rng = np.random.default_rng(3)
    n_feat, n_depth, C = 300, 9000, 64
    B, Z, Y, X = 2, 2, 5, 13          # 130 voxels per frame: partial last strip
    lens = np.array([5000, 97, 1, 1, 33, 96, 300, 1, 2, 64, 7], dtype=np.int32)
    vox = np.sort(rng.choice(B * Z * Y * X, size=lens.size, replace=False)).astype(np.int32)
    rb = np.repeat(vox, lens).astype(np.int32)
    P = rb.size
    rd = rng.integers(0, n_depth, P).astype(np.int32)
    rf = rng.integers(0, n_feat, P).astype(np.int32)
    st = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    depth = rng.random(n_depth, dtype=np.float32).reshape(1, 1, n_depth, 1, 1)
    feat = rng.standard_normal((1, 1, n_feat, 1, C), dtype=np.float32)
    shape = (B, Z, Y, X, C)
    exact = orc.bev_pool_v2_forward(depth, feat, rd, rf, rb, shape, st, lens, exact=True)
    t = [cu(a) for a in (depth, feat, rd, rf, rb)] + [shape, cu(st), cu(lens)]
    assert rel_to_max(pkg.QuickCumsumCuda.apply(*t).cpu().numpy(), exact) <= TOL
    junk = torch.full((B * C * Z * Y * X + 1024,), float("nan"), device=DEV)      # poison the allocator's next block
    del junk
    bev = pkg.bev_pool_v2(*t)
    assert torch.isfinite(bev).all()
    assert rel_to_max(bev.permute(0, 2, 3, 4, 1).cpu().numpy(), exact) <= TOL
    assert int((bev != 0).any(dim=1).sum()) == lens.size


def test_forward_pools_exactly_the_given_intervals(pkg, orc):
    """The reference kernel pools the intervals it is handed (bev_pool_cuda.cu:30-47): a SUBSET of the intervals, or
    intervals cut in two (the later piece overwrites the voxel), must give what the oracle gives for the same
    arrays — the fused voxel-walk must not be taken for them; out-of-range ranks never reach the fused path."""
    depth, feat_cl, (rb, rd, rf, st, ln), shape = _forward_cases(pkg, orc, "bevdet_r50_b8", 1)
    t = [cu(a) for a in (depth, feat_cl, rd, rf, rb)]
    # (a) every other interval only
    st_a, ln_a = st[::2].copy(), ln[::2].copy()
    ref = orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, shape, st_a, ln_a, exact=True)
    bev = pkg.bev_pool_v2(*t, shape, cu(st_a), cu(ln_a))
    assert rel_to_max(bev.permute(0, 2, 3, 4, 1).cpu().numpy(), ref) <= TOL
    trt = pkg.TRTBEVPoolv2.apply(t[0][0], t[1][0], t[2], t[3], t[4], cu(st_a), cu(ln_a), shape[2], shape[3])
    assert rel_to_max(trt.cpu().numpy()[0], ref[0, 0]) <= TOL
    # (b) the longest interval cut in two: both pieces write the same voxel (a race in the reference kernel too, one
    #     thread per interval and channel) — that voxel must hold ONE of the two partial sums, all others the full sum
    k = int(np.argmax(ln))
    assert ln[k] >= 2
    st_b = np.concatenate([st[:k + 1], [st[k] + ln[k] // 2], st[k + 1:]]).astype(np.int32)
    ln_b = np.concatenate([ln[:k], [ln[k] // 2, ln[k] - ln[k] // 2], ln[k + 1:]]).astype(np.int32)
    full = orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, shape, st, ln, exact=True)
    bev_b = pkg.bev_pool_v2(*t, shape, cu(st_b), cu(ln_b)).permute(0, 2, 3, 4, 1).cpu().numpy()
    vox = np.unravel_index(rb[st[k]], shape[:4])
    halves = [orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, shape, st_b[j:j + 1], ln_b[j:j + 1], exact=True)[vox]
              for j in (k, k + 1)]
    assert min(rel_to_max(bev_b[vox], h) for h in halves) <= TOL
    bev_b[vox] = full[vox]
    assert rel_to_max(bev_b, full) <= TOL                        # every other voxel is the full result
    # (c) canonical intervals passed as fresh tensors (no plan attached) still take the fused walk and agree
    bev_c = pkg.bev_pool_v2(*t, shape, cu(st), cu(ln))
    assert rel_to_max(bev_c.permute(0, 2, 3, 4, 1).cpu().numpy(), full) <= TOL


def test_forward_matches_reference_cuda_kernel_bitwise(pkg, orc):
    """The reference's unmodified bev_pool_cuda.cu compiled for sm_100a (oracle/_ref) on the same inputs."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_bevpool_v2.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    ref = ctypes.CDLL(path)
    depth, feat_cl, (rb, rd, rf, st, ln), shape = _forward_cases(pkg, orc, "bevdet_r50_b8", 2)
    t = [cu(a) for a in (depth, feat_cl, rd, rf, rb, st, ln)]
    out_ref = torch.zeros(shape, device=DEV)
    torch.cuda.synchronize()
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    rc = ref.ref_bev_pool_v2_fwd(shape[4], st.size, p(t[0]), p(t[1]), p(t[2]), p(t[3]), p(t[4]), p(t[5]), p(t[6]), p(out_ref))
    torch.cuda.synchronize()
    assert rc == 0
    out = pkg.QuickCumsumCuda.apply(t[0], t[1], t[2], t[3], t[4], shape, t[5], t[6])
    assert torch.equal(out, out_ref)
    # and the backward, regrouped on the host the way bev_pool.py:47-57 does
    gout = torch.randn(shape, device=DEV)
    order = t[3].argsort(stable=True)
    rf_s, rd_s, rb_s = t[3][order].contiguous(), t[2][order].contiguous(), t[4][order].contiguous()
    kept = torch.ones(rf_s.numel(), dtype=torch.bool, device=DEV)
    kept[1:] = rf_s[1:] != rf_s[:-1]
    st_bp = torch.where(kept)[0].int()
    ln_bp = torch.diff(torch.cat([st_bp, torch.tensor([rf_s.numel()], device=DEV, dtype=torch.int32)])).int()
    dg_ref, fg_ref = torch.zeros_like(t[0]), torch.zeros_like(t[1])
    torch.cuda.synchronize()
    rc = ref.ref_bev_pool_v2_bwd(shape[4], st_bp.numel(), p(gout), p(t[0]), p(t[1]), p(rd_s), p(rf_s), p(rb_s), p(st_bp),
                                 p(ln_bp), p(dg_ref), p(fg_ref))
    torch.cuda.synchronize()
    assert rc == 0
    d = t[0].clone().requires_grad_()
    f = t[1].clone().requires_grad_()
    pkg.QuickCumsumCuda.apply(d, f, t[2], t[3], t[4], shape, t[5], t[6]).backward(gout)
    assert rel_to_max(d.grad.cpu().numpy(), dg_ref.cpu().numpy()) <= TOL
    assert rel_to_max(f.grad.cpu().numpy(), fg_ref.cpu().numpy()) <= TOL


# --------------------------------------------------------------------------------------- backward
@pytest.mark.parametrize("case", ["tiny_bev_z1", "tiny_occ_z16", ("bevdet_r50_b8", 2, None), ("bevdet_r50_b8", 1, 256),
                                  ("bevdet_r50_b8", 1, 6)])
def test_backward_general_path_vs_oracle(pkg, orc, case):
    """Arbitrary rank tensors (no plan attached): device regroup by ranks_feat + pixel-major kernel."""
    args = (case,) if isinstance(case, str) else case
    depth, feat_cl, (rb, rd, rf, st, ln), shape = _forward_cases(pkg, orc, *args)
    rng = np.random.default_rng(1)
    gout_cl = rng.standard_normal(shape).astype(np.float32)
    gd, gf = orc.bev_pool_v2_backward(gout_cl, depth, feat_cl, rd, rf, rb, exact=True)
    for fn, g in ((pkg.QuickCumsumCuda.apply, cu(gout_cl)),
                  (pkg.bev_pool_v2, cu(gout_cl).permute(0, 4, 1, 2, 3).contiguous())):
        d, f = cu(depth).requires_grad_(), cu(feat_cl).requires_grad_()
        out = fn(d, f, cu(rd), cu(rf), cu(rb), shape, cu(st), cu(ln))
        out.backward(g)
        assert rel_to_max(d.grad.cpu().numpy(), gd) <= TOL
        assert rel_to_max(f.grad.cpu().numpy(), gf) <= TOL


def _view_modes_vs_oracle(pkg, orc, cfg, B, dt, modes, seed=0, roll=0.0, pitch=0.0):
    """Run the public view-transform paths in dtype `dt` and compare with the float64 oracle fed the SAME
    (dt-rounded) inputs. fp32: 1e-5 of max (north_star); bf16: 2^-8 of max (one output rounding,
    fp32 accumulation inside)."""
    view, rots, trans, coor, (rb, rd, rf, st, ln), depth, feat, gout = synth_pool_case(pkg, orc, cfg, B, seed, roll, pitch)
    tol = TOL if dt == torch.float32 else TOL_BF16
    depth, feat, gout = (t.to(dt) for t in (depth, feat, gout))
    dq, fq, gq = depth.float(), feat.float(), gout.float()           # what the kernels actually read
    X, Y, Z = (int(v) for v in view.nx)
    C = cfg.channels
    feat_cl = fq.permute(0, 1, 3, 4, 2).contiguous().numpy()
    ref = orc.bev_pool_v2_forward(dq.numpy(), feat_cl, rd, rf, rb, (B, Z, Y, X, C), st, ln, exact=True)
    gd, gf = orc.bev_pool_v2_backward(gq.permute(0, 2, 3, 4, 1).contiguous().numpy(), dq.numpy(), feat_cl,
                                      rd, rf, rb, exact=True)
    rank = orc.voxel_rank(coor, view.dx.numpy(), view.bx.numpy(), view.nx.numpy())
    dropped = torch.from_numpy(rank < 0).to(DEV)
    view = view.to(DEV)
    for mode in modes:
        d = depth.to(DEV).requires_grad_()
        f = feat.to(DEV).requires_grad_()
        view.frame_groups = 1
        view.deterministic = "sorted" in mode      # default: sort-free scatter forward; True: sorted streaming forward
        if mode == "api":
            bev = view.voxel_pooling_v2(view.get_geometry(rots.to(DEV), trans.to(DEV)), d, f)
        else:
            if mode.endswith("groups"):      # independent frame groups on concurrent streams
                view.frame_groups = 2 if B % 2 == 0 else 1
            bev = view(d, f, rots.to(DEV), trans.to(DEV))
        assert bev.shape == (B, C, Z, Y, X) and bev.dtype == dt
        assert rel_to_max(bev.detach().float().permute(0, 2, 3, 4, 1).cpu().numpy(), ref) <= tol, mode
        bev.backward(gout.to(DEV))
        assert d.grad.dtype == dt and f.grad.dtype == dt
        assert rel_to_max(d.grad.float().cpu().numpy(), gd) <= tol, mode
        assert rel_to_max(f.grad.float().permute(0, 1, 3, 4, 2).cpu().numpy(), gf) <= tol, mode
        # dropped points get exactly zero depth gradient
        assert float(d.grad.reshape(-1)[dropped].abs().max()) == 0.0


ALL_MODES = ("api", "fused", "fused_groups", "fused_sorted", "fused_sorted_groups")


@pytest.mark.parametrize("cfg_name,B", [("bevdet_r50_b8", 8), ("occ_200x200x16_b64", 2), ("bevdepth_hires_b16", 1),
                                        ("rcfusion_omnihd_b32", 1)])
def test_view_transform_fwd_bwd_vs_oracle(pkg, orc, cfg_name, B):
    """Public API end to end at BASELINE sizes: get_geometry -> voxel_pooling_v2 (prepare + bev_pool_v2 with
    the sort-free backward) and the fully fused module forward; both against the float64 oracle, in the dtype the
    config names (cfg 3 is bf16 — the kernels bench.py runs for it)."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    dt = torch.bfloat16 if cfg.dtype == "bf16" else torch.float32
    _view_modes_vs_oracle(pkg, orc, cfg, B, dt, ALL_MODES)


@pytest.mark.parametrize("cfg_name,dt,roll,pitch", [("bevdet_r50_b8", torch.float32, 0.08, 0.03),
                                                     ("bevdet_r50_b8", torch.float32, 0.6, 0.0),
                                                     ("bevdepth_hires_b16", torch.bfloat16, 0.05, -0.04)])
def test_rolled_cameras_mixed_column_bins(pkg, orc, cfg_name, dt, roll, pitch):
    """Cameras with roll / pitch (real calibrations): the 16 pixels of an image column no longer share one voxel per
    depth bin, so the column kernels' mixed-bin slow paths carry real work (the SURVEY ring never reaches them).
    Every public path against the oracle."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    view, rots, trans, coor, ranks, *_ = synth_pool_case(pkg, orc, cfg, 1, 6, roll, pitch)
    rank = np.asarray(orc.voxel_rank(coor, view.dx.numpy(), view.bx.numpy(), view.nx.numpy())).reshape(coor.shape[1:5])   # [N, D, H, W]
    kept = rank >= 0
    lead = np.where(kept, rank, -1).max(axis=2, keepdims=True)
    assert ((rank != lead) & kept).sum() > 0.02 * kept.sum()          # a real share of the points sits in mixed bins
    _view_modes_vs_oracle(pkg, orc, cfg, 2, dt, ALL_MODES, seed=6, roll=roll, pitch=pitch)


@pytest.mark.parametrize("cfg_name,B", [("occ_200x200x16_b64", 2), ("rcfusion_omnihd_small", 2)])
def test_column_backward_forced_on_z16_grids(pkg, orc, cfg_name, B, monkeypatch):
    """BEVPOOL_BWD_KERNEL=joint forces the column-GEMM backward onto Z = 16 grids, where almost every (column, bin)
    holds several voxels: its slow path (one extra item per additional voxel) must give the same gradients as the
    block kernels that are the default there."""
    import dataclasses
    C = pkg.synthetic.CONFIGS
    cfg = C[cfg_name] if cfg_name in C else dataclasses.replace(C["rcfusion_omnihd_b32"], final_dim=(136, 240))
    monkeypatch.setenv("BEVPOOL_BWD_KERNEL", "joint")
    _view_modes_vs_oracle(pkg, orc, cfg, B, torch.float32, ("fused", "api"), seed=2)


def _bf16_cases(pkg):
    import dataclasses
    C = pkg.synthetic.CONFIGS
    return {
        # cfg 3 shapes (C = 80, Z = 1, D = 118, 32 x 88 features): view_fwd<bf16,20>, acc_layout<bf16>, joint<bf16,20>
        "hires_b2": (C["bevdepth_hires_b16"], 2),
        # cfg 2 shapes in bf16, 4 frames (frame groups of 2)
        "r50_b4": (dataclasses.replace(C["bevdet_r50_b8"], dtype="bf16"), 4),
        # C = 64 on a Z = 16 grid (RCFusion grid, smaller image): block_half<bf16,64> backward, two lane groups forward
        "omnihd_small": (dataclasses.replace(C["rcfusion_omnihd_b32"], final_dim=(136, 240), dtype="bf16"), 2),
        # C = 32 on the occupancy grid: block_half<bf16,32>, four lane groups
        "occ_b2": (dataclasses.replace(C["occ_200x200x16_b64"], dtype="bf16"), 2),
        # C = 32 / 64 / 128 on a Z = 1 grid: the other joint-backward instantiations
        "r50_c32": (dataclasses.replace(C["bevdet_r50_b8"], channels=32, dtype="bf16"), 2),
        "r50_c64": (dataclasses.replace(C["bevdet_r50_b8"], channels=64, dtype="bf16"), 2),
        "r50_c128": (dataclasses.replace(C["bevdet_r50_b8"], channels=128, dtype="bf16"), 2),
    }


@pytest.mark.parametrize("case", ["hires_b2", "r50_b4", "omnihd_small", "occ_b2", "r50_c32", "r50_c64", "r50_c128"])
def test_fused_path_bf16_vs_oracle(pkg, orc, case):
    """bf16 in / bf16 out through the FUSED module path (the kernels bench.py times for cfg 3) and the API
    sequence, frame groups 1 and 2, both `deterministic` settings: forward and both gradients against the float64
    oracle fed the bf16-rounded inputs, tolerance 2^-8 of max|ref|."""
    cfg, B = _bf16_cases(pkg)[case]
    _view_modes_vs_oracle(pkg, orc, cfg, B, torch.bfloat16, ALL_MODES, seed=3)


@pytest.mark.parametrize("cfg_name,B,dt", [("bevdet_r50_b8", 2, torch.float32), ("occ_200x200x16_b64", 1, torch.float32),
                                           ("bevdet_r50_b8", 2, torch.bfloat16)])
def test_bev_pool_v2_takes_the_permuted_view_the_reference_passes(pkg, cfg_name, B, dt):
    """cam_stream_lss_bevpoolv2.py:282 passes `feat.permute(0, 1, 3, 4, 2)` (a view of the neck's [B,N,C,H,W]) and lets
    bev_pool.py:20 copy it. Our bev_pool_v2 recognises that view (own transpose kernel, gradient written back in
    [B,N,C,H,W]): values and gradients must equal the contiguous-copy route bit for bit."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
    rots, trans = pkg.synthetic.camera_ring(B, cfg.n_cams, cfg.final_dim, seed=4)
    depth, feat, gout = (t.to(DEV, dt) for t in pkg.synthetic.pool_inputs(cfg, batch=B, seed=4))
    ranks = view.voxel_pooling_prepare_v2(view.get_geometry(rots.to(DEV), trans.to(DEV)))
    rb, rd, rf, st, ln = ranks
    X, Y, Z = (int(v) for v in view.nx)
    shape = (B, Z, Y, X, cfg.channels)
    res = []
    for as_view in (True, False):
        d, f = depth.clone().requires_grad_(), feat.clone().requires_grad_()
        fcl = f.permute(0, 1, 3, 4, 2)
        bev = pkg.bev_pool_v2(d, fcl if as_view else fcl.contiguous(), rd, rf, rb, shape, st, ln)
        bev.backward(gout)
        assert f.grad.shape == feat.shape
        res.append((bev.detach(), d.grad, f.grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_linearity_and_checksum_full_size(pkg):
    """cfg 2 full size, no oracle: pool(a*d1 + d2, f) == a*pool(d1,f) + pool(d2,f); sum over the grid equals
    the sum over kept points of depth * sum_c feat (a checksum of checksums)."""
    cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
    B = cfg.batch
    view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
    rots, trans = pkg.synthetic.camera_ring(B, 6, cfg.final_dim, seed=11)
    depth, feat, _ = pkg.synthetic.pool_inputs(cfg, seed=11)
    d1, f = depth.to(DEV), feat.to(DEV)
    d2 = torch.rand_like(d1)
    coor = view.get_geometry(rots.to(DEV), trans.to(DEV))
    p1 = view.voxel_pooling_v2(coor, d1, f)
    p2 = view.voxel_pooling_v2(coor, d2, f)
    p3 = view.voxel_pooling_v2(coor, 0.5 * d1 + d2, f)
    assert rel_to_max((0.5 * p1 + p2).cpu().numpy(), p3.cpu().numpy()) <= TOL
    rb, rd, rf, st, ln = view.voxel_pooling_prepare_v2(coor)
    f_cl = f.permute(0, 1, 3, 4, 2).reshape(-1, cfg.channels)
    want = (d1.reshape(-1)[rd.long()].double() * f_cl.double().sum(1)[rf.long()]).sum()
    got = p1.double().sum()
    assert abs(float(got - want)) <= 1e-6 * float((d1.reshape(-1)[rd.long()].double() * f_cl.double().abs().sum(1)[rf.long()]).sum())


def test_bf16_io(pkg, orc):
    """bf16 in / bf16 out (explicit extension): compared with the fp32 oracle fed the bf16-rounded inputs."""
    depth, feat_cl, (rb, rd, rf, st, ln), shape = _forward_cases(pkg, orc, "bevdet_r50_b8", 2)
    d16, f16 = cu(depth).bfloat16(), cu(feat_cl).bfloat16()
    dq, fq = d16.float().cpu().numpy(), f16.float().cpu().numpy()
    ref = orc.bev_pool_v2_forward(dq, fq, rd, rf, rb, shape, st, ln, exact=True)
    rng = np.random.default_rng(2)
    g16 = cu(rng.standard_normal((shape[0], shape[4]) + tuple(shape[1:4])).astype(np.float32)).bfloat16()
    gq = g16.float().permute(0, 2, 3, 4, 1).contiguous().cpu().numpy()
    gd, gf = orc.bev_pool_v2_backward(gq, dq, fq, rd, rf, rb, exact=True)
    d = d16.clone().requires_grad_()
    f = f16.clone().requires_grad_()
    bev = pkg.bev_pool_v2(d, f, cu(rd), cu(rf), cu(rb), shape, cu(st), cu(ln))
    assert bev.dtype == torch.bfloat16
    assert rel_to_max(bev.detach().float().permute(0, 2, 3, 4, 1).cpu().numpy(), ref) <= TOL_BF16
    bev.backward(g16)
    assert d.grad.dtype == torch.bfloat16 and f.grad.dtype == torch.bfloat16
    assert rel_to_max(d.grad.float().cpu().numpy(), gd) <= TOL_BF16
    assert rel_to_max(f.grad.float().cpu().numpy(), gf) <= TOL_BF16


def test_grid_transpose_roundtrip(pkg):
    x = torch.randn(3, 80, 2, 7, 45, device=DEV)
    cl = torch.empty(3, 2, 7, 45, 80, device=DEV)
    pkg.bev_pool._launch_transpose(x, cl, 3, 80, 2 * 7 * 45, True)
    assert torch.equal(cl, x.permute(0, 2, 3, 4, 1).contiguous())
    back = torch.empty_like(x)
    pkg.bev_pool._launch_transpose(cl, back, 3, 80, 2 * 7 * 45, False)
    assert torch.equal(back, x)


def test_errors_are_loud(pkg):
    g = load("tiny_bev_z1")
    d, f = cu(g["depth"]), cu(np.ascontiguousarray(g["feat"].transpose(0, 1, 3, 4, 2)))
    rb, rd, rf, st, ln = (cu(g[k]) for k in ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths"))
    with pytest.raises(ValueError):
        pkg.bev_pool_v2(d, f, rd, rf[:-1], rb, (2, 1, 128, 128, 8), st, ln)
    with pytest.raises(ValueError):
        pkg.bev_pool_v2(d, f, rd, rf, rb, (2, 1, 128, 128, 16), st, ln)
    with pytest.raises(ValueError):
        pkg.bev_pool_v2(d, f, rd, rf, rb.cpu(), (2, 1, 128, 128, 8), st, ln)
    lib = pkg._lib.load()
    assert lib.bevpool_v2_forward(0, 0, 0, 0, 0, 0, 0, 0, 5, 5, 8, 0, 0) == -1
    assert lib.bevpool_v2_forward(0, 0, 0, 0, 0, 0, 0, 0, 5, 5, 0, 0, 0) == -2
    assert lib.bevpool_v2_forward_dense(16, 16, 16, 16, 0, 16, 16, 5, 0, 6, 1, 8, 8, 0, 0, 1, 0, 0, 0, 0) == -2     # C % 4
    assert lib.bevpool_v2_forward_dense(16, 16, 16, 16, 0, 16, 16, 5, 0, 8, 1, 8, 8, 0, 0, 1, 0, 0, 0, 0) == -1     # no ranks_feat, no dims


# --------------------------------------------------------------------------------------- v1 op
@pytest.mark.parametrize("C,N", [(8, 6000), (80, 50000), (6, 3000)])
def test_v1_bev_pool_vs_oracle_and_reference_kernel(pkg, orc, C, N):
    rng = np.random.default_rng(C)
    B, D, H, W = 2, 3, 40, 36
    feats = rng.standard_normal((N, C)).astype(np.float32)
    coords = np.stack([rng.integers(0, H, N), rng.integers(0, W, N), rng.integers(0, D, N), rng.integers(0, B, N)], 1)
    ref, order, geom, starts, lengths = orc.bev_pool_v1(feats, coords, B, D, H, W)
    x = cu(feats).requires_grad_()
    out = pkg.bev_pool_v1.bev_pool(x, cu(coords), B, D, H, W)
    assert out.shape == (B, C, D, H, W) and out.is_contiguous()
    assert np.array_equal(out.detach().cpu().numpy(), ref)        # same sequential summation order: bit-identical
    og = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(cu(og))
    assert np.array_equal(x.grad.cpu().numpy(), orc.bev_pool_v1_backward(og, order, geom, starts, lengths, D, H, W))
    # the reference's unmodified v1 kernels (oracle/_ref), same sorted inputs
    path = os.path.join(ROOT, "oracle", "_ref", "libref_bevpool_v1.so")
    if os.path.exists(path):
        lib = ctypes.CDLL(path)
        xs, gs, st, ln = cu(feats[order]), cu(geom), cu(starts), cu(lengths)
        o_ref = torch.zeros((B, D, H, W, C), device=DEV)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        torch.cuda.synchronize()
        assert lib.ref_bev_pool_v1_fwd(B, D, H, W, N, C, starts.size, p(xs), p(gs), p(st), p(ln), p(o_ref)) == 0
        torch.cuda.synchronize()
        assert torch.equal(o_ref.permute(0, 4, 1, 2, 3), out.detach())


# ------------------------------------------------------------------------------------------ lift head (§8(f) rank 2)
def test_lift_head_vs_reference_golden(pkg, orc):
    """get_depth_feat against the reference's CamEncode output and gradients, both feature layouts."""
    g = load("lift_head")
    D, C = (int(v) for v in g["dims"])
    for cl in (False, True):
        x = cu(g["x"]).requires_grad_()
        depth, feat = pkg.get_depth_feat(x, D, C, channels_last=cl)
        f_nchw = feat.permute(0, 3, 1, 2) if cl else feat
        assert rel_to_max(depth.detach().cpu().numpy(), g["depth"]) <= TOL
        assert np.array_equal(f_nchw.detach().cpu().numpy(), g["feat"])
        gf = cu(g["feat_grad"])
        torch.autograd.backward([depth, feat], [cu(g["depth_grad"]), gf.permute(0, 2, 3, 1).contiguous() if cl else gf])
        assert rel_to_max(x.grad.cpu().numpy(), g["x_grad"]) <= TOL
    assert rel_to_max(pkg.get_depth_dist(cu(g["x"])).cpu().numpy(), orc.lift_head(g["x"], g["x"].shape[1], 0)[0]) <= TOL


@pytest.mark.parametrize("D,C,H,W,dt", [(59, 80, 16, 44, torch.float32), (118, 80, 32, 88, torch.bfloat16),
                                         (88, 32, 9, 13, torch.float32), (256, 4, 3, 5, torch.float32),
                                         (1, 132, 2, 33, torch.float32)])
def test_lift_head_shapes_vs_oracle(pkg, orc, D, C, H, W, dt):
    """BASELINE lift shapes, ragged pixel tails (H*W % 32 != 0), the D limit, extra trailing channels, bf16 io."""
    rng = np.random.default_rng(D * 1000 + C)
    BN = 5
    x = (rng.standard_normal((BN, D + C + 3, H, W)) * 4).astype(np.float32)
    xt = cu(x).to(dt)
    xin = xt.float().cpu().numpy()                                     # what the kernel actually reads
    depth_ref, feat_ref = orc.lift_head(xin, D, C)
    tol = TOL if dt == torch.float32 else 2 ** -8                      # bf16: one rounding of outputs in [0, 1]
    xr = xt.clone().requires_grad_()
    depth, feat = pkg.get_depth_feat(xr, D, C, channels_last=True)
    assert depth.dtype == dt and feat.shape == (BN, H, W, C)
    assert rel_to_max(depth.detach().float().cpu().numpy(), depth_ref) <= tol
    assert np.array_equal(feat.detach().float().permute(0, 3, 1, 2).cpu().numpy(), feat_ref)
    gd = rng.standard_normal(depth_ref.shape).astype(np.float32)
    gf = rng.standard_normal(feat_ref.shape).astype(np.float32)
    gdt, gft = cu(gd).to(dt), cu(gf).to(dt)
    torch.autograd.backward([depth, feat], [gdt, gft.permute(0, 2, 3, 1).contiguous()])
    want = orc.lift_head_backward(depth.detach().float().cpu().numpy(), gdt.float().cpu().numpy(), gft.float().cpu().numpy())
    got = xr.grad.float().cpu().numpy()
    assert got.shape == x.shape and np.all(got[:, D + C:] == 0)        # unused trailing channels get zero gradient
    assert rel_to_max(got[:, :D + C], want) <= tol
    with pytest.raises(RuntimeError, match="bad argument"):
        pkg.get_depth_feat(cu(np.zeros((1, 300, 2, 2), np.float32)), 257, 4)


def test_lift_splat_equals_separate_stages(pkg):
    """lift_splat (fused head -> channels-last features -> fused pool, no transposes) gives bit-identical BEV grids
    and input gradients within fp32 summation noise of torch.softmax + slice + LSSViewTransform.forward."""
    cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
    B, N, C = 2, cfg.n_cams, cfg.channels
    view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=3)
    rots, trans = rots.to(DEV), trans.to(DEV)
    torch.manual_seed(5)
    x = torch.randn(B * N, view.D + C, view.fH, view.fW, device=DEV)
    gout = torch.randn(B, C, *[int(v) for v in view.nx.flip(0)], device=DEV)
    xa = x.clone().requires_grad_()
    bev_a, depth_a = view.lift_splat(xa, rots, trans, C)
    bev_a.backward(gout)
    xb = x.clone().requires_grad_()
    depth_b = xb[:, :view.D].softmax(dim=1)
    bev_b = view(depth_b.view(B, N, view.D, view.fH, view.fW), xb[:, view.D:].reshape(B, N, C, view.fH, view.fW), rots, trans)
    bev_b.backward(gout)
    assert rel_to_max(depth_a.detach().cpu().numpy(), depth_b.detach().cpu().numpy()) <= TOL
    assert rel_to_max(bev_a.detach().cpu().numpy(), bev_b.detach().cpu().numpy()) <= TOL
    assert rel_to_max(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) <= TOL


@pytest.mark.parametrize("cfg_name,B", [("occ_200x200x16_b64", 2), ("bevdet_r50_b8", 2)])
def test_fused_s2c_layout_is_bit_identical(pkg, cfg_name, B):
    """§8(f) rank 1: forward(..., s2c=True) writes [B, Z*C, Y, X] directly; it must equal s2c(forward(...)) bit for
    bit, and so must the gradients (same kernels, same summation order, only the layout pass differs)."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    view = pkg.LSSViewTransform.from_config(cfg).to(DEV)
    N, C = cfg.n_cams, cfg.channels
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=1)
    rots, trans = rots.to(DEV), trans.to(DEV)
    torch.manual_seed(2)
    depth = torch.randn(B, N, view.D, view.fH, view.fW, device=DEV).softmax(2)
    feat = torch.randn(B, N, C, view.fH, view.fW, device=DEV)
    X, Y, Z = (int(v) for v in view.nx)
    g = torch.randn(B, Z * C, Y, X, device=DEV)
    for det in (True, False):
        view.deterministic = det
        res = []
        for s2c in (False, True):
            d, f = depth.clone().requires_grad_(), feat.clone().requires_grad_()
            bev = view(d, f, rots, trans, s2c=s2c)
            if not s2c:
                bev = view.s2c(bev)
            assert bev.shape == (B, Z * C, Y, X)
            bev.backward(g)
            res.append((bev.detach(), d.grad, f.grad))
        for a, b in zip(*res):
            if det:
                assert torch.equal(a, b)
            else:   # sort-free forward: summation order across image columns is not fixed; backward is exact
                assert rel_to_max(a.cpu().numpy(), b.cpu().numpy()) <= TOL


@pytest.mark.parametrize("cfg_name,B", [("bevdet_r50_b8", 4), ("occ_200x200x16_b64", 2)])
def test_fused_channels_last_3d_is_bit_identical(pkg, cfg_name, B):
    """memory_format=torch.channels_last_3d: same values as the contiguous result (bit for bit), returned as a
    channels-last view; a channels-last gradient is consumed without a transpose, a contiguous one still works."""
    cfg = pkg.synthetic.CONFIGS[cfg_name]
    view = pkg.LSSViewTransform.from_config(cfg, frame_groups=2).to(DEV)
    N, C = cfg.n_cams, cfg.channels
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=4)
    rots, trans = rots.to(DEV), trans.to(DEV)
    torch.manual_seed(3)
    depth = torch.randn(B, N, view.D, view.fH, view.fW, device=DEV).softmax(2)
    feat = torch.randn(B, N, C, view.fH, view.fW, device=DEV)
    X, Y, Z = (int(v) for v in view.nx)
    g = torch.randn(B, C, Z, Y, X, device=DEV)
    for det in (True, False):
        view.deterministic = det
        d0, f0 = depth.clone().requires_grad_(), feat.clone().requires_grad_()
        ref = view(d0, f0, rots, trans)
        ref.backward(g)
        for g_in in (g.contiguous(memory_format=torch.channels_last_3d), g):
            d, f = depth.clone().requires_grad_(), feat.clone().requires_grad_()
            bev = view(d, f, rots, trans, memory_format=torch.channels_last_3d)
            assert bev.shape == ref.shape and bev.is_contiguous(memory_format=torch.channels_last_3d)
            if det:
                assert torch.equal(bev, ref)
            else:
                assert rel_to_max(bev.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TOL
            bev.backward(g_in)
            assert torch.equal(d.grad, d0.grad) and torch.equal(f.grad, f0.grad)


@pytest.mark.parametrize("name", ["tiny_bev_z1", "tiny_occ_z16", "tiny_omnihd", "tiny_hires"])
def test_scatter_forward_ranks_bit_exact_vs_reference_golden(pkg, orc, name):
    """Sort-free forward: the ranks it computes in-kernel (point_rank, the table the backward walks) must be the
    reference's voxel ranks bit for bit, and its BEV grid must match the float64 oracle on the golden geometry."""
    g = load(name)
    B, N, D, H, W = g["coor"].shape[:5]
    dx, bx, nx = (torch.from_numpy(g[k]) for k in ("dx", "bx", "nx"))
    X, Y, Z = (int(v) for v in nx)
    C = 8
    vt = pkg.view_transform

    class V:            # the attributes _view_forward_scatter reads
        pass
    v = V()
    v.dx, v.bx, v.nx, v.frustum = dx, bx, nx, cu(g["frustum"])
    rng = np.random.default_rng(1)
    depth = rng.random((B, N, D, H, W), dtype=np.float32)
    feat_cl = rng.standard_normal((B, N, H, W, C)).astype(np.float32)
    # the stored cumsum_pooled of the tiny goldens is the REFERENCE's own CPU pooling of its own depth/feat
    if "cumsum_pooled" in g.files and g["feat"].shape[2] == C:
        depth, feat_cl = g["depth"], np.ascontiguousarray(g["feat"].transpose(0, 1, 3, 4, 2))
    for layout, shape in ((pkg._lib.LAYOUT_BZYXC, (B, Z, Y, X, C)), (pkg._lib.LAYOUT_BCZYX, (B, C, Z, Y, X))):
        out = torch.full(shape, 7.0, device=DEV)
        pr = vt._view_forward_scatter(cu(depth), cu(feat_cl), out, v, cu(g["rots"]), cu(g["trans"]), B, N, D, H, W, C,
                                      B, Z * Y, layout)
        want_rank = orc.voxel_rank(g["coor"], g["dx"], g["bx"], g["nx"]).reshape(-1)
        assert np.array_equal(pr.point_rank.cpu().numpy().astype(np.int64), want_rank)
        rb, rd, rf, st, ln = orc.prepare_v2(g["coor"], g["dx"], g["bx"], g["nx"])
        ref = orc.bev_pool_v2_forward(depth, feat_cl, rd, rf, rb, (B, Z, Y, X, C), st, ln, exact=True)
        got = out if layout == pkg._lib.LAYOUT_BZYXC else out.permute(0, 2, 3, 4, 1)
        assert rel_to_max(got.cpu().numpy(), ref) <= TOL
        if "cumsum_pooled" in g.files and g["feat"].shape[2] == C:      # cumsum is the inexact party: 1e-4
            assert rel_to_max(got.permute(0, 4, 1, 2, 3).cpu().numpy(), g["cumsum_pooled"]) <= 1e-4


def test_deterministic_switch(pkg):
    """deterministic=True (or torch.use_deterministic_algorithms) runs the sorted forward: bit-identical from run to
    run; the default sort-free forward agrees with it to fp32 summation noise; the backward is the same kernel."""
    cfg = pkg.synthetic.CONFIGS["bevdet_r50_b8"]
    B, N, C = 4, cfg.n_cams, cfg.channels
    view = pkg.LSSViewTransform.from_config(cfg, deterministic=True).to(DEV)
    rots, trans = pkg.synthetic.camera_ring(B, N, cfg.final_dim, seed=9)
    rots, trans = rots.to(DEV), trans.to(DEV)
    torch.manual_seed(9)
    depth = torch.randn(B, N, view.D, view.fH, view.fW, device=DEV).softmax(2)
    feat = torch.randn(B, N, C, view.fH, view.fW, device=DEV)
    a, b = view(depth, feat, rots, trans), view(depth, feat, rots, trans)
    assert torch.equal(a, b)
    view.deterministic = False
    c = view(depth, feat, rots, trans)
    assert rel_to_max(c.cpu().numpy(), a.cpu().numpy()) <= TOL
    torch.use_deterministic_algorithms(True)
    try:
        d = view(depth, feat, rots, trans)
    finally:
        torch.use_deterministic_algorithms(False)
    assert torch.equal(d, a)


# ------------------------------------------------------------------------------------------ pillar scatter (§8(f) rank 4)
@pytest.mark.parametrize("B,C,ny,nx,P,dt", [(2, 64, 320, 480, 40000, torch.float32), (1, 64, 320, 480, 0, torch.float32),
                                            (3, 7, 5, 13, 150, torch.float32), (2, 64, 33, 65, 3000, torch.bfloat16)])
def test_pillar_scatter_vs_oracle(pkg, orc, B, C, ny, nx, P, dt):
    """PointPillarsScatter drop-in at the RCFusion config size (64 channels, 320x480, 40 000 pillars), empty input,
    duplicated cells (last pillar wins), out-of-range pillars (ignored), bf16 io; backward vs the oracle."""
    rng = np.random.default_rng(B * 100 + C)
    cells = rng.integers(0, B * ny * nx, size=P) if P and P > B * ny * nx // 2 else rng.permutation(B * ny * nx)[:P]
    coors = np.stack([cells // (ny * nx), rng.integers(0, 3, size=P), (cells % (ny * nx)) // nx, cells % nx], 1).astype(np.int32)
    feats = rng.standard_normal((P, C)).astype(np.float32)
    ft = cu(feats).to(dt).requires_grad_()
    fin = ft.detach().float().cpu().numpy()
    m = pkg.pillar_scatter.PointPillarsScatter(C, [ny, nx])
    out = m(ft, cu(coors), B)
    assert out.shape == (B, C, ny, nx) and out.dtype == dt
    assert np.array_equal(out.detach().float().cpu().numpy(), orc.pillar_scatter(fin, coors, B, ny, nx))
    g = rng.standard_normal((B, C, ny, nx)).astype(np.float32)
    gt = cu(g).to(dt)
    out.backward(gt)
    assert np.array_equal(ft.grad.float().cpu().numpy(), orc.pillar_scatter_backward(gt.float().cpu().numpy(), coors))
    if P:   # pillars outside the canvas are dropped, their gradient is zero
        bad = coors.copy()
        bad[0] = (B, 0, 0, 0)
        bad[-1] = (0, 0, ny, 0)
        f2 = cu(feats).to(dt).requires_grad_()
        out2 = m(f2, cu(bad), B)
        keep = np.ones(P, bool)
        keep[[0, P - 1]] = False
        assert np.array_equal(out2.detach().float().cpu().numpy(), orc.pillar_scatter(fin[keep], bad[keep], B, ny, nx))
        out2.backward(gt)
        assert float(f2.grad[0].abs().max()) == 0.0 and float(f2.grad[-1].abs().max()) == 0.0


@pytest.mark.parametrize("C", [4, 32, 64, 128, 132])
def test_fused_path_channel_counts_z1(pkg, orc, C):
    """Z = 1 grid with the channel counts that select the other lane mappings of the fused kernels (C <= 32: four lane
    groups, C <= 64: two, C = 128: one full-width group, C = 4: the generic block backward, C = 132 > 128: the sorted
    forward with the tile kernel and the channel-chunked block backward), fwd + bwd vs the oracle."""
    import dataclasses
    cfg = dataclasses.replace(pkg.synthetic.CONFIGS["bevdet_r50_b8"], channels=C)
    B = 2
    view, rots, trans, coor, (rb, rd, rf, st, ln), depth, feat, gout = synth_pool_case(pkg, orc, cfg, B, seed=5)
    X, Y, Z = (int(v) for v in view.nx)
    feat_cl = feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
    ref = orc.bev_pool_v2_forward(depth.numpy(), feat_cl, rd, rf, rb, (B, Z, Y, X, C), st, ln, exact=True)
    gd, gf = orc.bev_pool_v2_backward(gout.permute(0, 2, 3, 4, 1).contiguous().numpy(), depth.numpy(), feat_cl,
                                      rd, rf, rb, exact=True)
    view = view.to(DEV)
    for det in (False, True):
        view.deterministic = det
        d, f = depth.to(DEV).requires_grad_(), feat.to(DEV).requires_grad_()
        bev = view(d, f, rots.to(DEV), trans.to(DEV))
        assert rel_to_max(bev.detach().permute(0, 2, 3, 4, 1).cpu().numpy(), ref) <= TOL
        bev.backward(gout.to(DEV))
        assert rel_to_max(d.grad.cpu().numpy(), gd) <= TOL
        assert rel_to_max(f.grad.permute(0, 1, 3, 4, 2).cpu().numpy(), gf) <= TOL


@pytest.mark.parametrize("zb,C", [((-5.0, 3.0, 8.0), 80), ((-3.0, 5.0, 0.5), 64), ((-3.0, 5.0, 2.0), 12)])
def test_fused_path_ragged_feature_map(pkg, orc, zb, C):
    """Feature map 5 x 13 (neither a multiple of the 4 x 8 pixel block), D = 7 (padded bins), one camera: partial
    blocks, pad bins and idle warps of the scatter forward / joint / block backward kernels, vs the oracle."""
    cfg = pkg.synthetic.ViewConfig("ragged", (80, 208), 16, (1.0, 29.0, 4.0), (-25.6, 25.6, 0.8), (-25.6, 25.6, 0.8), zb, C, 3,
                                   n_cams=1)
    B = 3
    view, rots, trans, coor, (rb, rd, rf, st, ln), depth, feat, gout = synth_pool_case(pkg, orc, cfg, B, seed=8)
    assert (view.fH, view.fW, view.D) == (5, 13, 7)
    X, Y, Z = (int(v) for v in view.nx)
    feat_cl = feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
    ref = orc.bev_pool_v2_forward(depth.numpy(), feat_cl, rd, rf, rb, (B, Z, Y, X, C), st, ln, exact=True)
    gd, gf = orc.bev_pool_v2_backward(gout.permute(0, 2, 3, 4, 1).contiguous().numpy(), depth.numpy(), feat_cl,
                                      rd, rf, rb, exact=True)
    view = view.to(DEV)
    for det in (False, True):
        view.deterministic = det
        d, f = depth.to(DEV).requires_grad_(), feat.to(DEV).requires_grad_()
        bev = view(d, f, rots.to(DEV), trans.to(DEV))
        assert rel_to_max(bev.detach().permute(0, 2, 3, 4, 1).cpu().numpy(), ref) <= TOL
        bev.backward(gout.to(DEV))
        assert rel_to_max(d.grad.cpu().numpy(), gd) <= TOL
        assert rel_to_max(f.grad.permute(0, 1, 3, 4, 2).cpu().numpy(), gf) <= TOL


@pytest.mark.parametrize("fw,C,dt,roll", [(12, 80, torch.float32, 0.0), (28, 32, torch.float32, 0.3), (28, 64, torch.bfloat16, 0.0),
                                            (20, 80, torch.float32, 0.0)])
def test_column_backward_wide_and_narrow_tiles_small(pkg, orc, fw, C, dt, roll):
    """Small Z = 1 cases for both tile widths of the column backward: fW = 12 / 28 take the 16-column kernel (128-bit
    staging loads, one lane per (column, bin), raw ranks parked in the row ring; ragged last tile of 12 columns),
    fW = 20 the 8-column kernel (W % 16 leaves a tighter cover with 8-column tiles); fH = 6 (partial 16-row tile), D = 9
    (pad bins), a rolled camera (mixed bins -> slow path), bf16. BEVPOOL_BWD_TILE_W=8|16 (read once per process) forces
    either kernel onto every shape: profiles/sanitize.sh runs this test under both settings as well."""
    cfg = pkg.synthetic.ViewConfig("wide", (96, 16 * fw), 16, (1.0, 46.0, 5.0), (-25.6, 25.6, 0.8), (-25.6, 25.6, 0.8),
                                   (-5.0, 3.0, 8.0), C, 2, n_cams=2)
    B = 2
    view, rots, trans, coor, (rb, rd, rf, st, ln), depth, feat, gout = synth_pool_case(pkg, orc, cfg, B, seed=9, roll=roll)
    assert (view.fH, view.fW, view.D) == (6, fw, 9)
    depth, feat, gout = (t.to(dt).float() for t in (depth, feat, gout))       # the oracle sees the rounded inputs
    feat_cl = feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
    gd, gf = orc.bev_pool_v2_backward(gout.permute(0, 2, 3, 4, 1).contiguous().numpy(), depth.numpy(), feat_cl,
                                      rd, rf, rb, exact=True)
    tol = TOL if dt == torch.float32 else TOL_BF16
    view = view.to(DEV)
    d, f = depth.to(DEV, dt).requires_grad_(), feat.to(DEV, dt).requires_grad_()
    bev = view(d, f, rots.to(DEV), trans.to(DEV))
    bev.backward(gout.to(DEV, dt))
    assert rel_to_max(d.grad.float().cpu().numpy(), gd) <= tol
    assert rel_to_max(f.grad.float().permute(0, 1, 3, 4, 2).cpu().numpy(), gf) <= tol


# ------------------------------------------------------------------------------------------ cross-modal fusion (§8(f) rank 4)
def test_cross_modal_fusion_vs_reference_golden(pkg):
    """Cross_Modal_Fusion mirror with the reference's attention-conv weights: its `fuse` (everything before the final
    3x3 reduction conv) must reproduce the reference module's output and input gradients."""
    g = load("cross_modal")
    m = pkg.cross_modal.Cross_Modal_Fusion(kernel_size=3, img_channels=6, radar_channels=10, out_channels=4).to(DEV)
    assert sorted(k for k, _ in m.named_parameters()) == ['att_img.0.weight', 'att_radar.0.weight', 'reduce_mixBEV.conv.bias',
                                                           'reduce_mixBEV.conv.weight']
    with torch.no_grad():
        m.att_img[0].weight.copy_(cu(g["w_img"]))
        m.att_radar[0].weight.copy_(cu(g["w_radar"]))
    img, rad = cu(g["img"]).requires_grad_(), cu(g["radar"]).requires_grad_()
    out = m.fuse(img, rad)
    assert rel_to_max(out.detach().cpu().numpy(), g["out"]) <= TOL
    out.backward(cu(g["out_grad"]))
    assert rel_to_max(img.grad.cpu().numpy(), g["img_grad"]) <= TOL
    assert rel_to_max(rad.grad.cpu().numpy(), g["radar_grad"]) <= TOL
    assert m(img.detach(), rad.detach()).shape == (2, 4, 5, 9)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_cross_modal_glue_kernels_vs_oracle(pkg, orc, dt):
    """channel_avg_max / gate_concat at the RCFusion size (256 + 384 channels, 160 x 240 BEV), ties in the max
    (first index wins), fwd + bwd vs the oracle."""
    rng = np.random.default_rng(21)
    N, Ca, Cb, H, W = 2, 256, 384, 160, 240
    tol = TOL if dt == torch.float32 else 2 ** -7
    x = rng.standard_normal((N, Ca, H, W)).astype(np.float32)
    x[:, 7] = x[:, 3] = x.max(axis=1) + 1.0                                   # two maximal channels: 3 must win
    xt = cu(x).to(dt).requires_grad_()
    xin = xt.detach().float().cpu().numpy()
    out = pkg.cross_modal.channel_avg_max(xt)
    ref = orc.channel_avg_max(xin)
    assert rel_to_max(out.detach().float().cpu().numpy(), ref) <= tol
    gm = rng.standard_normal(ref.shape).astype(np.float32)
    gmt = cu(gm).to(dt)
    out.backward(gmt)
    assert rel_to_max(xt.grad.float().cpu().numpy(), orc.channel_avg_max_backward(xin, gmt.float().cpu().numpy())) <= tol
    b = rng.standard_normal((N, Cb, H, W)).astype(np.float32)
    wa, wb = rng.random((N, 1, H, W), dtype=np.float32), rng.random((N, 1, H, W), dtype=np.float32)
    ts = [cu(v).to(dt).requires_grad_() for v in (x, b, wa, wb)]
    ins = [t.detach().float().cpu().numpy() for t in ts]
    cat = pkg.cross_modal.gate_concat(*ts)
    assert rel_to_max(cat.detach().float().cpu().numpy(), orc.gate_concat(*ins)) <= tol
    gc = cu(rng.standard_normal((N, Ca + Cb, H, W)).astype(np.float32)).to(dt)
    cat.backward(gc)
    for t, want in zip(ts, orc.gate_concat_backward(gc.float().cpu().numpy(), *ins)):
        assert rel_to_max(t.grad.float().cpu().numpy(), want) <= (tol if dt == torch.float32 else 2 ** -6)
