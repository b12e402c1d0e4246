"""CPU (-m "not gpu"): the C-ABI library loads and exports every declared symbol (no compute
calls), host-side logic of the plugin mirror, frame sharding over gloo with world_size 2."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "bevpool_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bevpool_\w+)\s*\(", header))
    assert len(declared) >= 14
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bevpool_b200.h but not exported"
    assert declared == set(pkg._lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.bevpool_b200_abi_version() == 1
    assert lib.bevpool_b200_strerror(-3).decode() == "workspace too small"


def test_library_is_in_tree_and_sm100a_only(pkg):
    path = pkg._lib.library_path()
    assert path.startswith(ROOT) and os.path.exists(path)
    assert "compute_100a" in " ".join(pkg.build.FLAGS) and "-use_fast_math" not in pkg.build.FLAGS


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "omnihd-scenes_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), f"{fn} mentions the oracle"


def test_no_cpu_fallback(pkg):
    d = torch.rand(1, 1, 2, 2, 2)
    f = torch.ones(1, 1, 2, 2, 2)
    i = torch.zeros(4, dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.bev_pool_v2(d, f, i, i, i, (1, 1, 2, 2, 2), i[:1], i[:1])
    with pytest.raises(ValueError, match="CUDA"):
        pkg.voxel_pooling_prepare_v2(torch.zeros(1, 1, 2, 2, 2, 3), *pkg.gen_dx_bx((0, 4, 1), (0, 4, 1), (0, 1, 1)))
    with pytest.raises(ValueError, match="CUDA"):
        pkg.get_geometry(torch.zeros(2, 2, 2, 3), torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 3))


def test_view_transform_module_state(pkg):
    cfg = pkg.synthetic.CONFIGS["rcfusion_omnihd_b32"]
    v = pkg.LSSViewTransform.from_config(cfg)
    # frustum must stay in the state_dict with the reference's shape (checkpoint compatibility)
    sd = v.state_dict()
    assert list(sd) == ["frustum"] and tuple(sd["frustum"].shape) == (59, 136, 240, 3)
    assert v.nx.tolist() == [240, 160, 16] and v.D == 59 and (v.fH, v.fW) == (136, 240)
    v2 = pkg.LSSViewTransform.from_lss_args((544, 960), [1, 60, 1], [-60, -40, -3, 60, 40, 5], 4, 0.5)
    assert torch.equal(v2.frustum, v.frustum) and torch.equal(v2.bx, v.bx)
    x = torch.arange(2 * 3 * 4 * 5 * 6, dtype=torch.float32).view(2, 3, 4, 5, 6)
    assert torch.equal(pkg.LSSViewTransform.s2c(x), torch.cat(x.unbind(dim=2), 1))
    with pytest.raises(NotImplementedError):
        v.get_geometry(torch.zeros(1, 6, 3, 3), torch.zeros(1, 6, 3), post_rots=torch.zeros(1))


def test_config_shapes_match_survey(pkg):
    c = pkg.synthetic.CONFIGS
    assert (c["bevdet_r50_b8"].D, c["bevdet_r50_b8"].fH, c["bevdet_r50_b8"].fW) == (59, 16, 44)
    assert (c["bevdepth_hires_b16"].D, c["bevdepth_hires_b16"].fH, c["bevdepth_hires_b16"].fW) == (118, 32, 88)
    assert c["occ_200x200x16_b64"].D == 88
    rots, trans = pkg.synthetic.camera_ring(2, 6, (256, 704), seed=0)
    assert rots.shape == (2, 6, 3, 3) and trans.shape == (2, 6, 3) and rots.dtype == torch.float32
    r2, _ = pkg.synthetic.camera_ring(2, 6, (256, 704), seed=0)
    assert torch.equal(rots, r2)


def test_plugin_install_registers_reference_module_path(pkg):
    mod = pkg.plugin.install(force=True)
    import importlib
    m = importlib.import_module("projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool")
    assert m is mod and m.bev_pool_v2 is pkg.bev_pool_v2 and m.__all__ == ['bev_pool_v2', 'TRTBEVPoolv2']

    class FakeLSS:
        def get_geometry(self, *a, **k):
            return "orig"

        def voxel_pooling_prepare_v2(self, coor):
            return "orig"
    pkg.plugin.patch_lss_class(FakeLSS)
    assert FakeLSS._bevpool_b200_orig_prepare is not FakeLSS.voxel_pooling_prepare_v2
    # non-default transforms still go to the original implementation
    assert FakeLSS().get_geometry(None, None, post_rots=1) == "orig"
    del sys.modules["projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool"]


def test_plan_registry_is_identity_and_version_checked(pkg):
    bp = pkg.bev_pool
    t = [torch.zeros(4, dtype=torch.int32) for _ in range(5)]
    bp.register_plan(*t, point_rank=torch.zeros(8, dtype=torch.int32), bn=1, d=2, h=2, w=2)
    depth, feat = torch.zeros(8), torch.zeros(4, 3)
    assert bp._find_plan(*t, depth, feat) is not None
    assert bp._find_plan(t[0], t[1].clone(), t[2], t[3], t[4], depth, feat) is None      # different object
    t[1].add_(1)                                                                         # mutated in place
    assert bp._find_plan(*t, depth, feat) is None
    # the plan is an attribute of the ranks_bev tensor object: nothing global to leak
    assert not hasattr(bp, "_PLANS") and getattr(t[0], bp._PLAN_ATTR).point_rank.numel() == 8


def test_frame_shard_partition(pkg):
    sh = pkg.sharding
    for B in (1, 7, 8, 32, 64):
        for G in (1, 2, 4, 8):
            spans = [sh.frame_shard(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.frame_shard(8, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(4 * 3 * 2 * 5 * 5, dtype=torch.float32).view(4, 3, 2, 5, 5)
    (mine,) = pkg.sharding.shard_frames((full,), rank, world)
    got, work = pkg.sharding.all_gather_bev(mine, async_op=(rank == 0))
    if work is not None:
        work.wait()
    q.put((rank, bool(torch.equal(got, full)), tuple(mine.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_and_all_gather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True, (2, 3, 2, 5, 5)), (1, True, (2, 3, 2, 5, 5))]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU (it times the CPU port of the reference path) and prints one
    JSON line with the keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bev_pool_fwd_bwd_frames_per_s" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["workload"] == "bevdet_r50_b8"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_cross_modal_state_dict_keys_match_reference_construction(pkg):
    """The reference detector builds Cross_Modal_Fusion(kernel_size=3, norm_cfg=dict(type='BN', eps=1e-3,
    momentum=0.01)) (rcfusion_faster_rcnn.py:38,74), so `reduce_mixBEV` is mmcv ConvModule = conv(bias=False) + BN +
    ReLU with keys conv.weight / bn.* (mmcv v1.4.0 conv_module.py: bias='auto' -> not with_norm; norm name from
    build_norm_layer's abbreviation 'bn'). norm_cfg=None (the class default) keeps the conv bias."""
    cm = pkg.cross_modal
    m = cm.Cross_Modal_Fusion(kernel_size=3, norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01))
    assert sorted(m.state_dict()) == sorted([
        'att_img.0.weight', 'att_radar.0.weight', 'reduce_mixBEV.conv.weight', 'reduce_mixBEV.bn.weight',
        'reduce_mixBEV.bn.bias', 'reduce_mixBEV.bn.running_mean', 'reduce_mixBEV.bn.running_var',
        'reduce_mixBEV.bn.num_batches_tracked'])
    assert tuple(m.state_dict()['reduce_mixBEV.conv.weight'].shape) == (384, 640, 3, 3)
    assert m.reduce_mixBEV.bn.eps == 1e-3 and m.reduce_mixBEV.bn.momentum == 0.01
    plain = cm.Cross_Modal_Fusion(kernel_size=7)
    assert sorted(plain.state_dict()) == ['att_img.0.weight', 'att_radar.0.weight', 'reduce_mixBEV.conv.bias',
                                          'reduce_mixBEV.conv.weight']
    assert tuple(plain.att_img[0].weight.shape) == (1, 2, 7, 7)
    sync = cm.Cross_Modal_Fusion(norm_cfg=dict(type='SyncBN', requires_grad=False), img_channels=4, radar_channels=4,
                                 out_channels=4)
    assert isinstance(sync.reduce_mixBEV.bn, torch.nn.SyncBatchNorm) and not sync.reduce_mixBEV.bn.weight.requires_grad
    with pytest.raises(ValueError):
        cm.Cross_Modal_Fusion(norm_cfg=dict(type='LN'))
    # the BN module in eval mode is conv + affine + relu of the fused glue output (CPU-checkable part)
    x = torch.randn(2, 8, 5, 7)
    ref = torch.relu(torch.nn.functional.batch_norm(sync.reduce_mixBEV.conv(x), sync.reduce_mixBEV.bn.running_mean,
                                                    sync.reduce_mixBEV.bn.running_var, sync.reduce_mixBEV.bn.weight,
                                                    sync.reduce_mixBEV.bn.bias, False, 0.0, sync.reduce_mixBEV.bn.eps))
    assert torch.allclose(sync.reduce_mixBEV.eval()(x), ref)


def test_patch_real_reference_class_cpu(pkg):
    """The UNMODIFIED reference LiftSplatShoot (imported from the staged files / the reference tree) is patched and
    un-patched; on a CPU box the patched methods must refuse loudly (no CPU fallback), the cached fused view must not
    show up in the module's state_dict, and the reference's own op module must import on the ctypes ext binding."""
    sys.path.insert(0, ROOT)
    from oracle import refimport as ri
    if ri.ref_root() is None:
        pytest.skip("reference Python files not staged")
    pkg.plugin.install(force=True)
    ref = ri.import_reference_lss("bevfusion")
    assert ref.bev_pool_v2 is pkg.bev_pool_v2
    lss = ri.make_reference_lss(ref, (64, 176), 8, (1.0, 60.0, 1.0), (-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8), (-5.0, 3.0, 8.0))
    keys = sorted(lss.state_dict())
    orig = ref.LiftSplatShoot.get_voxels
    pkg.plugin.patch_lss_class(ref.LiftSplatShoot)
    assert ref.LiftSplatShoot.get_voxels is not orig
    view = pkg.plugin.fused_view_of(lss)
    assert view.frustum is lss.frustum and (view.D, view.fH, view.fW) == (59, 8, 22) and view.nx.tolist() == [128, 128, 1]
    assert pkg.plugin.fused_view_of(lss) is view and sorted(lss.state_dict()) == keys and list(view.state_dict()) == []
    lss.nx = lss.nx.clone()                                  # re-assigned constants invalidate the cached view
    assert pkg.plugin.fused_view_of(lss) is not view
    rots, trans = pkg.synthetic.camera_ring(1, 6, (64, 176), seed=0)
    with pytest.raises(ValueError, match="CUDA"):
        lss.get_geometry(rots, trans)
    with pytest.raises(ValueError, match="CUDA"):
        lss.get_voxels(torch.zeros(1, 6, 8, 8, 22), rots, trans)
    # non-default transforms still reach the reference's own implementation (CPU torch ops)
    coor = lss.get_geometry(rots, trans, post_rots=torch.eye(3).expand(1, 6, 3, 3), post_trans=torch.zeros(1, 6, 3))
    assert coor.shape == (1, 6, 59, 8, 22, 3)
    pkg.plugin.patch_lss_class(ref.LiftSplatShoot, fused=False)
    assert ref.LiftSplatShoot.get_voxels is orig
    pkg.plugin.unpatch_lss_class(ref.LiftSplatShoot)
    assert ref.LiftSplatShoot.get_voxels is orig and not hasattr(ref.LiftSplatShoot, "_bevpool_b200_orig_prepare")
    op = ri.import_reference_op(ext=pkg.plugin.install_ext())
    assert op.bev_pool_v2_ext is pkg.bev_pool_v2_ext and op.__all__ == ['bev_pool_v2', 'TRTBEVPoolv2']
    with pytest.raises(ValueError, match="CUDA"):
        op.bev_pool_v2(torch.rand(1, 1, 2, 2, 2), torch.ones(1, 1, 2, 2, 2), torch.zeros(4).int(), torch.zeros(4).int(),
                       torch.zeros(4).int(), (1, 1, 2, 2, 2), torch.zeros(1).int(), torch.ones(1).int())
    for name in list(sys.modules):
        if name.startswith("projects.mmdet3d_plugin"):
            del sys.modules[name]
