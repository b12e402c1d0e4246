#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference Python code on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

What is imported from the reference (never copied):
  projects/mmdet3d_plugin/bevfusion/detectors/cam_stream_lss_bevpoolv2.py
    LiftSplatShoot.create_frustum      :216-227
    LiftSplatShoot.get_geometry        :229-258
    LiftSplatShoot.voxel_pooling_prepare_v2  :294-351
    gen_dx_bx                          :77-82
    QuickCumsum                        :96-122   (CPU cumsum segmented sum)
mmcv / mmdet3d / matplotlib are absent here, so they are stubbed in sys.modules
and the package chain is registered as bare namespace modules so the heavy
plugin __init__ does not run.

Outputs: tests/golden/*.npz (small; committed). Each file holds the inputs and
the reference outputs of one case. The reference's only known-answer test
(ops/bev_pool_v2/bev_pool.py:145-176) is transcribed as kat_bev_pool_v2.npz.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


class _Permissive(types.ModuleType):
    """Any attribute is a pass-through decorator factory / dummy class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def factory(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f
        return factory


def import_reference_lss():
    for name in ["mmcv", "mmcv.runner", "mmcv.cnn", "mmdet", "mmdet.models",
                 "mmdet.models.backbones", "mmdet.models.backbones.resnet",
                 "mmdet3d", "mmdet3d.models", "mmdet3d.models.fusion_layers",
                 "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
        sys.modules.setdefault(name, _Permissive(name))
    chain = {
        "projects": "projects",
        "projects.mmdet3d_plugin": "projects/mmdet3d_plugin",
        "projects.mmdet3d_plugin.ops": "projects/mmdet3d_plugin/ops",
        "projects.mmdet3d_plugin.ops.bev_pool_v2": "projects/mmdet3d_plugin/ops/bev_pool_v2",
        "projects.mmdet3d_plugin.bevfusion": "projects/mmdet3d_plugin/bevfusion",
        "projects.mmdet3d_plugin.bevfusion.detectors": "projects/mmdet3d_plugin/bevfusion/detectors",
    }
    for mod, rel in chain.items():
        m = types.ModuleType(mod)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[mod] = m
    # the compiled ext is only needed so `from . import bev_pool_v2_ext` resolves
    sys.modules["projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool_v2_ext"] = _Permissive("bev_pool_v2_ext")
    sys.modules["projects.mmdet3d_plugin.ops.bev_pool_v2"].bev_pool_v2_ext = \
        sys.modules["projects.mmdet3d_plugin.ops.bev_pool_v2.bev_pool_v2_ext"]
    return importlib.import_module(
        "projects.mmdet3d_plugin.bevfusion.detectors.cam_stream_lss_bevpoolv2")


def make_lss(ref, final_dim, downsample, dbound, xb, yb, zb):
    lss = ref.LiftSplatShoot(lss=False, final_dim=final_dim, camera_depth_range=list(dbound),
                             pc_range=[xb[0], yb[0], zb[0], xb[1], yb[1], zb[1]],
                             downsample=downsample, grid=xb[2], inputC=8, camC=8)
    # the class forces one scalar step for x, y, z: overwrite with per-axis constants
    lss.dx, lss.bx, lss.nx = ref.gen_dx_bx(list(xb), list(yb), list(zb))
    return lss


def ref_cumsum_pool(ref, lss, coor, depth, feat):
    """Upstream-LSS glue around the reference's own QuickCumsum (CPU)."""
    B, N, D, H, W, _ = coor.shape
    C = feat.shape[2]
    x = depth.unsqueeze(-1) * feat.permute(0, 1, 3, 4, 2).unsqueeze(2)    # B,N,D,H,W,C
    x = x.reshape(-1, C)
    g = ((coor - (lss.bx - lss.dx / 2.)) / lss.dx).long().view(-1, 3)
    bidx = torch.cat([torch.full((N * D * H * W, 1), b, dtype=torch.long) for b in range(B)])
    g = torch.cat((g, bidx), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < lss.nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < lss.nx[1]) \
        & (g[:, 2] >= 0) & (g[:, 2] < lss.nx[2])
    x, g = x[kept], g[kept]
    ranks = g[:, 3] * (lss.nx[2] * lss.nx[1] * lss.nx[0]) + g[:, 2] * (lss.nx[1] * lss.nx[0]) \
        + g[:, 1] * lss.nx[0] + g[:, 0]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x, g = ref.QuickCumsum.apply(x, g, ranks)
    final = torch.zeros((B, C, int(lss.nx[2]), int(lss.nx[1]), int(lss.nx[0])))
    final[g[:, 3], :, g[:, 2], g[:, 1], g[:, 0]] = x
    return final


def main():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    syn = pkg.synthetic
    ref = import_reference_lss()
    torch.manual_seed(0)

    cases = {
        # name: final_dim, downsample, dbound, xb, yb, zb, B, C
        "tiny_bev_z1": ((64, 176), 16, (1.0, 60.0, 1.0), (-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8), (-5.0, 3.0, 8.0), 2, 8),
        "tiny_occ_z16": ((64, 176), 16, (1.0, 45.0, 0.5), (-40.0, 40.0, 0.4), (-40.0, 40.0, 0.4), (-1.0, 5.4, 0.4), 2, 4),
        "tiny_omnihd": ((96, 160), 8, (1.0, 60.0, 1.0), (-60.0, 60.0, 0.5), (-40.0, 40.0, 0.5), (-3.0, 5.0, 0.5), 1, 8),
        "tiny_hires": ((64, 176), 8, (1.0, 60.0, 0.5), (-51.2, 51.2, 0.512), (-51.2, 51.2, 0.512), (-5.0, 3.0, 8.0), 1, 8),
    }
    for name, (fd, ds, db, xb, yb, zb, B, C) in cases.items():
        lss = make_lss(ref, fd, ds, db, xb, yb, zb)
        rots, trans = syn.camera_ring(B, 6, fd, seed=0)
        with torch.no_grad():
            coor = lss.get_geometry(rots, trans)
            rb, rd, rf, st, ln = lss.voxel_pooling_prepare_v2(coor)
        D, fH, fW = lss.frustum.shape[:3]
        g = torch.Generator().manual_seed(1)
        depth = torch.randn(B, 6, D, fH, fW, generator=g).softmax(2)
        feat = torch.randn(B, 6, C, fH, fW, generator=g)
        with torch.no_grad():
            pooled = ref_cumsum_pool(ref, lss, coor, depth, feat)
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            final_dim=np.array(fd), downsample=np.array(ds), dbound=np.array(db, dtype=np.float64),
            xbound=np.array(xb, dtype=np.float64), ybound=np.array(yb, dtype=np.float64),
            zbound=np.array(zb, dtype=np.float64),
            dx=lss.dx.numpy(), bx=lss.bx.numpy(), nx=lss.nx.numpy(),
            frustum=lss.frustum.detach().numpy(), rots=rots.numpy(), trans=trans.numpy(),
            coor=coor.numpy(), ranks_bev=rb.numpy(), ranks_depth=rd.numpy(), ranks_feat=rf.numpy(),
            interval_starts=st.numpy(), interval_lengths=ln.numpy(),
            depth=depth.numpy(), feat=feat.numpy(), cumsum_pooled=pooled.numpy().astype(np.float32))
        stable = bool(((rd[1:] > rd[:-1]) | (rb[1:] != rb[:-1])).all())
        print(name, "P0", coor.numel() // 3, "P", rb.numel(), "I", st.numel(),
              "maxlen", int(ln.max()), "nx", lss.nx.tolist(), "reference tie order stable:", stable)

    # Larger cases: torch's CPU argsort is only stable (ties in ascending point index, the order the
    # CUDA radix sort gives at real sizes) above ~5e4 elements, so exact tie order is pinned on cases
    # with P >= 5e4. Only camera poses + reference outputs + a hash of the reference coor are stored.
    import hashlib
    mids = {
        "mid_bev_z1": ((64, 176), 16, (1.0, 60.0, 1.0), (-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8), (-5.0, 3.0, 8.0), 8),
        "mid_occ_z16": ((64, 176), 16, (1.0, 45.0, 0.5), (-40.0, 40.0, 0.4), (-40.0, 40.0, 0.4), (-1.0, 5.4, 0.4), 6),
        "mid_omnihd": ((96, 160), 8, (1.0, 60.0, 1.0), (-60.0, 60.0, 0.5), (-40.0, 40.0, 0.5), (-3.0, 5.0, 0.5), 2),
    }
    for name, (fd, ds, db, xb, yb, zb, B) in mids.items():
        lss = make_lss(ref, fd, ds, db, xb, yb, zb)
        rots, trans = syn.camera_ring(B, 6, fd, seed=7)
        with torch.no_grad():
            coor = lss.get_geometry(rots, trans)
            rb, rd, rf, st, ln = lss.voxel_pooling_prepare_v2(coor)
        assert rb.numel() >= 50000
        assert bool(((rd[1:] > rd[:-1]) | (rb[1:] != rb[:-1])).all()), \
            "reference order is not the stable order at this size"
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            final_dim=np.array(fd), downsample=np.array(ds), dbound=np.array(db, dtype=np.float64),
            dx=lss.dx.numpy(), bx=lss.bx.numpy(), nx=lss.nx.numpy(), rots=rots.numpy(), trans=trans.numpy(),
            coor_sha256=np.frombuffer(hashlib.sha256(coor.numpy().tobytes()).digest(), dtype=np.uint8),
            ranks_bev=rb.numpy(), ranks_depth=rd.numpy(), ranks_feat=rf.numpy(),
            interval_starts=st.numpy(), interval_lengths=ln.numpy())
        print(name, "P0", coor.numel() // 3, "P", rb.numel(), "I", st.numel(), "maxlen", int(ln.max()))

    # nothing in range -> five Nones (cam_stream_lss_bevpoolv2.py:344-345)
    lss = make_lss(ref, (64, 176), 16, (1.0, 60.0, 1.0), (500.0, 602.4, 0.8), (500.0, 602.4, 0.8), (-5.0, 3.0, 8.0))
    rots, trans = syn.camera_ring(1, 6, (64, 176), seed=0)
    out = lss.voxel_pooling_prepare_v2(lss.get_geometry(rots, trans))
    assert all(o is None for o in out)
    print("empty case -> 5 x None (reference behaviour confirmed)")

    # v1 op (ops/bev_pool/bev_pool.py): the reference's own CPU QuickCumsum on seeded random points
    v1_mod = types.ModuleType("projects.mmdet3d_plugin.ops.bev_pool")
    v1_mod.__path__ = [os.path.join(REF, "projects/mmdet3d_plugin/ops/bev_pool")]
    sys.modules["projects.mmdet3d_plugin.ops.bev_pool"] = v1_mod
    sys.modules["projects.mmdet3d_plugin.ops.bev_pool.bev_pool_ext"] = _Permissive("bev_pool_ext")
    v1_mod.bev_pool_ext = sys.modules["projects.mmdet3d_plugin.ops.bev_pool.bev_pool_ext"]
    v1 = importlib.import_module("projects.mmdet3d_plugin.ops.bev_pool.bev_pool")
    g = torch.Generator().manual_seed(5)
    Bv, Dv, Hv, Wv, Cv, Nv = 2, 3, 12, 10, 8, 6000
    feats = torch.randn(Nv, Cv, generator=g)
    coords = torch.stack([torch.randint(0, Hv, (Nv,), generator=g), torch.randint(0, Wv, (Nv,), generator=g),
                          torch.randint(0, Dv, (Nv,), generator=g), torch.randint(0, Bv, (Nv,), generator=g)], 1)
    ranks = coords[:, 0] * (Wv * Dv * Bv) + coords[:, 1] * (Dv * Bv) + coords[:, 2] * Bv + coords[:, 3]
    idx = ranks.argsort(stable=True)
    xo, go = v1.QuickCumsum.apply(feats[idx], coords[idx], ranks[idx])
    dense = torch.zeros(Bv, Dv, Hv, Wv, Cv)
    dense[go[:, 3], go[:, 2], go[:, 0], go[:, 1]] = xo       # index order of bev_pool_cuda.cu:36-38
    np.savez_compressed(os.path.join(HERE, "v1_bev_pool.npz"), feats=feats.numpy(), coords=coords.numpy(),
                        dims=np.array([Bv, Dv, Hv, Wv]), pooled=dense.permute(0, 4, 1, 2, 3).contiguous().numpy())
    print("v1 golden: intervals", xo.shape[0])

    # lift head: the reference's CamEncode.get_depth_feat (:134-141) with the 1x1 depthnet replaced by identity,
    # plus its autograd gradients for seeded upstream gradients
    torch.manual_seed(11)
    Dl, Cl, BNl, Hl, Wl = 13, 8, 3, 5, 7
    enc = ref.CamEncode(Dl, Cl, 4)
    enc.depthnet = torch.nn.Identity()
    xl = (torch.randn(BNl, Dl + Cl, Hl, Wl) * 3).requires_grad_()
    dl, fl = enc.get_depth_feat(xl)
    gd, gf = torch.randn_like(dl), torch.randn_like(fl)
    (dl * gd).sum().backward(retain_graph=True)
    gx_d = xl.grad.clone()
    xl.grad = None
    (fl * gf).sum().backward()
    np.savez_compressed(os.path.join(HERE, "lift_head.npz"), x=xl.detach().numpy(), depth=dl.detach().numpy(),
                        feat=fl.detach().numpy(), depth_grad=gd.numpy(), feat_grad=gf.numpy(),
                        x_grad=(gx_d + xl.grad).numpy(), dims=np.array([Dl, Cl]))
    print("lift golden", tuple(dl.shape), tuple(fl.shape))

    # cross-modal fusion glue: the reference's Cross_Modal_Fusion.forward (BEVCross_modal_attention.py:31-43) with
    # mmcv's ConvModule (the final 3x3 reduction conv) replaced by identity, so the output is the gated concat;
    # gradients from autograd
    class _IdentityConvModule(torch.nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()
    sys.modules["mmcv.cnn"].ConvModule = _IdentityConvModule
    sys.modules["mmcv.cnn"].xavier_init = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location(
        "ref_cross_modal", os.path.join(REF, "projects/mmdet3d_plugin/rcfusion/detectors/BEVCross_modal_attention.py"))
    cm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cm)
    torch.manual_seed(13)
    fus = cm.Cross_Modal_Fusion(kernel_size=3)
    img = torch.randn(2, 6, 5, 9, requires_grad=True)
    rad = torch.randn(2, 10, 5, 9, requires_grad=True)
    outf = fus(img, rad)
    gof = torch.randn_like(outf)
    outf.backward(gof)
    np.savez_compressed(os.path.join(HERE, "cross_modal.npz"), img=img.detach().numpy(), radar=rad.detach().numpy(),
                        w_img=fus.att_img[0].weight.detach().numpy(), w_radar=fus.att_radar[0].weight.detach().numpy(),
                        out=outf.detach().numpy(), out_grad=gof.numpy(), img_grad=img.grad.numpy(), radar_grad=rad.grad.numpy())
    print("cross-modal golden", tuple(outf.shape))

    # the reference's own KAT, transcribed (ops/bev_pool_v2/bev_pool.py:145-176)
    np.savez(os.path.join(HERE, "kat_bev_pool_v2.npz"),
             depth=np.array([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], dtype=np.float32).reshape(1, 1, 2, 2, 2),
             feat=np.ones((1, 1, 2, 2, 2), dtype=np.float32),
             ranks_depth=np.array([0, 4, 1, 6], dtype=np.int32),
             ranks_feat=np.array([0, 0, 1, 2], dtype=np.int32),
             ranks_bev=np.array([0, 0, 1, 1], dtype=np.int32),
             bev_feat_shape=np.array([1, 1, 2, 2, 2]),
             loss=np.array(4.4, dtype=np.float32),
             grad_depth=np.array([2., 2., 0., 0., 2., 0., 2., 0.], dtype=np.float32).reshape(1, 1, 2, 2, 2),
             grad_feat=np.array([1., 1., .4, .4, .8, .8, 0., 0.], dtype=np.float32).reshape(1, 1, 2, 2, 2))

    # gen_dx_bx goldens for every BASELINE config (cam_stream_lss_bevpoolv2.py:77-82)
    rec = {}
    for k, c in syn.CONFIGS.items():
        dx, bx, nx = ref.gen_dx_bx(list(c.xbound), list(c.ybound), list(c.zbound))
        rec[k + "_dx"], rec[k + "_bx"], rec[k + "_nx"] = dx.numpy(), bx.numpy(), nx.numpy()
    np.savez(os.path.join(HERE, "gen_dx_bx.npz"), **rec)
    print("done")


if __name__ == "__main__":
    main()
