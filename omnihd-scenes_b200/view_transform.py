"""Host-side mirror of the view-transform half of the reference's LSS necks
(projects/mmdet3d_plugin/bevfusion/detectors/cam_stream_lss_bevpoolv2.py and the two
`_depthnet` variants): grid constants, frustum, geometry, voxel_pooling_prepare_v2,
voxel_pooling_v2 and s2c — without the conv nets around them (out of scope).

    gen_dx_bx(xbound, ybound, zbound)                :77-82
    create_frustum(final_dim, downsample, dbound)    :216-227
    get_geometry(frustum, rots, trans)               :229-258 (callers only use the plain branch)
    voxel_pooling_prepare_v2(coor, dx, bx, nx)       :294-351  -> (ranks_bev, ranks_depth, ranks_feat,
                                                                  interval_starts, interval_lengths) | 5 x None
    LSSViewTransform                                 the module-shaped shim: same method names as
                                                     LiftSplatShoot, `frustum` kept as a Parameter so
                                                     released checkpoints load

Device work is done by csrc/prepare.cu and csrc/pool.cu through the C ABI.
"""
import ctypes
import os

import torch
from torch import nn

from . import _lib
from .bev_pool import bev_pool_v2, register_plan, _ptr, _stream, _launch_forward_dense, _launch_transpose, \
    _launch_voxel_table, _dtype_code, _column_hint


# ----------------------------------------------------------------------------- constants
def gen_dx_bx(xbound, ybound, zbound):
    """dx (step), bx (first voxel centre), nx (voxel counts, float divide then truncation)."""
    rows = (xbound, ybound, zbound)
    dx = torch.tensor([float(r[2]) for r in rows], dtype=torch.float32)
    bx = torch.tensor([r[0] + r[2] / 2.0 for r in rows], dtype=torch.float32)
    nx = torch.tensor([int((r[1] - r[0]) / r[2]) for r in rows], dtype=torch.int64)
    return dx, bx, nx


def create_frustum(final_dim, downsample, dbound):
    """[D, fH, fW, 3] fp32 with (x_pixel, y_pixel, depth) per cell — same torch ops as the
    reference so that the values are bit-identical to the checkpointed Parameter."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.arange(*dbound, dtype=torch.float)
    D = ds.numel()
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float)
    frustum = torch.empty(D, fH, fW, 3, dtype=torch.float)
    frustum[..., 0] = xs.view(1, 1, fW)
    frustum[..., 1] = ys.view(1, fH, 1)
    frustum[..., 2] = ds.view(D, 1, 1)
    return frustum


# ----------------------------------------------------------------------------- checks
def _check_f32_cuda(name, t, shape_tail=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (bevpool_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError(f"{name} must end with shape {tuple(shape_tail)}, got {tuple(t.shape)}")


_GRID_CACHE = {}


def _grid_struct(B, N, D, H, W, dx, bx, nx):
    """bevpool_grid_t for this call. The three constants are small CPU (sometimes CUDA) tensors that never change in
    practice: the converted values are cached per tensor identity + version (a CUDA tensor would cost a sync each call)."""
    key = None
    if all(isinstance(t, torch.Tensor) for t in (dx, bx, nx)):
        key = (id(dx), dx._version, id(bx), bx._version, id(nx), nx._version)
        hit = _GRID_CACHE.get(key)
        if hit is not None and hit[0]() is dx and hit[1]() is bx and hit[2]() is nx:
            g = _lib.GridT()
            g.b, g.n, g.d, g.h, g.w = B, N, D, H, W
            g.nx[:], g.lo[:], g.dx[:] = hit[3], hit[4], hit[5]
            return g
    g = _grid_struct_uncached(B, N, D, H, W, dx, bx, nx)
    if key is not None:
        import weakref
        if len(_GRID_CACHE) > 64:
            _GRID_CACHE.clear()
        _GRID_CACHE[key] = (weakref.ref(dx), weakref.ref(bx), weakref.ref(nx), list(g.nx), list(g.lo), list(g.dx))
    return g


def _grid_struct_uncached(B, N, D, H, W, dx, bx, nx):
    dx = torch.as_tensor(dx).detach().to("cpu", torch.float32)
    bx = torch.as_tensor(bx).detach().to("cpu", torch.float32)
    nx = torch.as_tensor(nx).detach().to("cpu", torch.int64)
    lo = bx - dx / 2.            # fp32 tensor arithmetic, as cam_stream_lss_bevpoolv2.py:317
    g = _lib.GridT()
    g.b, g.n, g.d, g.h, g.w = B, N, D, H, W
    for a in range(3):
        g.nx[a] = int(nx[a])
        g.lo[a] = float(lo[a])
        g.dx[a] = float(dx[a])
    return g


# ----------------------------------------------------------------------------- geometry
def get_geometry(frustum, rots, trans):
    """coor [B,N,D,fH,fW,3] fp32: (x,y,z) of every frustum point in the lidar/ego frame.
    rots/trans are img->lidar (inverse of lidar2img, bevf_faster_rcnn.py:119-128)."""
    _check_f32_cuda("frustum", frustum, (3,))
    _check_f32_cuda("rots", rots, (3, 3))
    _check_f32_cuda("trans", trans, (3,))
    if frustum.dim() != 4 or rots.dim() != 4 or trans.dim() != 3:
        raise ValueError("expected frustum [D,H,W,3], rots [B,N,3,3], trans [B,N,3]")
    B, N = trans.shape[:2]
    D, H, W, _ = frustum.shape
    frustum, rots, trans = frustum.contiguous(), rots.contiguous(), trans.contiguous()
    coor = torch.empty((B, N, D, H, W, 3), dtype=torch.float32, device=rots.device)
    lib = _lib.load()
    _lib.check(lib.bevpool_geometry(_ptr(frustum), _ptr(rots), _ptr(trans), _ptr(coor), B * N, D, H * W, _stream()),
               "bevpool_geometry")
    return coor


# ----------------------------------------------------------------------------- prepare
class _Prepared:
    """Worst-case-sized device outputs of one prepare call (no host sync yet)."""
    __slots__ = ("rb", "rd", "rf", "starts", "lengths", "counts", "point_rank", "bn", "d", "h", "w", "hw", "p0",
                 "host_counts")


def _pad64(n):
    return (max(n, 1) + 63) // 64 * 64            # rows stay 256-byte aligned (128-bit loads)


def _prepare_device(coor, frustum, rots, trans, B, N, D, H, W, dx, bx, nx, device, want_intervals=True,
                    host_counts=False):
    """want_intervals=False (fused path): only the sorted (ranks_bev, ranks_depth) lists, the kept count and
    point_rank are produced — ranks_feat is derivable and the interval arrays are replaced by the voxel table.
    host_counts=True (API path): `bevpool_prepare_v2_counts` — the call hands (P, I) back as soon as the rank
    kernel has run, with the sort still in flight; they are left in `out.host_counts`."""
    lib = _lib.load()
    g = _grid_struct(B, N, D, H, W, dx, bx, nx)
    p0 = B * N * D * H * W
    vtot = B * g.nx[0] * g.nx[1] * g.nx[2]
    if p0 >= 2 ** 30 or vtot >= 2 ** 31 - 1:
        raise ValueError("problem too large for int32 ranks: shard the frame batch")
    out = _Prepared()
    # one int32 block: [ranks_bev | ranks_depth | (ranks_feat) | (starts | lengths) | counts]; point_rank is its own
    # tensor (the fused path keeps only that one alive until the backward)
    pp, n1 = _pad64(p0), max(p0, 1)
    rows = 3 if want_intervals else 2
    n_int = max(min(p0, vtot), 1) if want_intervals else 0
    pi = _pad64(n_int) if want_intervals else 0
    o_int = rows * pp
    o_cnt = o_int + 2 * pi
    blk = torch.empty(o_cnt + 64, dtype=torch.int32, device=device)
    base = blk.data_ptr()
    p_rf = p_st = p_ln = None
    if want_intervals:
        p_rf, p_st, p_ln = base + 8 * pp, base + 4 * o_int, base + 4 * (o_int + pi)
    out.point_rank = torch.empty(n1, dtype=torch.int32, device=device)
    out.bn, out.d, out.h, out.w, out.hw, out.p0 = B * N, D, H, W, H * W, p0
    ws_bytes = max(lib.bevpool_prepare_v2_workspace_bytes(g), 256)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    args = (_ptr(coor), _ptr(frustum), _ptr(rots), _ptr(trans), ctypes.byref(g), base, base + 4 * pp, p_rf, p_st, p_ln,
            base + 4 * o_cnt, out.point_rank.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    n_r, n_i = n1, n_int                     # worst-case lengths unless the host learns the counts
    if host_counts:
        hc = (ctypes.c_int32 * 2)()
        _lib.check(lib.bevpool_prepare_v2_counts(*args, ctypes.byref(hc)), "bevpool_prepare_v2_counts")
        out.host_counts = n_r, n_i = hc[0], hc[1]
    else:
        _lib.check(lib.bevpool_prepare_v2(*args), "bevpool_prepare_v2")
        out.host_counts = None
    out.rb, out.rd = blk[:n_r], blk[pp:pp + n_r]
    out.rf = out.starts = out.lengths = None
    if want_intervals:
        out.rf = blk[2 * pp:2 * pp + n_r]
        out.starts, out.lengths = blk[o_int:o_int + n_i], blk[o_int + pi:o_int + pi + n_i]
    out.counts = blk[o_cnt:o_cnt + 2]
    return out


def _view_forward_scatter(depth, feat_cl, out, view, rots, trans, n, N, D, H, W, C, frames, rows, layout):
    """Sort-free forward (csrc/pool_scatter.cu) of `n` frames; returns the _Prepared stub the backward needs."""
    lib = _lib.load()
    g = _grid_struct(n, N, D, H, W, view.dx, view.bx, view.nx)
    p0 = n * N * D * H * W
    vtot = n * g.nx[0] * g.nx[1] * g.nx[2]
    if p0 >= 2 ** 31 - 1 or vtot >= 2 ** 31 - 1:
        raise ValueError("problem too large for int32 ranks: shard the frame batch")
    pr = _Prepared()
    pr.rb = pr.rd = pr.rf = pr.starts = pr.lengths = pr.counts = None
    pr.point_rank = torch.empty(max(p0, 1), dtype=torch.int32, device=out.device)
    pr.bn, pr.d, pr.h, pr.w, pr.hw, pr.p0 = n * N, D, H, W, H * W, p0
    code = _dtype_code(feat_cl)
    nbytes = lib.bevpool_view_forward_scratch_bytes(vtot, C, layout, code)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=out.device) if nbytes else None
    _lib.check(lib.bevpool_view_forward(_ptr(depth), _ptr(feat_cl), _ptr(view.frustum), _ptr(rots), _ptr(trans),
                                        ctypes.byref(g), C, _ptr(pr.point_rank), 1, _ptr(out), frames, rows, layout, code,
                                        _ptr(scratch), nbytes, _stream()), "bevpool_view_forward")
    return pr


def voxel_pooling_prepare_v2(coor, dx, bx, nx):
    """Free-function form of `LiftSplatShoot.voxel_pooling_prepare_v2(self, coor)`.

    coor [B,N,D,H,W,3] fp32 CUDA. Returns int32 contiguous
    (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths) — note the order —
    sorted by ranks_bev with ties in ascending point index, or five Nones if no point falls
    inside the grid. Exactly one device->host read (the two counts) is performed, because the
    API returns exact-length tensors; the reference performs at least three.
    """
    _check_f32_cuda("coor", coor, (3,))
    if coor.dim() != 6:
        raise ValueError("coor must be [B, N, D, H, W, 3]")
    B, N, D, H, W, _ = coor.shape
    if B * N * D * H * W == 0:
        return None, None, None, None, None
    coor = coor.contiguous()
    # BEVPOOL_LATE_COUNTS=1 (measurement only): read the counts back after the whole prepare, as round 1 did
    early = os.environ.get("BEVPOOL_LATE_COUNTS") != "1"
    pr = _prepare_device(coor, None, None, None, B, N, D, H, W, dx, bx, nx, coor.device, host_counts=early)
    P, I = pr.host_counts if early else (int(v) for v in pr.counts.tolist())
    if P == 0 or I == 0:
        return None, None, None, None, None
    # with early counts the views already have their exact lengths
    res = (pr.rb, pr.rd, pr.rf, pr.starts, pr.lengths) if early else \
        (pr.rb[:P], pr.rd[:P], pr.rf[:P], pr.starts[:I], pr.lengths[:I])
    register_plan(*res, pr.point_rank, pr.bn, pr.d, pr.h, pr.w)
    return res


# ----------------------------------------------------------------------------- fused module path
_SIDE_STREAMS = {}


def _side_streams(device, n):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


class _Fork:
    """Run frame groups on side streams that fork from / join into the current stream. Frames are
    independent, so the groups' kernels overlap and fill each other's tails and launch gaps; with one
    group nothing is forked. Works under CUDA-graph capture (parallel graph branches)."""

    def __init__(self, device, groups):
        self.cur = torch.cuda.current_stream(device)
        self.streams = _side_streams(device, groups) if groups > 1 else [self.cur]
        if groups > 1:
            for st in self.streams:
                st.wait_stream(self.cur)

    def join(self):
        if len(self.streams) > 1 or self.streams[0] is not self.cur:
            for st in self.streams:
                self.cur.wait_stream(st)


class _FusedViewPool(torch.autograd.Function):
    """geometry -> rank -> sort -> pool with no host synchronisation: point counts stay on the device
    (counts_dev), buffers are worst-case sized. Takes feat as the neck produces it ([B,N,C,H,W]); the
    NCHW->NHWC transpose (reference :282), the zero fill, the pooling and the output permute are our
    kernels, and the backward is the sort-free kernel writing feat_grad straight back in [B,N,C,H,W].
    The frame batch is cut into `groups` independent groups that run on concurrent streams."""

    @staticmethod
    def forward(ctx, depth, feat, rots, trans, view, groups, feat_channels_last=False, s2c=False):
        B, N = trans.shape[:2]
        D, H, W = view.D, view.fH, view.fW
        C = feat.shape[4] if feat_channels_last else feat.shape[2]
        X, Y, Z = (int(v) for v in view.nx)
        depth = depth.contiguous()
        feat = feat.contiguous()
        if depth.dtype != feat.dtype or depth.dtype not in (torch.float32, torch.bfloat16):
            depth, feat = depth.float(), feat.float()
        rots, trans = rots.contiguous(), trans.contiguous()
        # s2c (reference :363-365) folds Z into the channel axis: [B, Z*C, Y, X]. Same kernels, the layout pass
        # just treats every (frame, z) plane as a frame of Y rows.
        # s2c == "channels_last": the grid stays in the kernels' own [B,Z,Y,X,C] order and is returned as a
        # [B,C,Z,Y,X] VIEW of it (torch.channels_last_3d strides): no layout pass at all.
        cl_out = s2c == "channels_last"
        s2c = bool(s2c) and not cl_out
        out = feat.new_empty((B, Z, Y, X, C) if cl_out else (B, Z * C, Y, X) if s2c else (B, C, Z, Y, X))
        planes, rows = (Z, Y) if s2c else (1, Z * Y)
        n = B // groups
        saved = []
        # sorted (deterministic summation order, radix sort + streaming pool) or sort-free scatter (default)
        sorted_path = view.deterministic or torch.are_deterministic_algorithms_enabled() or C > 128
        fork = _Fork(feat.device, groups)
        for g, st in enumerate(fork.streams):
            sl = slice(g * n, (g + 1) * n)
            with torch.cuda.stream(st):
                if feat_channels_last:                                                # lift head already wrote NHWC
                    feat_cl = feat[sl]
                else:
                    feat_cl = feat.new_empty((n * N, H, W, C))
                    _launch_transpose(feat[sl], feat_cl, n * N, C, H * W, True)       # [BN,C,HW] -> [BN,HW,C]
                layout = _lib.LAYOUT_BZYXC if cl_out else _lib.LAYOUT_BCZYX
                if sorted_path:
                    pr = _prepare_device(None, view.frustum, rots[sl], trans[sl], n, N, D, H, W, view.dx, view.bx,
                                         view.nx, feat.device, want_intervals=False)
                    vox_pt = _launch_voxel_table(pr.rb, pr.p0, pr.counts, n * Z * Y * X)
                    _launch_forward_dense(depth[sl], feat_cl, out[sl], pr.rd, None, pr.rb, vox_pt, n * planes, rows, X,
                                          layout, dhw=D * pr.hw, hw=pr.hw, n_points=pr.p0, counts_dev=pr.counts)
                else:
                    pr = _view_forward_scatter(depth[sl], feat_cl, out[sl], view, rots[sl], trans[sl], n, N, D, H, W, C,
                                               n * planes, rows, layout)
            saved.append((pr, feat_cl))
        fork.join()
        ctx.dims, ctx.groups, ctx.feat_cl, ctx.s2c = (B, N, C, D, H, W, X, Y, Z), groups, feat_channels_last, s2c
        # everything the backward reads goes through save_for_backward: with feat_channels_last the per-group
        # feature tensors are VIEWS of the caller's input, and autograd's version check must catch an in-place edit
        ctx.save_for_backward(depth, *[fc for _, fc in saved], *[pr.point_rank for pr, _ in saved])
        ctx.bn = [pr.bn for pr, _ in saved]
        return out.permute(0, 4, 1, 2, 3) if cl_out else out

    @staticmethod
    def backward(ctx, out_grad):
        depth = ctx.saved_tensors[0]
        B, N, C, D, H, W, X, Y, Z = ctx.dims
        groups = ctx.groups
        feat_cls, point_ranks = ctx.saved_tensors[1:1 + groups], ctx.saved_tensors[1 + groups:1 + 2 * groups]
        n = B // groups
        dt = feat_cls[0].dtype
        # a channels_last_3d gradient ([B,Z,Y,X,C] in memory) is consumed as it is; anything else is transposed
        og_is_cl = (not ctx.s2c) and out_grad.dim() == 5 and out_grad.permute(0, 2, 3, 4, 1).is_contiguous()
        out_grad = (out_grad.permute(0, 2, 3, 4, 1) if og_is_cl else out_grad.contiguous()).to(dt)
        depth_grad = torch.empty_like(depth)
        feat_grad = depth.new_empty((B, N, H, W, C) if ctx.feat_cl else (B, N, C, H, W))
        lib = _lib.load()
        fork = _Fork(depth.device, groups)
        for g, st in enumerate(fork.streams):
            sl = slice(g * n, (g + 1) * n)
            feat_cl, point_rank, bn = feat_cls[g], point_ranks[g], ctx.bn[g]
            with torch.cuda.stream(st):
                og_cl = out_grad[sl] if og_is_cl else out_grad.new_empty((n, Z, Y, X, C))
                if og_is_cl:
                    pass
                elif ctx.s2c:
                    _launch_transpose(out_grad[sl], og_cl, n * Z, C, Y * X, True)
                else:
                    _launch_transpose(out_grad[sl], og_cl, n, C, Z * Y * X, True)
                _lib.check(lib.bevpool_v2_backward_dense(_ptr(og_cl), _ptr(depth_grad[sl]), _ptr(feat_grad[sl]),
                                                         _ptr(depth[sl]), _ptr(feat_cl), _ptr(point_rank), bn, D, H,
                                                         W, C, 0 if ctx.feat_cl else 1, _column_hint(Z),
                                                         _dtype_code(feat_cl), _stream()),
                           "bevpool_v2_backward_dense")
        fork.join()
        return depth_grad, feat_grad, None, None, None, None, None, None


class _GraphedForward:
    """Forward-only CUDA graph of a view transform over static input copies (LSSViewTransform.graphed without grads)."""

    def __init__(self, view, sample):
        self.static_in = [t.detach() for t in sample]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                view(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = view(*self.static_in)

    def __call__(self, depth, feat, rots, trans):
        for s, t in zip(self.static_in, (depth, feat, rots, trans)):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t)
        self.graph.replay()
        return self.out


class LSSViewTransform(nn.Module):
    """View-transform part of the reference's `LiftSplatShoot` (cam_stream_lss_bevpoolv2.py:149-375):
    same attribute names (`dx`, `bx`, `nx`, `frustum`, `D`, `fH`, `fW`) and method names, no conv nets.
    Per-axis bounds are accepted (the reference forces one scalar `grid` for x, y and z, :164-169)."""

    def __init__(self, final_dim, downsample, dbound, xbound, ybound, zbound, frame_groups=1, deterministic=False):
        super().__init__()
        self.frame_groups = frame_groups          # fused path: independent frame groups on concurrent streams
        # fused path: False = sort-free pixel-major forward (fp32 sums meet in L2 atomics: order across image
        # columns not fixed); True (or torch.use_deterministic_algorithms(True)) = sorted, fixed-order forward
        self.deterministic = deterministic
        self.final_dim = tuple(final_dim)
        self.downsample = downsample
        self.grid_conf = dict(xbound=list(xbound), ybound=list(ybound), zbound=list(zbound), dbound=list(dbound))
        self.dx, self.bx, self.nx = gen_dx_bx(xbound, ybound, zbound)       # plain CPU tensors, as the reference
        self.fH, self.fW = final_dim[0] // downsample, final_dim[1] // downsample
        self.frustum = nn.Parameter(create_frustum(final_dim, downsample, dbound), requires_grad=False)
        self.D = self.frustum.shape[0]

    @classmethod
    def from_config(cls, cfg, frame_groups=1, deterministic=False):
        return cls(cfg.final_dim, cfg.downsample, cfg.dbound, cfg.xbound, cfg.ybound, cfg.zbound, frame_groups,
                   deterministic)

    @classmethod
    def from_lss_args(cls, final_dim, camera_depth_range, pc_range, downsample, grid):
        """Constructor taking the reference's own keyword set (:150)."""
        return cls(final_dim, downsample, camera_depth_range, (pc_range[0], pc_range[3], grid),
                   (pc_range[1], pc_range[4], grid), (pc_range[2], pc_range[5], grid))

    @classmethod
    def adopt(cls, frustum, dx, bx, nx, frame_groups=1, deterministic=False):
        """View transform over an EXISTING module's state (plugin.patch_lss_class): shares the reference module's
        `frustum` Parameter ([D, fH, fW, 3], cam_stream_lss_bevpoolv2.py:216-227) and its `dx / bx / nx` tensors
        (:77-82) instead of rebuilding them, so values set by hand on the reference module are honoured."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.frame_groups, self.deterministic = frame_groups, deterministic
        self.final_dim = self.downsample = self.grid_conf = None
        self.dx = torch.as_tensor(dx).detach().float().cpu()
        self.bx = torch.as_tensor(bx).detach().float().cpu()
        self.nx = torch.as_tensor(nx).detach().long().cpu()
        self.__dict__["frustum"] = frustum          # shared, not re-registered: the owner's state_dict is unchanged
        self.D, self.fH, self.fW = (int(v) for v in frustum.shape[:3])
        return self

    # -- reference method names -------------------------------------------------------------
    def get_geometry(self, rots, trans, post_rots=None, post_trans=None, extra_rots=None, extra_trans=None):
        if any(v is not None for v in (post_rots, post_trans, extra_rots, extra_trans)):
            # never passed by any caller in the reference; refuse rather than be silently wrong
            raise NotImplementedError("post_/extra_ transforms are not used by the reference's callers")
        return get_geometry(self.frustum, rots, trans)

    def voxel_pooling_prepare_v2(self, coor):
        return voxel_pooling_prepare_v2(coor, self.dx, self.bx, self.nx)

    def _nx_ints(self):
        """(X, Y, Z) as Python ints, cached against the identity / version of `self.nx`."""
        nx = self.nx
        hit = self.__dict__.get("_nx_cache")
        if hit is None or hit[0] is not nx or hit[1] != nx._version:
            hit = self.__dict__["_nx_cache"] = (nx, nx._version, tuple(int(v) for v in nx))
        return hit[2]

    def voxel_pooling_v2(self, coor, depth, feat):
        """coor [B,N,D,H,W,3], depth [B,N,D,H,W], feat [B,N,C,H,W] -> [B,C,Z,Y,X] (or None)."""
        ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths = self.voxel_pooling_prepare_v2(coor)
        if ranks_feat is None:
            print('warning ---> no points within the predefined bev receptive field')
            return None
        feat = feat.permute(0, 1, 3, 4, 2)           # a view, as the reference passes it (:282)
        X, Y, Z = self._nx_ints()
        bev_feat_shape = (depth.shape[0], Z, Y, X, feat.shape[-1])
        return bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                           interval_lengths)

    @staticmethod
    def s2c(x):
        """[B,C,Z,Y,X] -> [B,Z*C,Y,X] (:363-365)."""
        B, C, Z, Y, X = x.shape
        return x.permute(0, 2, 1, 3, 4).reshape(B, Z * C, Y, X)

    def get_voxels(self, depth, feat, rots, trans):
        """Reference order of operations (:357-361) given already-computed depth / feat."""
        return self.voxel_pooling_v2(self.get_geometry(rots, trans), depth, feat)

    # -- fused path ---------------------------------------------------------------------------
    def forward(self, depth, feat, rots, trans, feat_channels_last=False, s2c=False, memory_format=None):
        """Fused view transform: geometry is never materialised, nothing synchronises with the host.
        depth [B,N,D,fH,fW], feat [B,N,C,fH,fW] ([B,N,fH,fW,C] with feat_channels_last, as `lift` returns it)
        -> [B,C,Z,Y,X] (all zeros if no point is in range), or [B,Z*C,Y,X] = s2c(...) directly with s2c=True.
        `memory_format=torch.channels_last_3d` returns the same [B,C,Z,Y,X] values as a channels-last view (what a
        channels-last conv stack consumes) and accepts a channels-last gradient: both layout passes disappear."""
        _check_f32_cuda("rots", rots, (3, 3))
        _check_f32_cuda("trans", trans, (3,))
        B, N = trans.shape[:2]
        D, H, W = self.D, self.fH, self.fW
        if tuple(depth.shape) != (B, N, D, H, W):
            raise ValueError(f"depth must be {(B, N, D, H, W)}, got {tuple(depth.shape)}")
        C = feat.shape[4] if feat_channels_last else feat.shape[2]
        want = (B, N, H, W, C) if feat_channels_last else (B, N, C, H, W)
        if tuple(feat.shape) != want:
            raise ValueError(f"feat must be {want}, got {tuple(feat.shape)}")
        if C % 4:
            raise ValueError("the fused path needs C % 4 == 0; use voxel_pooling_v2 for other channel counts")
        groups = self.frame_groups if (self.frame_groups > 1 and B % self.frame_groups == 0) else 1
        if memory_format not in (None, torch.contiguous_format, torch.channels_last_3d):
            raise ValueError("memory_format must be torch.contiguous_format or torch.channels_last_3d")
        if memory_format == torch.channels_last_3d:
            if s2c:
                raise ValueError("s2c and channels_last_3d are exclusive (s2c of a channels-last grid is not a view)")
            s2c = "channels_last"
        return _FusedViewPool.apply(depth, feat, rots, trans, self, groups, feat_channels_last, s2c)

    def graphed(self, depth, feat, rots, trans):
        """CUDA-graphed form of `forward` for fixed shapes (training or inference loops): the ~12 launches of a step
        and their Python become two graph replays (forward, backward). Built with `torch.cuda.make_graphed_callables`
        on sample tensors of the real shapes / dtypes / requires_grad flags; the returned callable takes
        `(depth, feat, rots, trans)` and is differentiable. The usual CUDA-graph contract applies: outputs (and input
        gradients) live in static buffers that the NEXT call overwrites — consume or clone them before calling again.
        Nothing in the fused path synchronises with the host, which is what makes it capturable."""
        sample = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in (depth, feat, rots, trans))
        if torch.is_grad_enabled() and any(t.requires_grad for t in sample):
            # a plain function, not `self`: for a Module make_graphed_callables swaps `forward` in place
            return torch.cuda.make_graphed_callables(lambda d, f, r, t: self.forward(d, f, r, t), sample)
        return _GraphedForward(self, sample)          # inference: nothing to differentiate, one forward graph

    def lift_splat(self, x, rots, trans, C, s2c=False):
        """Depthnet output x [B*N, D+C, fH, fW] -> BEV grid [B,C,Z,Y,X]: fused softmax/split/transpose head
        (`lift.get_depth_feat`) feeding the fused view transform. Returns (bev, depth [B*N, D, fH, fW])."""
        from .lift import get_depth_feat
        B, N = trans.shape[:2]
        depth, feat_cl = get_depth_feat(x, self.D, C, channels_last=True)
        bev = self.forward(depth.view(B, N, self.D, self.fH, self.fW), feat_cl.view(B, N, self.fH, self.fW, C), rots, trans,
                           feat_channels_last=True, s2c=s2c)
        return bev, depth
