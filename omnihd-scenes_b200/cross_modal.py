"""Cross-modal fusion: mirror of the reference's `Cross_Modal_Fusion`
(rcfusion/detectors/BEVCross_modal_attention.py:6-43). The three convolutions stay library calls (cuDNN through
torch); the bandwidth-bound glue around them — channel mean/max (:32-38) and the cross gating + concat (:40-42) —
are two fused kernels each way (csrc/fusion.cu) instead of seven elementwise / reduction / cat passes.
"""
import torch
from torch import nn

from . import _lib
from .bev_pool import _dtype_code, _ptr, _stream


def _canon(*ts):
    if not all(t.is_cuda for t in ts):
        raise ValueError("cross_modal: CUDA tensors only (this library has no CPU path)")
    dt = ts[0].dtype if ts[0].dtype in (torch.float32, torch.bfloat16) and all(t.dtype == ts[0].dtype for t in ts) \
        else torch.float32
    return [t.to(dt).contiguous() for t in ts]


class _ChannelAvgMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        (x,) = _canon(x)
        if x.dim() != 4:
            raise ValueError("expected [B, C, H, W]")
        B, C, H, W = x.shape
        out = x.new_empty((B, 2, H, W))
        arg = torch.empty((B, H, W), dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().bevpool_channel_avg_max_forward(_ptr(x), _ptr(out), _ptr(arg), B, C, H * W, _dtype_code(x),
                                                               _stream()), "bevpool_channel_avg_max_forward")
        ctx.save_for_backward(arg)
        ctx.dims = (B, C, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        B, C, H, W = ctx.dims
        (g,) = _canon(g)
        dx = g.new_empty((B, C, H, W))
        _lib.check(_lib.load().bevpool_channel_avg_max_backward(_ptr(g), _ptr(arg), _ptr(dx), B, C, H * W, _dtype_code(g),
                                                                _stream()), "bevpool_channel_avg_max_backward")
        return dx


class _GateConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, att_for_a, att_for_b):
        a, b, att_for_a, att_for_b = _canon(a, b, att_for_a, att_for_b)
        N, Ca, H, W = a.shape
        Cb = b.shape[1]
        if b.shape != (N, Cb, H, W) or att_for_a.shape != (N, 1, H, W) or att_for_b.shape != (N, 1, H, W):
            raise ValueError("expected a [N,Ca,H,W], b [N,Cb,H,W], attention maps [N,1,H,W]")
        out = a.new_empty((N, Ca + Cb, H, W))
        _lib.check(_lib.load().bevpool_gate_concat_forward(_ptr(a), _ptr(b), _ptr(att_for_a), _ptr(att_for_b), _ptr(out), N, Ca,
                                                           Cb, H * W, _dtype_code(a), _stream()), "bevpool_gate_concat_forward")
        ctx.save_for_backward(a, b, att_for_a, att_for_b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, wa, wb = ctx.saved_tensors
        (g,) = _canon(g.to(a.dtype))
        N, Ca, H, W = a.shape
        Cb = b.shape[1]
        da, db, dwa, dwb = torch.empty_like(a), torch.empty_like(b), torch.empty_like(wa), torch.empty_like(wb)
        _lib.check(_lib.load().bevpool_gate_concat_backward(_ptr(g), _ptr(a), _ptr(b), _ptr(wa), _ptr(wb), _ptr(da), _ptr(db),
                                                            _ptr(dwa), _ptr(dwb), N, Ca, Cb, H * W, _dtype_code(a), _stream()),
                   "bevpool_gate_concat_backward")
        return da, db, dwa, dwb


def channel_avg_max(x):
    """[B,C,H,W] -> [B,2,H,W] = cat([mean over C, max over C], 1) (BEVCross_modal_attention.py:32-34)."""
    return _ChannelAvgMax.apply(x)


def gate_concat(a, b, att_for_a, att_for_b):
    """cat([a * att_for_a, b * att_for_b], 1) (:40-42)."""
    return _GateConcat.apply(a, b, att_for_a, att_for_b)


def _build_norm(norm_cfg, channels):
    """(attribute name, module) the way mmcv's `build_norm_layer` names and builds it (mmcv/cnn/bricks/norm.py):
    BN / BN2d / SyncBN -> `bn`, GN -> `gn`; `requires_grad` (default True) freezes the affine parameters."""
    cfg = dict(norm_cfg)
    kind = cfg.pop("type")
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if kind in ("BN", "BN2d"):
        name, layer = "bn", nn.BatchNorm2d(channels, **cfg)
    elif kind == "SyncBN":
        name, layer = "bn", nn.SyncBatchNorm(channels, **cfg)
    elif kind == "GN":
        name, layer = "gn", nn.GroupNorm(num_channels=channels, **cfg)
    else:
        raise ValueError(f"unsupported norm_cfg type {kind!r} (BN, BN2d, SyncBN, GN)")
    for prm in layer.parameters():
        prm.requires_grad = requires_grad
    return name, layer


class _ConvModule(nn.Module):
    """What mmcv's ConvModule(cin, cout, k, padding, conv_cfg=None, norm_cfg=..., act_cfg=ReLU, inplace=False) is,
    with the same state_dict keys: `.conv` (bias only when there is no norm: mmcv's bias='auto'), the norm layer
    under mmcv's abbreviation (`.bn` / `.gn`), ReLU; order conv -> norm -> act."""

    def __init__(self, cin, cout, k, padding, norm_cfg=None):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=norm_cfg is None)
        self.norm_name = None
        if norm_cfg is not None:
            self.norm_name, layer = _build_norm(norm_cfg, cout)
            self.add_module(self.norm_name, layer)
        self.activate = nn.ReLU(inplace=False)

    def forward(self, x):
        x = self.conv(x)
        if self.norm_name is not None:
            x = getattr(self, self.norm_name)(x)
        return self.activate(x)


class Cross_Modal_Fusion(nn.Module):
    """Same constructor, parameter / buffer names and forward as the reference class. The reference detector builds it
    with `norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01)` (rcfusion_faster_rcnn.py:38,74), i.e. `reduce_mixBEV` is
    conv(bias=False) + BN + ReLU with keys `reduce_mixBEV.conv.weight`, `reduce_mixBEV.bn.*`; `norm_cfg=None` (the
    class default) gives conv(bias=True) + ReLU. Channel counts are the reference's 256 + 384 -> 384 by default."""

    def __init__(self, kernel_size=3, norm_cfg=None, img_channels=256, radar_channels=384, out_channels=384):
        super().__init__()
        assert kernel_size in (3, 7), 'kernel size must be 3 or 7'
        padding = 3 if kernel_size == 7 else 1
        self.att_img = nn.Sequential(nn.Conv2d(2, 1, kernel_size, padding=padding, bias=False), nn.Sigmoid())
        self.att_radar = nn.Sequential(nn.Conv2d(2, 1, kernel_size, padding=padding, bias=False), nn.Sigmoid())
        self.reduce_mixBEV = _ConvModule(img_channels + radar_channels, out_channels, 3, 1, norm_cfg)

    def fuse(self, img_bev, radar_bev):
        """Everything up to the 3x3 reduction conv: [N, Ci + Cr, H, W]."""
        img_att = self.att_img(channel_avg_max(img_bev))
        radar_att = self.att_radar(channel_avg_max(radar_bev))
        return gate_concat(img_bev, radar_bev, radar_att, img_att)

    def forward(self, img_bev, radar_bev):
        return self.reduce_mixBEV(self.fuse(img_bev, radar_bev))
