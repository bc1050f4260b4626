"""B200-native Video Swin Transformer (drop-in for the reference's visbackbone/video_swin.py).

Same constructor, parameter names / shapes and state_dict keys as `SwinTransformer3D`
(video_swin.py:408-480) and the same `get_vidswin_model(args)` entry point (:571-645), but forward and
backward run as sequences of lavender_b200 C-ABI kernels:

  patch embed   : im2col of (frame d, frame d+1) 4x4 patches -> tcgen05 GEMM (K=96) -> LayerNorm   (:388-405)
  block, part 1 : LayerNorm fused with roll(-shift)+window_partition (row map) -> QKV GEMM -> fused window
                  attention (dense rel-pos bias + shift-mask classes) -> proj GEMM whose epilogue does
                  window_reverse + roll(+shift) + DropPath + residual                               (:204-243,254)
  block, part 2 : LayerNorm -> fc1 GEMM + erf-GELU epilogue -> fc2 GEMM + DropPath + residual        (:245-246,259)
  patch merging : 2x2 gather + LayerNorm(4C) in one kernel -> reduction GEMM                         (:271-287)
The residual stream is fp32 token-major [B*D*H*W, C]; GEMM operands are fp16.
Not restated: F.pad to window multiples (:211-216, :274-276) — all BASELINE shapes (224^2 / 384^2) divide.
"""
import math
import os
from functools import lru_cache

import torch
import torch.nn as nn

from . import ops
from . import precision
from . import streams
from . import _lib as L
from .arena import arena_of
from .functional import (F16, F32, cast16, empty16, empty32, linear_dgrad, linear_fwd, linear_fwd_hp, linear_wgrad,
                         require_cuda)

# (embed_dim, depths, num_heads, window, patch) — the only keys get_vidswin_model reads from the mmcv config
# files (video_swin.py:616-634): swin_tiny.py:4-17, swin_base.py:3-5, swin_large.py:3-5 and the
# swin_*_patch244_window*_*.py files (patch_size=(2,4,4); window (8,7,7) or (8,12,12)).
SWIN_VARIANTS = {
    ("tiny", 224): dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(8, 7, 7)),
    ("base", 224): dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7)),
    ("large", 224): dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=(8, 7, 7)),
    ("large", 384): dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=(8, 12, 12)),
}


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """video_swin.py:18-43 (truncated normal by inverse CDF)."""
    def norm_cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.
    with torch.no_grad():
        lo, up = norm_cdf((a - mean) / std), norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * up - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


def get_window_size(x_size, window_size, shift_size=None):
    """video_swin.py:93-106."""
    ws = list(window_size)
    ss = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            ws[i] = x_size[i]
            if ss is not None:
                ss[i] = 0
    return tuple(ws) if ss is None else (tuple(ws), tuple(ss))


def relative_position_index(window_size):
    """video_swin.py:121-135."""
    wd, wh, ww = window_size
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


# ---------------------------------------------------------------------------------------------------------
# index maps (built once per shape with the reference's own roll / partition definition, cached on device)
# ---------------------------------------------------------------------------------------------------------
@lru_cache(maxsize=64)
def window_row_map(B, D, H, W, ws, ss, device):
    """map[r] = token-major source row of window-major row r: torch.roll(-shift) followed by window_partition
    (video_swin.py:218-227, 82-86) applied to an index tensor.  The same map scatters back (window_reverse +
    roll(+shift), :231-239)."""
    ids = torch.arange(B * D * H * W, dtype=torch.int32).view(B, D, H, W)
    if any(s > 0 for s in ss):
        ids = torch.roll(ids, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    x = ids.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2])
    return x.permute(0, 1, 3, 5, 2, 4, 6).contiguous().view(-1).to(device)


@lru_cache(maxsize=64)
def window_row_map_inv(B, D, H, W, ws, ss, device):
    """inv[token-major row] = window-major row (the map above is a permutation: the feature map divides the window)."""
    m = window_row_map(B, D, H, W, ws, ss, device).long()
    inv = torch.empty_like(m, dtype=torch.int32)
    inv[m] = torch.arange(m.numel(), dtype=torch.int32, device=m.device)
    return inv


FUSE_CASTS = os.environ.get("LAV_FUSE_CASTS", "1") != "0"   # gradient casts emitted by the LayerNorm backward kernels


@lru_cache(maxsize=64)
def merge_row_map(B, D, H, W, device):
    """map[r*4+g] for PatchMerging (video_swin.py:278-282): x0=(even h, even w), x1=(odd h, even w),
    x2=(even h, odd w), x3=(odd h, odd w)."""
    ids = torch.arange(B * D * H * W, dtype=torch.int32).view(B, D, H, W)
    parts = [ids[:, :, 0::2, 0::2], ids[:, :, 1::2, 0::2], ids[:, :, 0::2, 1::2], ids[:, :, 1::2, 1::2]]
    return torch.stack(parts, -1).contiguous().view(-1).to(device)


@lru_cache(maxsize=64)
def shift_mask_classes(D, H, W, ws, ss, device):
    """Region labels of compute_mask (video_swin.py:290-305) factorised into window classes.
    Returns (labels uint8 [ncls, NP], class_of_window int32 [nW]) or (None, None) when nothing is shifted.
    A window's mask pattern depends only on whether it is the last window along each shifted axis."""
    axes = [ax for ax in range(3) if ss[ax] > 0]
    if not axes:
        return None, None
    n = ws[0] * ws[1] * ws[2]
    NP = (n + 127) // 128 * 128
    nwin = (D // ws[0], H // ws[1], W // ws[2])
    tok = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij"), -1)
    tok = tok.view(-1, 3)
    ncls = 2 ** len(axes)
    labels = torch.zeros(ncls, NP, dtype=torch.uint8)
    for cls in range(ncls):
        lab = torch.zeros(n, dtype=torch.int64)
        for i, ax in enumerate(axes):
            lab = lab * 2
            if (cls >> i) & 1:  # last window on this axis: tokens split at ws - shift
                lab = lab + (tok[:, ax] >= ws[ax] - ss[ax]).long()
        labels[cls, :n] = lab.to(torch.uint8)
    win = torch.stack(torch.meshgrid(torch.arange(nwin[0]), torch.arange(nwin[1]), torch.arange(nwin[2]), indexing="ij"), -1)
    win = win.view(-1, 3)
    cls_of = torch.zeros(win.shape[0], dtype=torch.int32)
    for i, ax in enumerate(axes):
        cls_of += ((win[:, ax] == nwin[ax] - 1).int() << i)
    return labels.to(device), cls_of.to(device)


# ---------------------------------------------------------------------------------------------------------
# parameter containers (same names / shapes / init as the reference; their torch forward is never used)
# ---------------------------------------------------------------------------------------------------------
class WindowAttention3D(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        trunc_normal_(self.relative_position_bias_table, std=.02)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class SwinTransformerBlock3D(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop_path):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.drop_path_rate = float(drop_path)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention3D(dim, window_size, num_heads, qkv_bias, qk_scale)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop_path, downsample):
        super().__init__()
        self.window_size = window_size
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth = depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim, num_heads, window_size, (0, 0, 0) if i % 2 == 0 else self.shift_size,
                                   mlp_ratio, qkv_bias, qk_scale, drop_path[i]) for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim, patch_norm):
        super().__init__()
        self.patch_size, self.in_chans, self.embed_dim = patch_size, in_chans, embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=(1, 4, 4))
        self.norm = nn.LayerNorm(embed_dim) if patch_norm else None


class SwinTransformer3D(nn.Module):
    """Signature of video_swin.py:409-427.  forward(x[B,3,T,H,W]) -> [B, 8C, T, H/32, W/32] (a permuted view of
    the channels-last result, as the reference returns after :478)."""

    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(2, 4, 4), in_chans=3, embed_dim=128,
                 depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=True, frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        assert drop_rate == 0. and attn_drop_rate == 0., "get_vidswin_model hard-codes 0 (video_swin.py:628-629)"
        assert patch_norm and tuple(patch_size) == (2, 4, 4) and in_chans == 3 and qkv_bias
        self.pretrained, self.pretrained2d = pretrained, pretrained2d
        self.num_layers, self.embed_dim = len(depths), embed_dim
        self.patch_norm, self.frozen_stages = patch_norm, frozen_stages
        self.window_size, self.patch_size = tuple(window_size), tuple(patch_size)
        self.depths, self.num_heads = list(depths), list(num_heads)
        self.patch_embed = PatchEmbed3D(self.patch_size, in_chans, embed_dim, patch_norm)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]  # video_swin.py:445
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(int(embed_dim * 2 ** i), depths[i], num_heads[i], self.window_size, mlp_ratio,
                                          qkv_bias, qk_scale, dpr[sum(depths[:i]):sum(depths[:i + 1])],
                                          downsample=i < self.num_layers - 1))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = nn.LayerNorm(self.num_features)

    def inflate_weights(self):
        """video_swin.py:482-533: 2-D Swin checkpoint (`{'model': state_dict}`) -> 3-D: the patch-embed kernel is repeated
        over the temporal patch axis and divided by it, every [(2Wh-1)(2Ww-1), nH] relative-position table is (bicubically
        resized if the 2-D window differs, then) repeated 2*Wd-1 times; index / mask buffers are always re-created."""
        checkpoint = torch.load(self.pretrained, map_location="cpu")
        state_dict = dict(checkpoint["model"])
        for k in [k for k in state_dict if "relative_position_index" in k or "attn_mask" in k]:
            del state_dict[k]
        w = state_dict["patch_embed.proj.weight"]
        state_dict["patch_embed.proj.weight"] = w.unsqueeze(2).repeat(1, 1, self.patch_size[0], 1, 1) / self.patch_size[0]
        own = self.state_dict()
        wd, wh, ww = self.window_size
        for k in [k for k in state_dict if "relative_position_bias_table" in k]:
            tab = state_dict[k]
            L1, nH1 = tab.shape
            nH2 = own[k].shape[1]
            L2 = (2 * wh - 1) * (2 * ww - 1)
            if nH1 != nH2:
                print(f"Error in loading {k}, passing")
            elif L1 != L2:
                S1 = int(L1 ** 0.5)
                tab = torch.nn.functional.interpolate(tab.permute(1, 0).view(1, nH1, S1, S1), size=(2 * wh - 1, 2 * ww - 1),
                                                      mode="bicubic").view(nH2, L2).permute(1, 0)
            state_dict[k] = tab.repeat(2 * wd - 1, 1)
        msg = self.load_state_dict(state_dict, strict=False)
        print(msg)
        print(f"=> loaded successfully '{self.pretrained}'")

    def init_weights(self, pretrained=None):
        """video_swin.py:535-568: trunc_normal(0.02) Linear weights, zero biases, unit LayerNorm; with a checkpoint path
        (`pretrained` or self.pretrained) the 2-D weights are inflated (pretrained2d) or a 3-D state dict is loaded."""
        def _init(m):
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str):
            self.apply(_init)
            print(f"load model from: {self.pretrained}")
            if self.pretrained2d:
                self.inflate_weights()
            else:   # directly load a 3-D model (the reference goes through mmcv's load_checkpoint, strict=False)
                missing, unexpected = self.load_state_dict(load_checkpoint_3d(self.pretrained), strict=False)
                print(f"Missing keys: {missing}\nUnexpected keys: {unexpected}")
        elif self.pretrained is None:
            self.apply(_init)
        else:
            raise TypeError("pretrained must be a str or None")

    # -----------------------------------------------------------------------------------------------------
    def forward_features(self, x, keep=None):
        """x: [B,3,T,H,W] (any strides) -> channels-last [B,T,H/32,W/32,8C] fp32."""
        require_cuda(x, "SwinTransformer3D")
        B = x.shape[0]
        if keep is None and self.training:
            keep = self.sample_drop_path(B, x.device)
        params = list(self.parameters())
        return _SwinFn.apply(x, self, keep, *params)

    def forward(self, x):
        return self.forward_features(x).permute(0, 4, 1, 2, 3)

    def sample_drop_path(self, B, device):
        """Per-sample DropPath factors floor(keep_prob + U)/keep_prob (video_swin.py:46-54), [n_blocks, 2, B]."""
        c = self.__dict__.get("_lav_keep_prob")
        if c is None or c.device != device:   # cached: no host->device copy per step (CUDA-graph capturable)
            rates = torch.tensor([blk.drop_path_rate for layer in self.layers for blk in layer.blocks])
            c = self.__dict__["_lav_keep_prob"] = (1.0 - rates).view(-1, 1, 1).to(device)
        kp, rates = c, c
        u = torch.rand(rates.numel(), 2, B, device=device)
        return (torch.floor(kp + u) / kp).contiguous()


# ---------------------------------------------------------------------------------------------------------
# fused forward / backward over the whole backbone
# ---------------------------------------------------------------------------------------------------------
def _im2col(x, T, dtype=F16):
    """[B,3,T,H,W] -> fp16 [B*T*(H/4)*(W/4), 96]; column = c*32 + kt*16 + kh*4 + kw, frame d+1 is zeros for
    d = T-1 (F.pad at video_swin.py:396 + Conv3d weight order)."""
    B, Cin, _, H, W = x.shape
    xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, 1))
    pat = xp.unfold(2, 2, 1).unfold(3, 4, 4).unfold(4, 4, 4)  # [B,3,T,h,w,2,4,4]
    return pat.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B * T * (H // 4) * (W // 4), Cin * 32).to(dtype)


class _SwinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, keep, *params):
        dev = x.device
        ar = arena_of(mod)
        ar.refresh16()
        B, _, T, H, W = x.shape
        assert H % 32 == 0 and W % 32 == 0, "H, W must be multiples of 32"
        C = mod.embed_dim
        saved = {"blocks": [], "merges": []}
        # ---- patch embed (K1, K2)
        pe = mod.patch_embed
        hp = precision.high("swin")
        if hp:
            cols32 = _im2col(x, T, F32)
            cols16 = cols32.to(F16)
        else:
            cols16 = _im2col(x, T)
        M = cols16.shape[0]
        y = empty32(M, C, device=dev)
        if hp:
            linear_fwd_hp(ar, cols32, pe.proj.weight.data.view(C, 96), pe.proj.bias, y)
            del cols32
        else:
            linear_fwd(cols16, ar.w16(pe.proj.weight).view(C, 96), pe.proj.bias, y)
        xcur = empty32(M, C, device=dev)
        pm, pr = empty32(M, device=dev), empty32(M, device=dev)
        ops.layernorm_fwd(y, pe.norm.weight, pe.norm.bias, pe.norm.eps, rows=M, C=C, out32=xcur, mean=pm, rstd=pr)
        saved["pe"] = (cols16, y, pm, pr)
        D, Hc, Wc = T, H // 4, W // 4
        bi = 0
        for s, layer in enumerate(mod.layers):
            ws, ss_full = get_window_size((D, Hc, Wc), layer.window_size, layer.shift_size)
            assert D % ws[0] == 0 and Hc % ws[1] == 0 and Wc % ws[2] == 0, \
                "feature map must divide the window (the reference's F.pad path is not restated)"
            N = ws[0] * ws[1] * ws[2]
            NP = (N + 127) // 128 * 128
            nW = (D // ws[0]) * (Hc // ws[1]) * (Wc // ws[2])
            rps = D * Hc * Wc
            for b, blk in enumerate(layer.blocks):
                ss = ss_full if any(v > 0 for v in blk.shift_size) else (0, 0, 0)
                rmap = window_row_map(B, D, Hc, Wc, ws, ss, dev)
                rinv = window_row_map_inv(B, D, Hc, Wc, ws, ss, dev) if FUSE_CASTS else None
                labels, cls_of = shift_mask_classes(D, Hc, Wc, ws, ss, dev)
                k1 = keep[bi, 0].contiguous() if keep is not None and blk.drop_path_rate > 0 else None
                k2 = keep[bi, 1].contiguous() if keep is not None and blk.drop_path_rate > 0 else None
                xcur, sv = _block_fwd(ar, blk, xcur, M, C, N, NP, B * nW, rmap, labels, cls_of, k1, k2, rps, dev)
                sv["rinv"] = rinv
                saved["blocks"].append(sv)
                bi += 1
            if layer.downsample is not None:
                ds = layer.downsample
                mmap = merge_row_map(B, D, Hc, Wc, dev)
                M4 = M // 4
                y16 = empty16(M4, 4 * C, device=dev)
                y32 = empty32(M4, 4 * C, device=dev) if hp else None
                mm, mr = empty32(M4, device=dev), empty32(M4, device=dev)
                ops.layernorm_fwd(xcur, ds.norm.weight, ds.norm.bias, ds.norm.eps, rows=M4, C=C, G=4, row_map=mmap,
                                  out16=y16, out32=y32, mean=mm, rstd=mr)
                xn = empty32(M4, 2 * C, device=dev)
                if hp:
                    linear_fwd_hp(ar, y32, ds.reduction.weight.data, None, xn)
                    del y32
                else:
                    linear_fwd(y16, ar.w16(ds.reduction.weight), None, xn)
                saved["merges"].append((xcur, y16, mm, mr, mmap, M, C))
                xcur, M, C, Hc, Wc = xn, M4, 2 * C, Hc // 2, Wc // 2
        out = empty32(M, C, device=dev)
        fm, fr = empty32(M, device=dev), empty32(M, device=dev)
        ops.layernorm_fwd(xcur, mod.norm.weight, mod.norm.bias, mod.norm.eps, rows=M, C=C, out32=out, mean=fm, rstd=fr)
        saved["final"] = (xcur, fm, fr, M, C)
        ctx.mod, ctx.saved, ctx.keep = mod, saved, keep
        ctx.geom = (B, T, H, W)
        return out.view(B, D, Hc, Wc, C)

    @staticmethod
    def backward(ctx, gout):
        mod, saved = ctx.mod, ctx.saved
        ar = arena_of(mod)
        if ar.on_swin_backward is not None:
            ar.on_swin_backward()  # data parallel: BERT + head gradients are final -> reduce them under this backward
        ar.prepare_grads(list(mod.parameters()))
        dev = gout.device
        B, T, H, W = ctx.geom
        xin, fm, fr, M, C = saved["final"]
        g = empty32(M, C, device=dev)
        gout2 = gout.reshape(M, C)
        if not gout2.is_contiguous():
            gout2 = gout2.contiguous()
        blocks = saved["blocks"]
        bi = len(blocks)
        # fused gradient casts: every LayerNorm backward that produces a block-input gradient also emits the fp16 operand
        # its consumer starts with (the next block's MLP branch: DropPath-scaled; a PatchMerging reduction: plain)
        g16 = None
        if FUSE_CASTS and bi > 0:
            g16 = empty16(M, C, device=dev)
            ops.layernorm_bwd(gout2, xin, mod.norm.weight, fm, fr, rows=M, C=C, dx32=g, dx16=g16, dx16_at_src=True,
                              dx16_scale=blocks[-1]["k2"], dx16_rps=blocks[-1]["rps"], dgamma=ar.g(mod.norm.weight),
                              dbeta=ar.g(mod.norm.bias))
        else:
            ops.layernorm_bwd(gout2, xin, mod.norm.weight, fm, fr, rows=M, C=C, dx32=g, dgamma=ar.g(mod.norm.weight),
                              dbeta=ar.g(mod.norm.bias))
        ds_ws = {}
        for s in range(mod.num_layers - 1, -1, -1):
            layer = mod.layers[s]
            if ar.on_swin_stage is not None:
                ar.on_swin_stage(s)   # data parallel: the gradients of the deeper stages are final -> reduce them now
            if layer.downsample is not None:
                ds = layer.downsample
                xprev, y16, mm, mr, mmap, Mp, Cp = saved["merges"][s]
                M4 = Mp // 4
                if g16 is None:
                    g16 = ops.scale_cast(g, empty16(M4, 2 * Cp, device=dev), rows=M4, C=2 * Cp)
                linear_wgrad(g16, y16, ar.g(ds.reduction.weight))
                dy16 = empty16(M4, 4 * Cp, device=dev)
                linear_dgrad(g16, ar.w16(ds.reduction.weight), dy16)
                g = empty32(Mp, Cp, device=dev)
                ops.layernorm_bwd(dy16, xprev, ds.norm.weight, mm, mr, rows=M4, C=Cp, G=4, row_map=mmap, dx32=g,
                                  dgamma=ar.g(ds.norm.weight), dbeta=ar.g(ds.norm.bias))
                g16 = None   # (the G = 4 scatter kernel has no fused fp16 output)
            for b in range(layer.depth - 1, -1, -1):
                bi -= 1
                # consumer of this block's input gradient: the previous block (its k2), a PatchMerging (plain), or the
                # patch embedding's LayerNorm (fp32: nothing to emit)
                emit = None
                if FUSE_CASTS and b > 0:
                    emit = (blocks[bi - 1]["k2"], blocks[bi - 1]["rps"])
                elif FUSE_CASTS and s > 0:
                    emit = (None, 1)
                g, g16 = _block_bwd(ar, layer.blocks[b], g, blocks[bi], ds_ws, dev, g16, emit)
        for old in ds_ws.values():   # the side stream may still read the last dS workspaces
            streams.hold(dev, *old["bufs"])
        cols16, y, pm, pr = saved["pe"]
        pe = mod.patch_embed
        M0, C0 = y.shape
        dy16 = empty16(M0, C0, device=dev)
        ops.layernorm_bwd(g, y, pe.norm.weight, pm, pr, rows=M0, C=C0, dx16=dy16, dgamma=ar.g(pe.norm.weight),
                          dbeta=ar.g(pe.norm.bias))
        linear_wgrad(dy16, cols16, ar.g(pe.proj.weight).view(C0, 96), ar.g(pe.proj.bias))
        # the input video needs no gradient (PatchEmbed3D dgrad is never computed in the reference either)
        return (None, None, None) + (None,) * len(list(mod.parameters()))


def _block_fwd(ar, blk, x, M, C, N, NP, nprob, rmap, labels, cls_of, k1, k2, rps, dev):
    at = blk.attn
    nh = blk.num_heads
    hd = C // nh
    hp = precision.high("swin")   # parity mode: split-fp16 linears on fp32 activations, fp32 attention output (precision.py)
    y16 = empty16(M, C, device=dev)
    y32 = empty32(M, C, device=dev) if hp else None
    m1, r1 = empty32(M, device=dev), empty32(M, device=dev)
    ops.layernorm_fwd(x, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, rows=M, C=C, row_map=rmap, out16=y16, out32=y32,
                      mean=m1, rstd=r1)
    qkv16 = empty16(M, 3 * C, device=dev)
    if hp:
        linear_fwd_hp(ar, y32, at.qkv.weight.data, at.qkv.bias, qkv16)
    else:
        linear_fwd(y16, ar.w16(at.qkv.weight), at.qkv.bias, qkv16)
    rel = _rel_index(at, N, dev)
    ncls = labels.shape[0] if labels is not None else 1
    dense = empty16(ncls, nh, NP, NP, device=dev)
    ops.relpos_bias_expand(at.relative_position_bias_table, rel, N, labels, dense, at.scale)
    o16 = empty16(M, C, device=dev)
    o32 = empty32(M, C, device=dev) if hp else None
    lse = empty32(nh, M, device=dev)
    ops.attn_fwd(qkv16, o16, lse, q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nh, nprob=nprob, L_tok=N,
                 scale=at.scale, bias16=dense, prob_class=cls_of, out32=o32)
    x1 = empty32(M, C, device=dev)
    kw = dict(residual=x, row_map=rmap, row_scale=k1, rows_per_scale=rps)
    if hp:
        linear_fwd_hp(ar, o32, at.proj.weight.data, at.proj.bias, x1, **kw)
    else:
        linear_fwd(o16, ar.w16(at.proj.weight), at.proj.bias, x1, **kw)
    y2 = empty16(M, C, device=dev)
    y2_32 = empty32(M, C, device=dev) if hp else None
    m2, r2 = empty32(M, device=dev), empty32(M, device=dev)
    ops.layernorm_fwd(x1, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, rows=M, C=C, out16=y2, out32=y2_32, mean=m2,
                      rstd=r2)
    a16 = empty16(M, 4 * C, device=dev)
    x2 = empty32(M, C, device=dev)
    kw = dict(residual=x1, row_scale=k2, rows_per_scale=rps)
    if hp:
        h32 = empty32(M, 4 * C, device=dev)
        linear_fwd_hp(ar, y2_32, blk.mlp.fc1.weight.data, blk.mlp.fc1.bias, h32, act=L.ACT_GELU, aux=a16)
        linear_fwd_hp(ar, h32, blk.mlp.fc2.weight.data, blk.mlp.fc2.bias, x2, **kw)
        h16 = cast16(h32)
        del y32, o32, y2_32, h32
    else:
        h16 = empty16(M, 4 * C, device=dev)
        linear_fwd(y2, ar.w16(blk.mlp.fc1.weight), blk.mlp.fc1.bias, h16, act=L.ACT_GELU, aux=a16)
        linear_fwd(h16, ar.w16(blk.mlp.fc2.weight), blk.mlp.fc2.bias, x2, **kw)
    sv = dict(x=x, y16=y16, m1=m1, r1=r1, qkv16=qkv16, dense=dense, o16=o16, lse=lse, x1=x1, y2=y2, m2=m2, r2=r2,
              a16=a16, h16=h16, rmap=rmap, cls_of=cls_of, k1=k1, k2=k2, rps=rps, N=N, NP=NP, nprob=nprob, M=M, C=C,
              rel=rel)
    return x2, sv


def _block_bwd(ar, blk, g, sv, ds_ws, dev, g16=None, emit=None):
    """g16: fp16(k2 * g) if the producer of g already emitted it; emit = (row scale or None, rows per scale) asks norm1's
    backward for the fp16 operand of this block's consumer.  Returns (input gradient fp32, that operand or None)."""
    at = blk.attn
    M, C, N, NP, nprob, rps = sv["M"], sv["C"], sv["N"], sv["NP"], sv["nprob"], sv["rps"]
    nh = blk.num_heads
    hd = C // nh
    fc1, fc2 = blk.mlp.fc1, blk.mlp.fc2
    # ---- MLP branch: x2 = x1 + k2 * (fc2(gelu(fc1(LN2(x1)))))
    if g16 is None:
        g16 = ops.scale_cast(g, empty16(M, C, device=dev), rows=M, C=C, row_scale=sv["k2"], rows_per_scale=rps)
    linear_wgrad(g16, sv["h16"], ar.g(fc2.weight), ar.g(fc2.bias))
    da16 = empty16(M, 4 * C, device=dev)
    linear_dgrad(g16, ar.w16(fc2.weight), da16, act=L.ACT_GELU_BWD, aux=sv["a16"])   # a16 = gelu'(fc1 output), saved by fc1's epilogue
    linear_wgrad(da16, sv["y2"], ar.g(fc1.weight), ar.g(fc1.bias))
    dy2 = empty16(M, C, device=dev)
    linear_dgrad(da16, ar.w16(fc1.weight), dy2)
    g1 = empty32(M, C, device=dev)
    # ---- attention branch: x1[map] = x[map] + k1 * proj(attn(qkv(LN1(x)[map])))
    if sv.get("rinv") is not None:   # norm2's backward also writes go16 = fp16(k1 * g1) in window order
        go16 = empty16(M, C, device=dev)
        ops.layernorm_bwd(dy2, sv["x1"], blk.norm2.weight, sv["m2"], sv["r2"], rows=M, C=C, add32=g, dx32=g1, dx16=go16,
                          dx16_map=sv["rinv"], dx16_scale=sv["k1"], dx16_rps=rps,
                          dgamma=ar.g(blk.norm2.weight), dbeta=ar.g(blk.norm2.bias))
    else:
        ops.layernorm_bwd(dy2, sv["x1"], blk.norm2.weight, sv["m2"], sv["r2"], rows=M, C=C, add32=g, dx32=g1,
                          dgamma=ar.g(blk.norm2.weight), dbeta=ar.g(blk.norm2.bias))
        go16 = ops.scale_cast(g1, empty16(M, C, device=dev), rows=M, C=C, row_map=sv["rmap"], row_scale=sv["k1"],
                              rows_per_scale=rps)
    linear_wgrad(go16, sv["o16"], ar.g(at.proj.weight), ar.g(at.proj.bias))
    do16 = empty16(M, C, device=dev)
    linear_dgrad(go16, ar.w16(at.proj.weight), do16)
    dq_acc = empty32(M, C, device=dev)   # zeroed by the attention backward's pre-pass
    dqkv16 = empty16(M, 3 * C, device=dev)
    # dS workspace: two buffers per shape, alternated, because the table-gradient reduction that reads one of them runs
    # on the side stream (off the critical path, like the weight gradients) while the next block's attention backward
    # already fills the other (up to 2 x 268 MB at stage 0, B=8)
    key = (nprob, nh, NP)
    slot = ds_ws.get(key)
    if slot is None:
        for old in ds_ws.values():
            streams.hold(dev, *old["bufs"])
        ds_ws.clear()
        slot = ds_ws[key] = {"bufs": [empty16(nprob, nh, NP, NP, device=dev) for _ in range(2 if streams.enabled() else 1)],
                             "ev": [None, None], "i": 0}
    i = slot["i"]
    ds16 = slot["bufs"][i]
    if slot["ev"][i] is not None:
        torch.cuda.current_stream(dev).wait_event(slot["ev"][i])   # its previous reader (two blocks ago) is done
    ops.attn_bwd(sv["qkv16"], sv["o16"], do16, sv["lse"], dq_acc, dqkv16, q_off=0, k_off=C, v_off=2 * C, head_dim=hd,
                 nheads=nh, nprob=nprob, L_tok=N, scale=at.scale, bias16=sv["dense"], prob_class=sv["cls_of"], ds16=ds16)
    ops.scale_cast(dq_acc, dqkv16, rows=M, C=C)  # Q block of dqkv (columns 0..C)
    if streams.enabled():
        side = streams.fork(dev)
        with torch.cuda.stream(side):
            ops.relpos_bias_grad(ds16, sv["rel"], N, ar.g(at.relative_position_bias_table))
            ev = slot["ev"][i] or torch.cuda.Event()
            ev.record(side)
            slot["ev"][i] = ev
        slot["i"] = i ^ 1
    else:
        ops.relpos_bias_grad(ds16, sv["rel"], N, ar.g(at.relative_position_bias_table))
    linear_wgrad(dqkv16, sv["y16"], ar.g(at.qkv.weight), ar.g(at.qkv.bias))
    dy1 = empty16(M, C, device=dev)
    linear_dgrad(dqkv16, ar.w16(at.qkv.weight), dy1)
    gn16 = None
    if emit is not None:
        gn16 = empty16(M, C, device=dev)
        ops.layernorm_bwd(dy1, sv["x"], blk.norm1.weight, sv["m1"], sv["r1"], rows=M, C=C, row_map=sv["rmap"], add32=g1,
                          dx32=g1, dx16=gn16, dx16_at_src=True, dx16_scale=emit[0], dx16_rps=emit[1],
                          dgamma=ar.g(blk.norm1.weight), dbeta=ar.g(blk.norm1.bias))
    else:
        ops.layernorm_bwd(dy1, sv["x"], blk.norm1.weight, sv["m1"], sv["r1"], rows=M, C=C, row_map=sv["rmap"], add32=g1,
                          dx32=g1, dgamma=ar.g(blk.norm1.weight), dbeta=ar.g(blk.norm1.bias))
    return g1, gn16


def _rel_index(attn, N, dev):
    c = attn.__dict__.get("_lav_rel")
    if c is None or c[0] != N or c[1].device != dev:
        c = (N, attn.relative_position_index[:N, :N].contiguous().to(device=dev, dtype=torch.int32))
        attn.__dict__["_lav_rel"] = c
    return c[1]


# ---------------------------------------------------------------------------------------------------------
def get_vidswin_model(args):
    """Same contract as video_swin.py:571-645: picks the variant from args.size_img / args.vis_backbone_size and
    builds SwinTransformer3D with mlp_ratio=4, qkv_bias=True, drop_path_rate=0.2 (hard-coded there, :616-634).
    The mmcv Config loader (visbackbone/config.py, needs addict+yapf) is replaced by the static table above."""
    size_img = int(args.size_img)
    size = args.vis_backbone_size
    if size_img == 384:
        assert size == "large"
    if size == "tiny":
        assert size_img == 224
    key = (size, 384 if size_img == 384 else 224)
    if key not in SWIN_VARIANTS:
        raise ValueError(f"unsupported video swin variant {key}")
    init = getattr(args, "vis_backbone_init", "random")
    model_path = None
    if init != "random":
        model_path = vidswin_checkpoint_path(args)
        if not os.path.isfile(model_path):
            # the reference dies in torch.load here (video_swin.py:636-639 / :494); starting a "pre-trained" run from random
            # weights silently would be a results bug, so this raises as well (utils/args.py:195-196 already forces
            # vis_backbone_init = "random" whenever --path_ckpt is given)
            raise FileNotFoundError(f"vis_backbone_init={init!r} needs {model_path} (relative to the working directory); "
                                    f"use vis_backbone_init='random' (or --path_ckpt) to train from scratch")
    pretrained2d = model_path if init == "2d" else None
    args.vis_backbone_pretrained_weight = model_path
    print("video swin random initialized" if init == "random" else
          f"video swin with pre-trained {init} (model path): {model_path}")
    m = SwinTransformer3D(pretrained=pretrained2d, pretrained2d=True, patch_size=(2, 4, 4), in_chans=3, mlp_ratio=4.,
                          qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2,
                          patch_norm=True, frozen_stages=-1, use_checkpoint=False, **SWIN_VARIANTS[key])
    if init == "3d" and model_path is not None:
        missing, unexpected = m.load_state_dict(load_checkpoint_3d(model_path), strict=False)
        print(f"Missing keys in loaded video_swin_transformerr: {missing}")
        print(f"Unexpected keys in loaded video_swin_transformer: {unexpected}")
    else:
        m.init_weights()
    return m


def vidswin_checkpoint_path(args):
    """./_models/... path rules of video_swin.py:572-593 (CWD-relative)."""
    size, init = args.vis_backbone_size, getattr(args, "vis_backbone_init", "random")
    kin = getattr(args, "kinetics", 400)
    if int(args.size_img) == 384:
        if init == "2d":
            return f"./_models/swin_transformer/swin_{size}_patch4_window12_384_22k.pth"
        return f"./_models/video_swin_transformer/swin_{size}_384_patch244_window81212_kinetics{kin}_22k.pth"
    if size != "tiny":
        if init == "2d":
            return f"./_models/swin_transformer/swin_{size}_patch4_window7_224_22k.pth"
        return f"./_models/video_swin_transformer/swin_{size}_patch244_window877_kinetics{kin}_22k.pth"
    if init == "2d":
        return f"./_models/swin_transformer/swin_{size}_patch4_window7_224.pth"
    return "./_models/video_swin_transformer/swin_tiny_patch244_window877_kinetics400_1k.pth"


def load_checkpoint_3d(model_path):
    """video_swin.py:647-654: `{'state_dict': ...}` with the `backbone.` prefix stripped."""
    sd = torch.load(model_path, map_location="cpu")["state_dict"]
    return {k.replace("backbone.", ""): v for k, v in sd.items()}
