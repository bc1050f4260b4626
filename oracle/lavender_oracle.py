"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch fp32, no custom kernels) of the LAVENDER
data-parallel forward hot path; backward comes from torch autograd over this restatement.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file.  The product path (`lavender_b200/`) never does.

Pinning: the reference ships no tests / golden vectors and its BERT arithmetic lives in
the un-pinned third-party `transformers` (5.5.0 in this image).  This restatement is pinned against the
reference's own modules executed in the build container: `oracle/make_golden.py` runs the unmodified
`/root/reference` `LAVENDER_Pretrain_MLM` (under `oracle/ref_shims.py`) on seeded inputs/weights and
commits the outputs to `tests/golden/`; `tests/test_oracle_golden.py` checks this file against them.

All functions operate on a flat `state_dict` with the reference's key names so that
the same dict drives the reference, this oracle and the CUDA path.

`set_operand_rounding(dtype)` makes every contraction round its two operands to `dtype` first
(accumulating in fp32) — a model of tensor-core arithmetic used by tests to separate "kernel is wrong"
from "fp16 operands differ from fp32 by this much".
"""
import math
from dataclasses import dataclass, field
from functools import lru_cache
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_Q = None  # operand-rounding dtype (None = exact fp32)
_Q_ACT = None  # activation-storage rounding dtype (None = keep fp32)


def set_operand_rounding(dtype, act_dtype=None):
    global _Q, _Q_ACT
    _Q, _Q_ACT = dtype, act_dtype


def _q(x):
    return x if _Q is None else x.to(_Q).to(torch.float32)


def _qa(x):
    """round a stored 16-bit activation (what the device path writes to HBM between kernels)"""
    return x if _Q_ACT is None else x.to(_Q_ACT).to(torch.float32)


def linear(x, w, b=None):
    y = _q(x) @ _q(w).t()
    return y if b is None else y + b


def bmm(a, b):
    return _q(a) @ _q(b)


def gelu(x):  # exact erf GELU: nn.GELU() (video_swin.py:64) / HF "gelu"
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


# ------------------------------------------------------------------------------------------------
# configuration tables (visbackbone/swin_tiny.py:4-17, swin_base.py:3-5, swin_large.py:3-5;
# only these keys are consumed: video_swin.py:616-634)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class SwinCfg:
    embed_dim: int = 96
    depths: Tuple[int, ...] = (2, 2, 6, 2)
    num_heads: Tuple[int, ...] = (3, 6, 12, 24)
    window: Tuple[int, int, int] = (8, 7, 7)
    patch: Tuple[int, int, int] = (2, 4, 4)
    drop_path_rate: float = 0.2  # hard-coded video_swin.py:630

    @property
    def num_features(self):
        return self.embed_dim * 2 ** (len(self.depths) - 1)


SWIN = {
    "tiny": SwinCfg(96, (2, 2, 6, 2), (3, 6, 12, 24), (8, 7, 7)),
    "base": SwinCfg(128, (2, 2, 18, 2), (4, 8, 16, 32), (8, 7, 7)),
    "large384": SwinCfg(192, (2, 2, 18, 2), (6, 12, 24, 48), (8, 12, 12)),
}


@dataclass(frozen=True)
class ModelCfg:
    swin: SwinCfg = SWIN["tiny"]
    hidden: int = 768
    bert_layers: int = 2
    bert_heads: int = 12
    bert_ffn: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    max_size_frame: int = 6   # model.py:12
    max_size_patch: int = 14  # model.py:13
    size_patch: int = 32      # main_pretrain_mlm.py:45
    vtm_batch: int = 4        # min(size_batch, 4) main_pretrain_mlm.py:50
    true_id: int = 2995
    false_id: int = 6270
    enable_task_token: bool = True


# ------------------------------------------------------------------------------------------------
# Swin helpers
# ------------------------------------------------------------------------------------------------
def get_window_size(x_size, window_size, shift_size=None):
    """video_swin.py:93-106: an axis whose extent <= window gets window=extent, shift=0."""
    ws = list(window_size)
    ss = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            ws[i] = x_size[i]
            if ss is not None:
                ss[i] = 0
    return tuple(ws) if ss is None else (tuple(ws), tuple(ss))


def window_partition(x, ws):
    """video_swin.py:82-86: [B,D,H,W,C] -> [B*nW, wd*wh*ww, C]; windows ordered (B,i,j,k)."""
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], C)


def window_reverse(win, ws, B, D, H, W):
    """video_swin.py:88-91."""
    x = win.view(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


@lru_cache(maxsize=None)
def relative_position_index(window):
    """video_swin.py:121-135, built for the *configured* window (8,.,.) and sliced [:N,:N] by callers."""
    wd, wh, ww = window
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    cf = coords.flatten(1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


@lru_cache(maxsize=None)
def compute_mask(D, H, W, ws, ss):
    """video_swin.py:290-305: region ids from 3 slabs per axis on the shifted frame; 0 / -100."""
    img = torch.zeros((1, D, H, W, 1))
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    mw = window_partition(img, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def window_attention(sd, p, x, mask, num_heads, window_cfg):
    """WindowAttention3D.forward video_swin.py:145-170. x:[B_,N,C]; mask:[nW,N,N] or None."""
    B_, N, C = x.shape
    hd = C // num_heads
    qkv = linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = _qa(qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4))
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = bmm(q, k.transpose(-2, -1))
    idx = relative_position_index(tuple(window_cfg))[:N, :N].reshape(-1).to(x.device)
    bias = sd[p + "relative_position_bias_table"][idx].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        mask = mask.to(x.device)
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(-1)
    out = _qa(bmm(attn, v).transpose(1, 2).reshape(B_, N, C))
    return linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def swin_block(sd, p, x, mask_matrix, num_heads, window_cfg, shift_cfg, keep1=None, keep2=None):
    """SwinTransformerBlock3D.forward video_swin.py:204-261 (no padding: extents divide the window).
    keep1/keep2: optional per-sample DropPath factors (mask/keep_prob, video_swin.py:46-54)."""
    B, D, H, W, C = x.shape
    ws, ss = get_window_size((D, H, W), window_cfg, shift_cfg)
    assert D % ws[0] == 0 and H % ws[1] == 0 and W % ws[2] == 0, "padding path not restated"
    h = _qa(F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5))
    shifted = any(s > 0 for s in ss)
    if shifted:
        h = torch.roll(h, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
    win = window_partition(h, ws)
    win = window_attention(sd, p + "attn.", win, mask_matrix if shifted else None, num_heads, window_cfg)
    h = window_reverse(win.view(-1, *(ws + (C,))), ws, B, D, H, W)
    if shifted:
        h = torch.roll(h, shifts=ss, dims=(1, 2, 3))
    if keep1 is not None:
        h = h * keep1.view(B, 1, 1, 1, 1)
    x = x + h
    h = _qa(F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5))
    h = _qa(gelu(linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])))
    h = linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if keep2 is not None:
        h = h * keep2.view(B, 1, 1, 1, 1)
    return x + h


def patch_merging(sd, p, x):
    """PatchMerging.forward video_swin.py:271-287 (even H, W)."""
    x0 = x[:, :, 0::2, 0::2, :]
    x1 = x[:, :, 1::2, 0::2, :]
    x2 = x[:, :, 0::2, 1::2, :]
    x3 = x[:, :, 1::2, 1::2, :]
    x = torch.cat([x0, x1, x2, x3], -1)
    x = _qa(F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5))
    return linear(x, sd[p + "reduction.weight"])


def patch_embed(sd, p, x, cfg: SwinCfg):
    """PatchEmbed3D.forward video_swin.py:388-405. x:[B,3,T,H,W] -> [B,T,H/4,W/4,C] channels-last.
    Conv3d(k=(2,4,4), stride=(1,4,4)) after one zero frame appended == per-token linear over the
    (c, kt, kh, kw) patch of frames (d, d+1)."""
    B, Cin, T, H, W = x.shape
    pt, ph, pw = cfg.patch
    assert H % ph == 0 and W % pw == 0
    x = F.pad(x, (0, 0, 0, 0, 0, 1))
    pat = x.unfold(2, pt, 1).unfold(3, ph, ph).unfold(4, pw, pw)  # [B,3,T,h,w,pt,ph,pw]
    pat = pat.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B, T, H // ph, W // pw, Cin * pt * ph * pw)
    w = sd[p + "proj.weight"].reshape(cfg.embed_dim, -1)
    y = linear(_qa(pat), w, sd[p + "proj.bias"])
    return F.layer_norm(y, (cfg.embed_dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def swin_forward(sd, p, x, cfg: SwinCfg, keep=None):
    """SwinTransformer3D.forward video_swin.py:468-480 + BasicLayer.forward :350-368.
    x:[B,3,T,H,W] -> [B,T,H/32,W/32,8C] (channels-last; the reference returns the channels-first permute).
    keep: optional [n_blocks, 2, B] DropPath factors."""
    x = patch_embed(sd, p + "patch_embed.", x, cfg)
    bi = 0
    for s, depth in enumerate(cfg.depths):
        B, D, H, W, C = x.shape
        shift_cfg = tuple(i // 2 for i in cfg.window)
        ws, ss = get_window_size((D, H, W), cfg.window, shift_cfg)
        mask = compute_mask(D, H, W, ws, ss)
        for b in range(depth):
            k1 = keep[bi, 0] if keep is not None else None
            k2 = keep[bi, 1] if keep is not None else None
            x = swin_block(sd, f"{p}layers.{s}.blocks.{b}.", x, mask, cfg.num_heads[s], cfg.window,
                           (0, 0, 0) if b % 2 == 0 else shift_cfg, k1, k2)
            bi += 1
        if s < len(cfg.depths) - 1:
            x = patch_merging(sd, f"{p}layers.{s}.downsample.", x)
    return F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


# ------------------------------------------------------------------------------------------------
# EncVideo / EncTxt / fusion BERT / MLM head
# ------------------------------------------------------------------------------------------------
def enc_video(sd, img, cfg: ModelCfg, keep=None, odr=None, vt_mask=None):
    """EncVideo.forward model.py:37-93 (non-swinbert). img:[B,T,3,H,W]; odr: per-clip frame order (model.py:72-81:
    frame i keeps emb_len[i] only where odr[b][i] == i, else emb_odr); vt_mask [B,T,1+hw] multiplies the video key
    mask (model.py:87-91)."""
    B, T, _, H, W = img.shape
    h, w = H // 32, W // 32
    f = swin_forward(sd, "enc_img.swin.", img.transpose(1, 2), cfg.swin, keep)  # [B,T,h,w,8C]
    f = f.reshape(B, T, h * w, cfg.swin.num_features)
    if "enc_img.fc.weight" in sd:
        f = linear(_qa(f), sd["enc_img.fc.weight"], sd["enc_img.fc.bias"])
    f = torch.cat([sd["enc_img.emb_cls"].expand(B, T, -1, -1), f], dim=2)
    f = f + sd["enc_img.emb_pos"][:, :, :1 + h * w, :]
    if odr is not None:
        rows = [torch.cat([sd["enc_img.emb_len"][:, i:i + 1] if i == int(p) else sd["enc_img.emb_odr"]
                           for i, p in enumerate(odr[b])], dim=1) for b in range(B)]
        f = f + torch.cat(rows, dim=0)
    else:
        f = f + sd["enc_img.emb_len"][:, :T]
    f = F.layer_norm(f, (cfg.hidden,), sd["enc_img.norm.weight"], sd["enc_img.norm.bias"], 1e-5)
    f = f.view(B, T * (1 + h * w), cfg.hidden)
    m = torch.ones(B, T, 1 + h * w, dtype=torch.long, device=img.device)
    if vt_mask is not None:
        m = m * vt_mask
    return f, m.view(B, T * (1 + h * w))


def bert_embeddings(sd, ids, p="enc_txt.emb_txt."):
    """EncTxt.forward model.py:125-142 with txt_backbone_embed_only -> HF BertEmbeddings (eval)."""
    L = ids.shape[-1]
    e = sd[p + "word_embeddings.weight"][ids] + sd[p + "position_embeddings.weight"][:L] \
        + sd[p + "token_type_embeddings.weight"][0]
    return F.layer_norm(e, (e.shape[-1],), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)


def bert_layer(sd, p, x, ext_mask, n_heads):
    """HF BertLayer (post-LN), eval mode; formulas SURVEY §8a-A10."""
    B, L, Hd = x.shape
    hd = Hd // n_heads
    xq = _qa(x)

    def heads(t):
        return _qa(t).view(B, L, n_heads, hd).transpose(1, 2)

    q = heads(linear(xq, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]))
    k = heads(linear(xq, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]))
    v = heads(linear(xq, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]))
    s = bmm(q, k.transpose(-1, -2)) / math.sqrt(hd) + ext_mask
    ctx = _qa(bmm(s.softmax(-1), v).transpose(1, 2).reshape(B, L, Hd))
    a = linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    a = F.layer_norm(a + x, (Hd,), sd[p + "attention.output.LayerNorm.weight"],
                     sd[p + "attention.output.LayerNorm.bias"], 1e-12)
    i = _qa(gelu(linear(_qa(a), sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"])))
    o = linear(i, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return F.layer_norm(o + a, (Hd,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)


def extended_mask(mask):
    """HF get_extended_attention_mask: [B,L] -> [B,1,1,L]; [B,L,L] -> [B,1,L,L]; (1-m)*finfo.min."""
    m = mask[:, None, None, :] if mask.dim() == 2 else mask[:, None, :, :]
    return (1.0 - m.to(torch.float32)) * torch.finfo(torch.float32).min


def go_cross(sd, feat_img, mask_img, feat_txt, mask_txt, cfg: ModelCfg):
    """LAVENDER_Base.go_cross model.py:223-243 (attn_mask_type='full', no pretxt)."""
    x = torch.cat([feat_img, feat_txt], dim=1)
    ext = extended_mask(torch.cat([mask_img, mask_txt], dim=1))
    for l in range(cfg.bert_layers):
        x = bert_layer(sd, f"trsfr.layer.{l}.", x, ext, cfg.bert_heads)
    return x


def seq2seq_mask(mask_img, Lt, n_pre=0):
    """LAVENDER_Base.get_attn_mask model.py:194-221 (attn_mask_type='seq2seq'): [B, L, L] with L = Lv + n_pre + Lt;
    every query sees the (masked) video + prefix keys, the Lt text queries see the text keys causally (the text padding
    mask is NOT applied: model.py:213-214 builds the triangle from ones)."""
    B, Lv = mask_img.shape
    full = torch.cat([mask_img, torch.ones(B, n_pre, dtype=mask_img.dtype, device=mask_img.device)], dim=1)
    Lf, L = Lv + n_pre, Lv + n_pre + Lt
    m = torch.zeros(B, L, L, dtype=torch.long, device=mask_img.device)
    m[:, :, :Lf] = full.unsqueeze(1)
    m[:, Lf:, Lf:] = torch.tril(torch.ones(Lt, Lt, dtype=torch.long, device=mask_img.device))
    return m


def go_cross_masked(sd, feat, mask, cfg: ModelCfg, decoder=False):
    """go_cross on an already concatenated sequence with a [B, L] or [B, L, L] 0/1 mask.  decoder=True restates HF's
    get_extended_attention_mask for config.is_decoder (model_for_captioning.py:43 sets it on the shared config): a 2-D
    padding mask is AND-ed with a causal mask over the whole sequence."""
    if decoder and mask.dim() == 2:
        L = mask.shape[1]
        mask = torch.tril(torch.ones(L, L, dtype=mask.dtype, device=mask.device)).unsqueeze(0) * mask.unsqueeze(1)
    ext = extended_mask(mask)
    x = feat
    for l in range(cfg.bert_layers):
        x = bert_layer(sd, f"trsfr.layer.{l}.", x, ext, cfg.bert_heads)
    return x


TASK_TOK2ID = {"vtm": 0, "mc": 1, "oe": 2, "cap": 3}   # main_pretrain_mlm.py:51, model_for_captioning.py:48


def add_task_token(sd, cfg: ModelCfg, txt_like, mask_txt, feat_txt, task_name):
    """prepro_txt_inputs with enable_task_token (model.py:248-306): prefix (id 0, mask 1, emb_task[task]) on batched rows."""
    if not cfg.enable_task_token:
        return txt_like, mask_txt, feat_txt
    n = feat_txt.shape[0]
    e = sd["emb_task"][TASK_TOK2ID[task_name]].view(1, 1, -1).expand(n, -1, -1)
    z = torch.zeros(n, 1, dtype=txt_like.dtype, device=txt_like.device)
    o = torch.ones(n, 1, dtype=mask_txt.dtype, device=mask_txt.device)
    return torch.cat([z, txt_like], 1), torch.cat([o, mask_txt], 1), torch.cat([e, feat_txt], 1)


def multitask_forward(sd, batch, cfg: ModelCfg, task, task_name, decoder=False):
    """LAVENDER_Multi_Task.forward main_multi_task_mlm.py:82-225 and LAVENDER_Captioning.encode_forward
    model_for_captioning.py:61-93 (task-token prefix, no prompt).  Returns (logits, labels)."""
    img = batch["img"]
    B, T, _, H, W = img.shape
    Lv = (1 + (H // 32) * (W // 32)) * T
    n_pre = 1 if cfg.enable_task_token else 0
    if "captioning" in task:
        feat_img, mask_img = enc_video(sd, img, cfg)
        feat_txt = bert_embeddings(sd, batch["txt"])
        ans, _, feat_txt = add_task_token(sd, cfg, batch["ans_mtm"], batch["mask"], feat_txt, "cap")
        ans = ans.clone()
        ans[:, :n_pre] = -1
        m = seq2seq_mask(mask_img, batch["txt"].shape[1], n_pre)
        out = go_cross_masked(sd, torch.cat([feat_img, feat_txt], 1), m, cfg)
        return mlm_head(sd, out[:, Lv:]), ans
    if "retrieval" in task:
        txt, mask, vid = batch["txt"], batch["mask"], batch["vid"]
        feat_img, mask_img = enc_video(sd, img, cfg)
        feat_txt = bert_embeddings(sd, txt)
        vi = torch.arange(B).repeat_interleave(B)
        ti = torch.arange(B).repeat(B)
        t, mt, ft = add_task_token(sd, cfg, txt[ti], mask[ti], feat_txt[ti], task_name)
        ans = torch.full_like(t, -1)
        same = torch.tensor([vid[int(i)] == vid[int(j)] for i, j in zip(vi, ti)])
        ans[:, -1] = torch.where(same, torch.tensor(cfg.true_id), torch.tensor(cfg.false_id))
        out = go_cross_masked(sd, torch.cat([feat_img[vi], ft], 1), torch.cat([mask_img[vi], mt], 1), cfg, decoder)
        return mlm_head(sd, out[:, Lv:]), ans
    if "qamc" in task and "lsmdc-mc" in task:
        txt, mask, ans = batch["txt"], batch["mask"], batch["mask_ans"]
        O_ = txt.shape[1]
        feat_img, mask_img = enc_video(sd, img, cfg)
        feat_txt = bert_embeddings(sd, txt.flatten(0, 1))
        vi = torch.arange(B).repeat_interleave(O_)
        a, mt, ft = add_task_token(sd, cfg, ans.flatten(0, 1), mask.flatten(0, 1), feat_txt, task_name)
        a = a.clone()
        a[:, :n_pre] = -1
        out = go_cross_masked(sd, torch.cat([feat_img[vi], ft], 1), torch.cat([mask_img[vi], mt], 1), cfg, decoder)
        return mlm_head(sd, out[:, Lv:]), a.view(B, O_, -1)
    # qamc / qaoe
    txt, mask, ans = batch["txt"], batch["mask"], batch["mask_ans"]
    feat_img, mask_img = enc_video(sd, img, cfg)
    feat_txt = bert_embeddings(sd, txt)
    a, mt, ft = add_task_token(sd, cfg, ans, mask, feat_txt, task_name)
    a = a.clone()
    a[:, :n_pre] = -1
    out = go_cross_masked(sd, torch.cat([feat_img, ft], 1), torch.cat([mask_img, mt], 1), cfg, decoder)
    return mlm_head(sd, out[:, Lv:]), a


def make_multitask_batch(task, B, seed=0, X=25, O_=5, vocab=30522):
    """Seeded synthetic batches with the keys the multi-task forwards read (main_multi_task_mlm.py:108-225,
    model_for_captioning.py:61-70): padded captions / questions, a [MASK] answer slot for QA, duplicate video ids for
    retrieval, BERT-style masked caption tokens for captioning."""
    g = torch.Generator().manual_seed(7000 + seed)
    img = torch.randn(B, 5, 3, 224, 224, generator=g)

    def sent(n, X):
        t = torch.randint(1000, min(30000, vocab), (n, X), generator=g)
        m = torch.ones(n, X, dtype=torch.long)
        t[:, 0] = 101
        for r in range(n):
            ln = int(torch.randint(X // 2, X, (1,), generator=g))   # [CLS] w.. [SEP] pad..
            t[r, ln - 1] = 102
            t[r, ln:] = 0
            m[r, ln:] = 0
        return t, m
    if "captioning" in task:
        txt, mask = sent(B, X)
        ans = torch.full((B, X), -1, dtype=torch.long)
        sel = (torch.rand(B, X, generator=g) < 0.3) & (mask == 1) & (txt != 101)
        sel[:, 2] = True
        ans[sel] = txt[sel]
        txt[sel] = 103
        return {"img": img, "txt": txt, "mask": mask, "ans_mtm": ans}
    if "retrieval" in task:
        txt, mask = sent(B, X)
        txt[:, -1], mask[:, -1] = 103, 1            # appended [MASK] answer slot (Dataset_Retrieval_MLM)
        vid = list(range(B))
        if B > 2:
            vid[-1] = vid[0]                        # two captions of the same video: both pairs are positives
        return {"img": img, "txt": txt, "mask": mask, "vid": vid}
    if "qamc" in task and "lsmdc-mc" in task:
        txt, mask = sent(B * O_, X)
        txt[:, -1], mask[:, -1] = 103, 1
        ans = torch.full((B * O_, X), -1, dtype=torch.long)
        gt = torch.randint(0, O_, (B,), generator=g)
        for b in range(B):
            for o in range(O_):
                ans[b * O_ + o, -1] = 2995 if int(gt[b]) == o else 6270
        return {"img": img, "txt": txt.view(B, O_, X), "mask": mask.view(B, O_, X), "mask_ans": ans.view(B, O_, X)}
    txt, mask = sent(B, X)
    txt[:, -1], mask[:, -1] = 103, 1
    ans = torch.full((B, X), -1, dtype=torch.long)
    ans[:, -1] = torch.randint(1000, 5000, (B,), generator=g)
    return {"img": img, "txt": txt, "mask": mask, "mask_ans": ans}


def mlm_head(sd, x, p="fc_mtm.predictions."):
    """HF BertOnlyMLMHead (main_pretrain_mlm.py:46-48,69,115); decoder.bias aliases predictions.bias."""
    t = gelu(linear(_qa(x), sd[p + "transform.dense.weight"], sd[p + "transform.dense.bias"]))
    t = _qa(F.layer_norm(t, (t.shape[-1],), sd[p + "transform.LayerNorm.weight"],
                         sd[p + "transform.LayerNorm.bias"], 1e-12))
    return linear(t, sd[p + "decoder.weight"], sd[p + "bias"])


def draw_negatives(B, O):
    """Same numpy global-RNG call sequence as main_pretrain_mlm.py:90-91."""
    return [np.random.permutation([j for j in range(B) if j != i])[:max(O - 1, 0)].tolist()
            for i in range(B)]


def vtm_pairs(B, O, negs):
    """(video index, text index, is_positive) in the order of main_pretrain_mlm.py:74-106."""
    pairs = []
    for i in range(B):
        pairs.append((i, i, True))
        for j in range(O - 1):
            pairs.append((i, int(negs[i][j]), False))
    return pairs


def pretrain_forward(sd, batch, cfg: ModelCfg, negs=None, keep=None):
    """LAVENDER_Pretrain_MLM.forward main_pretrain_mlm.py:55-119 (eval-mode arithmetic).
    batch: img[B,T,3,H,W] f32, txt[B,X] i64, mask[B,X] i64, ans_mtm[B,X] i64."""
    img, txt, mask = batch["img"], batch["txt"], batch["mask"]
    B, T, _, H, W = img.shape
    Lv = (1 + (H // cfg.size_patch) * (W // cfg.size_patch)) * T
    O = min(B, cfg.vtm_batch)
    feat_img, mask_img = enc_video(sd, img, cfg, keep, vt_mask=batch.get("vt_mask"))
    feat_txt = bert_embeddings(sd, txt)
    out = go_cross(sd, feat_img, mask_img, feat_txt, mask, cfg)
    out_mtm = mlm_head(sd, out[:, Lv:])
    if negs is None:
        negs = draw_negatives(B, O)
    pairs = vtm_pairs(B, O, negs)
    vi = torch.tensor([p[0] for p in pairs], device=img.device)
    ti = torch.tensor([p[1] for p in pairs], device=img.device)
    ft, mt, tt = feat_txt[ti], mask[ti], txt[ti]
    if cfg.enable_task_token:  # model.py:248-265,292-306: emb_task[0] row prefixed, mask 1, txt id 0
        n = len(pairs)
        ft = torch.cat([sd["emb_task"][0].view(1, 1, -1).expand(n, -1, -1), ft], dim=1)
        mt = torch.cat([torch.ones(n, 1, dtype=mt.dtype, device=mt.device), mt], dim=1)
        tt = torch.cat([torch.zeros(n, 1, dtype=tt.dtype, device=tt.device), tt], dim=1)
    ans_vtm = torch.full_like(tt, -1)
    ans_vtm[:, -1] = torch.tensor([cfg.true_id if p[2] else cfg.false_id for p in pairs], device=tt.device)
    out = go_cross(sd, feat_img[vi], mask_img[vi], ft, mt, cfg)
    out_vtm = mlm_head(sd, out[:, Lv:])
    return {"out_mtm": out_mtm, "out_vtm": out_vtm, "ans_mtm": batch.get("ans_mtm"), "ans_vtm": ans_vtm}


def pretrain_loss(out):
    """Agent_Pretrain_MLM.step main_pretrain_mlm.py:158-163: CE(ignore_index=-1) x2, summed."""
    ce = torch.nn.CrossEntropyLoss(ignore_index=-1)
    ls_mtm = ce(out["out_mtm"].flatten(0, 1), out["ans_mtm"].flatten())
    ls_vtm = ce(out["out_vtm"].flatten(0, 1), out["ans_vtm"].flatten())
    return ls_mtm + ls_vtm, ls_mtm, ls_vtm


# ------------------------------------------------------------------------------------------------
# deterministic weights / inputs (shared by the golden generator, the CPU tests and the GPU tests)
# ------------------------------------------------------------------------------------------------
def state_dict_schema(cfg: ModelCfg):
    """[(key, shape)] in the reference's state_dict order (dump: SURVEY §8b; verified by make_golden)."""
    H, s = cfg.hidden, cfg.swin
    out = [("emb_task", (10, H))]
    e = "enc_txt.emb_txt."
    out += [(e + "word_embeddings.weight", (cfg.vocab, H)), (e + "position_embeddings.weight", (cfg.max_pos, H)),
            (e + "token_type_embeddings.weight", (2, H)), (e + "LayerNorm.weight", (H,)), (e + "LayerNorm.bias", (H,))]
    for l in range(cfg.bert_layers):
        p = f"trsfr.layer.{l}."
        for n in ("query", "key", "value"):
            out += [(p + f"attention.self.{n}.weight", (H, H)), (p + f"attention.self.{n}.bias", (H,))]
        out += [(p + "attention.output.dense.weight", (H, H)), (p + "attention.output.dense.bias", (H,)),
                (p + "attention.output.LayerNorm.weight", (H,)), (p + "attention.output.LayerNorm.bias", (H,)),
                (p + "intermediate.dense.weight", (cfg.bert_ffn, H)), (p + "intermediate.dense.bias", (cfg.bert_ffn,)),
                (p + "output.dense.weight", (H, cfg.bert_ffn)), (p + "output.dense.bias", (H,)),
                (p + "output.LayerNorm.weight", (H,)), (p + "output.LayerNorm.bias", (H,))]
    v = "enc_img."
    out += [(v + "emb_cls", (1, 1, 1, H)), (v + "emb_pos", (1, 1, 1 + cfg.max_size_patch ** 2, H)),
            (v + "emb_len", (1, cfg.max_size_frame, 1, H)), (v + "emb_odr", (1, 1, 1, H))]
    w = v + "swin."
    C = s.embed_dim
    out += [(w + "patch_embed.proj.weight", (C, 3) + tuple(s.patch)), (w + "patch_embed.proj.bias", (C,)),
            (w + "patch_embed.norm.weight", (C,)), (w + "patch_embed.norm.bias", (C,))]
    ntab = (2 * s.window[0] - 1) * (2 * s.window[1] - 1) * (2 * s.window[2] - 1)
    nwin = s.window[0] * s.window[1] * s.window[2]
    for st, depth in enumerate(s.depths):
        c = C * 2 ** st
        for b in range(depth):
            p = f"{w}layers.{st}.blocks.{b}."
            out += [(p + "norm1.weight", (c,)), (p + "norm1.bias", (c,)),
                    (p + "attn.relative_position_bias_table", (ntab, s.num_heads[st])),
                    (p + "attn.relative_position_index", (nwin, nwin)),
                    (p + "attn.qkv.weight", (3 * c, c)), (p + "attn.qkv.bias", (3 * c,)),
                    (p + "attn.proj.weight", (c, c)), (p + "attn.proj.bias", (c,)),
                    (p + "norm2.weight", (c,)), (p + "norm2.bias", (c,)),
                    (p + "mlp.fc1.weight", (4 * c, c)), (p + "mlp.fc1.bias", (4 * c,)),
                    (p + "mlp.fc2.weight", (c, 4 * c)), (p + "mlp.fc2.bias", (c,))]
        if st < len(s.depths) - 1:
            p = f"{w}layers.{st}.downsample."
            out += [(p + "reduction.weight", (2 * c, 4 * c)), (p + "norm.weight", (4 * c,)), (p + "norm.bias", (4 * c,))]
    out += [(w + "norm.weight", (s.num_features,)), (w + "norm.bias", (s.num_features,))]
    if s.num_features != H:  # model.py:16-20 (module order: swin, fc, norm)
        out += [(v + "fc.weight", (H, s.num_features)), (v + "fc.bias", (H,))]
    out += [(v + "norm.weight", (H,)), (v + "norm.bias", (H,))]
    f = "fc_mtm.predictions."
    out += [(f + "bias", (cfg.vocab,)), (f + "transform.dense.weight", (H, H)), (f + "transform.dense.bias", (H,)),
            (f + "transform.LayerNorm.weight", (H,)), (f + "transform.LayerNorm.bias", (H,)),
            (f + "decoder.weight", (cfg.vocab, H)), (f + "decoder.bias", (cfg.vocab,))]
    return out


def make_state_dict(cfg: ModelCfg, seed=0, requires_grad=False):
    """Deterministic non-trivial weights from a CPU generator (bit-reproducible for a given torch build):
    matrices/embeddings ~ 0.02*N(0,1) (x2.5 for rel-pos tables so the bias matters), norm weights
    1+0.1*N, biases 0.02*N — every affine term is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in state_dict_schema(cfg):
        if k.endswith("relative_position_index"):
            sd[k] = relative_position_index(tuple(cfg.swin.window)).clone()
            continue
        if k == "fc_mtm.predictions.decoder.bias":
            sd[k] = sd["fc_mtm.predictions.bias"]  # one parameter, two keys (Q24 / model.py:470)
            continue
        r = torch.randn(shape, generator=g)
        is_norm_w = k.endswith("weight") and ("norm" in k.lower()) and len(shape) == 1
        if is_norm_w:
            t = 1.0 + 0.1 * r
        elif k.endswith("relative_position_bias_table"):
            t = 0.05 * r
        else:
            t = 0.02 * r
        sd[k] = t.requires_grad_(requires_grad)
    return sd


def make_batch(B, T=5, H=224, W=224, Lt=33, seed=0, p_mask=0.15, vocab=30522):
    """Seeded synthetic batch of SURVEY §8d: randn frames, random ids with [CLS]/[SEP]/[MASK] placed as
    Dataset_Pretrain_MLM.str2txt does (main_pretrain_mlm.py:22-25), fixed MLM masking of ~p_mask."""
    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.randn(B, T, 3, H, W, generator=g)
    txt = torch.randint(1000, min(30000, vocab), (B, Lt), generator=g)
    txt[:, 0], txt[:, -2], txt[:, -1] = 101, 102, 103
    mask = torch.ones(B, Lt, dtype=torch.long)
    ans = torch.full((B, Lt), -1, dtype=torch.long)
    sel = torch.rand(B, Lt, generator=g) < p_mask
    sel[:, 0] = False
    sel[:, -2:] = False
    sel[:, 5] = True  # at least one labelled position per row
    ans[sel] = txt[sel]
    txt[sel] = 103
    return {"img": img, "txt": txt, "mask": mask, "ans_mtm": ans}


def sample_flat(t, n=64):
    """n evenly spaced entries of a flattened tensor (integer index arithmetic; shared with make_golden)."""
    f = t.detach().reshape(-1)
    n = min(n, f.numel())
    idx = (torch.arange(n, dtype=torch.int64) * (f.numel() - 1)) // max(n - 1, 1)
    return f[idx]
