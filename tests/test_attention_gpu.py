"""Fused tcgen05 attention (forward + backward) vs an fp32 PyTorch restatement of
WindowAttention3D.forward (video_swin.py:145-170) and HF BertSelfAttention, on the same fp16 operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _window_case(nimg, nW, nheads, shifted, seed=0):
    import lavender_oracle as O
    g = torch.Generator().manual_seed(seed)
    L, hd, C = 245, 32, nheads * 32
    nprob = nimg * nW
    qkv = (torch.randn(nprob * L, 3 * C, generator=g) * 0.7).half().cuda()
    table = (torch.randn(2535, nheads, generator=g) * 0.5).cuda()
    rel = O.relative_position_index((8, 7, 7))[:L, :L].contiguous().int().cuda()
    if shifted:  # classes: 0 interior, 1 last col, 2 last row, 3 corner  (3x3 regions of compute_mask restricted to a window)
        side = int(math.isqrt(nW))
        H = side * 7
        mask = O.compute_mask(5, H, H, (5, 7, 7), (0, 3, 3))  # [nW, L, L] 0 / -100
        hh = torch.arange(7).view(1, 7, 1).expand(5, 7, 7).reshape(-1)
        ww = torch.arange(7).view(1, 1, 7).expand(5, 7, 7).reshape(-1)
        lab = torch.zeros(4, 256, dtype=torch.uint8)
        lab[1, :L] = (ww >= 4).to(torch.uint8)
        lab[2, :L] = (hh >= 4).to(torch.uint8) * 2
        lab[3, :L] = (hh >= 4).to(torch.uint8) * 2 + (ww >= 4).to(torch.uint8)
        cls = torch.tensor([(2 if (w // side) == side - 1 else 0) + (1 if (w % side) == side - 1 else 0)
                            for w in range(nW)], dtype=torch.int32)
        # the class labelling must reproduce the reference mask exactly
        for w in range(nW):
            l = lab[cls[w], :L].int()
            assert torch.equal((l[:, None] != l[None, :]), mask[w] != 0)
        return qkv, table, rel, lab.cuda(), cls.cuda(), mask.cuda(), nprob, L, hd, C
    return qkv, table, rel, None, None, None, nprob, L, hd, C


@pytest.mark.parametrize("nimg,nW,nheads,shifted", [(2, 4, 3, False), (2, 4, 4, True), (1, 16, 2, True), (3, 1, 8, False)])
def test_window_attention_fwd_bwd(nimg, nW, nheads, shifted):
    from lavender_b200 import ops
    qkv, table, rel, lab, cls, mask, nprob, L, hd, C = _window_case(nimg, nW, nheads, shifted)
    ncls = 4 if shifted else 1
    dense = torch.zeros(ncls, nheads, 256, 256, device="cuda", dtype=torch.float16)
    scale = hd ** -0.5
    ops.relpos_bias_expand(table, rel, L, lab, dense, scale)
    out = torch.zeros(nprob * L, C, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nheads, nprob * L, device="cuda")
    ops.attn_fwd(qkv, out, lse, q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nheads, nprob=nprob, L_tok=L,
                 scale=scale, bias16=dense, prob_class=cls)
    torch.cuda.synchronize()

    # reference (fp32 math on the same fp16 inputs; bias rounded to fp16 like the kernel's dense table)
    t = table.clone().requires_grad_(True)
    x = qkv.float().view(nprob, L, 3, nheads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    bias = t[rel.long().view(-1)].view(L, L, nheads).permute(2, 0, 1)
    # the kernel's dense table holds fp16(bias / scale) and the sum is scaled afterwards
    s = (q @ k.transpose(-1, -2)) * scale + (bias / scale).half().float().detach() * scale + (bias - bias.detach())
    if shifted:
        s = s.view(nimg, nW, nheads, L, L) + mask.view(1, nW, 1, L, L)
        s = s.view(nprob, nheads, L, L)
    p = s.softmax(-1)
    ref = (p @ v).transpose(1, 2).reshape(nprob * L, C)
    assert (out.float() - ref).abs().max().item() < 4e-3
    lse_ref = torch.logsumexp(s, -1).permute(1, 0, 2).reshape(nheads, nprob * L)
    assert (lse - lse_ref).abs().max().item() < 2e-3

    dout = (torch.randn(nprob * L, C, device="cuda") * 0.5).half()
    ref.backward(dout.float())
    dq_acc = torch.zeros(nprob * L, C, device="cuda")
    dqkv = torch.zeros(nprob * L, 3 * C, device="cuda", dtype=torch.float16)
    ds = torch.zeros(nprob, nheads, 256, 256, device="cuda", dtype=torch.float16)
    ops.attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nheads,
                 nprob=nprob, L_tok=L, scale=scale, bias16=dense, prob_class=cls, ds16=ds)
    dtable = torch.zeros_like(table)
    ops.relpos_bias_grad(ds, rel, L, dtable)
    torch.cuda.synchronize()
    gx = x.grad.permute(1, 3, 0, 2, 4).reshape(nprob * L, 3 * C)  # [3,nprob,nh,L,hd] -> rows x (3, nh, hd)
    scl = gx.abs().max().item()
    assert (dq_acc - gx[:, :C]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, C:2 * C].float() - gx[:, C:2 * C]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, 2 * C:].float() - gx[:, 2 * C:]).abs().max().item() < 1e-2 * scl
    assert (dtable - t.grad).abs().max().item() < 2e-2 * t.grad.abs().max().item()


@pytest.mark.parametrize("nseq,L", [(3, 283), (2, 284), (2, 128), (1, 384), (2, 40), (2, 758), (1, 1000)])
def test_bert_attention_fwd_bwd(nseq, L):
    from lavender_b200 import ops
    g = torch.Generator().manual_seed(L)
    nheads, hd = 12, 64
    H = nheads * hd
    qkv = (torch.randn(nseq * L, 3 * H, generator=g) * 0.8).half().cuda()
    keep = torch.ones(nseq, L)
    keep[0, L - 7:] = 0  # padded tail of sequence 0
    NPk = max(384, (L + 127) // 128 * 128)   # 758 = configs[3]: 5 x (1 + 12 x 12) video tokens + 33 text tokens
    key_bias = torch.full((nseq, NPk), float("-inf"))
    key_bias[:, :L] = torch.where(keep > 0, 0.0, float("-inf"))
    key_bias = key_bias.cuda()
    out = torch.zeros(nseq * L, H, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nheads, nseq * L, device="cuda")
    scale = 1.0 / math.sqrt(hd)
    ops.attn_fwd(qkv, out, lse, q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nheads, nprob=nseq, L_tok=L,
                 scale=scale, key_bias=key_bias)
    x = qkv.float().view(nseq, L, 3, nheads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    ext = (1.0 - keep.cuda()) * torch.finfo(torch.float32).min
    s = (x[0] @ x[1].transpose(-1, -2)) * scale + ext[:, None, None, :]
    ref = (s.softmax(-1) @ x[2]).transpose(1, 2).reshape(nseq * L, H)
    assert (out.float() - ref).abs().max().item() < 4e-3

    dout = (torch.randn(nseq * L, H, device="cuda") * 0.5).half()
    ref.backward(dout.float())
    dq_acc = torch.zeros(nseq * L, H, device="cuda")
    dqkv = torch.zeros(nseq * L, 3 * H, device="cuda", dtype=torch.float16)
    ops.attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nheads,
                 nprob=nseq, L_tok=L, scale=scale, key_bias=key_bias)
    gx = x.grad.permute(1, 3, 0, 2, 4).reshape(nseq * L, 3 * H)
    scl = gx.abs().max().item()
    assert (dq_acc - gx[:, :H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, H:2 * H].float() - gx[:, H:2 * H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, 2 * H:].float() - gx[:, 2 * H:]).abs().max().item() < 1e-2 * scl


def test_long_window_attention_fwd_bwd():
    """swin_large_384_patch244_window81212 (configs[3]): windows of 5 x 12 x 12 = 720 tokens go through the blocked
    kernel with the dense relative-position bias added by the tensor core."""
    import lavender_oracle as O
    from lavender_b200 import ops
    g = torch.Generator().manual_seed(3)
    nprob, nheads, hd, L, NP = 2, 2, 32, 720, 768
    C = nheads * hd
    qkv = (torch.randn(nprob * L, 3 * C, generator=g) * 0.7).half().cuda()
    table = (torch.randn(15 * 23 * 23, nheads, generator=g) * 0.5).cuda()
    rel = O.relative_position_index((8, 12, 12))[:L, :L].contiguous().int().cuda()
    scale = hd ** -0.5
    dense = torch.zeros(1, nheads, NP, NP, device="cuda", dtype=torch.float16)
    ops.relpos_bias_expand(table, rel, L, None, dense, scale)
    out = torch.zeros(nprob * L, C, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nheads, nprob * L, device="cuda")
    ops.attn_fwd(qkv, out, lse, q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nheads, nprob=nprob, L_tok=L,
                 scale=scale, bias16=dense)
    t = table.clone().requires_grad_(True)
    x = qkv.float().view(nprob, L, 3, nheads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    bias = t[rel.long().view(-1)].view(L, L, nheads).permute(2, 0, 1)
    s = (x[0] @ x[1].transpose(-1, -2)) * scale + (bias / scale).half().float().detach() * scale + (bias - bias.detach())
    ref = (s.softmax(-1) @ x[2]).transpose(1, 2).reshape(nprob * L, C)
    assert (out.float() - ref).abs().max().item() < 4e-3
    lse_ref = torch.logsumexp(s, -1).permute(1, 0, 2).reshape(nheads, nprob * L)
    assert (lse - lse_ref).abs().max().item() < 2e-3
    dout = (torch.randn(nprob * L, C, device="cuda") * 0.5).half()
    ref.backward(dout.float())
    dq_acc = torch.zeros(nprob * L, C, device="cuda")
    dqkv = torch.zeros(nprob * L, 3 * C, device="cuda", dtype=torch.float16)
    ds = torch.zeros(nprob, nheads, NP, NP, device="cuda", dtype=torch.float16)
    ops.attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nheads,
                 nprob=nprob, L_tok=L, scale=scale, bias16=dense, ds16=ds)
    dtable = torch.zeros_like(table)
    ops.relpos_bias_grad(ds, rel, L, dtable)
    gx = x.grad.permute(1, 3, 0, 2, 4).reshape(nprob * L, 3 * C)
    scl = gx.abs().max().item()
    assert (dq_acc - gx[:, :C]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, C:2 * C].float() - gx[:, C:2 * C]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, 2 * C:].float() - gx[:, 2 * C:]).abs().max().item() < 1e-2 * scl
    assert (dtable - t.grad).abs().max().item() < 2e-2 * t.grad.abs().max().item()


@pytest.mark.parametrize("Lfull,Lt", [(250, 34), (255, 80), (120, 8)])
def test_bert_attention_seq2seq_mask(Lfull, Lt):
    """The captioning mask of LAVENDER_Base.get_attn_mask (model.py:208-218) passed as (key mask, causal_from): every
    query sees the kept video / prefix keys, text queries see text keys causally — vs the materialised [B, L, L] mask."""
    from lavender_b200 import ops
    L = Lfull + Lt
    nseq, nheads, hd = 2, 4, 64
    H = nheads * hd
    g = torch.Generator().manual_seed(L)
    qkv = (torch.randn(nseq * L, 3 * H, generator=g) * 0.8).half().cuda()
    full = torch.ones(nseq, Lfull)
    full[1, 7:19] = 0                                     # masked video keys (vt_mask) in sequence 1
    NPk = (L + 127) // 128 * 128
    key_bias = torch.full((nseq, NPk), float("-inf"))
    key_bias[:, :L] = 0.0
    key_bias[:, :Lfull] = torch.where(full > 0, 0.0, float("-inf"))
    key_bias = key_bias.cuda()
    m3 = torch.zeros(nseq, L, L)                          # model.py:208-218
    m3[:, :, :Lfull] = full.unsqueeze(1)
    m3[:, Lfull:, Lfull:] = torch.tril(torch.ones(Lt, Lt))
    ext = ((1.0 - m3) * torch.finfo(torch.float32).min).cuda()
    out = torch.zeros(nseq * L, H, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nheads, nseq * L, device="cuda")
    scale = 1.0 / math.sqrt(hd)
    kw = dict(q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nheads, nprob=nseq, L_tok=L, scale=scale,
              key_bias=key_bias, causal_from=Lfull)
    ops.attn_fwd(qkv, out, lse, **kw)
    x = qkv.float().view(nseq, L, 3, nheads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    s = (x[0] @ x[1].transpose(-1, -2)) * scale + ext[:, None]
    ref = (s.softmax(-1) @ x[2]).transpose(1, 2).reshape(nseq * L, H)
    assert (out.float() - ref).abs().max().item() < 4e-3
    dout = (torch.randn(nseq * L, H, device="cuda") * 0.5).half()
    ref.backward(dout.float())
    dq_acc = torch.zeros(nseq * L, H, device="cuda")
    dqkv = torch.zeros(nseq * L, 3 * H, device="cuda", dtype=torch.float16)
    ops.attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, **kw)
    gx = x.grad.permute(1, 3, 0, 2, 4).reshape(nseq * L, 3 * H)
    scl = gx.abs().max().item()
    assert (dq_acc - gx[:, :H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, H:2 * H].float() - gx[:, H:2 * H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, 2 * H:].float() - gx[:, 2 * H:]).abs().max().item() < 1e-2 * scl
