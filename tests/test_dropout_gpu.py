"""Dropout of the native BERT path (HF hidden / attention-probability dropout in train() mode, model.py:242 as run by
Agent_Pretrain_MLM.step, main_pretrain_mlm.py:147).  A torch RNG stream cannot be reproduced bit for bit, so parity is
stated on what dropout IS: with the mask the kernels use (read back through lav_dropout_mask), every fused site must
equal the fp32 torch formula `x * mask / (1 - p)`, forward and backward, and the mask must be Bernoulli(1 - p),
independent across sites / steps / heads and reproducible."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

P = 0.1
KEEP = 1.0 - round(P * 65536) / 65536.0


def _rng(seed=1234, step=0):
    return torch.tensor([seed, step], dtype=torch.int64, device="cuda")


def test_mask_statistics_and_streams():
    from lavender_b200 import ops
    rng = _rng()
    m0 = ops.dropout_mask(4096, 768, (rng, 7, P)).float()
    assert abs(m0.mean().item() - KEEP) < 2e-3
    assert abs(m0.mean(0).min().item() - KEEP) < 3e-2 and abs(m0.mean(1).min().item() - KEEP) < 6e-2
    assert torch.equal(m0, ops.dropout_mask(4096, 768, (rng, 7, P)).float())          # reproducible
    for other in (ops.dropout_mask(4096, 768, (rng, 8, P)),                              # another site
                  ops.dropout_mask(4096, 768, (_rng(step=1), 7, P)),                     # next step
                  ops.dropout_mask(4096, 768, (_rng(seed=99), 7, P)),                    # another seed
                  ops.dropout_mask(4096, 768, (rng, 7, P), head=3)):                     # attention index space, head 3
        agree = (other.float() == m0).float().mean().item()
        assert abs(agree - (KEEP * KEEP + (1 - KEEP) ** 2)) < 3e-3, agree              # independent masks
    # neighbouring elements are uncorrelated
    x = m0 - m0.mean()
    assert abs((x[:, 1:] * x[:, :-1]).mean().item()) < 1e-3 and abs((x[1:] * x[:-1]).mean().item()) < 1e-3


def test_elementwise_dropout_matches_mask():
    from lavender_b200 import ops
    rng = _rng()
    x = torch.randn(264, 768, device="cuda")
    y = ops.dropout_f32(x, torch.empty_like(x), (rng, 3, P))
    m = ops.dropout_mask(264, 768, (rng, 3, P)).float()
    assert torch.allclose(y, x * m / KEEP, rtol=1e-6, atol=0)


def test_gemm_epilogue_dropout_and_ln_bwd_mask():
    """BertSelfOutput / BertOutput: LN(residual + dropout(dense(x))) — the dropout is applied in the GEMM epilogue and
    its mask is re-applied to the fp16 gradient that LayerNorm backward hands to the dense layer's dgrad / wgrad."""
    from lavender_b200 import ops
    M, N, K = 1132, 768, 256
    g = torch.Generator().manual_seed(0)
    a = (torch.randn(M, K, generator=g) * 0.3).half().cuda()
    w = (torch.randn(N, K, generator=g) * 0.3).half().cuda()
    bias = torch.randn(N, device="cuda") * 0.1
    res = torch.randn(M, N, device="cuda")
    rng = _rng(5)
    spec = (rng, 11, P)
    out = torch.zeros(M, N, device="cuda")
    ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=res, drop=spec)
    m = ops.dropout_mask(M, N, spec).double()
    ref = res.double() + (a.double() @ w.double().t() + bias.double()) * m / KEEP
    assert (out.double() - ref).abs().max().item() < 1e-3
    # LayerNorm backward: dx32 unmasked, dx16 = mask * dx / keep
    gamma, x = torch.rand(N, device="cuda") + 0.5, out
    mean = x.mean(1)
    rstd = (x.var(1, unbiased=False) + 1e-12).rsqrt()
    dy = torch.randn(M, N, device="cuda")
    dx32, dx16 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda", dtype=torch.float16)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=M, C=N, dx32=dx32, dx16=dx16, drop16=spec)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (N,), gamma, None, 1e-12).backward(dy)
    assert (dx32 - xr.grad).abs().max().item() < 1e-4 * max(1.0, xr.grad.abs().max().item())
    want = (dx32.double() * m / KEEP).half()
    assert (dx16.float() - want.float()).abs().max().item() <= 2e-3 * max(1.0, want.float().abs().max().item())
    assert torch.equal(dx16 == 0, (m == 0) | (want == 0))


@pytest.mark.parametrize("nseq,L", [(2, 283), (1, 384), (2, 40)])
def test_bert_attention_dropout_fwd_bwd(nseq, L):
    from lavender_b200 import ops
    g = torch.Generator().manual_seed(L)
    nheads, hd = 4, 64
    H = nheads * hd
    qkv = (torch.randn(nseq * L, 3 * H, generator=g) * 0.8).half().cuda()
    key_bias = torch.full((nseq, 384), float("-inf"))
    key_bias[:, :L] = 0.0
    key_bias[0, L - 5:L] = float("-inf")
    key_bias = key_bias.cuda()
    spec = (_rng(77, 3), 21, P)
    out = torch.zeros(nseq * L, H, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nheads, nseq * L, device="cuda")
    scale = 1.0 / math.sqrt(hd)
    ops.attn_fwd(qkv, out, lse, q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nheads, nprob=nseq, L_tok=L,
                 scale=scale, key_bias=key_bias, drop=spec)
    # mask[h][global query row][key column]
    mask = torch.stack([ops.dropout_mask(nseq * L, 384, spec, head=h) for h in range(nheads)]).float()
    mask = mask.view(nheads, nseq, L, 384)[..., :L].permute(1, 0, 2, 3)                 # [nseq, nh, L, L]
    x = qkv.float().view(nseq, L, 3, nheads, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    s = (x[0] @ x[1].transpose(-1, -2)) * scale + key_bias[:, None, None, :L]
    p = s.softmax(-1)
    ref = ((p * mask / KEEP) @ x[2]).transpose(1, 2).reshape(nseq * L, H)
    assert (out.float() - ref).abs().max().item() < 5e-3
    lse_ref = torch.logsumexp(s, -1).permute(1, 0, 2).reshape(nheads, nseq * L)
    assert (lse - lse_ref).abs().max().item() < 2e-3
    frac = (mask.mean().item())
    assert abs(frac - KEEP) < 1e-2

    dout = (torch.randn(nseq * L, H, device="cuda") * 0.5).half()
    ref.backward(dout.float())
    dq_acc = torch.zeros(nseq * L, H, device="cuda")
    dqkv = torch.zeros(nseq * L, 3 * H, device="cuda", dtype=torch.float16)
    ops.attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nheads,
                 nprob=nseq, L_tok=L, scale=scale, key_bias=key_bias, drop=spec)
    gx = x.grad.permute(1, 3, 0, 2, 4).reshape(nseq * L, 3 * H)
    scl = gx.abs().max().item()
    assert (dq_acc - gx[:, :H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, H:2 * H].float() - gx[:, H:2 * H]).abs().max().item() < 1e-2 * scl
    assert (dqkv[:, 2 * H:].float() - gx[:, 2 * H:]).abs().max().item() < 1e-2 * scl


def test_bert_encoder_train_mode_dropout_gradients():
    """End to end through BertEmbeddings + BertEncoder in train() mode: finite, non-identity, reproducible for a fixed
    (seed, step, site sequence), and E[output] over masks approaches the eval output (inverted-dropout scaling)."""
    from lavender_b200.bert import BertConfig, BertEncoder
    from lavender_b200 import dropout as DR
    torch.manual_seed(0)
    cfg = BertConfig(num_hidden_layers=2)
    enc = BertEncoder(cfg).cuda()
    x = torch.randn(2, 72, 768, device="cuda")
    mask = torch.ones(2, 72, device="cuda")
    enc.eval()
    y_eval = enc(x, mask)["last_hidden_state"]
    enc.train()
    st = DR.reseed("cuda", 42)
    st._site = 0
    xr = x.clone().requires_grad_(True)
    y1 = enc(xr, mask)["last_hidden_state"]
    y1.square().mean().backward()
    g1 = xr.grad.clone()
    w = enc.layer[0].attention.output.dense.weight
    gw1 = w.grad.clone()
    assert torch.isfinite(y1).all() and torch.isfinite(g1).all() and torch.isfinite(gw1).all()
    assert (y1 - y_eval).abs().max().item() > 1e-2            # dropout is really on
    DR.reseed("cuda", 42)._site = 0                           # same seed / step / site ids -> same masks
    w.grad = None
    for p_ in enc.parameters():
        p_.grad = None
    xr2 = x.clone().requires_grad_(True)
    y2 = enc(xr2, mask)["last_hidden_state"]
    y2.square().mean().backward()
    assert torch.equal(y1, y2) and torch.allclose(xr2.grad, g1, rtol=1e-4, atol=1e-6)
    y3 = enc(x, mask)["last_hidden_state"]                    # fresh site ids -> fresh masks
    assert not torch.equal(y3, y1)
