"""Row kernels (LayerNorm fwd/bwd with gather maps, casts, column sums) vs plain PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,C,G", [(1000, 96, 1), (4096, 128, 1), (777, 768, 1), (980, 128, 4), (245, 512, 4),
                                      (300, 1024, 1), (7840, 512, 1), (11360, 768, 1), (5000, 1024, 1)])
def test_layernorm_fwd_bwd(rows, C, G):
    from lavender_b200 import ops
    torch.manual_seed(0)
    W = G * C
    nsrc = rows * G
    x = torch.randn(nsrc, C, device="cuda") * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(W, device="cuda")
    beta = 0.1 * torch.randn(W, device="cuda")
    row_map = torch.randperm(nsrc, device="cuda").to(torch.int32)
    y16 = torch.zeros(rows, W, device="cuda", dtype=torch.float16)
    y32 = torch.zeros(rows, W, device="cuda")
    mean = torch.zeros(rows, device="cuda")
    rstd = torch.zeros(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-5, rows=rows, C=C, G=G, row_map=row_map, out16=y16, out32=y32, mean=mean,
                      rstd=rstd)
    xg = x.clone().requires_grad_(True)
    gref, bref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    gathered = xg[row_map.long()].view(rows, W)
    yref = F.layer_norm(gathered, (W,), gref, bref, 1e-5)
    assert (y32 - yref).abs().max().item() < 2e-5
    assert (y16.float() - yref).abs().max().item() < 5e-3

    dy = torch.randn(rows, W, device="cuda")
    add = torch.randn(nsrc, C, device="cuda")
    yref.backward(dy)
    for dyt, tol in ((dy, 2e-4), (dy.half(), 2e-2)):
        dx32 = torch.zeros(nsrc, C, device="cuda")
        dgamma = torch.zeros(W, device="cuda")
        dbeta = torch.zeros(W, device="cuda")
        dx16 = torch.zeros(rows, W, device="cuda", dtype=torch.float16) if G == 1 else None
        ops.layernorm_bwd(dyt, x, gamma, mean, rstd, rows=rows, C=C, G=G, row_map=row_map, add32=add, dx32=dx32,
                          dx16=dx16, dgamma=dgamma, dbeta=dbeta)
        assert (dx32 - (xg.grad + add)).abs().max().item() < tol
        assert (dgamma - gref.grad).abs().max().item() < tol * rows ** 0.5 * 4
        assert (dbeta - bref.grad).abs().max().item() < tol * rows ** 0.5 * 4
        if dx16 is not None:
            assert (dx16.float() - (xg.grad + add)[row_map.long()]).abs().max().item() < 2e-2


def test_layernorm_bwd_wide_rows_in_place_and_without_add():
    """The shared-memory staged kernel (C >= 512) in the two other forms the path uses: in place (add32 is dx32, Swin
    norm1) and without a residual gradient (BERT), over enough rows for every warp's ring to wrap several times."""
    from lavender_b200 import ops
    torch.manual_seed(1)
    rows, C = 9000, 768
    x = torch.randn(rows, C, device="cuda") * 1.5 - 0.3
    gamma = 1 + 0.1 * torch.randn(C, device="cuda")
    beta = torch.zeros(C, device="cuda")
    y32 = torch.zeros(rows, C, device="cuda")
    mean, rstd = torch.zeros(rows, device="cuda"), torch.zeros(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-12, rows=rows, C=C, out32=y32, mean=mean, rstd=rstd)
    xg, gref = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True)
    dy = torch.randn(rows, C, device="cuda")
    F.layer_norm(xg, (C,), gref, beta, 1e-12).backward(dy)
    # BERT form: fp32 dy, no add, dx32 + dx16
    dx32, dx16 = torch.zeros(rows, C, device="cuda"), torch.zeros(rows, C, device="cuda", dtype=torch.float16)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=rows, C=C, dx32=dx32, dx16=dx16, dgamma=dg, dbeta=db)
    assert (dx32 - xg.grad).abs().max().item() < 2e-4
    assert (dx16.float() - xg.grad).abs().max().item() < 2e-2
    assert (dg - gref.grad).abs().max().item() < 0.1 and (db - dy.sum(0)).abs().max().item() < 0.1
    # Swin norm1 form: fp16 dy, row map, add32 and dx32 the same buffer
    rmap = torch.randperm(rows, device="cuda").to(torch.int32)
    g1 = torch.randn(rows, C, device="cuda")
    want = g1.clone()
    xs = torch.empty_like(x)
    xs[rmap.long()] = x   # x stored in source order: output row r reads xs[rmap[r]] = x[r]
    want[rmap.long()] += xg.grad
    dg.zero_(), db.zero_()
    ops.layernorm_bwd(dy.half(), xs, gamma, mean, rstd, rows=rows, C=C, row_map=rmap, add32=g1, dx32=g1, dgamma=dg, dbeta=db)
    assert (g1 - want).abs().max().item() < 2e-2


@pytest.mark.parametrize("rows,C,rps", [(245 * 8, 128, 245), (7840, 512, 980), (1960, 1024, 245)])
def test_layernorm_bwd_fused_gradient_casts(rows, C, rps):
    """lav_layernorm_bwd_ex: the fp16 operand of the consumer (window-gathered / DropPath-scaled) written by the LayerNorm
    backward itself equals the separate scale_cast kernel applied to its fp32 output."""
    from lavender_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(rows, C, device="cuda")
    gamma = 1 + 0.1 * torch.randn(C, device="cuda")
    mean, rstd = x.mean(1).contiguous(), (x.var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    dy = torch.randn(rows, C, device="cuda").half()
    add = torch.randn(rows, C, device="cuda")
    # window partition permutes rows inside a sample only (the scale is per sample on both sides of the map)
    rmap = torch.cat([torch.randperm(rps, device="cuda") + i * rps for i in range(rows // rps)]).to(torch.int32)
    rinv = torch.empty_like(rmap)
    rinv[rmap.long()] = torch.arange(rows, device="cuda", dtype=torch.int32)
    keep = (torch.rand(rows // rps, device="cuda") > 0.3).float() / 0.7
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    # (1) norm2 form: token-order rows, fp16 output scattered to window order, scaled per sample
    dx32 = torch.zeros(rows, C, device="cuda")
    fused = torch.zeros(rows, C, device="cuda", dtype=torch.float16)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=rows, C=C, add32=add, dx32=dx32, dx16=fused, dx16_map=rinv,
                      dx16_scale=keep, dx16_rps=rps, dgamma=dg, dbeta=db)
    want = ops.scale_cast(dx32, torch.zeros_like(fused), rows=rows, C=C, row_map=rmap, row_scale=keep, rows_per_scale=rps)
    assert (fused.float() - want.float()).abs().max().item() <= 1e-3 * want.float().abs().max().item()
    assert (fused != want).float().mean().item() < 1e-3
    # (2) norm1 form: window-order rows scattered to token order (in place on add32), fp16 output at the token rows
    g1 = add.clone()
    fused2 = torch.zeros(rows, C, device="cuda", dtype=torch.float16)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=rows, C=C, row_map=rmap, add32=g1, dx32=g1, dx16=fused2,
                      dx16_at_src=True, dx16_scale=keep, dx16_rps=rps, dgamma=dg, dbeta=db)
    want2 = ops.scale_cast(g1, torch.zeros_like(fused2), rows=rows, C=C, row_scale=keep, rows_per_scale=rps)
    assert (fused2 != want2).float().mean().item() < 1e-3
    # (3) plain (PatchMerging consumer): no scale
    fused3 = torch.zeros(rows, C, device="cuda", dtype=torch.float16)
    g2 = add.clone()
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=rows, C=C, row_map=rmap, add32=g2, dx32=g2, dx16=fused3,
                      dx16_at_src=True, dgamma=dg, dbeta=db)
    assert (fused3 != g2.half()).float().mean().item() < 1e-3


def test_layernorm_identity_no_map_eps12():
    from lavender_b200 import ops
    x = torch.randn(333, 768, device="cuda")
    g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
    y = torch.zeros_like(x)
    ops.layernorm_fwd(x, g, b, 1e-12, rows=333, C=768, out32=y)
    assert (y - F.layer_norm(x, (768,), g, b, 1e-12)).abs().max().item() < 3e-5


def test_scale_cast_and_colsum_and_flatcast():
    from lavender_b200 import ops
    rows, C = 980, 96
    x = torch.randn(rows, C, device="cuda")
    perm = torch.randperm(rows, device="cuda").to(torch.int32)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25], device="cuda")
    out = torch.zeros(rows, 104, device="cuda", dtype=torch.float16)
    ops.scale_cast(x, out, rows=rows, C=C, row_map=perm, row_scale=scale, rows_per_scale=245, alpha=2.0)
    ref = (x[perm.long()] * scale.repeat_interleave(245)[:, None] * 2.0).half()
    assert torch.equal(out[:, :C], ref)
    acc = torch.ones(C, device="cuda")
    ops.colsum(out, acc, rows=rows, N=C, alpha=0.5)
    assert (acc - (1 + 0.5 * ref.float().sum(0))).abs().max().item() < 1e-2
    src = torch.randn(1_000_003, device="cuda")
    dst = torch.zeros(1_000_003, device="cuda", dtype=torch.float16)
    ops.cast_f16(src, dst)
    assert torch.equal(dst, src.half())
