"""tcgen05 GEMM (lav_gemm_f16) vs fp64 matmul of the same fp16 operands.  Replaces the nn.Linear call sites
video_swin.py:74,77,147,168,285,398 / model.py:49 / HF BERT linears, forward + dgrad + wgrad."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).cuda()


def _ref(a, b):  # a:[M,K], b:[N,K] fp16 -> fp64
    return a.double() @ b.double().t()


def _tol(K):
    return 2e-3 * math.sqrt(K) * 1.0  # fp32 accumulation of K products of O(1) numbers; generous


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (128, 64, 256), (1000, 384, 128), (1960, 768, 1024),
                                   (4096, 512, 96), (130, 72, 200), (2264, 2304, 768), (777, 256, 3072),
                                   (7840, 2048, 512), (20000, 640, 192), (257, 130, 64)])
def test_forward_plain(M, N, K):
    from lavender_b200 import ops
    a, b = _mk((M, K), 1), _mk((N, K), 2)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out, M=M, N=N, K=K)
    torch.cuda.synchronize()
    err = (out.double() - _ref(a, b)).abs().max().item()
    assert err < 1e-4 * K, err


def test_forward_f16_out_bias_alpha():
    from lavender_b200 import ops
    M, N, K = 900, 384, 192
    a, b = _mk((M, K), 3, 0.5), _mk((N, K), 4, 0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, alpha=0.25)
    ref = 0.25 * _ref(a, b) + bias.double()
    assert (out.double() - ref).abs().max().item() < 2e-2  # fp16 output rounding of O(10) values
    assert (out.double() - ref.half().double()).abs().max().item() < 2e-2


def test_forward_unaligned_vocab_ld():
    """MLM decoder shape (main_pretrain_mlm.py:69): N=30522 fp32 logits, ld not a multiple of 4."""
    from lavender_b200 import ops
    M, N, K = 300, 30522, 768
    a, b = _mk((M, K), 5, 0.3), _mk((N, K), 6, 0.3)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
    ref = _ref(a, b) + bias.double()
    assert (out.double() - ref).abs().max().item() < 1e-3


def test_gelu_with_aux_and_backward():
    from lavender_b200 import ops, _lib as L
    M, N, K = 640, 512, 128
    a, b = _mk((M, K), 7, 0.3), _mk((N, K), 8, 0.3)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    aux = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, act=L.ACT_GELU, aux=aux)
    pre = (_ref(a, b) + bias.double()).requires_grad_(True)
    act = torch.nn.functional.gelu(pre)
    act.sum().backward()
    assert (aux.double() - pre.grad).abs().max().item() < 2e-3      # aux = gelu'(pre-activation), saved for the backward
    assert (out.double() - act.detach()).abs().max().item() < 5e-3
    # dgrad with GELU': dA = (dY @ W) * aux   [A=dY K-major, B=W MN-major]
    dy = _mk((M, K), 9, 0.3)          # pretend grad wrt fc2 input of width K... here: [M,K] x W2[K?]
    w = _mk((K, N), 10, 0.3)          # contraction over K: out[M,N] = dy[M,K] @ w[K,N]
    dout = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(dy, w, dout, M=M, N=N, K=K, b_major=L.MAJOR_MN, act=L.ACT_GELU_BWD, aux=aux)
    ref = (dy.double() @ w.double()) * pre.grad
    assert (dout.double() - ref).abs().max().item() < 1e-2


def test_residual_rowmap_rowscale():
    """proj epilogue of SwinTransformerBlock3D (video_swin.py:231-242,254): scatter rows + DropPath + residual."""
    from lavender_b200 import ops
    M, N, K = 980, 96, 96
    a, b = _mk((M, K), 11, 0.3), _mk((N, K), 12, 0.3)
    bias = torch.randn(N, device="cuda") * 0.1
    perm = torch.randperm(M, device="cuda").to(torch.int32)
    res = torch.randn(M, N, device="cuda")
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25], device="cuda")  # 4 samples of 245 rows
    out = torch.zeros(M, N, device="cuda")
    ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=res, row_map=perm, row_scale=scale, rows_per_scale=245)
    val = (_ref(a, b) + bias.double()) * scale.double().repeat_interleave(245)[:, None]
    ref = res.double().clone()
    ref[perm.long()] += val
    assert (out.double() - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(512, 256, 128), (1000, 384, 128), (1352, 768, 30528), (300, 96, 384),
                                   (7840, 512, 2048), (31000, 128, 512)])
def test_dgrad_mn_major_b(M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N] with W stored [K(out features), N(in features)] row-major (MN-major B)."""
    from lavender_b200 import ops, _lib as L
    dy, w = _mk((M, K), 13, 0.3), _mk((K, N), 14, 0.3)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(dy, w, out, M=M, N=N, K=K, b_major=L.MAJOR_MN)
    ref = dy.double() @ w.double()
    assert (out.double() - ref).abs().max().item() < 2e-5 * K + 1e-3


@pytest.mark.parametrize("T,N,K,split", [(1024, 256, 128, 1), (1960, 384, 128, 0), (125440 // 8, 384, 128, 0),
                                         (1352, 768, 3072, 0), (1000, 96, 288, 4), (1352, 30522, 768, 0),
                                         (7840, 2048, 512, 0), (9088, 768, 3072, 0), (7840, 512, 512, 0)])
def test_wgrad_mn_mn_accumulate(T, N, K, split):
    """dW[N,K] += dY[T,N]^T @ X[T,K]: both operands token-major (MN-major), fp32 accumulate, split-K atomics."""
    from lavender_b200 import ops, _lib as L
    ldn = (N + 7) // 8 * 8
    dy_full = _mk((T, ldn), 15, 0.3)
    dy = dy_full[:, :N]
    x = _mk((T, K), 16, 0.3)
    init = torch.randn(N, K, device="cuda")
    out = init.clone()
    binit = torch.randn(N, device="cuda")
    gb = binit.clone()   # the nn.Linear bias gradient (column sum of dY) rides on the same kernel
    ops.gemm(dy, x, out, M=N, N=K, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, accumulate=True, split_k=split,
             bias_grad=gb)
    ref = init.double() + dy.double().t() @ x.double()
    assert (out.double() - ref).abs().max().item() < 2e-5 * T + 1e-3
    bref = binit.double() + dy.double().sum(0)
    assert (gb.double() - bref).abs().max().item() < 2e-5 * T + 1e-3


def test_launch_count_increases():
    from lavender_b200 import ops, _lib as L
    n0 = L.launch_count()
    a, b = _mk((128, 64), 1), _mk((128, 64), 2)
    out = torch.zeros(128, 128, device="cuda")
    ops.gemm(a, b, out, M=128, N=128, K=64)
    assert L.launch_count() == n0 + 1


def test_split_fp16_gemm_restores_fp32_operand_precision():
    """High-precision mode building block: lav_split3_f16 + ONE lav_gemm_f16 over K' = 3K computes xh*wh + xl*wh + xh*wl;
    the result must agree with the fp64 product of the fp32 operands ~1000x better than the plain fp16-operand GEMM."""
    from lavender_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, N, K = 520, 392, 768
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = (x.double() @ w.double().t() + bias.double())
    xs = ops.split3(x, torch.empty(M, 3 * K, device="cuda", dtype=torch.float16), rows=M, C=K)
    ws = ops.split3(w, torch.empty(N, 3 * K, device="cuda", dtype=torch.float16), rows=N, C=K, weight=True)
    assert torch.equal(xs[:, :K], x.half()) and torch.equal(xs[:, 2 * K:], x.half()) and torch.equal(ws[:, K:2 * K], w.half())
    assert torch.equal(xs[:, K:2 * K], (x - x.half().float()).half()) and torch.equal(ws[:, 2 * K:], (w - w.half().float()).half())
    out = torch.empty(M, N, device="cuda")
    ops.gemm(xs, ws, out, M=M, N=N, K=3 * K, bias=bias)
    plain = torch.empty(M, N, device="cuda")
    ops.gemm(x.half(), w.half(), plain, M=M, N=N, K=K, bias=bias)
    torch.cuda.synchronize()
    e_hp = (out.double() - ref).abs().max().item()
    e_16 = (plain.double() - ref).abs().max().item()
    print(f"split-fp16 GEMM max abs err {e_hp:.2e} vs plain fp16 operands {e_16:.2e} (|ref| max {ref.abs().max().item():.1f})")
    assert e_hp < 1e-4 and e_hp * 50 < e_16   # fp32 accumulation over K' = 2304 terms, |ref| up to ~9
