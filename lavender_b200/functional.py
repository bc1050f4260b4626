"""autograd.Function wrappers for the stand-alone native ops (Linear, LayerNorm) and shared helpers for the
fused module functions in video_swin.py / bert.py.

Gradient convention: parameter gradients are accumulated by the kernels directly into the arena's flat fp32
gradient buffer (`p.grad` is a view of it); the Functions return None for parameter inputs.
"""
import torch

from . import ops
from . import _lib as L
from . import precision
from . import streams
from .arena import arena_of

F16, F32 = torch.float16, torch.float32


def empty16(*shape, device):
    return torch.empty(shape, dtype=F16, device=device)


def empty32(*shape, device):
    return torch.empty(shape, dtype=F32, device=device)


def require_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what}: lavender_b200 has no CPU path - tensors must live on a CUDA (sm_100a) device")


def linear_fwd(x16, w16, bias, out, **kw):
    """out = x16 @ w16.T (+ bias) with a fused epilogue; x16 [M,K], w16 [N,K]."""
    M, K = x16.shape
    N = w16.shape[0]
    return ops.gemm(x16, w16, out, M=M, N=N, K=K, bias=bias, **kw)


def linear_fwd_hp(ar, x32, w32, bias, out, **kw):
    """High-precision forward (precision.py): out = x32 @ w32.T (+ bias) as ONE tcgen05 GEMM over the split-fp16
    operands [xh | xl | xh] . [wh | wh | wl]^T (K' = 3K); x32 [M,K] fp32, w32 [N,K] fp32 master weight view."""
    M, K = x32.shape
    N = w32.shape[0]
    xs = ops.split3(x32, empty16(M, 3 * K, device=x32.device), rows=M, C=K)
    return ops.gemm(xs, ar.w16x3(w32), out, M=M, N=N, K=3 * K, bias=bias, **kw)


def cast16(x32):
    """fp16 copy of an fp32 activation (what the default mode would have stored; backward operand in high-precision mode)."""
    M, C = x32.shape
    return ops.scale_cast(x32, empty16(M, C, device=x32.device), rows=M, C=C)


def linear_dgrad(dy16, w16, out, **kw):
    """out[M,K] = dy16[M,N] @ w16[N,K]   (w16 consumed MN-major, no transposed copy)."""
    M, N = dy16.shape
    K = w16.shape[1]
    return ops.gemm(dy16, w16, out, M=M, N=K, K=N, b_major=L.MAJOR_MN, **kw)


def linear_wgrad(dy16, x16, gw, gb=None, n_valid=None):
    """gw[N,K] += dy16[T,N].T @ x16[T,K] ; gb[N] += colsum(dy16).  (token-major operands, split-K atomics)"""
    T = dy16.shape[0]
    N = n_valid if n_valid is not None else dy16.shape[1]
    K = x16.shape[1]
    # the bias gradient (column sum of dY) rides on the wgrad GEMM: one extra N=16 MMA per k-step against a ones tile
    if streams.enabled():
        # off the critical path: launched on the side stream, joined when the gradient arena is finalised (streams.py)
        dev = dy16.device
        with torch.cuda.stream(streams.fork(dev)):
            ops.gemm(dy16, x16, gw, M=N, N=K, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, accumulate=True, bias_grad=gb)
        streams.hold(dev, dy16, x16)
        return
    ops.gemm(dy16, x16, gw, M=N, N=K, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, accumulate=True, bias_grad=gb)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b for an nn.Linear container (fp32 in / fp32 out, fp16 tensor-core operands)."""

    @staticmethod
    def forward(ctx, x, mod, weight, bias):
        require_cuda(x, "LinearFn")
        ar = arena_of(mod)
        ar.refresh16()
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M, K = x2.shape
        x16 = ops.scale_cast(x2, empty16(M, K, device=x.device), rows=M, C=K)
        y = empty32(M, weight.shape[0], device=x.device)
        if precision.high("fc"):
            linear_fwd_hp(ar, x2.float(), weight.data, bias, y)
        else:
            linear_fwd(x16, ar.w16(weight), bias, y)
        ctx.mod, ctx.weight, ctx.bias, ctx.x16, ctx.shp = mod, weight, bias, x16, shp
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        ar = arena_of(ctx.mod)
        w, b = ctx.weight, ctx.bias
        ar.prepare_grads([w] + ([b] if b is not None else []))
        g2 = gy.reshape(-1, gy.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        M, N = g2.shape
        g16 = ops.scale_cast(g2, empty16(M, N, device=gy.device), rows=M, C=N)
        linear_wgrad(g16, ctx.x16, ar.g(w), ar.g(b) if b is not None else None)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = empty32(M, w.shape[1], device=gy.device)
            linear_dgrad(g16, ar.w16(w), gx)
            gx = gx.view(ctx.shp)
        return gx, None, None, None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim, fp32 in/out (model.py:85; HF BertEmbeddings.LayerNorm)."""

    @staticmethod
    def forward(ctx, x, mod, weight, bias, eps):
        require_cuda(x, "LayerNormFn")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M, C = x2.shape
        y = empty32(M, C, device=x.device)
        mean, rstd = empty32(M, device=x.device), empty32(M, device=x.device)
        ops.layernorm_fwd(x2, weight, bias, eps, rows=M, C=C, out32=y, mean=mean, rstd=rstd)
        ctx.mod, ctx.weight, ctx.bias, ctx.saved, ctx.shp = mod, weight, bias, (x2, mean, rstd), shp
        return y.view(shp)

    @staticmethod
    def backward(ctx, gy):
        ar = arena_of(ctx.mod)
        ar.prepare_grads([ctx.weight, ctx.bias])
        x2, mean, rstd = ctx.saved
        g2 = gy.reshape(-1, gy.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        M, C = x2.shape
        gx = empty32(M, C, device=gy.device)
        ops.layernorm_bwd(g2, x2, ctx.weight, mean, rstd, rows=M, C=C, dx32=gx, dgamma=ar.g(ctx.weight),
                          dbeta=ar.g(ctx.bias))
        return gx.view(ctx.shp), None, None, None, None


def native_linear(mod, x):
    return LinearFn.apply(x, mod, mod.weight, mod.bias)


def native_layernorm(mod, x):
    return LayerNormFn.apply(x, mod, mod.weight, mod.bias, mod.eps)
