"""Torch-tensor front ends of the C-ABI kernels (no autograd here; see functional.py).

Tensors are passed as raw device pointers + sizes; all launches go to torch's current CUDA stream.
"""
import ctypes

import torch

from . import _lib as L

F16 = torch.float16


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk16(t, name):
    assert t.is_cuda and t.dtype == F16 and t.stride(-1) == 1, f"{name}: need a row-major CUDA fp16 tensor"


def gemm(a, b, out, *, M, N, K, a_major=L.MAJOR_K, b_major=L.MAJOR_K, bias=None, act=L.ACT_NONE, aux=None,
         residual=None, row_map=None, row_scale=None, rows_per_scale=1, alpha=1.0, accumulate=False, split_k=0):
    """out[M,N] = epilogue(alpha * A·Bᵀ).  a/b are 2-D fp16 tensors stored per `*_major`
    (K-major: [rows, K]; MN-major: [K, rows]); `out` is fp16 or fp32 2-D."""
    _chk16(a, "a")
    _chk16(b, "b")
    assert out.is_cuda and out.stride(-1) == 1 and out.dtype in (F16, torch.float32)
    e = L.GemmEpilogue()
    e.out, e.ldo = out.data_ptr(), out.stride(0)
    e.out_dtype = L.OUT_F16 if out.dtype == F16 else L.OUT_F32
    e.act = act
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
        e.bias = bias.data_ptr()
    if aux is not None:
        _chk16(aux, "aux")
        e.aux, e.ldaux = aux.data_ptr(), aux.stride(0)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1
        e.residual, e.ldres = residual.data_ptr(), residual.stride(0)
    if row_map is not None:
        assert row_map.dtype == torch.int32 and row_map.numel() >= M
        e.row_map = row_map.data_ptr()
    if row_scale is not None:
        assert row_scale.dtype == torch.float32
        e.row_scale, e.rows_per_scale = row_scale.data_ptr(), rows_per_scale
    e.alpha = alpha
    e.accumulate = L.ACCUMULATE if accumulate else L.STORE
    rc = L.lib().lav_gemm_f16(_ptr(a), a.stride(0), a_major, _ptr(b), b.stride(0), b_major, M, N, K,
                              ctypes.byref(e), split_k, _stream())
    L.check(rc, "lav_gemm_f16")
    return out
