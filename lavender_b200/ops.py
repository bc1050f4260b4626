"""Torch-tensor front ends of the C-ABI kernels (no autograd here; see functional.py).

Tensors are passed as raw device pointers + sizes; all launches go to torch's current CUDA stream.
"""
import ctypes

import torch

from . import _lib as L

F16 = torch.float16
_CHECK_IDS = __import__("os").environ.get("LAV_CHECK_IDS", "0") == "1"

# Optional per-call timing (bench.py's roofline leg): when PROFILE is a list, every wrapper brackets its C-ABI call
# with CUDA events on the launching stream and appends (family, start_event, end_event, algorithmic_flops).
PROFILE = None
PROFILE_META = False  # also record a shape tag per call (tools/profile_step.py)


class _Timed:
    __slots__ = ("name", "flops", "ev", "meta", "bytes")

    def __init__(self, name, flops=0.0, meta=None, nbytes=0.0):
        self.name, self.flops, self.ev, self.meta, self.bytes = name, flops, None, meta, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and self.ev is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE.append((self.name, self.ev, end, self.flops, self.meta if PROFILE_META else None, self.bytes))
        return False


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk16(t, name):
    assert t.is_cuda and t.dtype == F16 and t.stride(-1) == 1, f"{name}: need a row-major CUDA fp16 tensor"


def _drop(d):
    """d: None or (rng int64[2] device tensor, site, p) -> ctypes LavDropout pointer (or None)."""
    if d is None or d[2] <= 0.0:
        return None
    rng, site, p = d
    assert rng.is_cuda and rng.dtype == torch.int64 and rng.numel() >= 2
    return ctypes.byref(L.Dropout(rng.data_ptr(), int(site) & 0xFFFFFFFF, float(p)))


def gemm(a, b, out, *, M, N, K, a_major=L.MAJOR_K, b_major=L.MAJOR_K, bias=None, act=L.ACT_NONE, aux=None,
         residual=None, row_map=None, row_scale=None, rows_per_scale=1, alpha=1.0, accumulate=False, split_k=0,
         drop=None, bias_grad=None):
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
    if bias_grad is not None:
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() >= M and accumulate and a_major == L.MAJOR_MN
        e.bias_grad = bias_grad.data_ptr()
    if drop is not None and drop[2] > 0.0:
        e.drop = L.Dropout(drop[0].data_ptr(), int(drop[1]) & 0xFFFFFFFF, float(drop[2]))
    # algorithmic HBM bytes of the launch: both operands once, the output once (+ what the epilogue reads / also writes)
    nbytes = 2.0 * K * (M + N) + M * N * ((2 if out.dtype == F16 else 4) + (4 if residual is not None else 0)
                                          + (2 if aux is not None else 0) + (4 if accumulate else 0))
    with _Timed("gemm", 2.0 * M * N * K, ("gemm", M, N, K, a_major, b_major, act, out.dtype == F16, accumulate), nbytes):
        rc = L.lib().lav_gemm_f16(_ptr(a), a.stride(0), a_major, _ptr(b), b.stride(0), b_major, M, N, K,
                                  ctypes.byref(e), split_k, _stream())
    L.check(rc, "lav_gemm_f16")
    return out


def _p(t):
    return t.data_ptr() if t is not None else None


def layernorm_fwd(x, gamma, beta, eps, *, rows, C, G=1, row_map=None, out16=None, out32=None, mean=None, rstd=None):
    """out[r] = LN(concat_g x[row_map[r*G+g]]) (width G*C); x fp32 2-D."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    with _Timed("layernorm_fwd", 0.0, ("ln_fwd", rows, C, G, row_map is not None)):
        rc = L.lib().lav_layernorm_fwd(_p(x), x.stride(0), _p(row_map), G, C, _p(gamma), _p(beta), eps,
                                       _p(out16), out16.stride(0) if out16 is not None else 0,
                                       _p(out32), out32.stride(0) if out32 is not None else 0,
                                       _p(mean), _p(rstd), rows, _stream())
    L.check(rc, "lav_layernorm_fwd")


def layernorm_bwd(dy, x, gamma, mean, rstd, *, rows, C, G=1, row_map=None, add32=None, dx32=None, dx16=None,
                  dgamma=None, dbeta=None, drop16=None, dx16_map=None, dx16_at_src=False, dx16_scale=None, dx16_rps=1):
    """dx16 lands at row dx16_map[src row] / the src row (dx16_at_src) / r, times dx16_scale[src row // dx16_rps]:
    the gradient cast (+ window gather, + DropPath scale) that would otherwise follow as a separate kernel."""
    assert dy.dtype in (F16, torch.float32) and x.dtype == torch.float32
    ws = None
    if dgamma is not None and G == 1 and C <= 1024:   # scratch for the per-block partial sums of dgamma / dbeta
        ws = torch.empty(8 * 148 * 2 * C, dtype=torch.float32, device=x.device)
    with _Timed("layernorm_bwd", 0.0, ("ln_bwd", rows, C, G, row_map is not None, dy.dtype == F16, add32 is not None,
                                       dx32 is not None, dx16 is not None)):
        rc = L.lib().lav_layernorm_bwd_ex(_p(dy), dy.stride(0), int(dy.dtype == torch.float32), _p(x), x.stride(0),
                                          _p(row_map), G, C, _p(gamma), _p(mean), _p(rstd),
                                          _p(add32), add32.stride(0) if add32 is not None else 0,
                                          _p(dx32), dx32.stride(0) if dx32 is not None else 0,
                                          _p(dx16), dx16.stride(0) if dx16 is not None else 0,
                                          _p(dx16_map), int(bool(dx16_at_src)), _p(dx16_scale), int(dx16_rps),
                                          _p(dgamma), _p(dbeta), _p(ws), ws.numel() if ws is not None else 0, rows,
                                          _drop(drop16), _stream())
    L.check(rc, "lav_layernorm_bwd_ex")


def scale_cast(x, out16, *, rows, C, row_map=None, row_scale=None, rows_per_scale=1, alpha=1.0):
    assert x.dtype == torch.float32 and out16.dtype == F16
    with _Timed("cast", 0.0, ("cast", rows, C, row_map is not None)):
        rc = L.lib().lav_scale_cast_f16(_p(x), x.stride(0), _p(row_map), _p(row_scale), rows_per_scale, alpha,
                                        _p(out16), out16.stride(0), rows, C, _stream())
    L.check(rc, "lav_scale_cast_f16")
    return out16


def split3(x, out16, *, rows, C, weight=False):
    """High-precision mode operand: out16[r] = [hi | lo | hi] (activation) or [hi | hi | lo] (weight=True) of the fp32
    rows x[r, :C] (hi = fp16(x), lo = fp16(x - hi)); out16 is [rows, >= 3C] fp16."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and out16.dtype == F16 and out16.stride(-1) == 1
    with _Timed("cast"):
        rc = L.lib().lav_split3_f16(_p(x), x.stride(0), _p(out16), out16.stride(0), rows, C, 1 if weight else 0, _stream())
    L.check(rc, "lav_split3_f16")
    return out16


def cast_f16(src, dst):
    assert src.dtype == torch.float32 and dst.dtype == F16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    with _Timed("cast"):
        L.check(L.lib().lav_cast_f32_to_f16(_p(src), _p(dst), src.numel(), _stream()), "lav_cast_f32_to_f16")
    return dst


def colsum(x16, out, *, rows, N, alpha=1.0):
    assert x16.dtype == F16 and out.dtype == torch.float32
    with _Timed("colsum"):
        L.check(L.lib().lav_colsum_f16(_p(x16), x16.stride(0), rows, N, _p(out), alpha, _stream()), "lav_colsum_f16")


def attn_fwd(qkv, out, lse, *, q_off, k_off, v_off, head_dim, nheads, nprob, L_tok, scale, bias16=None,
             prob_class=None, key_bias=None, drop=None, causal_from=-1, out32=None):
    """out (fp16) and / or out32 (fp32, high-precision mode) receive O."""
    _chk16(qkv, "qkv")
    if out is not None:
        _chk16(out, "out")
    assert out32 is None or (out32.dtype == torch.float32 and out32.stride(-1) == 1)
    fam = "win_attn_fwd" if head_dim == 32 else "bert_attn_fwd"
    with _Timed(fam, 4.0 * L_tok * L_tok * head_dim * nheads * nprob, (fam, L_tok, head_dim, nheads, nprob)):
      rc = L.lib().lav_attn_fwd_ex(_p(qkv), qkv.stride(0), qkv.shape[0], q_off, k_off, v_off, head_dim, nheads, nprob,
                                 L_tok, scale, _p(bias16), bias16.shape[-1] if bias16 is not None else 0,
                                 _p(prob_class), prob_class.numel() if prob_class is not None else 1, _p(key_bias),
                                 key_bias.shape[-1] if key_bias is not None else 0, int(causal_from),
                                 _p(out), out.stride(0) if out is not None else 0,
                                 _p(out32), out32.stride(0) if out32 is not None else 0, _p(lse), _drop(drop), _stream())
    L.check(rc, "lav_attn_fwd_ex")


def attn_bwd(qkv, out, dout, lse, dq_acc, dqkv, *, q_off, k_off, v_off, head_dim, nheads, nprob, L_tok, scale,
             bias16=None, prob_class=None, key_bias=None, ds16=None, drop=None, causal_from=-1):
    fam = "win_attn_bwd" if head_dim == 32 else "bert_attn_bwd"
    delta = torch.empty(nheads, qkv.shape[0], dtype=torch.float32, device=qkv.device)   # workspace: rowsum(dO * O)
    with _Timed(fam, 8.0 * L_tok * L_tok * head_dim * nheads * nprob, (fam, L_tok, head_dim, nheads, nprob)):
      rc = L.lib().lav_attn_bwd_f16(_p(qkv), qkv.stride(0), qkv.shape[0], q_off, k_off, v_off, head_dim, nheads, nprob,
                                  L_tok, scale, _p(bias16), bias16.shape[-1] if bias16 is not None else 0,
                                  _p(prob_class), prob_class.numel() if prob_class is not None else 1,
                                  _p(key_bias), key_bias.shape[-1] if key_bias is not None else 0, int(causal_from),
                                  _p(out), out.stride(0), _p(dout), dout.stride(0), _p(lse), _p(delta),
                                  _p(dq_acc), dq_acc.stride(0), _p(dqkv), dqkv.stride(0),
                                  _p(ds16), ds16.shape[-1] if ds16 is not None else 0, _drop(drop), _stream())
    L.check(rc, "lav_attn_bwd_f16")


def relpos_bias_expand(table, rel_index, L_tok, labels, dense16, scale):
    """dense16: [ncls, nheads, NP, NP] fp16 = (bias + shift mask) / scale; labels: uint8 [ncls, NP] or None;
    `scale` is the softmax scale later passed to attn_fwd / attn_bwd."""
    ncls, nheads, NP, _ = dense16.shape
    with _Timed("relpos_expand"):
        rc = L.lib().lav_relpos_bias_expand(_p(table), nheads, _p(rel_index), L_tok, _p(labels), ncls, _p(dense16), NP,
                                            1.0 / float(scale), _stream())
    L.check(rc, "lav_relpos_bias_expand")


def relpos_bias_grad(ds16, rel_index, L_tok, dtable):
    nprob, nheads, NP, _ = ds16.shape
    with _Timed("relpos_grad"):
        rc = L.lib().lav_relpos_bias_grad(_p(ds16), nprob, nheads, NP, L_tok, _p(rel_index), _p(dtable), _stream())
    L.check(rc, "lav_relpos_bias_grad")


def dropout_f32(x, out, drop):
    """out = x * keep / (1 - p), fp32 2-D [rows, C] (C % 8 == 0); `drop` = (rng, site, p)."""
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.dim() == 2 and x.shape == out.shape
    assert x.stride(1) == 1 and out.stride(1) == 1
    d = L.Dropout(drop[0].data_ptr(), int(drop[1]) & 0xFFFFFFFF, float(drop[2]))
    with _Timed("dropout"):
        rc = L.lib().lav_dropout_f32(_p(x), x.stride(0), _p(out), out.stride(0), x.shape[0], x.shape[1], ctypes.byref(d),
                                     _stream())
    L.check(rc, "lav_dropout_f32")
    return out


def dropout_mask(rows, C, drop, head=-1):
    """uint8 keep mask [rows, C] of a site (test helper; head >= 0: attention-probability index space)."""
    keep = torch.empty(rows, C, dtype=torch.uint8, device=drop[0].device)
    d = L.Dropout(drop[0].data_ptr(), int(drop[1]) & 0xFFFFFFFF, float(drop[2]))
    L.check(L.lib().lav_dropout_mask(_p(keep), rows, C, head, ctypes.byref(d), _stream()), "lav_dropout_mask")
    return keep


def gelu_bwd(dy16, pre16, out16):
    """out16 = dy16 * pre16, where pre16 is the `aux` tensor of an ACT_GELU GEMM: since ABI v3 that epilogue saves
    gelu'(pre-activation) rather than the pre-activation itself, so the backward needs no erf."""
    assert dy16.is_contiguous() and pre16.is_contiguous() and out16.is_contiguous()
    with _Timed("gelu_bwd"):
        L.check(L.lib().lav_gelu_bwd_f16(_p(dy16), _p(pre16), _p(out16), dy16.numel(), _stream()), "lav_gelu_bwd_f16")
    return out16


def bert_embed_ln_fwd(ids, pos_ids, type_ids, word, pos, typ, gamma, beta, eps, sum32, y32, mean, rstd, *, Lt):
    """ids (int64, contiguous, `rows` elements); tables fp32 [*, C]."""
    rows, C = ids.numel(), word.shape[1]
    for t in (ids, pos_ids, type_ids):
        assert t is None or (t.dtype == torch.int64 and t.is_contiguous() and t.numel() == rows)
    check_ids(ids, word.shape[0], "word_embeddings")
    check_ids(pos_ids, pos.shape[0], "position_embeddings")
    check_ids(type_ids, typ.shape[0], "token_type_embeddings")
    with _Timed("embed"):
        rc = L.lib().lav_bert_embed_ln_fwd(_p(ids), _p(pos_ids), _p(type_ids), rows, Lt, C, word.shape[0], pos.shape[0],
                                           typ.shape[0], _p(word), _p(pos), _p(typ), _p(gamma), _p(beta), eps, _p(sum32),
                                           _p(y32), _p(mean), _p(rstd), _stream())
    L.check(rc, "lav_bert_embed_ln_fwd")


def check_ids(ids, n, what):
    """torch's nn.Embedding raises on out-of-range indices; the kernels clamp (no device-side assert).  LAV_CHECK_IDS=1
    validates on the host (one sync per call; off by default on the hot path, on in the tests)."""
    if _CHECK_IDS and ids is not None and ids.numel():
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= n:
            raise IndexError(f"{what}: index out of range [{lo}, {hi}] for an embedding of {n} rows")


def bert_embed_bwd(dsum32, ids, pos_ids, type_ids, dword, dpos, dtyp, *, Lt, padding_idx=-1):
    rows, C = ids.numel(), dsum32.shape[-1]
    assert dsum32.is_contiguous() and dsum32.dtype == torch.float32
    with _Timed("embed"):
        rc = L.lib().lav_bert_embed_bwd(_p(dsum32), _p(ids), _p(pos_ids), _p(type_ids), rows, Lt, C, dword.shape[0],
                                        dpos.shape[0], dtyp.shape[0], _p(dword), _p(dpos), _p(dtyp), int(padding_idx),
                                        _stream())
    L.check(rc, "lav_bert_embed_bwd")


def vid_embed_ln_fwd(feat, emb_cls, emb_pos, emb_len, emb_odr, odr_swap, gamma, beta, eps, sum32, y32, mean, rstd, *,
                     B, T, hw):
    C = feat.shape[-1]
    assert feat.dtype == torch.float32 and feat.stride(-1) == 1 and feat.dim() == 2
    with _Timed("embed"):
        rc = L.lib().lav_vid_embed_ln_fwd(_p(feat), feat.stride(0), _p(emb_cls), _p(emb_pos), _p(emb_len), _p(emb_odr),
                                          _p(odr_swap), B, T, hw, C, _p(gamma), _p(beta), eps, _p(sum32), _p(y32), _p(mean),
                                          _p(rstd), _stream())
    L.check(rc, "lav_vid_embed_ln_fwd")


def vid_embed_bwd(dsum32, odr_swap, *, B, T, hw, C, dfeat16=None, dfeat32=None, demb_cls=None, demb_pos=None,
                  demb_len=None, demb_odr=None):
    assert dsum32.is_contiguous() and dsum32.dtype == torch.float32
    with _Timed("embed"):
        rc = L.lib().lav_vid_embed_bwd(_p(dsum32), B, T, hw, C, _p(odr_swap), _p(dfeat16),
                                       dfeat16.stride(0) if dfeat16 is not None else 0, _p(dfeat32),
                                       dfeat32.stride(0) if dfeat32 is not None else 0, _p(demb_cls), _p(demb_pos),
                                       _p(demb_len), _p(demb_odr), _stream())
    L.check(rc, "lav_vid_embed_bwd")


def xent_fwd(logits, labels, ignore_index, row_lse, row_loss, loss_sum, count):
    """logits fp32 [rows, V] (row stride >= V); labels int64 [rows]."""
    rows, V = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and labels.dtype == torch.int64
    with _Timed("xent"):
        rc = L.lib().lav_xent_fwd(_p(logits), logits.stride(0), _p(labels), rows, V, ignore_index, _p(row_lse),
                                  _p(row_loss), _p(loss_sum), _p(count), _stream())
    L.check(rc, "lav_xent_fwd")


def xent_bwd(logits, labels, ignore_index, row_lse, gout, count, d32=None, d16=None):
    rows, V = logits.shape
    with _Timed("xent"):
        rc = L.lib().lav_xent_bwd(_p(logits), logits.stride(0), _p(labels), rows, V, ignore_index, _p(row_lse), _p(gout),
                                  _p(count), _p(d32), d32.stride(0) if d32 is not None else 0, _p(d16),
                                  d16.stride(0) if d16 is not None else 0, _stream())
    L.check(rc, "lav_xent_bwd")


def grad_stats(grad, state, ws=None):
    """ws: zero-initialised fp32 scratch (>= 8 * SMs + 1) for the deterministic fixed-order reduction (see the header)."""
    assert grad.dtype == torch.float32 and grad.is_contiguous() and state.dtype == torch.float32
    with _Timed("optimizer"):
        rc = L.lib().lav_grad_stats(_p(grad), grad.numel(), _p(state), _p(ws), ws.numel() if ws is not None else 0,
                                    _stream())
    L.check(rc, "lav_grad_stats")


def adamw_step(param, grad, exp_avg, exp_avg_sq, group_of_block, group_lr, group_wd, state, *, beta1, beta2, eps,
               max_grad_norm, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, param16=None,
               group_step0=None, group_bc=None):
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n and group_of_block.numel() * 8 == n
    ngroups = group_lr.numel()
    if group_bc is None:
        group_bc = torch.empty(2 * ngroups, dtype=torch.float32, device=param.device)
    assert group_bc.numel() >= 2 * ngroups and (group_step0 is None or group_step0.numel() >= ngroups)
    with _Timed("optimizer"):
        rc = L.lib().lav_adamw_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, _p(group_of_block), _p(group_lr),
                                    _p(group_wd), beta1, beta2, eps, max_grad_norm, _p(state), growth_factor,
                                    backoff_factor, growth_interval, _p(param16), _p(group_step0), _p(group_bc), ngroups,
                                    _stream())
    L.check(rc, "lav_adamw_step")
