"""B200-native BERT pieces of the LAVENDER hot path, with HuggingFace-identical module / parameter names so the
reference's state-dict keys load unchanged:

  BertEmbeddings   <- transformers BertEmbeddings   as used by EncTxt.forward              (model.py:100-102,127-129)
  BertEncoder      <- transformers BertEncoder      as used by LAVENDER_Base.go_cross       (model.py:152-165,242)
  BertOnlyMLMHead  <- transformers BertOnlyMLMHead  as used for `fc_mtm`                    (main_pretrain_mlm.py:46-48,69,115)

Arithmetic (verified against transformers 5.5 by the oracle goldens): post-LN layers
    a = LN(x + drop(Wo . softmax(Q K^T / sqrt(hd) + mask) V)) ;  y = LN(a + drop(W2 . gelu_erf(W1 . a)))
with LN eps 1e-12.  The residual stream, LN / softmax statistics and every accumulator are fp32; GEMM and
attention operands are fp16 (the reference's GPU path is fp16 autocast, agent.py:219).

Dropout (hidden 0.1 after the embeddings / BertSelfOutput.dense / BertOutput.dense, attention-probability 0.1) is
applied in train() mode exactly where HF applies it, inside the producing kernels (GEMM epilogue, attention
softmax, embedding row kernel) from counter-based masks that the backward kernels regenerate (dropout.py,
csrc/rng.cuh).  `config.lav_eval_dropout = True` forces identity dropout in train() mode (the configuration the
bit-level parity tests against the reference goldens use, since a torch RNG stream cannot be reproduced).
"""
import math

import torch
import torch.nn as nn

from . import ops
from . import precision
from . import _lib as L
from .arena import arena_of
from .dropout import rng_for
from .functional import (F16, F32, cast16, empty16, empty32, linear_dgrad, linear_fwd, linear_fwd_hp, linear_wgrad,
                         require_cuda)

NEG_INF = float("-inf")


class BertConfig:
    """The subset of transformers.BertConfig the hot path reads (bert-base-uncased defaults)."""

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02, **unused):
        self.vocab_size, self.hidden_size = vocab_size, hidden_size
        self.num_hidden_layers, self.num_attention_heads = num_hidden_layers, num_attention_heads
        self.intermediate_size, self.max_position_embeddings = intermediate_size, max_position_embeddings
        self.type_vocab_size, self.layer_norm_eps = type_vocab_size, layer_norm_eps
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.initializer_range = initializer_range
        self.hidden_act = "gelu"
        self.model_type = "bert"


def _init_bert_weights(module, std):
    """transformers BertPreTrainedModel._init_weights: normal(0, 0.02) weights, zero biases, unit LayerNorm."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0.0, std=std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, mean=0.0, std=std)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


# ---------------------------------------------------------------------------------------------------------
# parameter containers (names == HF)
# ---------------------------------------------------------------------------------------------------------
class BertSelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)


class BertSelfOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class BertAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = BertSelfAttention(c)
        self.output = BertSelfOutput(c)


class BertIntermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.intermediate_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = BertAttention(c)
        self.intermediate = BertIntermediate(c)
        self.output = BertOutput(c)


def _drop_p(mod, p):
    """Effective drop probability of a site of `mod` (0 in eval() mode or under config.lav_eval_dropout)."""
    if mod.training and p > 0.0 and not getattr(mod.config, "lav_eval_dropout", False):
        return float(p)
    return 0.0


class BertEmbeddings(nn.Module):
    """forward(input_ids, token_type_ids=None, position_ids=None) -> [.., Lt, H] fp32."""

    def __init__(self, config):
        super().__init__()
        c = self.config = config
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        _init_bert_weights(self, c.initializer_range)
        with torch.no_grad():
            self.word_embeddings.weight[0].zero_()

    def forward(self, input_ids, token_type_ids=None, position_ids=None):
        return _BertEmbedFn.apply(input_ids, token_type_ids, position_ids, self, self.word_embeddings.weight,
                                  self.position_embeddings.weight, self.token_type_embeddings.weight,
                                  self.LayerNorm.weight, self.LayerNorm.bias)


class _BertEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, type_ids, pos_ids, mod, word, pos, typ, gamma, beta):
        require_cuda(ids, "BertEmbeddings")
        shp = ids.shape
        Lt = shp[-1]
        H = word.shape[1]

        def flat(t):
            if t is None:
                return None
            return t.expand(shp).reshape(-1).to(torch.int64).contiguous()
        ids1, tt1, ps1 = flat(ids), flat(type_ids), flat(pos_ids)
        rows = ids1.numel()
        dev = ids.device
        s32, y32 = empty32(rows, H, device=dev), empty32(rows, H, device=dev)
        mean, rstd = empty32(rows, device=dev), empty32(rows, device=dev)
        ops.bert_embed_ln_fwd(ids1, ps1, tt1, word, pos, typ, gamma, beta, mod.LayerNorm.eps, s32, y32, mean, rstd, Lt=Lt)
        ctx.drop = rng_for(dev).spec(_drop_p(mod, mod.config.hidden_dropout_prob))   # BertEmbeddings.dropout
        if ctx.drop is not None:
            ops.dropout_f32(y32, y32, ctx.drop)
        ctx.mod, ctx.saved, ctx.Lt = mod, (ids1, tt1, ps1, s32, mean, rstd), Lt
        ctx.params = (word, pos, typ, gamma, beta)
        return y32.view(*shp, H)

    @staticmethod
    def backward(ctx, gy):
        ar = arena_of(ctx.mod)
        word, pos, typ, gamma, beta = ctx.params
        ar.prepare_grads([word, pos, typ, gamma, beta])
        ids1, tt1, ps1, s32, mean, rstd = ctx.saved
        rows, H = s32.shape
        g2 = gy.reshape(rows, H)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        if ctx.drop is not None:
            g2 = ops.dropout_f32(g2, empty32(rows, H, device=gy.device), ctx.drop)
        d = empty32(rows, H, device=gy.device)
        ops.layernorm_bwd(g2, s32, gamma, mean, rstd, rows=rows, C=H, dx32=d, dgamma=ar.g(gamma), dbeta=ar.g(beta))
        pad = ctx.mod.word_embeddings.padding_idx
        ops.bert_embed_bwd(d, ids1, ps1, tt1, ar.g(word), ar.g(pos), ar.g(typ), Lt=ctx.Lt,
                           padding_idx=-1 if pad is None else pad)
        return (None,) * 9


class BertEncoder(nn.Module):
    """forward(hidden[B,L,H] fp32, attention_mask) -> dict(last_hidden_state=[B,L,H], attentions=None).

    attention_mask is what LAVENDER_Base.go_cross passes (model.py:239-242): HF's *extended additive* mask
    [B,1,1,L] (0 keep / finfo.min masked), or a plain [B,L] 0/1 key mask.  The seq2seq mask of
    LAVENDER_Base.get_attn_mask (model.py:208-218) is passed as its structure — the key mask of the video / prefix
    part plus `causal_from` = first text position — and applied inside the attention kernels; an arbitrary
    materialised [B,1,L,L] mask raises.
    `output_attentions` is accepted and ignored: every caller discards the maps."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])
        _init_bert_weights(self, config.initializer_range)

    def key_keep_mask(self, attention_mask, B, Lq):
        m = attention_mask
        if m.dim() == 4:
            if m.shape[1] != 1 or m.shape[2] != 1:
                raise NotImplementedError("BertEncoder: pass the seq2seq mask as (key mask, causal_from=first text "
                                          "position); arbitrary materialised [B,1,L,L] masks are not supported")
            return (m.reshape(B, Lq) >= -1.0)
        if m.dim() == 2:
            return m != 0
        raise ValueError(f"BertEncoder: unsupported attention_mask shape {tuple(m.shape)}")

    def forward(self, hidden_states, attention_mask=None, output_attentions=False, causal_from=None, **unused):
        require_cuda(hidden_states, "BertEncoder")
        B, Lq, H = hidden_states.shape
        NPk = (Lq + 127) // 128 * 128
        kb = torch.full((B, NPk), NEG_INF, dtype=F32, device=hidden_states.device)
        if attention_mask is None:
            kb[:, :Lq] = 0.0
        else:
            keep = self.key_keep_mask(attention_mask, B, Lq)
            kb[:, :Lq].masked_fill_(keep, 0.0)
        params = list(self.parameters())
        out = _BertEncoderFn.apply(hidden_states, kb, self, -1 if causal_from is None else int(causal_from), *params)
        return {"last_hidden_state": out, "attentions": None}


def _layer_views(ar, lyr, H, grad=False):
    sa = lyr.attention.self
    if grad:
        return (ar.span32(sa.query.weight, sa.value.weight, (3 * H, H), grad=True),
                ar.span32(sa.query.bias, sa.value.bias, (3 * H,), grad=True))
    return (ar.span16(sa.query.weight, sa.value.weight, (3 * H, H)),
            ar.span32(sa.query.bias, sa.value.bias, (3 * H,)))


class _BertEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kb, mod, causal_from, *params):
        dev = x.device
        ar = arena_of(mod)
        ar.refresh16()
        c = mod.config
        B, Lq, H = x.shape
        nh = c.num_attention_heads
        hd = H // nh
        M = B * Lq
        x32 = x.reshape(M, H)
        if not x32.is_contiguous():
            x32 = x32.contiguous()
        if x32.dtype != F32:
            x32 = x32.float()
        x16 = ops.scale_cast(x32, empty16(M, H, device=dev), rows=M, C=H)
        saved = []
        rng = rng_for(dev)
        p_hid, p_att = _drop_p(mod, c.hidden_dropout_prob), _drop_p(mod, c.attention_probs_dropout_prob)
        hp = precision.high("bert")   # parity mode: split-fp16 linears on fp32 activations, fp32 attention output
        for lyr in mod.layer:
            d_att, d_so, d_oo = rng.spec(p_att), rng.spec(p_hid), rng.spec(p_hid)   # three dropout sites per layer
            wqkv, bqkv = _layer_views(ar, lyr, H)
            sa = lyr.attention.self
            qkv16 = empty16(M, 3 * H, device=dev)
            if hp:
                linear_fwd_hp(ar, x32, ar.span32(sa.query.weight, sa.value.weight, (3 * H, H)), bqkv, qkv16)
            else:
                linear_fwd(x16, wqkv, bqkv, qkv16)
            ctx16 = empty16(M, H, device=dev)
            ctx32 = empty32(M, H, device=dev) if hp else None
            lse = empty32(nh, M, device=dev)
            ops.attn_fwd(qkv16, ctx16, lse, q_off=0, k_off=H, v_off=2 * H, head_dim=hd, nheads=nh, nprob=B, L_tok=Lq,
                         scale=1.0 / math.sqrt(hd), key_bias=kb, drop=d_att, causal_from=causal_from, out32=ctx32)
            so = lyr.attention.output
            a_pre = empty32(M, H, device=dev)
            if hp:
                linear_fwd_hp(ar, ctx32, so.dense.weight.data, so.dense.bias, a_pre, residual=x32, drop=d_so)
            else:
                linear_fwd(ctx16, ar.w16(so.dense.weight), so.dense.bias, a_pre, residual=x32, drop=d_so)
            a32, a16 = empty32(M, H, device=dev), empty16(M, H, device=dev)
            m1, r1 = empty32(M, device=dev), empty32(M, device=dev)
            ops.layernorm_fwd(a_pre, so.LayerNorm.weight, so.LayerNorm.bias, so.LayerNorm.eps, rows=M, C=H, out16=a16,
                              out32=a32, mean=m1, rstd=r1)
            FF = lyr.intermediate.dense.weight.shape[0]
            pre16 = empty16(M, FF, device=dev)
            oo = lyr.output
            o_pre = empty32(M, H, device=dev)
            if hp:
                i32 = empty32(M, FF, device=dev)
                linear_fwd_hp(ar, a32, lyr.intermediate.dense.weight.data, lyr.intermediate.dense.bias, i32,
                              act=L.ACT_GELU, aux=pre16)
                linear_fwd_hp(ar, i32, oo.dense.weight.data, oo.dense.bias, o_pre, residual=a32, drop=d_oo)
                i16 = cast16(i32)
                del i32, ctx32
            else:
                i16 = empty16(M, FF, device=dev)
                linear_fwd(a16, ar.w16(lyr.intermediate.dense.weight), lyr.intermediate.dense.bias, i16, act=L.ACT_GELU,
                           aux=pre16)
                linear_fwd(i16, ar.w16(oo.dense.weight), oo.dense.bias, o_pre, residual=a32, drop=d_oo)
            y32, y16 = empty32(M, H, device=dev), empty16(M, H, device=dev)
            m2, r2 = empty32(M, device=dev), empty32(M, device=dev)
            ops.layernorm_fwd(o_pre, oo.LayerNorm.weight, oo.LayerNorm.bias, oo.LayerNorm.eps, rows=M, C=H, out16=y16,
                              out32=y32, mean=m2, rstd=r2)
            saved.append(dict(x16=x16, qkv16=qkv16, ctx16=ctx16, lse=lse, a_pre=a_pre, m1=m1, r1=r1, a16=a16,
                              pre16=pre16, i16=i16, o_pre=o_pre, m2=m2, r2=r2, d_att=d_att, d_so=d_so, d_oo=d_oo))
            x32, x16 = y32, y16
        ctx.mod, ctx.saved, ctx.kb, ctx.geom, ctx.causal_from = mod, saved, kb, (B, Lq, H, nh, hd), causal_from
        return x32.view(B, Lq, H)

    @staticmethod
    def backward(ctx, gy):
        mod, saved, kb = ctx.mod, ctx.saved, ctx.kb
        B, Lq, H, nh, hd = ctx.geom
        M = B * Lq
        dev = gy.device
        ar = arena_of(mod)
        ar.prepare_grads(list(mod.parameters()))
        g = gy.reshape(M, H)
        if not g.is_contiguous():
            g = g.contiguous()
        for li in range(len(mod.layer) - 1, -1, -1):
            lyr, sv = mod.layer[li], saved[li]
            so, oo, it = lyr.attention.output, lyr.output, lyr.intermediate
            FF = it.dense.weight.shape[0]
            # y = LN2(o_pre), o_pre = W2 i + b2 + a
            go32, go16 = empty32(M, H, device=dev), empty16(M, H, device=dev)
            ops.layernorm_bwd(g, sv["o_pre"], oo.LayerNorm.weight, sv["m2"], sv["r2"], rows=M, C=H, dx32=go32, dx16=go16,
                              dgamma=ar.g(oo.LayerNorm.weight), dbeta=ar.g(oo.LayerNorm.bias), drop16=sv["d_oo"])
            linear_wgrad(go16, sv["i16"], ar.g(oo.dense.weight), ar.g(oo.dense.bias))
            dpre16 = empty16(M, FF, device=dev)
            linear_dgrad(go16, ar.w16(oo.dense.weight), dpre16, act=L.ACT_GELU_BWD, aux=sv["pre16"])   # pre16 = gelu'(pre-activation)
            linear_wgrad(dpre16, sv["a16"], ar.g(it.dense.weight), ar.g(it.dense.bias))
            da32 = empty32(M, H, device=dev)  # grad wrt a = residual path + FFN path
            linear_dgrad(dpre16, ar.w16(it.dense.weight), da32, residual=go32)
            # a = LN1(a_pre), a_pre = Wo ctx + bo + x
            ga32, ga16 = empty32(M, H, device=dev), empty16(M, H, device=dev)
            ops.layernorm_bwd(da32, sv["a_pre"], so.LayerNorm.weight, sv["m1"], sv["r1"], rows=M, C=H, dx32=ga32,
                              dx16=ga16, dgamma=ar.g(so.LayerNorm.weight), dbeta=ar.g(so.LayerNorm.bias),
                              drop16=sv["d_so"])
            linear_wgrad(ga16, sv["ctx16"], ar.g(so.dense.weight), ar.g(so.dense.bias))
            dctx16 = empty16(M, H, device=dev)
            linear_dgrad(ga16, ar.w16(so.dense.weight), dctx16)
            dq_acc = empty32(M, H, device=dev)   # zeroed by the attention backward's pre-pass
            dqkv16 = empty16(M, 3 * H, device=dev)
            ops.attn_bwd(sv["qkv16"], sv["ctx16"], dctx16, sv["lse"], dq_acc, dqkv16, q_off=0, k_off=H, v_off=2 * H,
                         head_dim=hd, nheads=nh, nprob=B, L_tok=Lq, scale=1.0 / math.sqrt(hd), key_bias=kb,
                         drop=sv["d_att"], causal_from=ctx.causal_from)
            ops.scale_cast(dq_acc, dqkv16, rows=M, C=H)
            gw, gb = _layer_views(ar, lyr, H, grad=True)
            linear_wgrad(dqkv16, sv["x16"], gw, gb)
            wqkv, _ = _layer_views(ar, lyr, H)
            gx = empty32(M, H, device=dev)
            linear_dgrad(dqkv16, wqkv, gx, residual=ga32)
            g = gx
        return (g.view(B, Lq, H), None, None, None) + (None,) * len(list(mod.parameters()))


# ---------------------------------------------------------------------------------------------------------
# MLM head
# ---------------------------------------------------------------------------------------------------------
class BertPredictionHeadTransform(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class BertLMPredictionHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = BertPredictionHeadTransform(c)
        self.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=True)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))
        self.decoder.bias = self.bias  # one parameter, two state-dict keys (HF; model.py:470)


class BertOnlyMLMHead(nn.Module):
    """forward(x[..., H]) -> logits [..., vocab] fp32 (the last dim is a view of a buffer padded to a multiple of 8
    columns so that the rows stay 16-byte aligned for the kernels)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.predictions = BertLMPredictionHead(config)
        _init_bert_weights(self, config.initializer_range)

    def forward(self, sequence_output):
        require_cuda(sequence_output, "BertOnlyMLMHead")
        p = self.predictions
        return _MLMHeadFn.apply(sequence_output, self, p.transform.dense.weight, p.transform.dense.bias,
                                p.transform.LayerNorm.weight, p.transform.LayerNorm.bias, p.decoder.weight, p.bias)

    def forward_split(self, rows, split):
        """One head pass over `rows` [M, H] returned as two logits tensors (rows [:split] and [split:]): the MLM and VTM rows
        of a pre-training step share the transform / decoder GEMMs, the decoder weight gradient and the logit-gradient
        cast instead of running the 30522-wide head twice on 128 + 32 rows (main_pretrain_mlm.py:69,115 call it twice)."""
        require_cuda(rows, "BertOnlyMLMHead")
        assert rows.dim() == 2 and 0 < split < rows.shape[0]
        p = self.predictions
        return _MLMHeadFn.apply(rows, self, p.transform.dense.weight, p.transform.dense.bias, p.transform.LayerNorm.weight,
                                p.transform.LayerNorm.bias, p.decoder.weight, p.bias, split)


class _MLMHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, wd, bd, gamma, beta, wdec, bdec, split=0):
        dev = x.device
        ar = arena_of(mod)
        ar.refresh16()
        shp = x.shape
        H = shp[-1]
        V = wdec.shape[0]
        x2 = x.reshape(-1, H)
        M = x2.shape[0]
        x16 = ops.scale_cast(x2, empty16(M, H, device=dev), rows=M, C=H)  # handles row-strided x2
        pre16, t16 = empty16(M, H, device=dev), empty16(M, H, device=dev)
        t32 = empty32(M, H, device=dev)
        hp = precision.high("head")   # parity mode: split-fp16 transform / decoder products on fp32 activations (precision.py)
        if hp:
            xf = x2.float() if x2.dtype != F32 else x2
            linear_fwd_hp(ar, xf, wd.data, bd, t32, act=L.ACT_GELU, aux=pre16)
        else:
            linear_fwd(x16, ar.w16(wd), bd, t32, act=L.ACT_GELU, aux=pre16)
        mean, rstd = empty32(M, device=dev), empty32(M, device=dev)
        eps = mod.predictions.transform.LayerNorm.eps
        tn32 = empty32(M, H, device=dev) if hp else None
        ops.layernorm_fwd(t32, gamma, beta, eps, rows=M, C=H, out16=t16, out32=tn32, mean=mean, rstd=rstd)
        Vp = (V + 7) // 8 * 8
        logits = empty32(M, Vp, device=dev)
        if hp:
            linear_fwd_hp(ar, tn32, wdec.data, bdec, logits[:, :V])
        else:
            linear_fwd(t16, ar.w16(wdec), bdec, logits[:, :V])
        ctx.mod, ctx.params = mod, (wd, bd, gamma, beta, wdec, bdec)
        ctx.saved, ctx.shp, ctx.split = (x16, pre16, t32, mean, rstd, t16), shp, split
        if split:
            return logits[:split, :V], logits[split:, :V]
        return logits[:, :V].view(*shp[:-1], V)

    @staticmethod
    def backward(ctx, *gls):
        mod = ctx.mod
        wd, bd, gamma, beta, wdec, bdec = ctx.params
        x16, pre16, t32, mean, rstd, t16 = ctx.saved
        ar = arena_of(mod)
        ar.prepare_grads([wd, bd, gamma, beta, wdec, bdec])
        dev = x16.device
        M, H = x16.shape
        V = wdec.shape[0]
        Vp = (V + 7) // 8 * 8
        gl16 = torch.zeros(M, Vp, dtype=F16, device=dev) if Vp != V else empty16(M, Vp, device=dev)
        r0 = 0
        for gl, n in zip(gls, (ctx.split, M - ctx.split) if ctx.split else (M,)):
            if gl is not None:
                ops.scale_cast(gl.reshape(n, V), gl16[r0:r0 + n], rows=n, C=V)
            elif Vp == V:
                gl16[r0:r0 + n].zero_()
            r0 += n
        linear_wgrad(gl16, t16, ar.g(wdec), ar.g(bdec), n_valid=V)
        if M <= 1024:
            # skinny rows x K = vocab: 3-12 output tiles for 148 SMs -> split-K into a zeroed fp32 buffer (r1: one CTA per
            # tile walked all 477 k-blocks: 180-280 us per call for 6 GFLOP)
            dt = torch.zeros(M, H, dtype=F32, device=dev)
            ops.gemm(gl16, ar.w16(wdec), dt, M=M, N=H, K=V, b_major=L.MAJOR_MN, accumulate=True)
        else:   # enough row blocks to fill the GPU: plain fp32 store (same precision as the split-K path above)
            dt = empty32(M, H, device=dev)
            ops.gemm(gl16, ar.w16(wdec), dt, M=M, N=H, K=V, b_major=L.MAJOR_MN)
        dpre_g = empty16(M, H, device=dev)  # grad wrt gelu output (fp16), then through gelu'
        ops.layernorm_bwd(dt, t32, gamma, mean, rstd, rows=M, C=H, dx16=dpre_g, dgamma=ar.g(gamma), dbeta=ar.g(beta))
        dpre16 = ops.gelu_bwd(dpre_g, pre16, empty16(M, H, device=dev))
        linear_wgrad(dpre16, x16, ar.g(wd), ar.g(bd))
        gx = None
        if ctx.needs_input_grad[0]:
            gx = empty32(M, H, device=dev)
            linear_dgrad(dpre16, ar.w16(wd), gx)
            gx = gx.view(ctx.shp)
        return (gx,) + (None,) * 8


# ---------------------------------------------------------------------------------------------------------
# cross entropy (agent.py:73: nn.CrossEntropyLoss(ignore_index=-1))
# ---------------------------------------------------------------------------------------------------------
class CrossEntropyLoss(nn.Module):
    """Drop-in for torch.nn.CrossEntropyLoss(ignore_index=...) (mean reduction) on [rows, V] fp32 CUDA logits."""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits, target):
        return _XentFn.apply(logits, target, self.ignore_index)


class _XentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        require_cuda(logits, "CrossEntropyLoss")
        assert logits.dim() == 2 and target.dim() == 1 and logits.shape[0] == target.shape[0]
        if logits.dtype != F32 or logits.stride(1) != 1:
            logits = logits.float().contiguous()
        target = target.to(torch.int64).contiguous()
        rows = logits.shape[0]
        dev = logits.device
        row_lse = empty32(rows, device=dev)
        acc = torch.zeros(2, dtype=F32, device=dev)  # [loss_sum, count]
        ops.xent_fwd(logits, target, ignore_index, row_lse, None, acc[0:1], acc[1:2])
        ctx.saved = (logits, target, row_lse, acc)
        ctx.ignore_index = ignore_index
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, gout):
        logits, target, row_lse, acc = ctx.saved
        rows, V = logits.shape
        # rows padded to a multiple of 8 columns: 16-byte aligned rows for the vectorised kernels that consume the
        # gradient (the MLM head's fp16 cast); the returned gradient is the [rows, V] view
        Vp = (V + 7) // 8 * 8
        d = empty32(rows, Vp, device=logits.device)[:, :V]
        g = gout.reshape(1).to(F32).contiguous()
        ops.xent_bwd(logits, target, ctx.ignore_index, row_lse, g, acc[1:2], d32=d)
        return d, None, None
