"""Host-side mirror of the reference's model API (model.py) on top of the native kernels.

  EncVideo        model.py:5-93     Video Swin -> fc -> [cls ; patches] + emb_pos + emb_len -> LayerNorm
  EncTxt          model.py:96-142   BERT embeddings (txt_backbone_embed_only) or embeddings + text encoder
  LAVENDER_Base   model.py:145-473  go_feat / go_cross / get_attn_mask / get_pretxt / prepro_txt_inputs / load_ckpt

Module attribute names, parameter names and state-dict keys are the reference's, so released
checkpoints load and `agent.Agent_Base.build_optimizer`'s name-based grouping (agent.py:98-119) is unchanged.
The HuggingFace dependency is replaced by lavender_b200.bert (same parameter names); weights come from a local
HF directory when one exists (utils/args.py:216-237 maps names to ./_models/huggingface_transformers/...), else
random init in HF's scheme.
"""
import json
import os
import warnings

import torch
import torch.nn as nn

from . import ops
from .arena import arena_of, set_arena_root
from .bert import BertConfig, BertEmbeddings, BertEncoder, BertOnlyMLMHead
from .functional import F32, empty16, empty32, native_linear, require_cuda
from .video_swin import get_vidswin_model


# ---------------------------------------------------------------------------------------------------------
# pretrained BERT lookup (replaces transformers.AutoModel*.from_pretrained at model.py:100,152 and
# main_pretrain_mlm.py:46)
# ---------------------------------------------------------------------------------------------------------
def _bert_config_for(name_or_path, args=None):
    cfg = {}
    if isinstance(name_or_path, str) and os.path.isfile(os.path.join(name_or_path, "config.json")):
        with open(os.path.join(name_or_path, "config.json")) as f:
            cfg = json.load(f)
        if cfg.get("model_type", "bert") != "bert":
            raise NotImplementedError(f"only BERT fusion / text backbones are supported, got {cfg.get('model_type')}")
    elif isinstance(name_or_path, str) and "roberta" in name_or_path.lower():
        raise NotImplementedError("RoBERTa backbones (model.py:160-161) are outside the native hot path")
    override = getattr(args, "bert_config", None) if args is not None else None
    if override:
        cfg.update(dict(override))
    return BertConfig(**cfg)


def _hf_state_dict(name_or_path):
    """State dict of a local HF checkpoint directory, or None (offline: no hub access)."""
    if not isinstance(name_or_path, str) or not os.path.isdir(name_or_path):
        return None
    st = os.path.join(name_or_path, "model.safetensors")
    if os.path.isfile(st):
        try:
            from safetensors.torch import load_file
            return load_file(st)
        except ImportError:
            pass
    pt = os.path.join(name_or_path, "pytorch_model.bin")
    if os.path.isfile(pt):
        return torch.load(pt, map_location="cpu")
    return None


def _load_prefixed(module, sd, prefixes, what, source):
    """Copies the entries of `sd` under any of `prefixes` into `module` (HF checkpoints use `bert.` for the
    trunk and `cls.` for the head; old ones store LayerNorm as gamma/beta).  Never silent: without a checkpoint it
    warns that `what` keeps its random (HF-scheme) init — the reference would fetch the weights from the hub or fail
    (model.py:100-102,152-157) —, a checkpoint that matches nothing raises, a partial match reports the missing keys."""
    if sd is None:
        warnings.warn(f"lavender_b200: no local HuggingFace checkpoint at {source!r} (offline: no hub access) - {what} is "
                      f"RANDOMLY initialised, not pre-trained.  Point txt_backbone / fusion_encoder / tokenizer at a local "
                      f"directory (utils/args.py:216-237 maps names to ./_models/huggingface_transformers/...) or load a "
                      f"LAVENDER checkpoint with load_ckpt().", stacklevel=3)
        return 0
    own = module.state_dict()
    picked = {}
    for k, v in sd.items():
        k = k.replace("LayerNorm.gamma", "LayerNorm.weight").replace("LayerNorm.beta", "LayerNorm.bias")
        for p in prefixes:
            if k.startswith(p) and k[len(p):] in own and own[k[len(p):]].shape == v.shape:
                picked[k[len(p):]] = v
    if not picked:
        raise RuntimeError(f"lavender_b200: the checkpoint at {source!r} has no tensor matching {what} "
                           f"(prefixes {prefixes}); refusing to continue with random weights")
    module.load_state_dict(picked, strict=False)
    missing = sorted(set(own) - set(picked))
    if missing:
        warnings.warn(f"lavender_b200: {what} loaded {len(picked)} tensors from {source!r}, {len(missing)} keep their "
                      f"init: {missing[:6]}{' ...' if len(missing) > 6 else ''}", stacklevel=3)
    return len(picked)


def build_bert_embeddings(name_or_path, args=None):
    cfg = _bert_config_for(name_or_path, args)
    emb = BertEmbeddings(cfg)
    _load_prefixed(emb, _hf_state_dict(name_or_path), ("bert.embeddings.", "embeddings."), "the text embeddings",
                   name_or_path)
    return emb, cfg


def build_bert_encoder(name_or_path, args=None, rand_init=False):
    cfg = _bert_config_for(name_or_path, args)
    enc = BertEncoder(cfg)
    if not rand_init:
        _load_prefixed(enc, _hf_state_dict(name_or_path), ("bert.encoder.", "encoder."), "the BERT encoder", name_or_path)
    return enc, cfg


def build_mlm_head(name_or_path, args=None):
    cfg = _bert_config_for(name_or_path, args)
    head = BertOnlyMLMHead(cfg)
    sd = _hf_state_dict(name_or_path)
    if sd is not None and not any(k.startswith("cls.") for k in sd):
        sd = {**sd, **{"cls.predictions.decoder.weight": sd[k] for k in
                       ("bert.embeddings.word_embeddings.weight", "embeddings.word_embeddings.weight") if k in sd}}
    _load_prefixed(head, sd, ("cls.",), "the MLM head (fc_mtm)", name_or_path)
    if sd is not None and "cls.predictions.decoder.weight" not in sd:  # tied checkpoints store the decoder only as word embeddings
        for k in ("bert.embeddings.word_embeddings.weight", "embeddings.word_embeddings.weight"):
            if k in sd and sd[k].shape == head.predictions.decoder.weight.shape:
                with torch.no_grad():
                    head.predictions.decoder.weight.copy_(sd[k])
    return head, cfg


def extended_attention_mask(mask, shape=None, device=None, dtype=torch.float32):
    """HF get_extended_attention_mask (`LAVENDER_Base.mask_ext`, model.py:158,239): [B,L] -> [B,1,1,L],
    [B,L,L] -> [B,1,L,L]; kept keys 0, masked keys finfo.min."""
    if isinstance(device, torch.dtype):  # transformers >= 5 passes dtype third
        dtype = device
    m = mask[:, None, None, :] if mask.dim() == 2 else mask[:, None, :, :]
    return (1.0 - m.to(dtype)) * torch.finfo(dtype).min


# ---------------------------------------------------------------------------------------------------------
class _VidEmbedFn(torch.autograd.Function):
    """[cls ; feat] + emb_pos + emb_len|emb_odr -> LayerNorm (model.py:69-85), one kernel each way."""

    @staticmethod
    def forward(ctx, feat, swap, mod, geom, emb_cls, emb_pos, emb_len, emb_odr, gamma, beta):
        B, T, hw = geom
        C = feat.shape[-1]
        dev = feat.device
        rows = B * T * (1 + hw)
        f2 = feat.reshape(B * T * hw, C)
        if not f2.is_contiguous():
            f2 = f2.contiguous()
        s32, y32 = empty32(rows, C, device=dev), empty32(rows, C, device=dev)
        mean, rstd = empty32(rows, device=dev), empty32(rows, device=dev)
        ops.vid_embed_ln_fwd(f2, emb_cls, emb_pos, emb_len, emb_odr, swap, gamma, beta, mod.norm.eps, s32, y32, mean,
                             rstd, B=B, T=T, hw=hw)
        ctx.mod, ctx.geom, ctx.saved, ctx.feat_shape = mod, geom, (s32, mean, rstd, swap), feat.shape
        ctx.params = (emb_cls, emb_pos, emb_len, emb_odr, gamma, beta)
        return y32.view(B, T * (1 + hw), C)

    @staticmethod
    def backward(ctx, gy):
        B, T, hw = ctx.geom
        s32, mean, rstd, swap = ctx.saved
        emb_cls, emb_pos, emb_len, emb_odr, gamma, beta = ctx.params
        ar = arena_of(ctx.mod)
        touched = [emb_cls, emb_pos, emb_len, gamma, beta] + ([emb_odr] if swap is not None else [])
        ar.prepare_grads(touched)
        rows, C = s32.shape
        g2 = gy.reshape(rows, C)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        d = empty32(rows, C, device=gy.device)
        ops.layernorm_bwd(g2, s32, gamma, mean, rstd, rows=rows, C=C, dx32=d, dgamma=ar.g(gamma), dbeta=ar.g(beta))
        dfeat = empty32(B * T * hw, C, device=gy.device)
        ops.vid_embed_bwd(d, swap, B=B, T=T, hw=hw, C=C, dfeat32=dfeat, demb_cls=ar.g(emb_cls), demb_pos=ar.g(emb_pos),
                          demb_len=ar.g(emb_len), demb_odr=ar.g(emb_odr) if swap is not None else None)
        return (dfeat.view(ctx.feat_shape),) + (None,) * 9


class EncVideo(nn.Module):
    """model.py:5-93.  forward(img[B,T,3,H,W], odr=None, vt_mask=None) -> (f_img [B,T(1+hw),hidden], m_img [B,T(1+hw)])."""

    def __init__(self, args, hidden_size):
        super().__init__()
        self.swin = get_vidswin_model(args)
        self.latent_feat_size = self.swin.norm.normalized_shape[0]
        self.img_feature_dim = hidden_size
        self.swinbert = getattr(args, "swinbert", False)
        if self.swinbert:
            raise NotImplementedError("the SwinBERT-initialised variant (model.py:52-67) is outside the native hot path")
        self.max_size_frame = getattr(args, "max_size_frame", 6)
        self.max_size_patch = getattr(args, "max_size_patch", 14)
        self.fc = nn.Linear(self.latent_feat_size, hidden_size) if self.latent_feat_size != hidden_size else None
        self.emb_cls = nn.Parameter(0.02 * torch.randn(1, 1, 1, hidden_size))
        self.emb_pos = nn.Parameter(0.02 * torch.randn(1, 1, 1 + self.max_size_patch ** 2, hidden_size))
        self.emb_len = nn.Parameter(0.02 * torch.randn(1, self.max_size_frame, 1, hidden_size))
        self.emb_odr = nn.Parameter(0.02 * torch.randn(1, 1, 1, hidden_size))
        self.norm = nn.LayerNorm(hidden_size)
        self.transform_normalize = None

    def forward(self, img, odr=None, vt_mask=None):
        require_cuda(img, "EncVideo")
        B, T, _, H, W = img.shape
        h, w = H // 32, W // 32
        if T > self.max_size_frame or h * w > self.max_size_patch ** 2:
            raise ValueError(f"clip of {T} frames / {h}x{w} patches exceeds emb_len / emb_pos "
                             f"({self.max_size_frame} frames, {self.max_size_patch}^2 patches)")
        if self.transform_normalize is not None:
            img = self.transform_normalize(img)
        f = self.swin.forward_features(img.transpose(1, 2))          # [B,T,h,w,8C] channels-last fp32
        f = f.reshape(B * T * h * w, self.latent_feat_size)
        if self.fc is not None:
            f = native_linear(self.fc, f)
        swap = None
        if odr is not None:  # model.py:72-81: frame i keeps emb_len[i] only where odr[b][i] == i
            o = torch.as_tensor(odr, device=img.device).reshape(B, T)
            swap = (o != torch.arange(T, device=img.device)).to(torch.uint8).reshape(-1).contiguous()
        f_img = _VidEmbedFn.apply(f, swap, self, (B, T, h * w), self.emb_cls, self.emb_pos, self.emb_len, self.emb_odr,
                                  self.norm.weight, self.norm.bias)
        m_img = torch.ones(B, T, 1 + h * w, dtype=torch.long, device=img.device)
        if vt_mask is not None:
            m_img = m_img * vt_mask
        return f_img, m_img.view(B, T * (1 + h * w))


class EncTxt(nn.Module):
    """model.py:96-142.  forward(txt, mask_txt=None, token_type_ids=None, position_ids=None, attn_mask_type)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.emb_txt, cfg = build_bert_embeddings(args.txt_backbone, args)
        if args.txt_backbone_embed_only:
            self.txt_trsfr, self.mask_ext = None, None
        else:
            self.txt_trsfr, _ = build_bert_encoder(args.txt_backbone, args)
            self.mask_ext = extended_attention_mask
        self.size_vocab = cfg.vocab_size

    def get_attn_mask(self, mask_txt, attn_mask_type="full"):
        if attn_mask_type == "seq2seq":
            B, Lt = mask_txt.shape
            return torch.tril(torch.ones((B, Lt, Lt), dtype=torch.long, device=mask_txt.device))
        return mask_txt

    def forward(self, txt, mask_txt=None, token_type_ids=None, position_ids=None, attn_mask_type="full"):
        f_txt = self.emb_txt(txt, token_type_ids=token_type_ids, position_ids=position_ids)
        if self.txt_trsfr is None:
            return f_txt
        if mask_txt is None:
            mask_txt = torch.ones_like(txt)
        m = self.get_attn_mask(mask_txt, attn_mask_type=attn_mask_type)
        m = self.mask_ext(m, m.shape, m.device).to(dtype=f_txt.dtype)
        return self.txt_trsfr(f_txt, m, output_attentions=False)["last_hidden_state"]


class LAVENDER_Base(nn.Module):
    """model.py:145-473 (the class the north-star calls VIOLET_Base)."""

    def __init__(self, args, tokzr=None):
        super().__init__()
        self.args = args
        self.enc_txt = EncTxt(args)
        self.trsfr, self.config = build_bert_encoder(args.fusion_encoder, args,
                                                     rand_init=getattr(args, "fusion_encoder_rand_init", False))
        self.hidden_size = self.config.hidden_size
        self.mask_ext = extended_attention_mask
        self.enc_img = EncVideo(args, self.hidden_size)
        # args.use_checkpoint (fairscale checkpoint_wrapper + CPU offload, model.py:167-169) exists to fit 16-32 GB
        # GPUs; with 180 GB of HBM3e the activations stay resident and the flag is accepted as a no-op.
        self.tokzr = tokzr
        if tokzr is not None:
            (self.cls_token_id, self.sep_token_id, self.pad_token_id, self.mask_token_id,
             self.unk_token_id) = tokzr.convert_tokens_to_ids(
                [tokzr.cls_token, tokzr.sep_token, tokzr.pad_token, tokzr.mask_token, tokzr.unk_token])
            self.true_token_id = tokzr.convert_tokens_to_ids(["true"])[0]
            self.false_token_id = tokzr.convert_tokens_to_ids(["false"])[0]

    # ---- shared flat parameter arena -----------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)     # .cuda() / .to(): parameters moved -> arena is rebuilt lazily
        set_arena_root(self)
        return out

    def arena(self):
        """The flat fp32 parameter / gradient buffers shared by every native sub-module of this model."""
        if "_lav_root_set" not in self.__dict__:
            set_arena_root(self)
            self.__dict__["_lav_root_set"] = True
        return arena_of(self)

    # ---- features ------------------------------------------------------------------------------------
    def go_feat(self, img, txt, mask, odr=None, vt_mask=None, attn_mask_type="full"):
        feat_img, mask_img = self.enc_img(img, odr, vt_mask)
        feat_txt = self.enc_txt(txt, mask_txt=mask, attn_mask_type=attn_mask_type)
        return feat_img, mask_img, feat_txt, mask

    def get_attn_mask(self, mask_img, mask_txt, attn_mask_type="full", mask_pretxt=None):
        full = mask_img if mask_pretxt is None else torch.cat([mask_img, mask_pretxt], dim=1)
        if attn_mask_type != "seq2seq":
            return torch.cat([full, mask_txt], dim=1)
        # model.py:208-218: every query sees the (masked) video+prefix keys; text queries see text keys causally.
        B, Lfull = full.shape
        Lt = mask_txt.shape[1]
        L = Lfull + Lt
        m = torch.zeros((B, L, L), dtype=torch.long, device=mask_img.device)
        m[:, :, :Lfull] = full.unsqueeze(1)
        m[:, Lfull:, Lfull:] = torch.tril(torch.ones((Lt, Lt), dtype=torch.long, device=mask_img.device))
        return m

    def go_cross(self, feat_img, mask_img, feat_txt, mask_txt, attn_mask_type="full", feat_pretxt=None,
                 mask_pretxt=None):
        if feat_pretxt is not None:
            assert mask_pretxt is None
            feat = torch.cat([feat_img, feat_pretxt, feat_txt], dim=1)
        else:
            feat = torch.cat([feat_img, feat_txt], dim=1)
        if attn_mask_type == "seq2seq":
            # model.py:208-218 without the [B, L, L] tensor (and its per-call CPU build): every query sees
            # the (masked) video + prefix keys, text queries see text keys causally -> key mask + first text position
            full = mask_img if mask_pretxt is None else torch.cat([mask_img, mask_pretxt], dim=1)
            kmask = torch.cat([full, torch.ones_like(mask_txt)], dim=1)
            assert feat.shape[1] == kmask.shape[1], \
                f"mask and feat must have the same length, got {feat.shape[1]} vs. {kmask.shape[1]}"
            out = self.trsfr(feat, kmask, output_attentions=True, causal_from=full.shape[1])
            return out["last_hidden_state"], out["attentions"]
        mask = self.get_attn_mask(mask_img, mask_txt, attn_mask_type=attn_mask_type, mask_pretxt=mask_pretxt)
        assert feat.shape[1] == mask.shape[1], \
            f"mask and feat must have the same length, got {feat.shape[1]} vs. {mask.shape[1]}"
        # config.is_decoder (set by LAVENDER_Captioning(is_decoder=True), model_for_captioning.py:43, on the config object the
        # reference's mask_ext = bert.get_extended_attention_mask reads): HF turns a 2-D padding mask into padding AND a
        # causal mask over the whole sequence -> key mask + causal_from = 0 in the attention kernels.
        causal = 0 if getattr(self.config, "is_decoder", False) else None
        out = self.trsfr(feat, mask, output_attentions=True, causal_from=causal)
        return out["last_hidden_state"], out["attentions"]

    # ---- task token / prompt prefix (model.py:245-306) -------------------------------------------------
    def prepro_pretxt(self, task_or_prompt_txt):
        return task_or_prompt_txt

    def get_pretxt(self, mask_txt, task_name=None, prompt=None):
        batched = mask_txt.dim() > 1
        if self.args.enable_task_token:
            assert task_name is not None and task_name in self.task_tok2id
            feat = self.emb_task[self.task_tok2id[task_name], :].unsqueeze(0)
            msk = torch.ones(1, device=mask_txt.device, dtype=mask_txt.dtype)
            txt = torch.zeros(1, device=mask_txt.device, dtype=mask_txt.dtype)
            if batched:
                B = mask_txt.shape[0]
                txt, msk, feat = txt.unsqueeze(0).expand(B, -1), msk.unsqueeze(0).expand(B, -1), \
                    feat.unsqueeze(0).expand(B, -1, -1)
            return self.prepro_pretxt(txt), msk, feat
        if prompt is not None and self.args.enable_prompt:
            p_txt, p_mask = prompt
            p_batched = p_txt.dim() > 1
            feat = self.enc_txt(p_txt if p_batched else p_txt.unsqueeze(0))
            if batched and not p_batched:
                B = mask_txt.shape[0]
                p_txt, p_mask, feat = p_txt.unsqueeze(0).expand(B, -1), p_mask.unsqueeze(0).expand(B, -1), \
                    feat.expand(B, -1, -1)
            elif not batched and not p_batched:
                feat = feat[0]
            elif batched and p_batched:
                assert mask_txt.shape[0] == p_txt.shape[0]
            else:
                raise ValueError(f"txt dim: {mask_txt.dim()}, prompt_txt dim {p_txt.dim()}")
            return self.prepro_pretxt(p_txt), p_mask, feat
        return None, None, None

    def prepro_txt_inputs(self, txt, mask_txt, feat_txt, task_name=None, prompt=None):
        p_txt, p_mask, p_feat = self.get_pretxt(mask_txt, task_name, prompt)
        if p_txt is not None:
            mask_txt = torch.cat([p_mask, mask_txt], dim=-1)
            txt = torch.cat([p_txt, txt], dim=-1)
            feat_txt = torch.cat([p_feat, feat_txt], dim=-2)
        return txt, mask_txt, feat_txt

    # ---- checkpoints (model.py:352-473) -----------------------------------------------------------------
    def load_ckpt(self, ckpt):
        if ckpt == "":
            print("===== Finished Init LAVENDER  =====")
            return
        if not os.path.exists(ckpt):
            print(f"Try to load pre-trained weights from {ckpt}, but file does not exists...")
            return
        print(f"Loading pre-trained weights from {ckpt}")
        sd = torch.load(ckpt, map_location="cpu")
        name = os.path.splitext(os.path.basename(ckpt))[0]
        if "SwinBERT" in name:
            sd = self.remap_swinbert_keys(sd)
        self.__load_ckpt__(sd)

    def __load_ckpt__(self, loaded):
        """Shape-tolerant load: same-shape keys are copied, everything else is reported (model.py:370-404)."""
        own = self.state_dict()
        toload = {k: v for k, v in loaded.items() if k in own and own[k].shape == v.shape}
        mism = sorted((k, tuple(loaded[k].shape), tuple(own[k].shape)) for k in loaded
                      if k in own and own[k].shape != loaded[k].shape)
        unexpected, missing = sorted(set(loaded) - set(own)), sorted(set(own) - set(loaded))
        for title, lst in (("Unexpected", unexpected), ("Missing", missing), ("Shape Mismatched", mism)):
            if lst:
                print(f"===== {title}: {len(lst)} =====\n\t{lst}")
        self.load_state_dict(toload, strict=not (unexpected or missing or mism))
        # emb_len / emb_pos saved with other maxima (model.py:405-429; the loaded maxima are always read as 6 / 14)
        for key, dim, own_max in (("enc_img.emb_len", 1, self.enc_img.max_size_frame),
                                  ("enc_img.emb_pos", 2, 1 + self.enc_img.max_size_patch ** 2)):
            if key in loaded and loaded[key].shape != own[key].shape:
                n = min(loaded[key].shape[dim], own[key].shape[dim])
                with torch.no_grad():
                    getattr(self.enc_img, key.split(".")[1]).data.narrow(dim, 0, n).copy_(loaded[key].narrow(dim, 0, n))

    @staticmethod
    def remap_swinbert_keys(loaded):
        """Key renames of load_SwinBERT_weight (model.py:431-473)."""
        rules = (("swin.backbone", "enc_img.swin"), ("trans_encoder.bert.encoder", "trsfr"),
                 ("trans_encoder.bert.embeddings", "enc_txt.emb_txt"),
                 ("trans_encoder.bert.img_embedding", "enc_img.img_embedding"))
        out = {}
        for k, v in loaded.items():
            for old, new in rules:
                if old in k:
                    out[k.replace(old, new)] = v
                    break
            else:
                if k.startswith("fc."):
                    out["enc_img." + k] = v
                elif k.startswith("trans_encoder.cls."):
                    out[k.replace("trans_encoder.cls.", "fc_mtm.")] = v
        if "fc_mtm.predictions.bias" in out:
            out["fc_mtm.predictions.decoder.bias"] = out["fc_mtm.predictions.bias"]
        return out
