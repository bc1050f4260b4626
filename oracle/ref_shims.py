"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (microsoft/LAVENDER) from
/root/reference under a set of stub modules so that its own `LAVENDER_Pretrain_MLM` can run on CPU
in the build container.  Used only by `oracle/make_golden.py` (golden-vector generation) and by
`bench.py --impl reference` when the reference tree is present.  `/root/reference` does not exist on
the GPU box; nothing in `-m gpu` tests / smoke / the product path imports this file.

Shim list follows SURVEY.md §8c (1-7):
  * stub modules for packages missing in this image (easydict, skimage, fairscale, toolz, tensorboardX,
    deepspeed, addict, yapf, dataset) — the reference star-imports them in utils/lib.py:5-23
  * HF `from_pretrained` -> random-init `BertConfig` models (no weights offline; model.py:100-102,152-153,
    main_pretrain_mlm.py:46-47)
  * transformers>=5 drift: 2-arg `get_extended_attention_mask` (model.py:136,239) and encoder output dict
    with an 'attentions' key (model.py:242-243)
  * no-GPU box: `.cuda()` no-ops (model.py:59,87; agent.py:72,150)
Nothing under /root/reference is modified or copied.
"""
import os
import sys
import types
import inspect

import torch
import transformers

REF_ROOT = os.environ.get("LAVENDER_REF_ROOT", "/root/reference")


class EDict(dict):
    """easydict.EasyDict stand-in: attribute access that raises AttributeError on a missing key
    (model.py:11-13 relies on getattr(args, 'swinbert', False))."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


class FakeTokenizer:
    """bert-base-uncased special ids; 'true'/'false' ids are arbitrary distinct ids."""
    cls_token, sep_token, pad_token, mask_token, unk_token = "[CLS]", "[SEP]", "[PAD]", "[MASK]", "[UNK]"
    V = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103, "true": 2995, "false": 6270}

    def convert_tokens_to_ids(self, toks):
        return [self.V[t] for t in toks]


SWIN_SIZES = {  # visbackbone/swin_tiny.py:4-17, swin_base.py:3-5, swin_large.py:3-5
    "tiny": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]),
    "base": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32]),
    "large": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48]),
}

_installed = False


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(num_bert_layers=2):
    """Install the stubs and put the reference on sys.path. Idempotent."""
    global _installed
    os.environ["LAV_NUM_BERT_LAYERS"] = str(num_bert_layers)
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub("easydict", EasyDict=EDict)
    _stub("skimage")
    _stub("skimage.feature", hog=lambda *a, **k: None)
    _stub("fairscale")
    _stub("fairscale.nn")
    _stub("fairscale.nn.misc", checkpoint_wrapper=lambda m, **k: m)
    _stub("toolz")
    _stub("toolz.sandbox", unzip=lambda seq: zip(*seq))
    _stub("tensorboardX", SummaryWriter=object)
    _stub("deepspeed")

    class _DS(torch.utils.data.Dataset):
        def __init__(self, *a, **k):
            pass

    _stub("dataset", Dataset_Base=_DS, get_dl=lambda *a, **k: None, move_to_cuda=lambda b: b,
          get_tsv_dls=None, MetaLoader=object)
    _stub("addict", Dict=type("Dict", (dict,), {}))
    _stub("yapf")
    _stub("yapf.yapflib")
    _stub("yapf.yapflib.yapf_api", FormatCode=lambda *a, **k: ("", False))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec

    def bert_cfg():
        return transformers.BertConfig(num_hidden_layers=int(os.environ["LAV_NUM_BERT_LAYERS"]),
                                       attn_implementation="eager")

    transformers.AutoModel.from_pretrained = classmethod(
        lambda cls, *a, **k: transformers.BertModel(bert_cfg()))
    transformers.AutoModelForMaskedLM.from_pretrained = classmethod(
        lambda cls, *a, **k: transformers.BertForMaskedLM(bert_cfg()))
    sys.path.insert(0, REF_ROOT)
    _installed = True


def default_args(size="tiny", size_img=224, size_batch=2, enable_task_token=True):
    return EDict(vis_backbone_size=size, size_img=size_img, txt_backbone="bert-base-uncased",
                 tokenizer="bert-base-uncased", fusion_encoder="bert-base-uncased",
                 fusion_encoder_rand_init=False, txt_backbone_embed_only=True, use_checkpoint=False,
                 size_patch=32, size_batch=size_batch, enable_task_token=enable_task_token,
                 enable_prompt=False, vis_backbone_init="random", kinetics=600)


def build_reference_model(size="tiny", num_bert_layers=2, size_img=224, size_batch=2,
                          enable_task_token=True):
    """Returns the reference's own LAVENDER_Pretrain_MLM (main_pretrain_mlm.py:42-119), random init."""
    install(num_bert_layers)
    os.environ["LAV_NUM_BERT_LAYERS"] = str(num_bert_layers)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)  # config paths are CWD-relative (video_swin.py:574-593)
    try:
        from main_pretrain_mlm import LAVENDER_Pretrain_MLM
        import visbackbone.video_swin as vs
        import model as ref_model

        window = (8, 12, 12) if size_img == 384 else (8, 7, 7)

        def get_vidswin_model(args):  # bypasses only the mmcv Config loader (needs addict/yapf)
            m = vs.SwinTransformer3D(pretrained=None, patch_size=(2, 4, 4), window_size=window,
                                     drop_path_rate=0.2, patch_norm=True, **SWIN_SIZES[args.vis_backbone_size])
            m.init_weights()
            return m

        ref_model.get_vidswin_model = get_vidswin_model
        args = default_args(size, size_img, size_batch, enable_task_token)
        m = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
    finally:
        os.chdir(cwd)
    _ext = m.mask_ext
    m.mask_ext = lambda mask, shape, device=None: _ext(mask, shape)
    _fw = m.trsfr.forward

    def _enc_forward(feat, mask, output_attentions=False, **kw):
        o = _fw(feat, mask, **kw)
        return {"last_hidden_state": o[0] if isinstance(o, tuple) else o.last_hidden_state,
                "attentions": None}

    m.trsfr.forward = _enc_forward
    return m


def install_multitask():
    """Extra shims of SURVEY §8c-8 for main_multi_task_mlm.py: absent `evalcap`, the TSV writer helpers (their real
    module drags in utils/qd_common.py), `dataset.TsvCompositeDataset`, and the alias for the reference's own broken
    import (main_multi_task_mlm.py:13 asks for Dataset_QAOE_LSMDC_TSV, main_qaoe_mlm_lsmdc_fib.py:12 defines
    Dataset_QAOE_MLM_LSMDC)."""
    install(int(os.environ.get("LAV_NUM_BERT_LAYERS", "2")))
    if "evalcap" not in sys.modules:
        _stub("evalcap")
        _stub("evalcap.utils_caption_evaluate", evaluate_on_coco_caption=lambda *a, **k: {})
        _stub("utils.tsv_file_ops", tsv_writer=lambda *a, **k: None, reorder_tsv_keys=lambda *a, **k: None)
        sys.modules["dataset"].TsvCompositeDataset = object
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        import main_qaoe_mlm_lsmdc_fib as fib
        if not hasattr(fib, "Dataset_QAOE_LSMDC_TSV"):
            fib.Dataset_QAOE_LSMDC_TSV = fib.Dataset_QAOE_MLM_LSMDC
        import main_multi_task_mlm  # noqa: F401
    finally:
        os.chdir(cwd)


def build_reference_multitask(size="tiny", num_bert_layers=2, size_img=224, size_batch=2, enable_task_token=True,
                              is_decoder=False):
    """The reference's own LAVENDER_Multi_Task (main_multi_task_mlm.py:77-225), random init, constructed as the script
    does (main_multi_task_mlm.py:502-504: is_decoder = getattr(args, 'is_decoder', False) -> False; the constructor's own
    default True would make HF's get_extended_attention_mask turn every 2-D "full" mask into a CAUSAL one)."""
    install(num_bert_layers)
    os.environ["LAV_NUM_BERT_LAYERS"] = str(num_bert_layers)
    install_multitask()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        import main_multi_task_mlm as mt
        import visbackbone.video_swin as vs
        import model as ref_model
        window = (8, 12, 12) if size_img == 384 else (8, 7, 7)

        def get_vidswin_model(args):
            m = vs.SwinTransformer3D(pretrained=None, patch_size=(2, 4, 4), window_size=window,
                                     drop_path_rate=0.2, patch_norm=True, **SWIN_SIZES[args.vis_backbone_size])
            m.init_weights()
            return m

        ref_model.get_vidswin_model = get_vidswin_model
        args = default_args(size, size_img, size_batch, enable_task_token)
        m = mt.LAVENDER_Multi_Task(args, FakeTokenizer(), is_decoder=is_decoder)
    finally:
        os.chdir(cwd)
    _ext = m.mask_ext
    m.mask_ext = lambda mask, shape, device=None: _ext(mask, shape)
    _fw = m.trsfr.forward

    def _enc_forward(feat, mask, output_attentions=False, **kw):
        o = _fw(feat, mask, **kw)
        return {"last_hidden_state": o[0] if isinstance(o, tuple) else o.last_hidden_state, "attentions": None}

    m.trsfr.forward = _enc_forward
    return m
