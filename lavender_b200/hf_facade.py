"""`transformers` as the reference's scripts see it through utils/lib.py.

The unchanged scripts call (model.py:100,152-165; main_pretrain_mlm.py:46-48; agent.py:80):
    transformers.AutoModel.from_pretrained(name).embeddings
    transformers.AutoModelForMaskedLM.from_pretrained(name)  ->  .bert.encoder / .cls / .config / .get_extended_attention_mask
    transformers.AutoConfig.from_pretrained(name) ; AutoModelForMaskedLM.from_config(config)
    transformers.AutoTokenizer.from_pretrained(name)
This facade answers those with lavender_b200's native modules (HF-identical parameter names; weights from a local HF
directory when present) and forwards every other attribute to the real package when it is installed."""
import os
import types

from .model import (_bert_config_for, build_bert_embeddings, build_bert_encoder, build_mlm_head,
                    extended_attention_mask)
from .pretrain import FakeTokenizer

try:
    import transformers as _real
except ImportError:  # pragma: no cover
    _real = None


class _MaskedLM:
    def __init__(self, name, config=None, rand_init=False):
        self.bert = types.SimpleNamespace()
        self.bert.encoder, self.config = build_bert_encoder(name, rand_init=rand_init)
        self.bert.embeddings, _ = build_bert_embeddings(name)
        self.cls, _ = build_mlm_head(name)
        self.get_extended_attention_mask = extended_attention_mask


class _Model:
    def __init__(self, name):
        self.embeddings, self.config = build_bert_embeddings(name)
        self.encoder, _ = build_bert_encoder(name)
        self.get_extended_attention_mask = extended_attention_mask


class AutoModelForMaskedLM:
    @staticmethod
    def from_pretrained(name, *a, **k):
        return _MaskedLM(name)

    @staticmethod
    def from_config(config):
        return _MaskedLM(getattr(config, "_name_or_path", None), rand_init=True)


class AutoModel:
    @staticmethod
    def from_pretrained(name, *a, **k):
        return _Model(name)


class AutoConfig:
    @staticmethod
    def from_pretrained(name, *a, **k):
        c = _bert_config_for(name)
        c._name_or_path = name
        return c


class AutoTokenizer:
    @staticmethod
    def from_pretrained(name, *a, **k):
        if _real is not None and isinstance(name, str) and os.path.isdir(name):
            return _real.AutoTokenizer.from_pretrained(name, *a, **k)
        print(f"[lavender_b200] no local tokenizer files for {name!r}: using bert-base-uncased special-token ids")
        return FakeTokenizer()


class _Facade(types.ModuleType):
    AutoModelForMaskedLM, AutoModel, AutoConfig, AutoTokenizer = AutoModelForMaskedLM, AutoModel, AutoConfig, AutoTokenizer
    RobertaForMaskedLM = type("RobertaForMaskedLM", (), {})   # isinstance() probe at model.py:160

    def __getattr__(self, name):
        if _real is None:
            raise AttributeError(name)
        return getattr(_real, name)


transformers = _Facade("transformers")
