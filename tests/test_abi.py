"""The C-ABI library loads and exports exactly what include/lavender_b200.h declares (no compute calls: no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "lavender_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(lav_[a-z0-9_]+)\s*\(", src)


def test_header_matches_ctypes_table():
    from lavender_b200 import _lib
    declared = _header_functions()
    assert len(declared) == len(set(declared))
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from lavender_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), name
    l = _lib.lib()
    assert l.lav_abi_version() == 3
    assert l.lav_launch_count() == 0


def test_argument_counts_match_header():
    from lavender_b200 import _lib
    src = open(os.path.join(ROOT, "include", "lavender_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (res, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_no_cpu_fallback():
    """The product path fails loudly instead of computing on the CPU."""
    import pytest
    import torch
    from lavender_b200.video_swin import SwinTransformer3D, SWIN_VARIANTS
    from lavender_b200.bert import BertConfig, BertEncoder
    m = SwinTransformer3D(**SWIN_VARIANTS[("tiny", 224)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 5, 224, 224))
    enc = BertEncoder(BertConfig(num_hidden_layers=1)).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 8, 768), torch.ones(1, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lavender_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "lavender_oracle" not in txt and "import oracle" not in txt and "ref_shims" not in txt, f
