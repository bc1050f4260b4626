"""Operand precision of the native path.

default  fp16 tensor-core operands, fp32 accumulation / statistics / residual stream — what the reference's own GPU
         path computes (fp16 autocast, agent.py:219; DeepSpeed fp16, utils/deepspeed.py:21-24).  Against the fp32 CPU
         reference the logits differ by 1-2e-3 max-abs, which is exactly the error of rounding the GEMM operands to
         fp16 (profiles/PARITY.md: the oracle with fp16-rounded operands shows the same distance).
high     the stated PARITY MODE (`LAV_PRECISION=high`, or `precision.set_high(True)`): every nn.Linear on the path runs
         as a split-fp16 product — x = hi + lo, w = hi + lo, one tcgen05 GEMM over K' = 3K computing
         xh*wh + xl*wh + xh*wl with fp32 accumulation (operand error ~2^-22 instead of 2^-11) — biases are added in
         fp32, GELU / LayerNorm / softmax statistics in fp32, and the attention output O is handed to the projection
         un-rounded (fp32).  Q/K/V and the softmax probabilities stay fp16 tensor-core operands (their contribution is
         < 1.5e-4 on the logits, PARITY.md).  Forward values then agree with the fp32 reference to <= 1e-3 max-abs on
         the logits (north-star tolerance); backward uses the same fp16 kernels as the default mode on fp16 casts of
         the saved activations.  ~3x the GEMM work: a validation mode, not the benchmarked one.
"""
import os

SCOPES = ("swin", "fc", "bert", "head")   # Video Swin | EncVideo.fc (+ stand-alone Linear) | fusion BERT | MLM head


def _parse(v):
    v = v.strip().lower()
    if v in ("high", "1", "true", "all"):
        return frozenset(SCOPES)
    if v in ("", "0", "default", "false", "fp16"):
        return frozenset()
    got = frozenset(x for x in v.replace("high:", "").split(",") if x)
    bad = got - set(SCOPES)
    if bad:
        raise ValueError(f"LAV_PRECISION: unknown scope(s) {sorted(bad)}; use 'high' or a list of {SCOPES}")
    return got


_HIGH = _parse("high" if os.environ.get("LAV_PARITY", "0") == "1" else os.environ.get("LAV_PRECISION", ""))


def high(scope=None):
    """True when the high-precision mode is on (for `scope`, one of SCOPES; any scope when None)."""
    return bool(_HIGH) if scope is None else scope in _HIGH


def set_high(flag=True):
    """flag: bool, or an iterable of SCOPES (per-module bisection: profiles/PARITY.md).  Returns the previous setting."""
    global _HIGH
    prev = _HIGH
    if isinstance(flag, (bool, int)) or flag is None:
        _HIGH = frozenset(SCOPES) if flag else frozenset()
    elif isinstance(flag, str):
        _HIGH = _parse(flag)
    else:
        _HIGH = _parse(",".join(flag))
    return prev


class high_precision:
    """with precision.high_precision(): ... — scoped switch (tests / smoke)."""

    def __init__(self, flag=True):
        self.flag = flag

    def __enter__(self):
        self.prev = set_high(self.flag)
        return self

    def __exit__(self, *exc):
        set_high(self.prev)
        return False
