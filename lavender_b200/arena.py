"""Flat parameter / gradient arena.

All parameters of a model live in ONE fp32 buffer (nn.Parameter.data are views into it), with
  * an fp16 shadow copy of the whole buffer refreshed by a single cast kernel whenever a parameter
    changed (the reference gets its fp16 weights from autocast / DeepSpeed-fp16: agent.py:219,
    utils/deepspeed.py:21-24), and
  * a flat fp32 gradient buffer whose slices are the parameters' `.grad` — the wgrad kernels accumulate
    straight into it, and data-parallel training all-reduces it in one NCCL call (agent.py:252-265's
    DDP / DeepSpeed replaced by lavender_b200.dist.allreduce_gradients).
BERT's query/key/value weights (and biases) are placed adjacently so one [2304,768] GEMM serves all three.
"""
import weakref

import torch

from . import ops

_ALIGN = 8  # elements: 16 B for the fp16 shadow (TMA base alignment), 32 B for fp32


def _ordered_params(root):
    """named_parameters order, except that each BERT self-attention block is emitted as
    q.w, k.w, v.w, q.b, k.b, v.b so the fused QKV views are contiguous."""
    named = list(root.named_parameters())
    byname = dict(named)
    out, done = [], set()
    for n, p in named:
        if id(p) in done:
            continue
        if n.endswith("attention.self.query.weight"):
            pre = n[: -len("query.weight")]
            grp = [pre + s for s in ("query.weight", "key.weight", "value.weight", "query.bias", "key.bias", "value.bias")]
            if all(g in byname for g in grp):
                for g in grp:
                    out.append((g, byname[g]))
                    done.add(id(byname[g]))
                continue
        out.append((n, p))
        done.add(id(p))
    return out


class ParamArena:
    def __init__(self, root):
        named = _ordered_params(root)
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params), \
            "lavender_b200 keeps fp32 master parameters on one device"
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat16 = torch.empty(off, dtype=torch.float16, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in self.params:
                o = self.offsets[id(p)]
                self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
        self._ver16 = None
        self._x3, self._ver_x3 = {}, None   # high-precision mode: split-fp16 [hi | hi | lo] copies of the weights
        self._grad_views = {}
        self._clean = set()
        self.on_swin_backward = None  # set by dist.GradSync: called when the video encoder's backward begins
        self.on_swin_stage = None     # ... and when that backward enters stage s (3, 2, 1, 0)

    # ---- validity / shadow ------------------------------------------------------------------------
    def valid(self):
        base = self.flat.data_ptr()
        return all(p.data_ptr() == base + 4 * self.offsets[id(p)] for p in self.params)

    def _version(self):
        return sum(p._version for p in self.params)

    def refresh16(self):
        v = self._version()
        if v != self._ver16:
            ops.cast_f16(self.flat, self.flat16)
            self._ver16 = v

    # ---- views ------------------------------------------------------------------------------------
    def w16(self, p):
        o = self.offsets[id(p)]
        return self.flat16[o:o + p.numel()].view(p.shape)

    def w16x3(self, w32):
        """High-precision mode (precision.py): the [N, 3K] fp16 split operand [hi | hi | lo] of the fp32 master weight view
        `w32` ([N, K], a view of `flat`), cached until a parameter changes."""
        v = self._version()
        if v != self._ver_x3:
            self._x3.clear()
            self._ver_x3 = v
        key = (w32.data_ptr(), tuple(w32.shape))
        t = self._x3.get(key)
        if t is None:
            N, K = w32.shape
            t = torch.empty(N, 3 * K, dtype=torch.float16, device=w32.device)
            ops.split3(w32, t, rows=N, C=K, weight=True)
            self._x3[key] = t
        return t

    def span16(self, first, last, shape):
        """fp16 view covering `first`..`last` (adjacent in the arena), e.g. fused [Wq;Wk;Wv]."""
        o0, o1 = self.offsets[id(first)], self.offsets[id(last)] + last.numel()
        t = self.flat16[o0:o1]
        assert t.numel() == int(torch.Size(shape).numel()), "parameters are not adjacent in the arena"
        return t.view(shape)

    def span32(self, first, last, shape, grad=False):
        o0, o1 = self.offsets[id(first)], self.offsets[id(last)] + last.numel()
        t = (self.grad if grad else self.flat)[o0:o1]
        assert t.numel() == int(torch.Size(shape).numel()), "parameters are not adjacent in the arena"
        return t.view(shape)

    def g(self, p):
        """fp32 gradient slice of p (the tensor the kernels accumulate into)."""
        v = self._grad_views.get(id(p))
        if v is None:
            o = self.offsets[id(p)]
            v = self.grad[o:o + p.numel()].view(p.shape)
            self._grad_views[id(p)] = v
        return v

    # ---- gradients --------------------------------------------------------------------------------
    def prepare_grads(self, params):
        """Make `p.grad` of every p in `params` the arena slice (zeroed when it had no gradient yet), so that
        backward kernels can accumulate in place with the usual `.grad +=` semantics."""
        if all(p.grad is None for p in self.params):
            self.grad.zero_()  # one memset instead of one per tensor
            self._clean = {id(p) for p in self.params}
        for p in params:
            gv = self.g(p)
            if p.grad is None:
                if id(p) in self._clean:
                    self._clean.discard(id(p))
                else:
                    gv.zero_()
                p.grad = gv
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
                p.grad = gv

    def finalize_grads(self):
        """After backward: gradients that torch autograd produced for glue ops (e.g. `emb_task` indexing) are moved
        into the arena so that the flat buffer holds every gradient (one all-reduce, one clip, one optimizer pass).
        Also the join point of the weight-gradient side stream (streams.py)."""
        from . import streams
        streams.join(self.grad.device)
        for p in self.params:
            if p.grad is not None:
                gv = self.g(p)
                if p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
                    p.grad = gv

    def zero_grads(self, set_to_none=True):
        for p in self.params:
            p.grad = None
        if not set_to_none:
            self.grad.zero_()


def arena_of(module):
    """The arena that owns `module`'s parameters: the root model's (set by LAVENDER_Base) or its own."""
    ref = module.__dict__.get("_lav_root")
    root = ref() if ref is not None else None
    if root is None:
        root = module
    a = root.__dict__.get("_lav_arena")
    if a is None or not a.valid():
        a = ParamArena(root)
        root.__dict__["_lav_arena"] = a
    return a


def set_arena_root(root):
    """Call on the top-level model after construction: every native sub-module then shares root's arena."""
    for m in root.modules():
        if m is not root:
            m.__dict__["_lav_root"] = weakref.ref(root)
    root.__dict__.pop("_lav_arena", None)
