"""CUDA-graph capture of one training step's device work (forward of MLM + VTM, both cross-entropies, backward).

The eager step issues ~2000 kernels from Python (one ctypes call per C-ABI kernel plus torch glue) and is bound by
the host (~60 ms/step on the B200 box) rather than by the ~55 ms of kernels.  The hot path has static shapes
(fixed batch / frames / caption length), so the whole launch sequence is captured once per input signature and
replayed with one `cudaGraphLaunch`:

  * inputs live in static device buffers (`img, txt, mask, ans_mtm` and the VTM pair indices, which the reference
    draws from the numpy RNG on the host: main_pretrain_mlm.py:90) that are refreshed by async copies;
  * the graph starts by zeroing the flat gradient arena, so `p.grad` stays bound to the arena views between steps
    (no zero_grad(set_to_none) in this mode);
  * DropPath draws come from torch's graph-safe Philox generator (fresh numbers on every replay); BERT dropout masks
    are a function of a device-resident step counter that the graph increments first (dropout.py);
  * what stays eager after the replay: gradient all-reduce, unscale / clip / AdamW / LR schedule (agent.py).
TMA descriptors are kernel parameters encoded on the host at capture time; they stay valid because every buffer the
kernels touch belongs to the graph's private memory pool or to the model / arena.
"""
from collections import defaultdict

import numpy as np
import torch

from .dropout import rng_for

_STATIC_KEYS = ("img", "txt", "mask", "ans_mtm", "vt_mask", "mtm_rows")


def batch_signature(batch):
    return tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in _STATIC_KEYS
                 if batch.get(k) is not None and isinstance(batch[k], torch.Tensor))


class GraphedPretrainStep:
    """Captures `agent`'s forward + loss + scaled backward for batches with the signature of `example` (a dict of
    CUDA tensors as produced by Agent.prepare_batch)."""

    def __init__(self, agent, example, warmup=2):
        self.agent = agent
        model = agent.model
        dev = example["img"].device
        self.sig = batch_signature(example)
        self.static = {k: example[k].clone() for k in _STATIC_KEYS
                       if example.get(k) is not None and isinstance(example[k], torch.Tensor)}
        self.B = example["img"].shape[0]
        self.O = min(self.B, model.vtm_batch)
        vi, ti, lab = model.build_vtm_pairs(self.B, self.O, device=dev)
        self.static.update(vtm_vid_idx=vi, vtm_txt_idx=ti, vtm_labels=lab)
        # pinned staging for the per-step VTM text indices: a ring, because the pipelined input path lets the host run
        # up to 3 steps ahead of the device (a single buffer would be refilled before its async copy has executed)
        from .optim import PinnedRing   # each slot is guarded by an event: never refilled under a pending copy
        self._txt_idx_ring = PinnedRing(lambda: torch.empty(ti.shape, dtype=ti.dtype).pin_memory())
        model.train()
        ar = model.arena()
        # Data parallel: the gradient all-reduce is captured INSIDE the step's graph (LAV_GRAPH_NCCL=0: after the replay,
        # un-overlapped, as in round 1).  GradSync's pieces run on a side stream forked from the capture stream when the
        # Swin backward starts / enters stage 1, i.e. as parallel branches of the graph, joined by finish() below.
        import os as _os
        self.sync_in_graph = agent.grad_sync is not None and _os.environ.get("LAV_GRAPH_NCCL", "1") != "0"
        hooks = (ar.on_swin_backward, ar.on_swin_stage)
        if not self.sync_in_graph:
            ar.on_swin_backward = ar.on_swin_stage = None
        # eager warm-up on a side stream (lazy kernel attribute setup, index-map caches, allocator)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._fwd_bwd()
                agent.optzr.zero_grad(set_to_none=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        agent.optzr.zero_grad(set_to_none=True)   # first prepare_grads inside the capture records one arena memset
        self._shadow_by_optimizer = bool(getattr(agent, "fused", False))
        if self._shadow_by_optimizer:
            ar.refresh16()                         # FlatAdamW keeps the fp16 shadow current: no cast inside the graph
        else:
            ar._ver16 = None                       # ... and the first refresh16 records the fp32 -> fp16 weight cast
        # the critical path is captured on a high-priority stream: the side-stream branches (weight gradients, default
        # priority) then only get SMs the critical-path kernels leave idle
        import os
        cap = torch.cuda.Stream(priority=-1) if os.environ.get("LAV_GRAPH_PRIORITY", "1") != "0" else None
        with torch.cuda.graph(self.graph, stream=cap):
            self.l_mtm, self.l_vtm = self._fwd_bwd()
        self.native_launches = _lib.launch_count() - n0   # kernels of the C-ABI library inside one replay
        ar.on_swin_backward, ar.on_swin_stage = hooks

    def _fwd_bwd(self):
        ag = self.agent
        rng_for(self.static["img"].device).advance()   # captured: every replay draws new BERT dropout masks
        out = ag.forward_step(dict(self.static))
        l1 = ag.loss_func(out["out_mtm"].flatten(0, out["out_mtm"].dim() - 2), out["ans_mtm"].flatten())
        l2 = ag.loss_func(out["out_vtm"].flatten(0, out["out_vtm"].dim() - 2), out["ans_vtm"].flatten())
        ag.scaler.scale(l1 + l2).backward()
        if self.sync_in_graph:
            ag.grad_sync.finish()          # joins the overlapped pieces, reduces the rest: all inside the capture
        else:
            ag.model.arena().finalize_grads()
        return l1.detach(), l2.detach()

    def load(self, batch):
        """Async copies of the step's inputs into the static buffers (+ fresh VTM negatives from the numpy RNG)."""
        for k in _STATIC_KEYS:
            if k in self.static:
                self.static[k].copy_(batch[k], non_blocking=True)
        negs = self.agent.model.draw_negatives(self.B, self.O)
        host = self._txt_idx_ring.next()
        ti = host.view(self.B, self.O)
        for i in range(self.B):
            ti[i, 0] = i
            for j in range(self.O - 1):
                ti[i, 1 + j] = int(negs[i][j])
        self.static["vtm_txt_idx"].copy_(host, non_blocking=True)
        self._txt_idx_ring.copied()

    def __call__(self, batch):
        self.load(batch)
        if self._shadow_by_optimizer:
            self.agent.model.arena().refresh16()   # no-op unless the weights were changed outside the optimizer
        self.graph.replay()
        return self.l_mtm, self.l_vtm


class GraphCache:
    """One captured step per batch signature (Agent keeps this; unseen signatures are captured on first use)."""

    def __init__(self, agent):
        self.agent, self.graphs = agent, {}

    def get(self, batch):
        sig = batch_signature(batch)
        g = self.graphs.get(sig)
        if g is None:
            g = self.graphs[sig] = GraphedPretrainStep(self.agent, batch)
        return g
