"""Dropout state of the native BERT modules (HF hidden / attention-probability dropout in train() mode).

The kernels derive every mask bit from (seed, step, site, element index) with a counter-based generator
(csrc/rng.cuh), so nothing mask-shaped is stored for backward.  This module owns the per-device `[seed, step]`
tensor the kernels read and hands out site ids:

  * eager mode: every dropout call takes a fresh site id (a Python counter), so masks never repeat;
  * CUDA-graph mode: site ids are frozen into the captured launches; `advance()` (an in-place add on the device
    tensor, captured at the top of the graph) makes every replay draw new masks.
The seed comes from torch's default generator, so `torch.manual_seed` makes runs reproducible.
"""
import torch

_STATES = {}


class DropoutRNG:
    def __init__(self, device):
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self._site = 0

    def site(self):
        self._site = (self._site + 1) & 0x7FFFFFFF
        return self._site

    def advance(self):
        """New masks for the same site ids (graph-capturable)."""
        self.state[1:2].add_(1)

    def spec(self, p):
        """(rng tensor, fresh site id, p) as the ops.* wrappers take it, or None when p == 0."""
        return (self.state, self.site(), float(p)) if p > 0.0 else None


def rng_for(device):
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _STATES.get(key)
    if st is None:
        st = _STATES[key] = DropoutRNG(device)
    return st


def reseed(device, seed):
    st = rng_for(device)
    st.state[0] = int(seed)
    st.state[1] = 0
    return st
