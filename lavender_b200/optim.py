"""Fused optimizer over the flat parameter arena.

`FlatAdamW` is a torch.optim.Optimizer (so torch LR schedulers and `param_groups` logging keep working) whose step()
is three kernels over the arena's flat buffers instead of ~540 per-tensor foreach updates: gradient statistics
(sum of squares + non-finite check), a one-thread "prepare" (unscale factor, clip coefficient, bias corrections,
GradScaler growth / back-off) and the AdamW update.  `DeviceGradScaler` is the matching loss scaler: its state lives
in the same device array, so neither unscale_ nor step needs the host `.item()` syncs of torch.amp.GradScaler
(agent.py:240-248 semantics are kept: skip the step on inf/nan, halve the scale, double it every 2000 good steps).
"""
import torch

from . import ops


class DeviceGradScaler:
    def __init__(self, device, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.state = torch.zeros(16, dtype=torch.float32, device=device)
        self.state[0] = init_scale
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval

    def scale(self, loss):
        return loss * self.state[0]

    def get_scale(self):
        return float(self.state[0].item())

    # torch.amp.GradScaler API used by agent.backward_step; everything happens inside FlatAdamW.step
    def unscale_(self, optimizer):
        pass

    def step(self, optimizer):
        return optimizer.step()

    def update(self):
        pass


class FlatAdamW(torch.optim.Optimizer):
    """AdamW(betas, eps, per-group lr / weight_decay) + clip_grad_norm_(max_grad_norm) + loss-scale handling on a
    ParamArena.  `params` are the usual param-group dicts; every parameter must live in `arena`."""

    def __init__(self, params, arena, scaler, lr=2e-5, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-3, max_grad_norm=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        assert len(self.param_groups) < 255
        self.arena, self.scaler, self.max_grad_norm = arena, scaler, float(max_grad_norm)
        dev = arena.flat.device
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self._group_cpu = torch.full((arena.total // 8,), 255, dtype=torch.uint8)
        self._param_group = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if id(p) not in arena.offsets:
                    raise ValueError("FlatAdamW: parameter is not part of the arena")
                self._param_group[id(p)] = gi
        self._active = None
        self.group_of_block = torch.empty(arena.total // 8, dtype=torch.uint8, device=dev)
        n = len(self.param_groups)
        # ring of pinned staging buffers: the host may run several steps ahead of the device (pipelined input path)
        self._hyper_ring = [torch.zeros(2, n, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(2, n)
                            for _ in range(4)]
        self._hyper_i = 0
        self.hyper = torch.zeros(2, n, dtype=torch.float32, device=dev)

    def _refresh_active(self):
        """Blocks of parameters whose .grad is None are skipped (torch semantics), e.g. emb_odr / unused heads."""
        ar = self.arena
        active = tuple(p.grad is not None for p in ar.params)
        if active != self._active:
            self._group_cpu.fill_(255)
            for p, a in zip(ar.params, active):
                if a and id(p) in self._param_group:
                    o = ar.offsets[id(p)] // 8
                    self._group_cpu[o:o + (p.numel() + 7) // 8] = self._param_group[id(p)]
            self.group_of_block.copy_(self._group_cpu)
            self._active = active

    @torch.no_grad()
    def step(self, closure=None):
        ar = self.arena
        if not ar.valid():
            raise RuntimeError("FlatAdamW: the parameter arena was rebuilt (model moved?) - recreate the optimizer")
        ar.finalize_grads()
        self._refresh_active()
        host = self._hyper_ring[self._hyper_i % len(self._hyper_ring)]
        self._hyper_i += 1
        for gi, g in enumerate(self.param_groups):
            host[0, gi] = g["lr"]
            host[1, gi] = g["weight_decay"]
        self.hyper.copy_(host, non_blocking=True)
        b1, b2 = self.param_groups[0]["betas"]
        sc = self.scaler
        fresh16 = ar._ver16 is not None and ar._ver16 == ar._version()
        ops.grad_stats(ar.grad, sc.state)
        ops.adamw_step(ar.flat, ar.grad, self.exp_avg, self.exp_avg_sq, self.group_of_block, self.hyper[0], self.hyper[1],
                       sc.state, beta1=b1, beta2=b2, eps=self.param_groups[0]["eps"], max_grad_norm=self.max_grad_norm,
                       growth_factor=sc.growth_factor, backoff_factor=sc.backoff_factor,
                       growth_interval=sc.growth_interval, param16=ar.flat16)
        # the kernel rewrote the fp16 shadow of every element it updated: the shadow stays valid if it was valid
        # before (parameters without a gradient are untouched on both sides); otherwise the next forward re-casts
        if not fresh16:
            ar._ver16 = None
        return None

    def grad_norm(self):
        """Unscaled gradient norm of the last step (device scalar)."""
        return self.scaler.state[5]
