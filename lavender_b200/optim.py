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


class PinnedRing:
    """Ring of pinned host staging buffers for async H2D copies issued while the host runs ahead of the device (CUDA-graph
    replay enqueues a step in a few ms against ~25 ms on the device).  Each slot carries a CUDA event recorded after its
    copy; `next()` waits for that event before handing the slot out again, so a buffer is never refilled while a copy
    that reads it is still pending - however far ahead the host runs."""

    def __init__(self, make, n=4, cuda=True):
        self.slots = [make() for _ in range(n)]
        self.events = [torch.cuda.Event() if cuda else None for _ in range(n)]
        self.used = [False] * n
        self.i = 0

    def next(self):
        k = self.i % len(self.slots)
        self.i += 1
        if self.used[k] and self.events[k] is not None:
            self.events[k].synchronize()
        self._k = k
        return self.slots[k]

    def copied(self, stream=None):
        """Call right after enqueueing the async copy out of the slot returned by the last next()."""
        k = self._k
        if self.events[k] is not None:
            self.events[k].record(stream if stream is not None else torch.cuda.current_stream())
            self.used[k] = True


class FlatAdamW(torch.optim.Optimizer):
    """AdamW(betas, eps, per-group lr / weight_decay) + clip_grad_norm_(max_grad_norm) + loss-scale handling on a
    ParamArena.  `params` are the usual param-group dicts; every parameter must live in `arena`.

    torch.optim.AdamW semantics kept: a parameter whose .grad is None is skipped, and its step count (bias correction)
    only advances on steps where it has a gradient — parameters that receive their first gradient late (emb_odr, task
    heads) form a "cohort" of their param group with its own first-step offset (kernel groups = param group x cohort)."""

    MAX_KERNEL_GROUPS = 64

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
        self.steps_done = 0                 # step() calls (skipped inf/nan steps included)
        self._cohort = {}                   # id(p) -> cohort (how many times the active set had grown when p joined it)
        self._ncohorts = 0
        self._kgroups = []                  # kernel group id -> (param group index, cohort)
        n = self.MAX_KERNEL_GROUPS
        cuda = dev.type == "cuda"
        self._hyper_ring = PinnedRing(lambda: torch.zeros(2, n, dtype=torch.float32).pin_memory() if cuda
                                      else torch.zeros(2, n), cuda=cuda)
        self.hyper = torch.zeros(2, n, dtype=torch.float32, device=dev)   # rows: lr, weight decay per kernel group
        # per kernel group: non-skipped steps taken before its first update (-1 = not started; set by the kernel)
        self.group_step0 = torch.full((n,), -1.0, dtype=torch.float32, device=dev)
        self.group_bc = torch.zeros(2 * n, dtype=torch.float32, device=dev)
        # scratch of the deterministic gradient-norm reduction (per-block partials + ticket counter; zeroed once)
        self.stats_ws = torch.zeros(8 * 160 + 1, dtype=torch.float32, device=dev) if cuda else None

    def _refresh_active(self):
        """Blocks of parameters whose .grad is None are skipped (torch semantics), e.g. emb_odr / unused heads."""
        ar = self.arena
        active = tuple(p.grad is not None for p in ar.params)
        if active != self._active:
            self._group_cpu.fill_(255)
            if any(a and id(p) in self._param_group and id(p) not in self._cohort for p, a in zip(ar.params, active)):
                self._ncohorts += 1
            for p, a in zip(ar.params, active):
                if a and id(p) in self._param_group:
                    first = self._cohort.setdefault(id(p), self._ncohorts - 1)
                    key = (self._param_group[id(p)], first)
                    if key not in self._kgroups:
                        if len(self._kgroups) >= self.MAX_KERNEL_GROUPS:
                            raise RuntimeError("FlatAdamW: too many (param group, first-step) cohorts")
                        self._kgroups.append(key)
                    o = ar.offsets[id(p)] // 8
                    self._group_cpu[o:o + (p.numel() + 7) // 8] = self._kgroups.index(key)
            self.group_of_block.copy_(self._group_cpu)
            self._active = active

    @torch.no_grad()
    def step(self, closure=None):
        ar = self.arena
        if not ar.valid():
            raise RuntimeError("FlatAdamW: the parameter arena was rebuilt (model moved?) - recreate the optimizer")
        ar.finalize_grads()
        self._refresh_active()
        host = self._hyper_ring.next()
        for ki, (gi, _cohort) in enumerate(self._kgroups):
            g = self.param_groups[gi]
            host[0, ki] = g["lr"]
            host[1, ki] = g["weight_decay"]
        self.hyper.copy_(host, non_blocking=True)
        self._hyper_ring.copied()
        b1, b2 = self.param_groups[0]["betas"]
        sc = self.scaler
        fresh16 = ar._ver16 is not None and ar._ver16 == ar._version()
        ng = max(1, len(self._kgroups))
        ops.grad_stats(ar.grad, sc.state, self.stats_ws)
        ops.adamw_step(ar.flat, ar.grad, self.exp_avg, self.exp_avg_sq, self.group_of_block, self.hyper[0, :ng],
                       self.hyper[1, :ng], sc.state, beta1=b1, beta2=b2, eps=self.param_groups[0]["eps"],
                       max_grad_norm=self.max_grad_norm, growth_factor=sc.growth_factor,
                       backoff_factor=sc.backoff_factor, growth_interval=sc.growth_interval, param16=ar.flat16,
                       group_step0=self.group_step0, group_bc=self.group_bc)
        self.steps_done += 1
        # the kernel rewrote the fp16 shadow of every element it updated: the shadow stays valid if it was valid
        # before (parameters without a gradient are untouched on both sides); otherwise the next forward re-casts
        if not fresh16:
            ar._ver16 = None
        return None

    def grad_norm(self):
        """Unscaled gradient norm of the last step (device scalar)."""
        return self.scaler.state[5]

    # ---- exact resume: moments, step counts and the loss-scale state travel with the optimizer -------------------
    def state_dict(self):
        ar = self.arena
        name_of = {id(p): n for n, p in zip(ar.names, ar.params)}
        return {"param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                "scaler_state": self.scaler.state.detach().cpu(), "steps_done": int(self.steps_done),
                "cohort": {name_of[i]: c for i, c in self._cohort.items() if i in name_of}, "ncohorts": self._ncohorts,
                "kgroups": [list(k) for k in self._kgroups], "group_step0": self.group_step0.detach().cpu(),
                "arena_names": list(ar.names), "arena_total": int(ar.total)}

    def load_state_dict(self, sd):
        ar = self.arena
        if sd.get("arena_total") != ar.total or list(sd.get("arena_names", [])) != list(ar.names):
            raise ValueError("FlatAdamW.load_state_dict: the state was saved for a different parameter arena layout")
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.scaler.state.copy_(sd["scaler_state"])
        self.steps_done = int(sd["steps_done"])
        by_name = {n: p for n, p in zip(ar.names, ar.params)}
        self._cohort = {id(by_name[n]): int(c) for n, c in sd["cohort"].items() if n in by_name}
        self._ncohorts = int(sd["ncohorts"])
        self._kgroups = [tuple(k) for k in sd["kgroups"]]
        self.group_step0.copy_(sd["group_step0"])
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})
        self._active = None
