"""Fused flat-arena AdamW + device GradScaler == torch.optim.AdamW + clip_grad_norm_ + torch.amp.GradScaler
(the sequence of Agent_Base.backward_step, agent.py:240-250), including a skipped step on an inf gradient."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(37, 53)
        self.swin = torch.nn.Linear(53, 29)          # "swin." in the name -> its own lr group, like agent.py:108-119
        self.LayerNorm = torch.nn.LayerNorm(29)
        self.unused = torch.nn.Parameter(torch.randn(11))

    def forward(self, x):
        return self.LayerNorm(self.swin(torch.tanh(self.a(x))))


def _groups(m, lr, wd, mul):
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    g = {(d, s): [] for d in (True, False) for s in (True, False)}
    for n, p in m.named_parameters():
        g[(not any(nd in n for nd in no_decay), "swin." in n)].append(p)
    return [{"params": g[(True, True)], "weight_decay": wd, "lr": lr * mul}, {"params": g[(True, False)], "weight_decay": wd},
            {"params": g[(False, True)], "weight_decay": 0.0, "lr": lr * mul}, {"params": g[(False, False)], "weight_decay": 0.0}]


def test_flat_adamw_matches_torch():
    from lavender_b200.arena import ParamArena
    from lavender_b200.optim import DeviceGradScaler, FlatAdamW
    torch.manual_seed(0)
    ref = Toy().cuda()
    mine = copy.deepcopy(ref)
    ar = ParamArena(mine)
    lr, wd, mul, max_norm = 1e-2, 1e-1, 0.5, 0.3
    # eps well above the fp32 noise of near-zero gradients (saturated tanh units): with eps = 1e-8 the update of such
    # elements amplifies 1e-9-level differences in exp_avg by 1/eps and the comparison is ill-conditioned
    eps = 1e-5
    opt_r = torch.optim.AdamW(_groups(ref, lr, wd, mul), lr=lr, betas=(0.9, 0.98), weight_decay=wd, eps=eps)
    sc_r = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=3)
    sc_m = DeviceGradScaler("cuda", init_scale=1024.0, growth_interval=3)
    opt_m = FlatAdamW(_groups(mine, lr, wd, mul), ar, sc_m, lr=lr, betas=(0.9, 0.98), eps=eps, weight_decay=wd,
                      max_grad_norm=max_norm)
    sched_r = torch.optim.lr_scheduler.LambdaLR(opt_r, lambda s: 1.0 / (1 + s))
    sched_m = torch.optim.lr_scheduler.LambdaLR(opt_m, lambda s: 1.0 / (1 + s))
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(8):
        x = torch.randn(16, 37, device="cuda", generator=g) * (3.0 if it != 4 else float("inf"))   # step 4: inf loss
        for model, opt, sc, fused in ((ref, opt_r, sc_r, False), (mine, opt_m, sc_m, True)):
            loss = model(x).pow(2).mean()
            sc.scale(loss).backward()
            if fused:
                ar.finalize_grads()
            else:
                sc.unscale_(opt)
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
            sc.step(opt)
            sc.update()
            opt.zero_grad()
        sched_r.step()
        sched_m.step()
        torch.cuda.synchronize()
        assert abs(sc_r.get_scale() - sc_m.get_scale()) < 1e-6, (it, sc_r.get_scale(), sc_m.get_scale())
        for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=5e-6), (it, n, (p - q).abs().max().item())
    assert sc_m.get_scale() != 1024.0            # the scale moved (back-off at the inf step, growth afterwards)
    assert torch.equal(ref.unused, mine.unused)  # parameter without gradient: untouched by both


class ToyLate(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(19, 24)
        self.late = torch.nn.Parameter(torch.randn(24) * 0.1)   # like emb_odr: first gradient arrives at step 3
        self.use_late = False

    def forward(self, x):
        h = torch.tanh(self.a(x))
        return h * (1.0 + self.late) if self.use_late else h


def test_flat_adamw_late_parameter_uses_its_own_step_and_state_roundtrip():
    """torch.optim.AdamW keeps a step count PER PARAMETER: a parameter that gets its first gradient at step 3 is
    bias-corrected with t = 1 there, not with the global step.  Also: state_dict -> load_state_dict
    into a fresh optimizer continues identically (exact resume)."""
    from lavender_b200.arena import ParamArena
    from lavender_b200.optim import DeviceGradScaler, FlatAdamW
    torch.manual_seed(1)
    ref = ToyLate().cuda()
    mine = copy.deepcopy(ref)
    ar = ParamArena(mine)
    lr, wd, eps = 1e-2, 1e-2, 1e-5
    opt_r = torch.optim.AdamW([{"params": list(ref.parameters())}], lr=lr, betas=(0.9, 0.98), weight_decay=wd, eps=eps)
    sc_m = DeviceGradScaler("cuda", init_scale=256.0)
    opt_m = FlatAdamW([{"params": list(mine.parameters())}], ar, sc_m, lr=lr, betas=(0.9, 0.98), eps=eps, weight_decay=wd)
    g = torch.Generator(device="cuda").manual_seed(2)

    def one_step(it, opt_mine, scaler):
        x = torch.randn(8, 19, device="cuda", generator=g)
        ref.use_late = mine.use_late = it >= 3
        opt_r.zero_grad(set_to_none=True)
        ref(x).pow(2).mean().backward()
        opt_r.step()
        scaler.scale(mine(x).pow(2).mean()).backward()
        ar.finalize_grads()
        scaler.step(opt_mine)
        opt_mine.zero_grad()
        torch.cuda.synchronize()
        for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=5e-6), (it, n, (p - q).abs().max().item())

    for it in range(6):
        one_step(it, opt_m, sc_m)
    sd = opt_m.state_dict()
    sc2 = DeviceGradScaler("cuda", init_scale=1.0)
    opt2 = FlatAdamW([{"params": list(mine.parameters())}], ar, sc2, lr=lr, betas=(0.9, 0.98), eps=eps, weight_decay=wd)
    opt2.load_state_dict(sd)
    for it in range(6, 9):
        one_step(it, opt2, sc2)
