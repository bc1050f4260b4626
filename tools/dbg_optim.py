import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_optim_gpu import Toy, _groups
from lavender_b200.arena import ParamArena
from lavender_b200.optim import DeviceGradScaler, FlatAdamW
torch.manual_seed(0)
ref = Toy().cuda(); mine = copy.deepcopy(ref); ar = ParamArena(mine)
lr, wd, mul, max_norm = 1e-2, 1e-1, 0.5, 0.3
opt_r = torch.optim.AdamW(_groups(ref, lr, wd, mul), lr=lr, betas=(0.9, 0.98), weight_decay=wd)
sc_m = DeviceGradScaler("cuda", init_scale=1024.0, growth_interval=1000)
opt_m = FlatAdamW(_groups(mine, lr, wd, mul), ar, sc_m, lr=lr, betas=(0.9, 0.98), weight_decay=wd, max_grad_norm=max_norm)
g = torch.Generator(device="cuda").manual_seed(1)
for it in range(3):
    x = torch.randn(16, 37, device="cuda", generator=g) * 3.0
    loss = ref(x).pow(2).mean(); (loss * 1024).backward()
    for p in ref.parameters():
        if p.grad is not None: p.grad.mul_(1 / 1024)
    gr = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
    tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
    opt_r.step(); opt_r.zero_grad()
    loss = mine(x).pow(2).mean(); sc_m.scale(loss).backward(); ar.finalize_grads()
    gm = {n: p.grad.clone() / 1024 for n, p in mine.named_parameters() if p.grad is not None}
    opt_m.step(); opt_m.zero_grad()
    torch.cuda.synchronize()
    print(it, "norm torch", tn.item(), "mine", sc_m.state[5].item(), "state", sc_m.state[:10].tolist())
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        gd = (gr[n] - gm[n]).abs().max().item() if n in gr else -1
        print("   ", n, "param diff", (p - q).abs().max().item(), "grad diff", gd, "m diff",
              (opt_r.state[p]["exp_avg"] - opt_m.exp_avg[ar.offsets[id(q)]:ar.offsets[id(q)] + q.numel()].view(q.shape)).abs().max().item() if p in opt_r.state else -1,
              "v rel diff", ((opt_r.state[p]["exp_avg_sq"] - opt_m.exp_avg_sq[ar.offsets[id(q)]:ar.offsets[id(q)] + q.numel()].view(q.shape)).abs().max() / opt_r.state[p]["exp_avg_sq"].abs().max()).item() if p in opt_r.state else -1)
