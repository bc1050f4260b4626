"""Forward determinism stress (root cause hunt for the r1e flake: test_bert_encoder_train_mode_dropout_gradients once saw
2e-4-different outputs for identical seeds in a shared process).  Runs the BERT encoder (train mode, dropout on, fixed
seed / step / site ids) and a Swin block stack repeatedly, with backward passes and allocator churn in between, and
reports the first layer whose forward output is not bit-identical to the first run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(iters=40):
    from lavender_b200.bert import BertConfig, BertEncoder
    from lavender_b200 import dropout as DR
    torch.manual_seed(0)
    cfg = BertConfig(num_hidden_layers=4)
    enc = BertEncoder(cfg).cuda().train()
    x = torch.randn(3, 283, 768, device="cuda")
    mask = torch.ones(3, 283, device="cuda")
    mask[1, 250:260] = 0
    ref = None
    bad = 0
    for it in range(iters):
        st = DR.reseed("cuda", 42)
        st._site = 0
        xr = x.clone().requires_grad_(True)
        y = enc(xr, mask)["last_hidden_state"]
        if ref is None:
            ref = y.detach().clone()
        elif not torch.equal(ref, y.detach()):
            bad += 1
            d = (ref - y.detach()).abs()
            print(f"iter {it}: forward differs: max {d.max().item():.3e}, {int((d > 0).sum())} elements, rows "
                  f"{sorted(set((d > 0).nonzero()[:, 1].tolist()))[:10]}", flush=True)
        y.square().mean().backward()          # side-stream wgrads + allocator churn between forwards
        junk = [torch.empty(int(torch.randint(1, 64, (1,))) << 20, device="cuda") for _ in range(4)]
        del junk
        for p in enc.parameters():
            p.grad = None
    torch.cuda.synchronize()
    print(f"determinism_stress: {iters} forwards, {bad} differed from the first")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
