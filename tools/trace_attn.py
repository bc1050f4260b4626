"""Phase timeline of the one-shot window attention forward kernel (clock64 stamps per CTA, lav_debug_set_trace):
   0 entry | 1 prologue done (barriers, TMEM alloc, __syncthreads) | 2 TMA data landed (MMA thread) | 3 S ready (softmax)
   4 pass 1 (row maximum) done | 5 P published | 6 PV done | 7 all threads done (before TMEM dealloc)"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavender_b200 import ops, _lib as L  # noqa: E402


def main():
    torch.manual_seed(0)
    hd, nh, nprob, Ltok = 32, 16, 32, 245
    C = nh * hd
    rows = nprob * Ltok
    qkv = (torch.randn(rows, 3 * C, device="cuda") * 0.7).half()
    out = torch.zeros(rows, C, device="cuda", dtype=torch.float16)
    lse = torch.zeros(nh, rows, device="cuda")
    dense = (torch.randn(4, nh, 256, 256, device="cuda") * 0.5).half()
    dense[..., Ltok:] = -30000.0
    cls = torch.zeros(4, dtype=torch.int32, device="cuda")
    kw = dict(q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nh, nprob=nprob, L_tok=Ltok, scale=1 / math.sqrt(hd),
              bias16=dense, prob_class=cls)
    for _ in range(3):
        ops.attn_fwd(qkv, out, lse, **kw)
    nct = 2 * nh * nprob
    buf = torch.zeros(nct * 8, dtype=torch.int64, device="cuda")
    L.check(L.lib().lav_debug_set_trace(buf.data_ptr(), buf.numel()))
    ops.attn_fwd(qkv, out, lse, **kw)
    torch.cuda.synchronize()
    L.lib().lav_debug_set_trace(None, 0)
    t = buf.view(nct, 8).cpu().double()
    d = t[:, 1:] - t[:, :-1]
    names = ["prologue", "TMA wait", "S MMA -> softmax", "pass 1 (max)", "pass 2 (exp, P)", "PV MMA", "O store + sync"]
    print(f"window attention forward, {nct} CTAs: mean cycles per phase (clock64, per-SM clock)")
    for n, m, md in zip(names, d.mean(0).tolist(), d.median(0).values.tolist()):
        print(f"  {n:22s} mean {m:9.0f}  median {md:9.0f}")
    life = t[:, 7] - t[:, 0]
    print(f"  CTA lifetime           mean {life.mean().item():9.0f}  median {life.median().item():9.0f}")


if __name__ == "__main__":
    main()
