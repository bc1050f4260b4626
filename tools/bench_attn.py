"""Micro-benchmark of the fused attention kernels on the hot path's shapes (CUDA events, L2 flushed between runs)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from lavender_b200 import ops  # noqa: E402
from bench_gemm import timeit  # noqa: E402


def main():
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    drop_on = "--dropout" in sys.argv
    rng = torch.tensor([1, 0], dtype=torch.int64, device="cuda")
    cases = [("bert_vtm", 64, 12, 32, 284), ("bert_mlm", 64, 12, 8, 283), ("win_s2", 32, 16, 32, 245),
             ("win_s0", 32, 4, 512, 245), ("win_s1", 32, 8, 128, 245)]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for tag, hd, nh, nprob, L in cases:
        if only and tag not in only:
            continue
        C = nh * hd
        rows = nprob * L
        qkv = (torch.randn(rows, 3 * C, device="cuda") * 0.7).half()
        out = torch.zeros(rows, C, device="cuda", dtype=torch.float16)
        lse = torch.zeros(nh, rows, device="cuda")
        dout = (torch.randn(rows, C, device="cuda") * 0.5).half()
        dq = torch.zeros(rows, C, device="cuda")
        dqkv = torch.zeros(rows, 3 * C, device="cuda", dtype=torch.float16)
        scale = 1.0 / math.sqrt(hd)
        kw = dict(q_off=0, k_off=C, v_off=2 * C, head_dim=hd, nheads=nh, nprob=nprob, L_tok=L, scale=scale)
        if hd == 64:
            NPk = (L + 127) // 128 * 128
            kb = torch.full((nprob, NPk), float("-inf"), device="cuda")
            kb[:, :L] = 0
            kw["key_bias"] = kb
            if drop_on:
                kw["drop"] = (rng, 5, 0.1)
            ds = None
        else:
            dense = (torch.randn(4, nh, 256, 256, device="cuda") * 0.5).half()
            dense[..., L:] = -30000.0
            kw["bias16"] = dense
            kw["prob_class"] = torch.zeros(max(1, nprob // 8), dtype=torch.int32, device="cuda")
            ds = torch.zeros(nprob, nh, 256, 256, device="cuda", dtype=torch.float16)
        fwd = lambda: ops.attn_fwd(qkv, out, lse, **kw)
        bwd = lambda: ops.attn_bwd(qkv, out, dout, lse, dq, dqkv, ds16=ds, **kw)
        tf = timeit(fwd, flush=flush)
        tb = timeit(bwd, flush=flush)
        fl = 4.0 * L * L * hd * nh * nprob
        print(f"{tag:10s} hd={hd} nh={nh} nprob={nprob} L={L}  fwd {tf * 1e3:7.1f} us ({fl / tf / 1e9:6.1f} TF)  "
              f"bwd {tb * 1e3:7.1f} us ({2 * fl / tb / 1e9:6.1f} TF)", flush=True)


if __name__ == "__main__":
    main()
