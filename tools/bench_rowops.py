"""Micro-benchmark of the LayerNorm kernels on the hot path's shapes (CUDA events, L2 flushed between runs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from lavender_b200 import ops  # noqa: E402
from bench_gemm import timeit  # noqa: E402


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rows, C in [(7840, 512), (125440, 128), (31360, 256), (9088, 768), (1960, 1024)]:
        x = torch.randn(rows, C, device="cuda")
        g, b = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
        y16 = torch.empty(rows, C, device="cuda", dtype=torch.float16)
        mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
        perm = torch.randperm(rows, device="cuda").to(torch.int32)
        fwd = lambda: ops.layernorm_fwd(x, g, b, 1e-5, rows=rows, C=C, row_map=perm, out16=y16, mean=mean, rstd=rstd)
        tf = timeit(fwd, flush=flush)
        dy = torch.randn(rows, C, device="cuda").half()
        add = torch.randn(rows, C, device="cuda")
        dx = torch.empty(rows, C, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        bwd = lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, rows=rows, C=C, row_map=perm, add32=add, dx32=dx, dgamma=dg,
                                        dbeta=db)
        tb = timeit(bwd, flush=flush)
        bf, bb = rows * C * 6, rows * C * 14
        print(f"rows={rows:6d} C={C:4d}  fwd {tf * 1e3:6.1f} us ({bf / tf / 1e6:6.0f} GB/s)  bwd {tb * 1e3:6.1f} us "
              f"({bb / tb / 1e6:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
