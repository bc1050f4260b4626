"""Row-kernel micro-benchmark: LayerNorm forward / backward and the gradient casts at the step's shapes (B = 8, base),
one launch per measurement between L2 flushes, CUDA events; reports us and algorithmic GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavender_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [("swin_s0", 125440, 128, True), ("swin_s1", 31360, 256, True), ("swin_s2", 7840, 512, True),
          ("swin_s3", 1960, 1024, True), ("bert", 40 * 284, 768, False)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = []
for tag, rows, C, mapped in SHAPES:
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(rows, C, generator=g).to(dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rmap = torch.randperm(rows, generator=g).to(dev, torch.int32) if mapped else None
    if mapped:   # window partition keeps runs of 7 tokens contiguous; a random permutation is the worst case, so use runs
        base = torch.arange(rows, dtype=torch.int32).view(-1, 7) if rows % 7 == 0 else torch.arange(rows, dtype=torch.int32).view(-1, 1)
        rmap = base[torch.randperm(base.shape[0], generator=g)].reshape(-1).to(dev)
    y16 = torch.empty(rows, C, dtype=torch.float16, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    t_f = timeit(lambda: ops.layernorm_fwd(x, gamma, beta, 1e-5, rows=rows, C=C, row_map=rmap, out16=y16, mean=mean, rstd=rstd))
    dy = torch.randn(rows, C, generator=g).to(dev, torch.float16)
    add = torch.randn(rows, C, generator=g).to(dev)
    dx = torch.empty(rows, C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    t_b = timeit(lambda: ops.layernorm_bwd(dy, x, gamma, mean, rstd, rows=rows, C=C, row_map=rmap, add32=add, dx32=dx,
                                           dgamma=dg, dbeta=db))
    if tag == "bert":   # the two forms BERT's layer backward uses: fp32 dy, no residual gradient, dx32 + dx16 (+ dropout mask)
        dy32 = dy.float()
        dx16 = torch.empty(rows, C, dtype=torch.float16, device=dev)
        rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
        for name, drop in (("bert_form", None), ("bert_form_dropout", (rng, 3, 0.1))):
            t = timeit(lambda: ops.layernorm_bwd(dy32, x, gamma, mean, rstd, rows=rows, C=C, dx32=dx, dx16=dx16, dgamma=dg,
                                                 dbeta=db, drop16=drop))
            print(json.dumps({"tag": name, "rows": rows, "C": C, "ln_bwd_us": round(t, 1),
                              "ln_bwd_GBps": round(rows * C * 14 / t / 1e3, 0)}))
    o16 = torch.empty(rows, C, dtype=torch.float16, device=dev)
    t_c = timeit(lambda: ops.scale_cast(x, o16, rows=rows, C=C))
    rec = {"tag": tag, "rows": rows, "C": C,
           "ln_fwd_us": round(t_f, 1), "ln_fwd_GBps": round(rows * C * 6 / t_f / 1e3, 0),
           "ln_bwd_us": round(t_b, 1), "ln_bwd_GBps": round(rows * C * 14 / t_b / 1e3, 0),
           "cast_us": round(t_c, 1), "cast_GBps": round(rows * C * 6 / t_c / 1e3, 0)}
    print(json.dumps(rec))
    out.append(rec)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
