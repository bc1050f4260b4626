"""Throughput of BASELINE configs[3] (swin_large_384_patch244_window81212 + BERT-base, batch 4, 5x384x384) on one GPU:
eager fwd + 2 CE + bwd + fused AdamW steps, CUDA-event timed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavender_b200.agent import Agent_Pretrain_MLM  # noqa: E402
from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args  # noqa: E402


def main():
    B = 4
    graph = "--graph" in sys.argv
    args = default_args(vis_backbone_size="large", size_img=384, size_batch=B, cuda_graph=graph, max_iter=100000)
    torch.manual_seed(0)
    np.random.seed(0)
    m = LAVENDER_Pretrain_MLM(args, FakeTokenizer()).cuda()
    ag = Agent_Pretrain_MLM(args, m)
    g = torch.Generator().manual_seed(0)
    txt = torch.randint(1000, 30000, (B, 33), generator=g)
    txt[:, 0], txt[:, -2], txt[:, -1] = 101, 102, 103
    host = {"img": torch.randn(B, 5, 3, 384, 384, generator=g), "txt": txt, "mask": torch.ones(B, 33, dtype=torch.long)}
    host.update(ag.masking(host["txt"], host["mask"], 0.15))
    batch = ag.prepare_batch(host)
    for _ in range(3):
        r = ag.step(dict(batch), True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        r = ag.step(dict(batch), True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"configs[3] large-384 B={B} graph={graph}: {ms:.1f} ms/step = {B / ms * 1e3:.1f} clips/s, losses {r}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
