import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lavender_b200.agent import Agent_Pretrain_MLM
from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
args = default_args(vis_backbone_size="base", size_batch=8)
model = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
for cfg in (model.trsfr.config, model.enc_txt.emb_txt.config): cfg.lav_eval_dropout = True
model.cuda()
agent = Agent_Pretrain_MLM(args, model)
host = bench.make_host_batch(8, 0, True)
b = {"img": host["img"], "txt": host["txt"].clone(), "mask": host["mask"]}
b.update(agent.masking(b["txt"], b["mask"], 0.15))
dev = agent.prepare_batch(b)
def step():
    model.train()
    out = agent.forward_step(dev)
    l = agent.loss_func(out["out_mtm"].flatten(0, 1), out["ans_mtm"].flatten()) + agent.loss_func(out["out_vtm"].flatten(0, 1), out["ans_vtm"].flatten())
    agent.backward_step(l)
for mode in ("sync", "nosync", "sync", "nosync"):
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    t0 = time.perf_counter(); host_t = []
    evs[0].record()
    for i in range(12):
        h0 = time.perf_counter()
        step()
        evs[i + 1].record()
        if mode == "sync": torch.cuda.synchronize()
        host_t.append((time.perf_counter() - h0) * 1e3)
    torch.cuda.synchronize()
    print(mode, "dev ms:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(12)])
    print(mode, "host ms:", [round(x, 1) for x in host_t], "mem GB", torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
