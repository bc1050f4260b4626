"""Per-call device-time breakdown of one training step (CUDA events around every C-ABI call), grouped by kernel
family and GEMM shape.  Writes gpurun_out/step_profile.json.   python tools/profile_step.py [--size base] [--B 8]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="base")
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.json"))
    a = ap.parse_args()
    import numpy as np
    import torch
    from lavender_b200 import ops
    from lavender_b200.agent import Agent_Pretrain_MLM
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
    import bench
    args = default_args(vis_backbone_size=a.size, size_batch=a.B, bert_config={"num_hidden_layers": a.layers})
    model = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
    for cfg in (model.trsfr.config, model.enc_txt.emb_txt.config):
        cfg.lav_eval_dropout = True
    model.cuda()
    agent = Agent_Pretrain_MLM(args, model)
    host = bench.make_host_batch(a.B, 0, True)
    b = {"img": host["img"], "txt": host["txt"].clone(), "mask": host["mask"]}
    b.update(agent.masking(b["txt"], b["mask"], 0.15))
    dev = agent.prepare_batch(b)

    def step():
        model.train()
        out = agent.forward_step(dev)
        l = agent.loss_func(out["out_mtm"].flatten(0, 1), out["ans_mtm"].flatten()) + \
            agent.loss_func(out["out_vtm"].flatten(0, 1), out["ans_vtm"].flatten())
        agent.backward_step(l)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    # wall / device time of plain steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    host_issue = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / 5
    ops.PROFILE = []
    ops.PROFILE_META = True
    step()
    torch.cuda.synchronize()
    fam, shapes = {}, {}
    for rec in ops.PROFILE:
        name, s, e, fl = rec[:4]
        meta = rec[4] if len(rec) > 4 else None
        ms = s.elapsed_time(e)
        f = fam.setdefault(name, [0.0, 0, 0.0])
        f[0] += ms; f[1] += 1; f[2] += fl
        if meta:
            g = shapes.setdefault(str(meta), [0.0, 0, 0.0])
            g[0] += ms; g[1] += 1; g[2] += fl
    ops.PROFILE = None
    res = {"plain_step_ms": plain_ms, "host_issue_ms_per_step": host_issue * 1e3,
           "families": {k: {"ms": round(v[0], 3), "n": v[1], "tflops": round(v[2] / v[0] / 1e9, 1) if v[2] else None}
                        for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
           "shapes": {k: {"ms": round(v[0], 3), "n": v[1], "us_each": round(1e3 * v[0] / v[1], 1),
                          "tflops": round(v[2] / v[0] / 1e9, 1) if v[2] else None}
                      for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:120]}}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("plain_step_ms", "host_issue_ms_per_step", "families")}, indent=1))
    for k, v in res["shapes"].items():
        print(f'{v["ms"]:7.3f} ms  n={v["n"]:3d}  {v["us_each"]:7.1f} us  {str(v["tflops"]):>7s} TF  {k}')


if __name__ == "__main__":
    main()
