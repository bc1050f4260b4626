"""Parity decomposition for profiles/PARITY.md.

    python tools/parity_report.py cpu      # here (no GPU): error of fp16 operand rounding in the ORACLE itself
    python tools/parity_report.py gpu      # on the B200 box: native (default / per-module high precision) vs the goldens

Three numbers per golden case (max-abs on the logits, sub-sampled every 61st vocabulary column like the goldens):
  (a) native vs fp32 reference             -- GPU
  (b) native vs fp16-rounded oracle        -- GPU (tiny cases; the CPU oracle of the base case takes ~1 min)
  (c) fp16-rounded oracle vs fp32 oracle   -- CPU: what rounding every GEMM operand / stored activation to fp16 costs in
                                              the reference's OWN arithmetic, i.e. the error of its autocast GPU path
and the per-module bisection (a) with the high-precision mode switched on for growing sets of modules.
Writes gpurun_out/parity_{cpu,gpu}.json.  Test infrastructure: imports oracle/ (never used by the product path).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lavender_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [("tiny_l2_b2", "tiny", 2, 2, True, 0), ("tiny_l1_b3_notask", "tiny", 1, 3, False, 3),
         ("base_l12_b2", "base", 12, 2, True, 5)]


def _inputs(name, size, layers, B, task, seed):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.ModelCfg(swin=O.SWIN[size], bert_layers=layers, enable_task_token=task, vtm_batch=min(B, 4))
    sd = O.make_state_dict(cfg, seed)
    batch = O.make_batch(B, seed=seed)
    if "vt_mask" in gold.files:
        batch["vt_mask"] = torch.from_numpy(gold["vt_mask"])
    return gold, cfg, sd, batch


def _err(out, gold):
    e1 = (out["out_mtm"].detach().float().cpu()[..., ::61] - torch.from_numpy(gold["out_mtm_s"])).abs()
    e2 = (out["out_vtm"].detach().float().cpu()[..., ::61] - torch.from_numpy(gold["out_vtm_s"])).abs()
    return {"mtm_max": e1.max().item(), "vtm_max": e2.max().item(), "mtm_rms": e1.pow(2).mean().sqrt().item(),
            "vtm_rms": e2.pow(2).mean().sqrt().item()}


def cpu():
    torch.set_num_threads(os.cpu_count() or 8)
    res = {}
    for case in CASES:
        name, size, layers, B, task, seed = case
        gold, cfg, sd, batch = _inputs(*case)
        r = {}
        for tag, q, qa in (("fp32", None, None), ("fp16 operands", torch.float16, None),
                           ("fp16 operands + fp16 stored activations", torch.float16, torch.float16),
                           ("bf16 operands + bf16 stored activations", torch.bfloat16, torch.bfloat16)):
            O.set_operand_rounding(q, qa)
            try:
                np.random.seed(1 + seed)
                with torch.no_grad():
                    out = O.pretrain_forward(sd, batch, cfg)
            finally:
                O.set_operand_rounding(None, None)
            r[tag] = _err(out, gold)
            print(name, tag, r[tag], flush=True)
        res[name] = r
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_cpu.json"), "w"), indent=1)


def gpu():
    from lavender_b200 import precision
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
    res = {}
    for case in CASES:
        name, size, layers, B, task, seed = case
        gold, cfg, sd, batch = _inputs(*case)
        args = default_args(vis_backbone_size=size, size_batch=B, bert_config={"num_hidden_layers": layers},
                            enable_task_token=task)
        m = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
        m.load_state_dict(sd, strict=True)
        m.cuda().eval()
        dev = {k: v.cuda() for k, v in batch.items()}
        r = {}
        for tag, scopes in (("default (fp16 operands)", ()), ("high: head", ("head",)), ("high: head+bert", ("head", "bert")),
                            ("high: head+bert+fc", ("head", "bert", "fc")), ("high: all (LAV_PRECISION=high)", True)):
            prev = precision.set_high(scopes if scopes else False)
            try:
                np.random.seed(1 + seed)
                with torch.no_grad():
                    out = m(dict(dev))
                torch.cuda.synchronize()
            finally:
                precision.set_high(prev)
            r[tag] = _err(out, gold)
            if tag.startswith("default"):
                native = {k: out[k].detach().float().cpu() for k in ("out_mtm", "out_vtm")}
            print(name, tag, r[tag], flush=True)
        if size == "tiny" or "--all" in sys.argv:   # (b): against the oracle under the same operand rounding
            O.set_operand_rounding(torch.float16, torch.float16)
            try:
                np.random.seed(1 + seed)
                with torch.no_grad():
                    ref16 = O.pretrain_forward(sd, batch, cfg)
            finally:
                O.set_operand_rounding(None, None)
            r["default vs fp16-rounded oracle"] = {
                "mtm_max": (native["out_mtm"] - ref16["out_mtm"]).abs().max().item(),
                "vtm_max": (native["out_vtm"] - ref16["out_vtm"]).abs().max().item()}
            print(name, "default vs fp16-rounded oracle", r["default vs fp16-rounded oracle"], flush=True)
        res[name] = r
        del m
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_gpu.json"), "w"), indent=1)


if __name__ == "__main__":
    (gpu if (len(sys.argv) > 1 and sys.argv[1] == "gpu") else cpu)()
