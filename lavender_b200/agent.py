"""Train-loop mirror of the reference's agent.py (`Agent_Base`, `WarmupLinearLR`, `NormSoftmaxLoss`) and of
`Agent_Pretrain_MLM` (main_pretrain_mlm.py:122-232) for the native model.

Same method names and step semantics: CE(ignore -1) x2 -> scaled backward -> unscale -> clip(max_grad_norm) ->
AdamW(betas 0.9/0.98, name-based decay / lr groups) -> warm-up-linear LR -> zero_grad.  What changes underneath:
  * DDP / DeepSpeed ZeRO-1 (agent.py:252-265) -> lavender_b200.dist.GradSync: one all-reduce of the flat gradient
    arena, the BERT + head part overlapped with the Swin backward;
  * torch.autocast / DeepSpeed fp16 casts -> the native kernels pick fp16 operands themselves (fp32 master
    weights, fp32 accumulation), so `prepare_batch` only moves tensors; GradScaler is kept because activation
    gradients travel as fp16 exactly like under the reference's autocast;
  * the loss is the native cross-entropy kernel.
"""
import inspect
import json
import math
import os
from collections import defaultdict

import numpy as np
import torch

from . import dist as D
from .bert import CrossEntropyLoss


class WarmupLinearLR(torch.optim.lr_scheduler.LRScheduler):
    """agent.py:13-44: linear warm-up over warmup_ratio*max_iter steps, then linear decay to min_lr."""

    def __init__(self, optimizer, max_iter, min_lr=1e-8, warmup_ratio=0.1, last_epoch=-1):
        self.max_iter, self.min_lr, self.warmup_ratio = max_iter, min_lr, warmup_ratio
        self.warmup_iters = int(warmup_ratio * max_iter)
        super().__init__(optimizer, last_epoch)

    def get_lr_factor(self):
        step, tot, warm = self.last_epoch, self.max_iter, self.warmup_iters
        if step < warm:
            return max(0, step / warm)
        return max(0, (tot - min(step, tot)) / (tot - warm))

    def get_lr(self):
        f = self.get_lr_factor()
        return [max(self.min_lr, base * f) for base in self.base_lrs]


class NormSoftmaxLoss(torch.nn.Module):
    """agent.py:47-65: symmetric InfoNCE over a similarity matrix (retrieval baselines; not on the MLM path)."""

    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, x):
        i = torch.log_softmax(x / self.temperature, dim=1).diag()
        j = torch.log_softmax(x.t() / self.temperature, dim=1).diag()
        return -i.sum() / len(i) - j.sum() / len(j)


def move_to_cuda(batch):
    """dataset.py:333-344."""
    if isinstance(batch, torch.Tensor):
        return batch.cuda(non_blocking=True)
    if isinstance(batch, list):
        return [move_to_cuda(t) for t in batch]
    if isinstance(batch, tuple):
        return tuple(move_to_cuda(t) for t in batch)
    if isinstance(batch, dict):
        return {n: move_to_cuda(t) for n, t in batch.items()}
    return batch


def humanbytes(n):
    for unit in ("B", "KB", "MB", "GB", "TB"):
        if abs(n) < 1024 or unit == "TB":
            return f"{n:.2f} {unit}"
        n /= 1024.0


class Agent_Base:
    def __init__(self, args, model):
        self.args, self.model = args, model
        self.loss_func = CrossEntropyLoss(ignore_index=-1)
        # fused flat-arena optimizer (optim.py) unless args.fused_optimizer is False or the model has no arena
        self.fused = bool(getattr(args, "fused_optimizer", True)) and hasattr(model, "arena") and \
            next(model.parameters()).is_cuda
        if self.fused:
            from .optim import DeviceGradScaler
            self.scaler = DeviceGradScaler(next(model.parameters()).device)
        else:
            self.scaler = torch.amp.GradScaler("cuda", enabled=torch.cuda.is_available())
        self.optzr = self.build_optimizer()
        self.lr_scheduler = WarmupLinearLR(self.optzr, args.max_iter)
        self.log = None
        self.grad_sync = None
        self.graphs = None  # graph.GraphCache when args.cuda_graph
        self.tokzr = getattr(model, "tokzr", None)
        if self.tokzr is None:
            raise ValueError("Agent_Base needs model.tokzr (no tokenizer files are fetched offline)")
        tk = self.tokzr
        (self.cls_token_id, self.sep_token_id, self.pad_token_id, self.mask_token_id,
         self.unk_token_id) = tk.convert_tokens_to_ids([tk.cls_token, tk.sep_token, tk.pad_token, tk.mask_token,
                                                        tk.unk_token])
        self.true_token_id = tk.convert_tokens_to_ids(["true"])[0]
        self.false_token_id = tk.convert_tokens_to_ids(["false"])[0]
        self.global_step = 0

    # ---- optimizer (agent.py:96-140): groups decided by parameter NAME ---------------------------------
    def build_optimizer(self):
        no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
        groups = {(d, s): [] for d in (True, False) for s in (True, False)}
        for n, p in self.model.named_parameters():
            groups[(not any(nd in n for nd in no_decay), "swin." in n)].append(p)
        wd, lr, mul = self.args.decay, self.args.lr, self.args.vis_backbone_lr_mul
        spec = [{"params": groups[(True, True)], "weight_decay": wd, "lr": lr * mul},
                {"params": groups[(True, False)], "weight_decay": wd},
                {"params": groups[(False, True)], "weight_decay": 0.0, "lr": lr * mul},
                {"params": groups[(False, False)], "weight_decay": 0.0}]
        if getattr(self, "fused", False):
            from .optim import FlatAdamW
            return FlatAdamW(spec, self.model.arena(), self.scaler, lr=lr, betas=(0.9, 0.98), weight_decay=wd,
                             max_grad_norm=self.args.max_grad_norm)
        return torch.optim.AdamW(spec, lr=lr, betas=(0.9, 0.98), weight_decay=wd)

    # ---- metrics / checkpoints -------------------------------------------------------------------------
    def reduce_dict(self, data):
        return D.reduce_dict(data)

    def reduce_mean(self, v):
        world = D.get_world_size()
        if world < 2 or not torch.distributed.is_initialized():
            return v
        t = torch.tensor(float(v), device="cuda" if torch.cuda.is_available() else "cpu")
        torch.distributed.all_reduce(t)
        return t.item() / world

    def save_training_meta(self):
        if D.is_main_process():
            os.makedirs(self.args.path_output, exist_ok=True)
            with open(f"{self.args.path_output}/args.json", "w") as f:
                json.dump(self.args, f, indent=2)
            self.save_model(0)

    def save_model(self, ep):
        """Weights-only state_dict with the reference's keys and file name (agent.py:164-180)."""
        if D.is_main_process():
            out = self.args.path_output
            os.makedirs(out, exist_ok=True)
            sd = {k: v.detach().cpu().clone() if isinstance(v, torch.Tensor) else v
                  for k, v in self.model.state_dict().items()}
            torch.save(sd, f"{out}/ckpt_violet_{self.args.task}_{ep}.pt")
            if self.log is not None:
                with open(f"{out}/log.json", "w") as f:
                    json.dump(self.log, f, indent=2)

    def log_memory(self, ep=-1, step=-1):
        where = f"global step: {self.global_step}," if ep == -1 and step == -1 else f"ep: {ep}, step: {step},"
        mem = humanbytes(torch.cuda.max_memory_allocated()) if torch.cuda.is_available() else "n/a"
        g = self.optzr.param_groups
        return f"{where} lr_swin: {g[0]['lr']:.2e}, lr_bert: {g[1]['lr']:.2e}, max memory: {mem}"

    # ---- step --------------------------------------------------------------------------------------------
    def prepare_batch(self, batch):
        return move_to_cuda(batch)

    def forward_step(self, batch):
        if isinstance(batch, dict):
            names = inspect.getfullargspec(self.model.forward).args
            return self.model(batch) if "batch" in names else self.model(**batch)
        if isinstance(batch, tuple):
            return self.model(*batch)
        raise TypeError(f"batch is either dict or tuple, {type(batch)}")

    def backward_step(self, loss, graphed=False):
        """agent.py:235-250.  graphed=True: the scaled backward already ran inside a CUDA-graph replay (graph.py),
        which also zeroes the gradient arena at its start, so neither backward nor zero_grad happens here."""
        if not graphed:
            self.scaler.scale(loss).backward()
        if graphed and self.grad_sync is not None and graphed == "synced":
            pass                            # the gradient all-reduce was captured inside the replayed graph (graph.py)
        elif self.grad_sync is not None:
            self.grad_sync.finish()
        elif hasattr(self.model, "arena"):
            self.model.arena().finalize_grads()
        if self.args.max_grad_norm > 0 and not self.fused:     # the fused optimizer clips inside its own kernels
            self.scaler.unscale_(self.optzr)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.args.max_grad_norm)
        self.scaler.step(self.optzr)
        self.scaler.update()
        self.lr_scheduler.step()
        if not graphed:
            self.optzr.zero_grad()
        self.global_step += 1

    def prepare_dist_model(self):
        """agent.py:252-265.  No module wrapper: ranks start from rank 0's weights and gradients are averaged by
        GradSync; `args.deepspeed` is accepted and means the same thing here (fp16 operands are always on)."""
        ar = self.model.arena()
        if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            D.broadcast_parameters(ar)
            self.grad_sync = D.GradSync(ar)


class Agent_Pretrain_MLM(Agent_Base):
    """main_pretrain_mlm.py:122-232 (+ the pieces of Agent_Pretrain it inherits)."""

    def cal_vtm_loss(self, txt, out, ans, is_train=True):
        if is_train:
            return self.loss_func(out.flatten(0, out.dim() - 2), ans.flatten())
        B = txt.shape[0]
        p_true, p_false = out[:, :, self.true_token_id], out[:, :, self.false_token_id]
        score = (p_true / (p_true + p_false))[ans != -1].view(B, -1)
        lab = ans[ans != -1].view(B, -1)
        pred = torch.argmax(score, dim=-1)
        gt = (lab == self.true_token_id).nonzero()[:, 1]
        return float((pred == gt).float().sum() / B)

    def _train_step_device(self, batch):
        """One training step, everything enqueued, nothing read back: returns the two losses as device scalars."""
        self.model.train(True)
        if getattr(self.args, "cuda_graph", False) and batch.get("vtm_prompt") is None:
            if self.graphs is None:
                from .graph import GraphCache
                self.graphs = GraphCache(self)
            g = self.graphs.get(batch)
            ls_mtm, ls_vtm = g(batch)
            self.backward_step(None, graphed="synced" if g.sync_in_graph else True)
            return ls_mtm, ls_vtm
        out = self.forward_step(batch)
        out_mtm, out_vtm, ans_mtm, ans_vtm = out["out_mtm"], out["out_vtm"], out["ans_mtm"], out["ans_vtm"]
        ls_mtm = self.loss_func(out_mtm.flatten(0, out_mtm.dim() - 2), ans_mtm.flatten())
        ls_vtm = self.cal_vtm_loss(batch["txt"], out_vtm, ans_vtm, True)
        self.backward_step(ls_mtm + ls_vtm)
        return ls_mtm.detach(), ls_vtm.detach()

    # ---- pipelined input path: the next batch's host work and H2D copy run under the current step's kernels ----
    def prefetch(self, host_batch):
        """Starts the host->device copy of a (masked, ideally pinned) host batch on a copy stream and returns a handle
        for `step_async`.  Called right after `step_async` of the previous batch, it overlaps that step's compute."""
        if self.__dict__.get("_copy_stream") is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            dev = self.prepare_batch(host_batch)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return dev, ev

    def step_async(self, handle):
        """Enqueues one training step on the batch of `prefetch`; returns the pending losses (see `finish`)."""
        dev, ev = handle
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for v in dev.values():   # allocated on the copy stream, consumed here: keep the allocator from recycling early
            if isinstance(v, torch.Tensor) and v.is_cuda:
                v.record_stream(cur)
        l1, l2 = self._train_step_device(dev)
        # the losses of a graphed step live in static tensors that the next replay overwrites: park this step's pair
        # in pinned host memory (async D2H + event), so `finish` may be called after later steps were enqueued
        # (a slot is reused after 4 steps: if the host got that far ahead without reading it, wait for its copy first, so
        #  one step's losses are never overwritten by another's while a reader may still pick them up)
        slots = self.__dict__.setdefault("_loss_slots", [])
        if not slots:
            slots.extend([torch.empty(2, dtype=torch.float32).pin_memory(), torch.cuda.Event(), False] for _ in range(4))
            self._loss_i = 0
        slot = slots[self._loss_i % len(slots)]
        self._loss_i += 1
        buf, done = slot[0], slot[1]
        if slot[2]:
            done.synchronize()
        buf.copy_(torch.stack([l1.detach().float(), l2.detach().float()]), non_blocking=True)
        done.record(cur)
        slot[2] = True
        return buf, done

    @staticmethod
    def finish(pending):
        """Device->host read of a step's two losses: waits for THAT step only (later steps may already be queued; at
        most 3 steps may be outstanding)."""
        if isinstance(pending[1], torch.cuda.Event):
            buf, done = pending
            done.synchronize()
            return {"mtm": float(buf[0]), "vtm": float(buf[1])}
        return {"mtm": pending[0].item(), "vtm": pending[1].item()}

    def step(self, batch, is_train=True):
        self.model.train(is_train)
        if is_train and getattr(self.args, "cuda_graph", False) and batch.get("vtm_prompt") is None:
            return self.finish(self._train_step_device(batch))
        with torch.set_grad_enabled(is_train):
            out = self.forward_step(batch)
            out_mtm, out_vtm, ans_mtm, ans_vtm = out["out_mtm"], out["out_vtm"], out["ans_mtm"], out["ans_vtm"]
            ls_mtm = self.loss_func(out_mtm.flatten(0, out_mtm.dim() - 2), ans_mtm.flatten())
            ls_vtm = self.cal_vtm_loss(batch["txt"], out_vtm, ans_vtm, is_train)
        if is_train:
            self.backward_step(ls_mtm + ls_vtm)
            return {"mtm": ls_mtm.item(), "vtm": ls_vtm.item()}
        n = (ans_mtm != -1).sum()
        ac = float((torch.argmax(out_mtm, dim=-1) == ans_mtm).sum() / n) if n > 0 else -1
        return {"mtm": ac, "vtm": ls_vtm}

    def masking(self, txt, mask, p_mask=0.15):
        """BERT-style [MASK] replacement on the host (main_pretrain_mlm.py:178-200), vectorised; consumes the torch
        RNG once per row like the reference (T.rand(_X) per sample)."""
        B, X = txt.shape
        ans = torch.full(txt.shape, -1, dtype=torch.long)
        if p_mask <= 0:
            return {"txt": txt, "mask": mask, "ans_mtm": ans}
        special = (txt == self.cls_token_id) | (txt == self.sep_token_id) | (txt == self.pad_token_id) | \
                  (txt == self.mask_token_id)
        draw = torch.stack([torch.rand(X) for _ in range(B)]) < p_mask
        sel = draw & ~special
        ans[sel] = txt[sel]
        txt[sel] = self.mask_token_id
        out = {"txt": txt, "mask": mask, "ans_mtm": ans}
        rows = self.labelled_rows(ans)
        if rows is not None:
            out["mtm_rows"] = rows
        return out

    def labelled_rows(self, ans, capacity=None):
        """fixed-capacity list of the flat indices (b * Lt + t) of the labelled MLM rows, padded with the
        index of an unlabelled row (its label is -1, so it adds nothing to the loss).  Capacity = one 128-row GEMM tile
        (or B*Lt when smaller); returns None when the batch has more labelled rows (the model then computes full logits).
        `LAV_MLM_LABELLED_ONLY=0` disables it."""
        if os.environ.get("LAV_MLM_LABELLED_ONLY", "1") == "0":
            return None
        flat = ans.reshape(-1)
        cap = capacity or min(128, flat.numel())
        idx = (flat != -1).nonzero().reshape(-1)
        un = (flat == -1).nonzero().reshape(-1)
        if idx.numel() > cap or un.numel() == 0:
            return None
        pad = un[:1].expand(cap - idx.numel())
        return torch.cat([idx, pad]).contiguous()

    def go_dl(self, ep, dl, is_train):
        self.model.train(is_train)
        ret = defaultdict(list)
        idx = 0
        for idx, batch in enumerate(dl):
            batch = defaultdict(lambda: None, batch)
            batch.update(self.masking(batch["txt"], batch["mask"], getattr(self.args, "p_mask", 0.15)))
            if self.args.enable_prompt:
                batch["vtm_prompt"] = dl.dataset.get_vtm_prompt()
                batch["cap_prompt"] = dl.dataset.get_cap_prompt()
            r = self.step(self.prepare_batch(dict(batch)), is_train)
            for k, v in r.items():
                ret[k].append(v)
        return {k: self.reduce_mean(float(np.average([v for v in l if not math.isnan(v)]))) for k, l in ret.items()}
