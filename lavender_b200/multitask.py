"""Multi-task MLM training forwards (BASELINE configs[4]) on the native encoders: video-text retrieval (B^2 pairs),
multiple-choice QA (as MLM over the answer token, or as retrieval over the options), open-ended QA and captioning
(seq2seq-masked MLM) — all through the same three native entry points `go_feat` / `go_cross` / `fc_mtm`.

  LAVENDER_Captioning   model_for_captioning.py:40-93    (`encode_forward`, the training path; decoding is inference-time
                                                          code outside the fwd+bwd hot path)
  LAVENDER_Multi_Task   main_multi_task_mlm.py:77-225    (`forward` dispatch on the task name + the four forwards)
  CaptioningLoss        model_for_captioning.py:10-37
  train_step            main_multi_task_mlm.py:369-390   (Agent_Multi_Task.train_step)

Same batch keys, same output dict ({"out": logits, "ans": labels}) and the same pair ORDER as the reference; the
per-pair Python loops of forward_retrieval (main_multi_task_mlm.py:119-141, B^2 iterations of ~10 tiny ops) are replaced
by two index gathers.
"""
from collections import defaultdict

import torch
import torch.nn as nn

from .bert import CrossEntropyLoss
from .model import LAVENDER_Base, build_mlm_head


class CaptioningLoss(nn.Module):
    """model_for_captioning.py:10-37: KL(log_softmax(logits), smoothed one-hot) summed over the vocabulary, optional
    drop-worst, mean.  With label_smoothing = 0 and drop_worst_ratio = 0 (the defaults of every shipped JSON) this is
    exactly cross-entropy and runs on the native CE kernel; the smoothed / drop-worst variants use the torch formula."""

    def __init__(self, config=None):
        super().__init__()
        get = (lambda k, d: config.get(k, d)) if isinstance(config, dict) else (lambda k, d: getattr(config, k, d))
        self.label_smoothing = get("label_smoothing", 0) if config is not None else 0
        self.drop_worst_ratio = get("drop_worst_ratio", 0) if config is not None else 0
        self.drop_worst_after = get("drop_worst_after", 0) if config is not None else 0
        self.iter = 0
        self.ce = CrossEntropyLoss(ignore_index=-1)

    def forward(self, logits, target):
        self.iter += 1
        eps = self.label_smoothing
        if eps == 0 and not (self.drop_worst_ratio > 0 and self.iter > self.drop_worst_after):
            return self.ce(logits, target)
        n_class = logits.size(1)
        one_hot = torch.zeros_like(logits).scatter(1, target.view(-1, 1), 1)
        one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
        loss = torch.nn.functional.kl_div(torch.log_softmax(logits, dim=1), one_hot, reduction="none").sum(1)
        if self.drop_worst_ratio > 0 and self.iter > self.drop_worst_after:
            loss, _ = torch.topk(loss, k=int(loss.shape[0] * (1 - self.drop_worst_ratio)), largest=False)
        return loss.mean()


class LAVENDER_Captioning(LAVENDER_Base):
    def __init__(self, args, tokzr, is_decoder=True):
        super().__init__(args, tokzr)
        self.config.is_decoder = is_decoder
        self.fc_mtm, _ = build_mlm_head(args.tokenizer, args)
        self.task_tok2id = {"vtm": 0, "mc": 1, "oe": 2, "cap": 3}
        self.emb_task = nn.Parameter(0.02 * torch.randn(10, self.hidden_size))
        self.cap_prompt_txt_L = 0

    def forward(self, batch, is_decode=False):
        batch = defaultdict(lambda: None, batch)
        if is_decode:
            return self.generate(batch)
        return self.encode_forward(batch)

    def generate(self, batch):
        raise NotImplementedError("autoregressive caption decoding (model_for_captioning.py:94-534) is inference-time "
                                  "code outside the fwd+bwd hot path; only encode_forward is native")

    def _prefix_len(self, prompt):
        if prompt is not None and self.args.enable_prompt:
            return len(prompt[0])
        return 1 if self.args.enable_task_token else 0

    def encode_forward(self, batch):
        """model_for_captioning.py:61-93 (the `input_ids is None` branch): MLM over the caption under the seq2seq mask —
        every query sees the video (+ task token / prompt), caption tokens see the caption causally (model.py:208-218)."""
        if batch["input_ids"] is not None:
            raise NotImplementedError("the incremental-decoding branch of encode_forward belongs to generate()")
        img, txt, mask = batch["img"], batch["txt"], batch["mask"]
        ans_mtm, prompt = batch["ans_mtm"], batch["prompt"]
        _B, _T, _, _H, _W = img.shape
        Lv = (1 + (_H // 32) * (_W // 32)) * _T
        feat_img, mask_img, feat_txt, mask_txt = self.go_feat(img, txt, mask)
        ans_mtm, _, feat_txt = self.prepro_txt_inputs(ans_mtm, mask_txt, feat_txt, task_name="cap", prompt=prompt)
        _L = self._prefix_len(prompt)
        if not (prompt is not None and self.args.enable_prompt) and not self.args.enable_task_token:
            assert self.cap_prompt_txt_L == _L
        self.cap_prompt_txt_L = _L
        ans_mtm = ans_mtm.clone()
        ans_mtm[:, :_L] = -1
        mask_pretxt = torch.ones_like(mask_txt)[:, :_L] if _L > 0 else None
        out, _ = self.go_cross(feat_img, mask_img, feat_txt, mask_txt, attn_mask_type=batch["attn_mask_type"] or "full",
                               mask_pretxt=mask_pretxt)
        return {"out": self.fc_mtm(out[:, Lv:]), "ans": ans_mtm}


class LAVENDER_Multi_Task(LAVENDER_Captioning):
    def forward(self, batch, is_decode=False):
        batch = defaultdict(lambda: None, batch)
        task = batch["task"]
        batch["attn_mask_type"] = "full"
        if "captioning" in task:
            batch["attn_mask_type"] = "seq2seq"
            return self.forward_captioning(batch, is_decode=is_decode)
        if "retrieval" in task:
            out = self.forward_retrieval(batch)
        elif "qamc" in task:
            out = self.forward_qamc_ret(batch) if "lsmdc-mc" in task else self.forward_qamc(batch)
        elif "qaoe" in task:
            out = self.forward_qaoe(batch)
        else:
            raise NotImplementedError(f"forward() for {task}")
        return {"out": out[0], "ans": out[1]}

    def forward_captioning(self, batch, is_decode=False):
        return LAVENDER_Captioning.forward(self, batch, is_decode=is_decode)

    def forward_retrieval(self, batch):
        """main_multi_task_mlm.py:108-146: every clip i against every caption j (pair i*B + j), label `true` at the last
        text position when vid[i] == vid[j], else `false`."""
        img, txt, mask, vid = batch["img"], batch["txt"], batch["mask"], batch["vid"]
        B, _T, _, _H, _W = img.shape
        Lv = (1 + (_H // 32) * (_W // 32)) * _T
        feat_img, mask_img, feat_txt, mask_txt = self.go_feat(img, txt, mask)
        dev = img.device
        vi = torch.arange(B, device=dev).repeat_interleave(B)      # clip index of pair i*B + j
        ti = torch.arange(B, device=dev).repeat(B)                 # caption index
        p_txt, p_mask, p_feat = self.prepro_txt_inputs(txt[ti], mask_txt[ti], feat_txt[ti], task_name=batch["task_name"],
                                                       prompt=batch["prompt"])
        if torch.is_tensor(vid):
            same = (vid.reshape(-1)[vi] == vid.reshape(-1)[ti]).to(dev)
        else:
            same = torch.tensor([vid[i] == vid[j] for i in range(B) for j in range(B)], device=dev)
        ans = torch.full_like(p_txt, -1)
        ans[:, -1] = torch.where(same, torch.full_like(ans[:, -1], self.true_token_id),
                                 torch.full_like(ans[:, -1], self.false_token_id))
        out, _ = self.go_cross(feat_img[vi], mask_img[vi], p_feat, p_mask)
        return self.fc_mtm(out[:, Lv:]), ans

    def forward_qamc_ret(self, batch):
        """main_multi_task_mlm.py:148-177: multiple choice as retrieval — txt [B, O, X], each option paired with its clip."""
        img, txt, mask, ans = batch["img"], batch["txt"], batch["mask"], batch["mask_ans"]
        (B, _T, _, _H, _W), (_, O, _X) = img.shape, txt.shape
        Lv = (1 + (_H // 32) * (_W // 32)) * _T
        feat_img, mask_img, feat_txt, mask_txt = self.go_feat(img, txt.flatten(0, 1), mask.flatten(0, 1))
        vi = torch.arange(B, device=img.device).repeat_interleave(O)
        ans = ans.flatten(0, 1)
        prompt = batch["prompt"]
        ans, mask_txt, feat_txt = self.prepro_txt_inputs(ans, mask_txt, feat_txt, task_name=batch["task_name"], prompt=prompt)
        ans = ans.clone()
        ans[:, :self._prefix_len(prompt)] = -1
        out, _ = self.go_cross(feat_img[vi], mask_img[vi], feat_txt, mask_txt)
        return self.fc_mtm(out[:, Lv:]), ans.view(B, O, -1)

    def _forward_qa(self, batch):
        """main_multi_task_mlm.py:179-225 (forward_qamc and forward_qaoe are the same arithmetic)."""
        img, txt, mask, ans = batch["img"], batch["txt"], batch["mask"], batch["mask_ans"]
        _B, _T, _, _H, _W = img.shape
        Lv = (1 + (_H // 32) * (_W // 32)) * _T
        feat_img, mask_img, feat_txt, mask_txt = self.go_feat(img, txt, mask)
        prompt = batch["prompt"]
        ans, mask_txt, feat_txt = self.prepro_txt_inputs(ans, mask_txt, feat_txt, task_name=batch["task_name"], prompt=prompt)
        ans = ans.clone()
        ans[:, :self._prefix_len(prompt)] = -1
        out, _ = self.go_cross(feat_img, mask_img, feat_txt, mask_txt)
        return self.fc_mtm(out[:, Lv:]), ans

    forward_qamc = _forward_qa
    forward_qaoe = _forward_qa


TASK_NAME = (("retrieval", "vtm"), ("lsmdc-mc", "vtm"), ("qamc", "mc"), ("qaoe", "oe"), ("captioning", "cap"))


def add_task_token(batch):
    """Agent_Multi_Task.add_prompt_or_task_token (main_multi_task_mlm.py:254-275), task-token branch."""
    task = batch["task"]
    if "qamc" in task and "lsmdc-mc" in task:
        batch["task_name"] = "vtm"
        return batch
    for key, name in TASK_NAME:
        if key in task:
            batch["task_name"] = name
            return batch
    raise NotImplementedError(f"no task name for {task}")


def train_step(agent, batch, cap_loss=None):
    """Agent_Multi_Task.train_step (main_multi_task_mlm.py:369-390) on a native Agent_Base: forward, the task's loss
    (CaptioningLoss on the labelled caption rows, CE(ignore -1) otherwise), backward_step; returns device scalars."""
    agent.model.train()
    out = agent.forward_step(batch)
    logits, ans = out["out"], out["ans"]
    if "captioning" in batch["task"]:
        sel = ans != -1
        ls = (cap_loss or agent.__dict__.setdefault("_cap_loss", CaptioningLoss()))(logits[sel].float(), ans[sel])
    else:
        logits = logits.flatten(0, logits.dim() - 2)
        ans = ans.flatten(0, ans.dim() - 1)
        ls = agent.loss_func(logits, ans)
    agent.backward_step(ls)
    return ls.detach()
