"""Unified-MLM pre-training model: MLM + video-text-matching-as-MLM (main_pretrain_mlm.py:42-119).

`LAVENDER_Pretrain_MLM` keeps the reference's constructor, attributes (`fc_mtm`, `emb_task`, `task_tok2id`,
`vtm_batch`) and forward contract (dict batch in -> dict of logits / labels out) but builds the B*_O VTM pairs
with index gathers instead of the reference's per-sample Python loop; the pair order, the numpy
RNG call sequence (main_pretrain_mlm.py:90) and therefore the outputs are identical.
The reference's own `LAVENDER_Pretrain_MLM` (the unchanged script) also runs on top of lavender_b200.model —
see INTEGRATION.md.
"""
import os
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from .model import LAVENDER_Base, build_mlm_head


class LAVENDER_Pretrain_MLM(LAVENDER_Base):
    def __init__(self, args, tokzr=None):
        super().__init__(args, tokzr)
        self.patch_size = args.size_patch
        self.fc_mtm, _ = build_mlm_head(args.tokenizer, args)
        self.vtm_batch = min(args.size_batch, 4)
        self.task_tok2id = {"vtm": 0, "mc": 1, "oe": 2, "cap": 3}
        self.emb_task = nn.Parameter(0.02 * torch.randn(10, self.hidden_size))
        # one merged fusion-encoder pass for MLM + VTM (see forward); LAV_MERGE_PASSES=0 keeps the two passes
        self.merge_passes = os.environ.get("LAV_MERGE_PASSES", "1") != "0"
        self.vtm_last_token_only = os.environ.get("LAV_VTM_FULL_LOGITS", "0") != "1"
        self.merge_heads = os.environ.get("LAV_MERGE_HEADS", "1") != "0"   # MLM + VTM rows through one head pass

    @staticmethod
    def draw_negatives(B, O):
        """One np.random.permutation per clip, exactly as main_pretrain_mlm.py:90-91 consumes the numpy RNG."""
        return [np.random.permutation([j for j in range(B) if j != i])[:max(O - 1, 0)] for i in range(B)]

    def build_vtm_pairs(self, B, O, negs=None, device=None):
        """(video index, text index, label token) of the B*O pairs, in the reference's order."""
        if negs is None:
            negs = self.draw_negatives(B, O)
        vid_idx, txt_idx, label = [], [], []
        for i in range(B):
            vid_idx += [i] * O
            txt_idx += [i] + [int(j) for j in negs[i][:O - 1]]
            label += [self.true_token_id] + [self.false_token_id] * (O - 1)
        return (torch.tensor(vid_idx, device=device), torch.tensor(txt_idx, device=device),
                torch.tensor(label, device=device))

    def forward(self, batch):
        batch = defaultdict(lambda: None, batch)
        img, txt, mask = batch["img"], batch["txt"], batch["mask"]
        B, T, _, H, W = img.shape
        Lv = (1 + (H // self.patch_size) * (W // self.patch_size)) * T
        O = min(B, self.vtm_batch)

        feat_img, mask_img, feat_txt, mask_txt = self.go_feat(img, txt, mask, vt_mask=batch["vt_mask"])

        # VTM: clip i paired with its own caption (label "true") then with O-1 other captions ("false")
        dev = img.device
        if batch["vtm_vid_idx"] is not None:     # pre-built on the device (CUDA-graph replay: see graph.py)
            vi, ti, lab = batch["vtm_vid_idx"], batch["vtm_txt_idx"], batch["vtm_labels"]
        else:
            vi, ti, lab = self.build_vtm_pairs(B, O, batch["vtm_negatives"], dev)
        p_txt, p_mask, p_feat = self.prepro_txt_inputs(txt[ti], mask_txt[ti], feat_txt[ti], task_name="vtm",
                                                       prompt=batch["vtm_prompt"])
        ans_vtm = torch.full_like(p_txt, -1)
        ans_vtm[:, -1] = lab.to(ans_vtm.dtype)

        if not self.merge_passes:   # the reference's two fusion-encoder passes (main_pretrain_mlm.py:68-69, 112-115)
            out, _ = self.go_cross(feat_img, mask_img, feat_txt, mask_txt)
            out_mtm = self.fc_mtm(out[:, Lv:])
            out, _ = self.go_cross(feat_img[vi], mask_img[vi], p_feat, p_mask)
            out_vtm = self.fc_mtm(out[:, Lv:])
            return {"out_vtm": out_vtm, "out_mtm": out_mtm, "ans_vtm": ans_vtm, "ans_mtm": batch["ans_mtm"]}

        # ONE fusion-encoder pass over the B MLM sequences and the B*O VTM sequences (same weights, independent
        # sequences): the MLM sequences are padded behind the video tokens with as many key-masked dummy tokens as
        # the VTM sequences carry task / prompt tokens (one with enable_task_token).  A masked key has attention
        # weight exactly 0 and the dummy rows' outputs are dropped, so every real token sees what it sees in the
        # reference's separate pass — but the small MLM pass (8 x 283 rows: GEMMs of 18 row blocks) no longer runs
        # its own 12 layers of under-filled kernels.
        Lt, Lp = feat_txt.shape[1], p_feat.shape[1]
        d = Lp - Lt
        Hh = feat_txt.shape[-1]
        pad_f = feat_txt.new_zeros(B, d, Hh)
        pad_m = mask_txt.new_zeros(B, d)
        feat = torch.cat([torch.cat([feat_img, pad_f, feat_txt], dim=1), torch.cat([feat_img[vi], p_feat], dim=1)], dim=0)
        amask = torch.cat([torch.cat([mask_img, pad_m, mask_txt], dim=1), torch.cat([mask_img[vi], p_mask], dim=1)], dim=0)
        out = self.trsfr(feat, amask, output_attentions=True)["last_hidden_state"]
        ans_mtm = batch["ans_mtm"]
        rows = batch["mtm_rows"]
        if self.training and rows is not None:
            # the 30522-wide head only on the MLM rows that carry a label (ans != -1, ~15 % of B*Lt;
            # main_pretrain_mlm.py:158-160 ignores the rest).  `mtm_rows` is a fixed-capacity list of flat row indices
            # (b*Lt + t) built on the host by Agent_Pretrain_MLM.masking next to ans_mtm (padding entries point at an
            # unlabelled row), so the gather has a static shape and the step stays CUDA-graph capturable; loss and
            # gradients are identical because unlabelled rows contribute nothing to the cross-entropy.
            h = out[:B, Lv + d:].reshape(B * Lt, Hh)
            ans_mtm = ans_mtm.reshape(-1).index_select(0, rows).unsqueeze(0)     # [1, capacity] (-1 on padding)
            if self.vtm_last_token_only and self.merge_heads:
                # both heads in ONE pass over the capacity + B*O labelled rows (BertOnlyMLMHead.forward_split)
                hm = torch.cat([h.index_select(0, rows), out[B:, -1]], 0)
                out_mtm, out_vtm = self.fc_mtm.forward_split(hm, rows.numel())
                return {"out_vtm": out_vtm.unsqueeze(1), "out_mtm": out_mtm.unsqueeze(0), "ans_vtm": ans_vtm[:, -1:],
                        "ans_mtm": ans_mtm}
            out_mtm = self.fc_mtm(h.index_select(0, rows)).unsqueeze(0)          # [1, capacity, vocab]
        else:
            out_mtm = self.fc_mtm(out[:B, Lv + d:])   # (two head calls: one merged logits tensor would make autograd
            #                                            materialise two zero-padded [rows, vocab] gradients and add them)
        if self.training and self.vtm_last_token_only:
            # in training only the last text position of a VTM sequence carries a label (ans_vtm is -1
            # elsewhere, main_pretrain_mlm.py:103-105), so the 30522-wide head runs on those B*O rows instead of
            # B*O*34; loss and gradients are identical (unlabelled rows contribute nothing to the cross-entropy).
            # eval() keeps the reference's full [B*O, Lt, vocab] logits.
            out_vtm = self.fc_mtm(out[B:, -1:])
            ans_vtm = ans_vtm[:, -1:]
        else:
            out_vtm = self.fc_mtm(out[B:, Lv:])
        return {"out_vtm": out_vtm, "out_mtm": out_mtm, "ans_vtm": ans_vtm, "ans_mtm": ans_mtm}


class FakeTokenizer:
    """Stand-in for AutoTokenizer.from_pretrained('bert-base-uncased') when no vocabulary file is available
    offline: the special-token ids are bert-base-uncased's; 'true' / 'false' are 2995 / 6270."""
    cls_token, sep_token, pad_token, mask_token, unk_token = "[CLS]", "[SEP]", "[PAD]", "[MASK]", "[UNK]"
    vocab = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103, "true": 2995, "false": 6270}

    def convert_tokens_to_ids(self, toks):
        return [self.vocab[t] for t in toks]


def default_args(**kw):
    """The hot-path-relevant keys of utils/args.py (SURVEY §5 'Config / flags') with the reference's defaults."""
    from .config import Args
    a = Args(vis_backbone_size="base", size_img=224, size_frame=5, size_txt=32, size_batch=8, size_patch=32,
             max_size_frame=6, max_size_patch=14, vis_backbone_init="random", kinetics=400,
             txt_backbone="bert-base-uncased", tokenizer="bert-base-uncased", fusion_encoder="bert-base-uncased",
             fusion_encoder_rand_init=False, txt_backbone_embed_only=True, use_checkpoint=False,
             enable_task_token=True, enable_prompt=False, deepspeed=False, max_grad_norm=1.0, lr=2e-5, decay=1e-3,
             vis_backbone_lr_mul=1.0, max_iter=1000, p_mask=0.15, seed=0, logging_steps=100, distributed=False,
             task="pretrain", path_output="./_snapshot")
    a.update(kw)
    return a
