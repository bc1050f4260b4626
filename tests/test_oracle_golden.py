"""Pins the CPU oracle (oracle/lavender_oracle.py) against vectors produced by the UNMODIFIED reference
(oracle/make_golden.py ran /root/reference's LAVENDER_Pretrain_MLM and visbackbone.video_swin helpers in the build
container and committed the sub-sampled outputs to tests/golden/)."""
import os

import numpy as np
import pytest
import torch

import lavender_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_kat_partition_mask_relidx():
    g = np.load(os.path.join(GOLD, "kat_index.npz"))
    for tag, (D, H, W), win in (("w877_s0", (5, 56, 56), (8, 7, 7)), ("w877_s2", (5, 14, 14), (8, 7, 7)),
                                ("w81212_s1", (5, 48, 48), (8, 12, 12))):
        shift = tuple(i // 2 for i in win)
        ws, ss = O.get_window_size((D, H, W), win, shift)
        assert tuple(g[f"{tag}/ws"]) == ws and tuple(g[f"{tag}/ss"]) == ss
        ids = torch.arange(D * H * W, dtype=torch.float32).view(1, D, H, W, 1)
        rolled = torch.roll(ids, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        part = O.window_partition(rolled, ws).squeeze(-1).long().numpy().astype(np.int32)
        assert np.array_equal(part, g[f"{tag}/part_src"])                                   # T1 / T4
        m = O.compute_mask(D, H, W, ws, ss)
        assert np.array_equal((m != 0).numpy().reshape(m.shape[0], -1)[:, ::97], g[f"{tag}/mask_nz"])   # T3
        assert np.array_equal(m.sum((1, 2)).numpy(), g[f"{tag}/mask_sum"])
        N = ws[0] * ws[1] * ws[2]
        rel = O.relative_position_index(win)[:N, :N].numpy().astype(np.int32)
        assert np.array_equal(rel[::5, ::3], g[f"{tag}/relidx"])                              # T2
        # closed form of SURVEY T2
        a = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), torch.arange(ws[2]), indexing="ij"), -1).view(-1, 3)
        d = a[:, None, :] - a[None, :, :]
        closed = ((d[..., 0] + win[0] - 1) * (2 * win[1] - 1) + (d[..., 1] + win[1] - 1)) * (2 * win[2] - 1) + d[..., 2] + win[2] - 1
        assert np.array_equal(closed.numpy(), rel)


def test_window_partition_reverse_roundtrip():
    x = torch.randn(2, 5, 14, 14, 6)
    ws = (5, 7, 7)
    assert torch.equal(O.window_reverse(O.window_partition(x, ws), ws, 2, 5, 14, 14), x)


CASES = [("tiny_l2_b2", "tiny", 2, 2, True, 0), ("tiny_l1_b3_notask", "tiny", 1, 3, False, 3),
         # the benchmarked architecture (swin_base + 12-layer BERT-base), vt_mask on the last clip
         ("base_l12_b2", "base", 12, 2, True, 5)]


@pytest.mark.parametrize("name,size,layers,B,task,seed", CASES)
def test_oracle_reproduces_reference_outputs(name, size, layers, B, task, seed):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.ModelCfg(swin=O.SWIN[size], bert_layers=layers, enable_task_token=task, vtm_batch=min(B, 4))
    sd = O.make_state_dict(cfg, seed)
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    sd["fc_mtm.predictions.decoder.bias"] = sd["fc_mtm.predictions.bias"]
    batch = O.make_batch(B, seed=seed)
    if "vt_mask" in gold.files:
        batch["vt_mask"] = torch.from_numpy(gold["vt_mask"])
    np.random.seed(1 + seed)
    out = O.pretrain_forward(sd, batch, cfg)
    loss, l1, l2 = O.pretrain_loss(out)
    loss.backward()
    assert np.array_equal(out["ans_vtm"].numpy(), gold["ans_vtm"])
    assert np.abs(out["out_mtm"].detach()[..., ::61].numpy() - gold["out_mtm_s"]).max() < 2e-4
    assert np.abs(out["out_vtm"].detach()[..., ::61].numpy() - gold["out_vtm_s"]).max() < 2e-4
    assert abs(l1.item() - float(gold["ls_mtm"])) < 1e-4 and abs(l2.item() - float(gold["ls_vtm"])) < 1e-4
    for k, v in sd.items():
        if "gn/" + k not in gold.files:
            continue
        gn = float(gold["gn/" + k])
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        assert abs(g.double().norm().item() - gn) <= 1e-3 * gn + 1e-6, k
        assert np.abs(O.sample_flat(g).numpy() - gold["gs/" + k]).max() <= 1e-3 * max(gn, 1e-6) + 1e-7, k


def test_feature_goldens():
    gold = np.load(os.path.join(GOLD, "tiny_l2_b2.npz"))
    cfg = O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=2)
    sd = O.make_state_dict(cfg, 0)
    batch = O.make_batch(2, seed=0)
    with torch.no_grad():
        sw = O.swin_forward(sd, "enc_img.swin.", batch["img"].transpose(1, 2), cfg.swin)
        fi, _ = O.enc_video(sd, batch["img"], cfg)
        ft = O.bert_embeddings(sd, batch["txt"])
    assert np.abs(sw[..., ::7].numpy() - gold["swin_out_s"]).max() < 2e-4
    assert np.abs(fi[:, ::5, ::3].numpy() - gold["feat_img_s"]).max() < 2e-4
    assert np.abs(ft[..., ::3].numpy() - gold["feat_txt_s"]).max() < 1e-5


def test_enc_video_odr_vt_mask_golden():
    """EncVideo with a frame-order list and a video key mask (model.py:72-91) on the base backbone (covers EncVideo.fc)."""
    gold = np.load(os.path.join(GOLD, "base_l12_b2.npz"))
    cfg = O.ModelCfg(swin=O.SWIN["base"], bert_layers=12, vtm_batch=2)
    sd = O.make_state_dict(cfg, 5)
    batch = O.make_batch(2, seed=5)
    odr = [[0, 2, 1, 3, 4], [4, 1, 2, 3, 0]]
    with torch.no_grad():
        f, m = O.enc_video(sd, batch["img"], cfg, odr=odr, vt_mask=torch.from_numpy(gold["vt_mask"]))
    assert np.abs(f[:, ::5, ::3].numpy() - gold["feat_img_odr_s"]).max() < 2e-4
    assert np.array_equal(m.numpy(), gold["mask_img_odr"])


MT_TASKS = (("msrvtt-retrieval", "vtm", dict(B=3, X=25)), ("msvd-qaoe", "oe", dict(B=2, X=30)),
            ("tgif-qamc", "mc", dict(B=2, X=40)), ("lsmdc-mc-qamc", "vtm", dict(B=2, X=25, O_=5)),
            ("msrvtt-captioning", "cap", dict(B=2, X=20)))


def test_oracle_multitask_forwards_reproduce_reference():
    """BASELINE configs[4]: the five LAVENDER_Multi_Task forward variants (retrieval B^2, QA-OE, QA-MC, QA-MC as
    retrieval, seq2seq captioning) against outputs of the unmodified reference (oracle/make_golden.py multitask)."""
    gold = np.load(os.path.join(GOLD, "mt_tiny_l2.npz"))
    cfg = O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=2, enable_task_token=True)
    sd = O.make_state_dict(cfg, 11)
    ce = torch.nn.CrossEntropyLoss(ignore_index=-1)
    for ti, (task, tname, kw) in enumerate(MT_TASKS):
        batch = O.make_multitask_batch(task, seed=11 + ti, **kw)
        with torch.no_grad():
            lo, an = O.multitask_forward(sd, batch, cfg, task, tname)
        assert np.array_equal(an.numpy(), gold[f"{task}/ans"]), task
        assert np.abs(lo[..., ::61].numpy() - gold[f"{task}/out_s"]).max() < 2e-4, task
        loss = ce(lo.flatten(0, lo.dim() - 2), an.flatten())
        assert abs(loss.item() - float(gold[f"{task}/loss"])) < 1e-4, task
    batch = O.make_multitask_batch("msvd-qaoe", seed=12, B=2, X=30)
    with torch.no_grad():
        lo, _ = O.multitask_forward(sd, batch, cfg, "msvd-qaoe", "oe", decoder=True)
    assert np.abs(lo[..., ::61].numpy() - gold["msvd-qaoe/decoder/out_s"]).max() < 2e-4
