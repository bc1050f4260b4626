"""Host-side logic of the native path that needs no GPU: index maps, shift-mask classes, state-dict schema,
optimizer grouping, parameter arena, VTM pair construction."""
import os

import numpy as np
import pytest
import torch

import lavender_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CPU = torch.device("cpu")


@pytest.mark.parametrize("tag,dims,win", [("w877_s0", (5, 56, 56), (8, 7, 7)), ("w877_s2", (5, 14, 14), (8, 7, 7)),
                                          ("w81212_s1", (5, 48, 48), (8, 12, 12))])
def test_window_row_map_matches_reference_partition(tag, dims, win):
    from lavender_b200.video_swin import get_window_size, window_row_map
    g = np.load(os.path.join(GOLD, "kat_index.npz"))
    D, H, W = dims
    ws, ss = get_window_size(dims, win, tuple(i // 2 for i in win))
    rmap = window_row_map(1, D, H, W, ws, ss, CPU)
    assert np.array_equal(rmap.numpy().reshape(-1, ws[0] * ws[1] * ws[2]), g[f"{tag}/part_src"])
    # batch index is the slowest dimension
    r2 = window_row_map(2, D, H, W, ws, ss, CPU)
    n = D * H * W
    assert torch.equal(r2[:n], rmap) and torch.equal(r2[n:], rmap + n)


@pytest.mark.parametrize("dims,win", [((5, 56, 56), (8, 7, 7)), ((5, 14, 14), (8, 7, 7)), ((5, 24, 24), (8, 12, 12))])
def test_shift_mask_classes_reproduce_compute_mask(dims, win):
    from lavender_b200.video_swin import get_window_size, shift_mask_classes
    D, H, W = dims
    ws, ss = get_window_size(dims, win, tuple(i // 2 for i in win))
    labels, cls_of = shift_mask_classes(D, H, W, ws, ss, CPU)
    mask = O.compute_mask(D, H, W, ws, ss)          # [nW, N, N] in {0, -100}
    N = ws[0] * ws[1] * ws[2]
    assert cls_of.numel() == mask.shape[0]
    for w in range(mask.shape[0]):
        l = labels[cls_of[w], :N].int()
        assert torch.equal(l[:, None] != l[None, :], mask[w] != 0), w


def test_unshifted_stage_has_no_mask_classes():
    from lavender_b200.video_swin import get_window_size, shift_mask_classes
    ws, ss = get_window_size((5, 7, 7), (8, 7, 7), (4, 3, 3))
    assert ss == (0, 0, 0) and shift_mask_classes(5, 7, 7, ws, ss, CPU) == (None, None)


def test_merge_row_map_matches_patch_merging_order():
    from lavender_b200.video_swin import merge_row_map
    B, D, H, W = 2, 3, 4, 6
    x = torch.arange(B * D * H * W, dtype=torch.float32).view(B, D, H, W, 1)
    ref = torch.cat([x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 1::2]], -1)
    assert torch.equal(merge_row_map(B, D, H, W, CPU).view(-1, 4).float(), ref.view(-1, 4))


def _model(layers=1, B=2, size="tiny"):
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
    args = default_args(vis_backbone_size=size, size_batch=B, bert_config={"num_hidden_layers": layers})
    return LAVENDER_Pretrain_MLM(args, FakeTokenizer()), args


def test_state_dict_schema_is_the_references():
    m, _ = _model(layers=2)
    cfg = O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=2)
    want = {k: tuple(s) for k, s in O.state_dict_schema(cfg)}
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    m.load_state_dict(O.make_state_dict(cfg, 0), strict=True)
    assert m.fc_mtm.predictions.decoder.bias is m.fc_mtm.predictions.bias      # one parameter, two keys


def test_optimizer_groups_follow_reference_name_rules():
    """agent.py:98-119 — group sizes probed on the reference for tiny + 2 layers: 81 / 23 / 90 / 27 tensors."""
    from lavender_b200.agent import Agent_Base
    m, args = _model(layers=2)
    ag = Agent_Base.__new__(Agent_Base)
    ag.model, ag.args = m, args
    opt = ag.build_optimizer()
    assert [len(g["params"]) for g in opt.param_groups] == [81, 23, 90, 27]
    assert [g["weight_decay"] for g in opt.param_groups] == [args.decay, args.decay, 0.0, 0.0]


def test_arena_views_and_fused_qkv_adjacency():
    from lavender_b200.arena import ParamArena
    m, _ = _model(layers=1)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    ar = ParamArena(m)
    assert ar.valid()
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n]), n
        assert p.data_ptr() == ar.flat.data_ptr() + 4 * ar.offsets[id(p)]
        assert ar.offsets[id(p)] % 8 == 0
    sa = m.trsfr.layer[0].attention.self
    w = ar.span32(sa.query.weight, sa.value.weight, (3 * 768, 768))
    assert torch.equal(w[:768], sa.query.weight) and torch.equal(w[768:1536], sa.key.weight) and torch.equal(w[1536:], sa.value.weight)
    b = ar.span32(sa.query.bias, sa.value.bias, (3 * 768,))
    assert torch.equal(b[768:1536], sa.key.bias)
    # in-place optimizer-style updates are visible through the flat buffer
    with torch.no_grad():
        sa.key.bias.add_(1.0)
    assert torch.equal(ar.span32(sa.query.bias, sa.value.bias, (3 * 768,))[768:1536], sa.key.bias)
    # gradient slices: prepare / finalize
    p = m.emb_task
    ar.prepare_grads([p])
    assert p.grad.data_ptr() == ar.g(p).data_ptr() and float(p.grad.abs().sum()) == 0.0
    q = m.enc_img.emb_cls
    q.grad = torch.ones_like(q)
    ar.finalize_grads()
    assert q.grad.data_ptr() == ar.g(q).data_ptr() and float(q.grad.sum()) == q.numel()


def test_vtm_pair_order_matches_reference_loop():
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM
    np.random.seed(5)
    negs = LAVENDER_Pretrain_MLM.draw_negatives(4, 3)
    np.random.seed(5)
    ref = O.draw_negatives(4, 3)
    assert [list(map(int, n)) for n in negs] == ref


def test_masking_matches_reference_semantics():
    from lavender_b200.agent import Agent_Pretrain_MLM
    ag = Agent_Pretrain_MLM.__new__(Agent_Pretrain_MLM)
    ag.cls_token_id, ag.sep_token_id, ag.pad_token_id, ag.mask_token_id = 101, 102, 0, 103
    txt = torch.randint(1000, 30000, (4, 33))
    txt[:, 0], txt[:, -2], txt[:, -1] = 101, 102, 103
    orig = txt.clone()
    torch.manual_seed(3)
    out = ag.masking(txt, torch.ones_like(txt), 0.5)
    # reference: per row i, positions where rand(X) < p and not special
    torch.manual_seed(3)
    for i in range(4):
        sel = (torch.rand(33) < 0.5) & ~((orig[i] == 101) | (orig[i] == 102) | (orig[i] == 0) | (orig[i] == 103))
        assert torch.equal(out["ans_mtm"][i][sel], orig[i][sel]) and (out["ans_mtm"][i][~sel] == -1).all()
        assert (out["txt"][i][sel] == 103).all() and torch.equal(out["txt"][i][~sel], orig[i][~sel])


def test_warmup_linear_lr():
    from lavender_b200.agent import WarmupLinearLR
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = WarmupLinearLR(opt, max_iter=100)
    lrs = []
    for _ in range(100):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert lrs[0] == pytest.approx(1e-8) and lrs[5] == pytest.approx(0.5) and lrs[10] == pytest.approx(1.0)
    assert lrs[55] == pytest.approx(0.5) and lrs[99] == pytest.approx(1 / 90, rel=1e-3)


def test_dropout_rng_sites_and_specs():
    """Host side of the kernel dropout: fresh site id per call, None for p == 0, reseed resets the step counter."""
    import torch
    from lavender_b200 import dropout as DR
    st = DR.DropoutRNG(torch.device("cpu"))
    assert st.spec(0.0) is None
    a, b = st.spec(0.1), st.spec(0.1)
    assert a[0] is st.state and a[2] == 0.1 and b[1] == a[1] + 1
    step0 = int(st.state[1])
    st.advance()
    assert int(st.state[1]) == step0 + 1
    torch.manual_seed(3)
    s1 = DR.DropoutRNG(torch.device("cpu")).state[0].item()
    torch.manual_seed(3)
    s2 = DR.DropoutRNG(torch.device("cpu")).state[0].item()
    assert s1 == s2   # torch.manual_seed makes the mask stream reproducible


def test_side_stream_helpers_are_noops_without_cuda():
    import torch
    from lavender_b200 import streams
    streams.join(torch.device("cpu"))   # nothing forked, no CUDA: must not raise
    assert isinstance(streams.enabled(), bool)
