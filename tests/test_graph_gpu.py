"""CUDA-graph replay of the training step == the eager step (same kernels, same inputs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(graph):
    import lavender_oracle as O
    from lavender_b200.agent import Agent_Pretrain_MLM
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
    cfg = O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=1)
    args = default_args(vis_backbone_size="tiny", size_batch=3, bert_config={"num_hidden_layers": 1}, cuda_graph=graph,
                        lr=1e-3, max_iter=10)
    m = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
    m.load_state_dict(O.make_state_dict(cfg, 0), strict=True)
    for c in (m.trsfr.config, m.enc_txt.emb_txt.config):
        c.lav_eval_dropout = True
    for layer in m.enc_img.swin.layers:          # DropPath off: the two runs must see the same arithmetic
        for blk in layer.blocks:
            blk.drop_path_rate = 0.0
    m.cuda()
    ag = Agent_Pretrain_MLM(args, m)
    batch = {k: v.cuda() for k, v in O.make_batch(3, seed=0).items()}
    return m, ag, batch


def test_graph_replay_matches_eager():
    from lavender_b200.graph import GraphCache
    m1, a1, b1 = _setup(False)
    m2, a2, b2 = _setup(True)
    a2.graphs = GraphCache(a2)
    g = a2.graphs.get(b2)            # capture now (warm-up + capture leave the weights untouched)
    assert g.native_launches > 100
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p, q), n
    # 1) gradients of one forward/backward: eager autograd vs one replay
    np.random.seed(3)
    m1.train()
    out = a1.forward_step(dict(b1))
    ls = a1.loss_func(out["out_mtm"].flatten(0, 1), out["ans_mtm"].flatten()) + \
        a1.loss_func(out["out_vtm"].flatten(0, 1), out["ans_vtm"].flatten())
    a1.scaler.scale(ls).backward()
    m1.arena().finalize_grads()
    np.random.seed(3)
    l_mtm, l_vtm = g(b2)
    torch.cuda.synchronize()
    assert abs((l_mtm + l_vtm).item() - ls.item()) < 1e-4
    worst = 0.0
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        if p.grad is None:
            assert q.grad is None or float(q.grad.abs().max()) == 0.0, n
            continue
        e = ((p.grad - q.grad).norm() / (p.grad.norm() + 1e-3 * 65536)).item()   # grads carry the 65536 loss scale
        worst = max(worst, e)
        assert e < 1e-3, (n, e)
    print("worst relative gradient difference eager vs graph:", worst)
    a1.optzr.zero_grad()
    # 2) three optimizer steps: same loss trajectory
    for it in range(3):
        np.random.seed(10 + it)
        r1 = a1.step(dict(b1), True)
        np.random.seed(10 + it)
        r2 = a2.step(dict(b2), True)
        assert abs(r1["mtm"] - r2["mtm"]) < 5e-3 and abs(r1["vtm"] - r2["vtm"]) < 5e-3, (it, r1, r2)
    assert len(a2.graphs.graphs) == 1


def test_pipelined_input_path_matches_step():
    """Agent.prefetch / step_async / finish (next batch copied under the running step, losses read one step late) gives
    the loss trajectory of plain Agent.step on the same host batches."""
    import lavender_oracle as O
    m1, a1, _ = _setup(True)
    m2, a2, _ = _setup(True)
    host = [{k: v.pin_memory() for k, v in O.make_batch(3, seed=s).items()} for s in range(4)]
    ref = []
    for it, hb in enumerate(host):
        np.random.seed(20 + it)
        ref.append(a1.step(a1.prepare_batch({k: v.clone() for k, v in hb.items()}), True))
    got, pend = [], None
    np.random.seed(20)
    h = a2.prefetch({k: v.clone().pin_memory() for k, v in host[0].items()})
    for it in range(len(host)):
        np.random.seed(20 + it)          # the VTM negatives are drawn from the numpy RNG when the step is enqueued
        cur = a2.step_async(h)
        if it + 1 < len(host):
            h = a2.prefetch({k: v.clone().pin_memory() for k, v in host[it + 1].items()})
        if pend is not None:
            got.append(a2.finish(pend))
        pend = cur
    got.append(a2.finish(pend))
    assert len(got) == len(ref)
    for it, (r, g) in enumerate(zip(ref, got)):
        assert abs(r["mtm"] - g["mtm"]) < 5e-3 and abs(r["vtm"] - g["vtm"]) < 5e-3, (it, r, g)


def test_training_loop_reduces_loss_on_a_fixed_batch():
    """End-to-end sanity of the whole step (graph replay, side-stream weight gradients, dropout, clip + fused AdamW,
    loss scaling): 30 optimizer steps on one fixed batch must drive both losses down from ~ln(vocab)."""
    import lavender_oracle as O
    from lavender_b200.agent import Agent_Pretrain_MLM
    from lavender_b200.pretrain import LAVENDER_Pretrain_MLM, FakeTokenizer, default_args
    torch.manual_seed(0)
    cfg = O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=1)
    args = default_args(vis_backbone_size="tiny", size_batch=3, bert_config={"num_hidden_layers": 1}, cuda_graph=True,
                        lr=2e-4, max_iter=1000)
    m = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
    m.load_state_dict(O.make_state_dict(cfg, 0), strict=True)
    m.cuda()
    ag = Agent_Pretrain_MLM(args, m)
    batch = {k: v.cuda() for k, v in O.make_batch(3, seed=0).items()}
    np.random.seed(0)
    first = ag.step(dict(batch), True)
    for _ in range(29):
        last = ag.step(dict(batch), True)
    print("loss trajectory:", first, "->", last)
    assert np.isfinite(last["mtm"]) and np.isfinite(last["vtm"])
    assert last["mtm"] < first["mtm"] - 1.0 and last["vtm"] < first["vtm"] - 1.0, (first, last)
