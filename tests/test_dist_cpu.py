"""Data-parallel plumbing on the gloo backend, world_size 2 (CPU): the flat-arena gradient all-reduce (early BERT/head
piece + late piece) averages gradients exactly like DDP, parameters are broadcast from rank 0, env parsing."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.enc_img = torch.nn.Linear(8, 8)
        self.trsfr = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 8))
        self.fc_mtm = torch.nn.Linear(8, 4)
        self.emb_unused = torch.nn.Parameter(torch.ones(3))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from lavender_b200 import dist as D
    from lavender_b200.arena import ParamArena
    from lavender_b200.config import Args
    args = Args(seed=3)
    D.dist_init(args, distributed=True, backend="gloo")
    assert args.distributed and args.num_gpus == world and args.local_rank == rank
    assert D.get_rank() == rank and D.get_world_size() == world and D.is_main_process() == (rank == 0)
    torch.manual_seed(100 + rank)          # different weights per rank before the broadcast
    m = Toy()
    ar = ParamArena(m)
    D.broadcast_parameters(ar)
    sync = D.GradSync(ar, early_prefixes=("trsfr.", "fc_mtm."))
    assert len(sync.early) == 1 and len(sync.late) == 1      # trsfr + fc_mtm are adjacent, at the end of the arena
    w0 = ar.flat.clone()
    # per-rank data -> per-rank gradients
    torch.manual_seed(7 + rank)
    x = torch.randn(5, 8)
    loss = m.fc_mtm(m.trsfr(m.enc_img(x))).pow(2).mean()
    loss.backward()
    local = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    sync.start_early()          # what the Swin backward hook does
    n = sync.finish()
    assert n == ar.total
    out = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    gathered = D.all_gather({"rank": rank, "local": local, "w0": w0})
    assert [g["rank"] for g in gathered] == list(range(world))
    for k in out:
        mean = sum(g["local"][k] for g in gathered) / world
        assert torch.allclose(out[k], mean, atol=1e-6), k
        pk = dict(m.named_parameters())[k]
        assert pk.grad.data_ptr() == ar.g(pk).data_ptr()   # grads live in the arena
    assert m.emb_unused.grad is None
    assert torch.equal(gathered[0]["w0"], gathered[1]["w0"])                           # broadcast worked
    # second step without the early piece (graph mode): one reduction of the whole arena
    for p in m.parameters():
        p.grad = None
    loss = m.fc_mtm(m.trsfr(m.enc_img(x))).pow(2).mean()
    loss.backward()
    sync.finish()
    for k in out:
        assert torch.allclose(dict(m.named_parameters())[k].grad, out[k], atol=1e-6), k
    red = D.reduce_dict({"a": torch.tensor(float(rank)), "b": torch.tensor(2.0)})
    if rank == 0:
        assert abs(red["a"].item() - 0.5) < 1e-6 and abs(red["b"].item() - 2.0) < 1e-6
    D.synchronize()
    torch.distributed.destroy_process_group()
    q.put((rank, "ok"))


def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, "ok"), (1, "ok")]


def test_env_parsing_openmpi(monkeypatch):
    from lavender_b200 import dist as D
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("OMPI_COMM_WORLD_SIZE", "8")
    monkeypatch.setenv("OMPI_COMM_WORLD_RANK", "5")
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_RANK", "5")
    assert (D.get_world_size(), D.get_rank(), D.get_local_rank()) == (8, 5, 5) and not D.is_main_process()
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    assert (D.get_world_size(), D.get_rank()) == (2, 0)
