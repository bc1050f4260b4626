"""Data-parallel plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Replaces utils/dist.py (process-group bootstrap + control-plane helpers, same function names) and the
DDP / DeepSpeed ZeRO-1 wrap of agent.py:252-265 + utils/deepspeed.py.  The forward/backward of LAVENDER shards by
batch with no data-path exchange; the only collective is ONE all-reduce of the flat fp32 gradient
arena per step (`GradSync`), issued in three pieces so that all but ~2 % of it overlaps the Video-Swin backward:
  early  fusion BERT + MLM head + text embeddings + EncVideo parameters - final once autograd reaches the Swin backward
  mid    Swin stages 2-3                                                 - final when the backward enters stage 1
  late   Swin stages 0-1, patch embed, the rest                          - after backward (the only exposed part)
In CUDA-graph mode the collectives are captured inside the step's graph as parallel branches (graph.py).
NCCL picks NVLS (in-switch reduction) or ring on the NVSwitch domain; nothing here depends on the link count.
The same code runs on the `gloo` backend with CPU tensors, which is how the CPU tests cover world_size 2.
"""
import datetime
import os
import pickle
import random

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------
# environment (utils/dist.py:81-111): torchrun variables first, then OpenMPI's
# ---------------------------------------------------------------------------------------------------------
def _env_int(names, default):
    for n in names:
        if n in os.environ:
            return int(os.environ[n])
    return default


def get_world_size():
    return _env_int(("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE"), 1)


def get_rank():
    return _env_int(("RANK", "OMPI_COMM_WORLD_RANK"), 0)


def get_local_rank():
    return _env_int(("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"), 0)


def get_local_size():
    return _env_int(("LOCAL_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE"), 1)


def is_main_process():
    return get_rank() == 0


def set_seed(seed, n_gpu=0):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if n_gpu > 0 and torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def dist_init(args, distributed=True, backend=None):
    """utils/dist.py:20-75: fills args.num_gpus / distributed / local_rank, binds the GPU, creates the process group
    (env:// rendezvous; MASTER_ADDR defaults to 127.0.0.1 because container hostnames may not resolve), seeds."""
    world = get_world_size() if distributed else 1
    has_env = any(k in os.environ for k in ("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE"))
    if distributed and has_env:
        args.num_gpus = world
        args.local_rank = get_local_rank()
        args.distributed = True if "WORLD_SIZE" in os.environ else world > 1
        if "OMPI_COMM_WORLD_SIZE" in os.environ:
            args.num_nodes = max(1, world // 8)
        if args.distributed and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "12345")
            os.environ.setdefault("RANK", str(get_rank()))
            os.environ.setdefault("WORLD_SIZE", str(world))
            use_cuda = torch.cuda.is_available()
            if use_cuda:
                torch.cuda.set_device(args.local_rank)
            dist.init_process_group(backend=backend or ("nccl" if use_cuda else "gloo"), init_method="env://",
                                    timeout=datetime.timedelta(hours=5))
    else:
        args.num_gpus = torch.cuda.device_count() if not distributed else 1
        args.distributed = False
    set_seed(getattr(args, "seed", 0), args.num_gpus)


def synchronize():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()


def _comm_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather(data):
    """utils/dist.py:187-227: gathers arbitrary picklable objects (evaluation only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [data]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, data)
    return out


def reduce_dict(input_dict, average=True):
    """utils/dist.py:230-257: reduces a dict of scalar tensors onto rank 0."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k] for k in names], dim=0)
        dist.reduce(values, dst=0)
        if dist.get_rank() == 0 and average:
            values /= dist.get_world_size()
        return dict(zip(names, values))


class NoOp(object):
    """utils/dist.py:260-266."""

    def __getattr__(self, name):
        return self.noop

    def noop(self, *args, **kwargs):
        return


def iter_tqdm(item):
    if is_main_process():
        try:
            from tqdm import tqdm
            return tqdm(item, ascii=True)
        except ImportError:
            pass
    return item


# ---------------------------------------------------------------------------------------------------------
# gradient all-reduce on the flat arena
# ---------------------------------------------------------------------------------------------------------
class GradSync:
    """Averages the flat gradient buffer of a ParamArena across ranks (DDP semantics: mean over ranks,
    agent.py:261-265) in three pieces, each issued on a communication side stream as soon as its gradients are final, so
    that only the last (small) one is exposed:
      early  fusion BERT + MLM head + text embeddings + EncVideo's own parameters - final when autograd reaches the video
             encoder's backward (`arena.on_swin_backward`, called by _SwinFn.backward);      ~134 M parameters (base)
      mid    Swin stages 2-3 + the final Swin norm - final when the backward enters stage 1
             (`arena.on_swin_stage`);                                                        ~83 M parameters
      late   everything else (Swin stages 0-1, patch embed, emb_task ...) after backward;       ~4 M parameters
    The spans are contiguous arena ranges.  Under CUDA-graph capture (graph.py) the side-stream collectives become
    parallel branches of the captured step; `finish()` joins them."""

    def __init__(self, arena, group=None,
                 early_prefixes=("trsfr.", "fc_mtm.", "enc_txt.", "enc_img.fc.", "enc_img.norm.", "enc_img.emb_"),
                 mid_prefixes=("enc_img.swin.layers.2.", "enc_img.swin.layers.3.", "enc_img.swin.norm.")):
        self.arena, self.group = arena, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.early = self._ranges(early_prefixes)
        self.mid = self._ranges(mid_prefixes)
        self.late = self._complement(self.early + self.mid)
        self._early_done = self._mid_done = False
        self._stream = None
        self._event = None
        arena.on_swin_backward = self.start_early
        arena.on_swin_stage = self.on_stage

    def _ranges(self, prefixes):
        a = self.arena
        spans = []
        for n, p in zip(a.names, a.params):
            if any(n.startswith(pf) for pf in prefixes):
                o = a.offsets[id(p)]
                e = o + (p.numel() + 7) // 8 * 8
                if spans and spans[-1][1] == o:
                    spans[-1][1] = e
                else:
                    spans.append([o, e])
        return [tuple(s) for s in spans]

    def _complement(self, spans):
        out, pos = [], 0
        for o, e in sorted(spans):
            if o > pos:
                out.append((pos, o))
            pos = max(pos, e)
        if pos < self.arena.total:
            out.append((pos, self.arena.total))
        return out

    def _reduce(self, spans):
        g = self.arena.grad
        for o, e in spans:
            t = g[o:e]
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.div_(self.world)

    def _side(self, spans):
        """Reduce `spans` on the communication stream, ordered after everything enqueued so far on the current stream AND
        on the weight-gradient side stream.  The critical path itself does not wait for the weight gradients here (joining
        them into it at both overlap points cost ~0.5 ms per step); the span's gradients are all written by native
        kernels straight into the arena, so no autograd-side copy (finalize_grads) is needed before the reduction."""
        if self.arena.grad.is_cuda:
            from . import streams
            if self._stream is None:
                self._stream = torch.cuda.Stream()
                self._event = torch.cuda.Event()
            cur = torch.cuda.current_stream()
            self._stream.wait_stream(cur)
            wg = streams.pending_stream(self.arena.grad.device)
            if wg is not None:
                self._stream.wait_stream(wg)
            with torch.cuda.stream(self._stream):
                self._reduce(spans)
                self._event.record(self._stream)
        else:
            self.arena.finalize_grads()
            self._reduce(spans)

    def start_early(self):
        """Called when the Swin backward begins: BERT / head / text-embedding gradients are complete."""
        if self.world < 2 or self._early_done or not self.early:
            return
        self._side(self.early)
        self._early_done = True

    def on_stage(self, s):
        """Called by the Swin backward when it enters stage `s` (3, 2, 1, 0): at s == 1 stages 2-3 are complete."""
        if self.world < 2 or self._mid_done or not self.mid or s != 1:
            return
        self._side(self.mid)
        self._mid_done = True

    def finish(self):
        """After backward: reduce what is left and join the side stream.  Returns the number of elements reduced."""
        self.arena.finalize_grads()
        if self.world < 2:
            self._early_done = self._mid_done = False
            return 0
        rest = list(self.late) + ([] if self._early_done else list(self.early)) + ([] if self._mid_done else list(self.mid))
        if self._event is not None and (self._early_done or self._mid_done):
            torch.cuda.current_stream().wait_event(self._event)
        if not (self._early_done or self._mid_done):
            rest = [(0, self.arena.total)]
        self._reduce(sorted(rest))
        self._early_done = self._mid_done = False
        return self.arena.total


def broadcast_parameters(arena, src=0, group=None):
    """All ranks start from rank `src`'s weights (what DDP's constructor does, agent.py:261)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(arena.flat, src=src, group=group)
        arena._ver16 = None
