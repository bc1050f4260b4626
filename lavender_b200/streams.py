"""Side stream for the weight-gradient GEMMs.

In backward every nn.Linear needs two products of the same dY: dgrad (on the critical path: the next layer waits for
it) and wgrad (+ bias gradient; nobody waits for it until the optimizer).  The persistent GEMM kernels leave SMs idle
in their last wave (3.35 waves of tiles on 148 SMs) and every kernel has a serial prologue / tail; wgrad launched on
a second stream fills those holes instead of extending the critical path.  Under CUDA-graph capture the fork / join
become parallel branches of the graph.

    fork(dev)        -> stream on which to launch (it waits for everything enqueued on the current stream so far)
    hold(*tensors)      keeps operands alive until join() (they were allocated on the main stream: the caching
                        allocator must not hand their memory out while the side stream still reads it)
    join(dev)           the current stream waits for the side stream; queued as an end-of-backward callback of the
                        autograd engine and called by ParamArena.finalize_grads(), i.e. before anything reads the
                        gradient arena (all-reduce, clipping, optimizer, end of a graph capture)
LAV_WGRAD_STREAM=0 disables it (everything on the current stream).
"""
import os

import torch

_ENABLED = os.environ.get("LAV_WGRAD_STREAM", "1") != "0"
_STATE = {}


class _Side:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.held = []
        self.dirty = False
        self.cb_queued = False


def enabled():
    return _ENABLED


def _get(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _STATE.get(key)
    if st is None:
        st = _STATE[key] = _Side(device)
    return st


def fork(device):
    st = _get(device)
    st.stream.wait_stream(torch.cuda.current_stream(device))
    st.dirty = True
    if not st.cb_queued:
        try:   # inside autograd's backward: join automatically when this backward pass ends, so that code which reads
            #    p.grad right after loss.backward() (a plain torch optimizer, clip_grad_norm_) is ordered after wgrad
            torch.autograd.Variable._execution_engine.queue_callback(lambda: join(device))
            st.cb_queued = True
        except RuntimeError:
            pass   # not in a backward pass (forward-side use, direct functional calls): the caller synchronises
    return st.stream


def hold(device, *tensors):
    _get(device).held.extend(t for t in tensors if t is not None)


def join(device):
    if not _ENABLED or not torch.cuda.is_available():
        return
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _STATE.get(key)
    if st is None or not st.dirty:
        return
    torch.cuda.current_stream(device).wait_stream(st.stream)
    st.held.clear()
    st.dirty = False
    st.cb_queued = False


def pending_stream(device):
    """The weight-gradient side stream if work has been forked onto it since the last join, else None.  Lets another
    stream (the gradient all-reduce's communication stream) wait for the weight gradients WITHOUT joining them into the
    critical path."""
    if not _ENABLED or not torch.cuda.is_available():
        return None
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _STATE.get(key)
    return st.stream if st is not None and st.dirty else None
