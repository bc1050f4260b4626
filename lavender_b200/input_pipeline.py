"""GPU input pipeline: the pixel work of the reference's loader on the device.

The reference decodes base64 JPEG frames and resizes / crops / normalises them with PIL + torchvision in 4 DataLoader
worker processes (dataset.py:118-186, 288-291): at the native path's > 270 clips/s per GPU that is ~1400 frames/s of PIL
resampling per GPU.  Here only the JPEG entropy decode stays on the host (cv2.imdecode, exactly as `Dataset_Base.str2img`,
dataset.py:177-186); the decoded uint8 frames go to the GPU through a pinned, double-buffered staging area and ONE kernel
(`lav_frames_resize_crop_norm_u8`, csrc/frames.cu) does Resize(size_img) -> crop -> ToTensor -> Normalize with Pillow's
exact two-pass antialiased bilinear arithmetic, writing the `[T, 3, S, S]` fp32 clip the model consumes.

    tf = GpuClipTransform(size_img=224)
    clip = tf(frames)                       # frames: list of T uint8 HWC RGB arrays of one size -> cuda fp32 [T,3,224,224]
    loader = GpuBatchLoader(tf, source)     # source yields lists of clips (each a list of JPEG byte strings)
    for img in loader: ...                  # [B, T, 3, S, S] on the GPU; the next batch is decoded + copied meanwhile
"""
import base64
import ctypes
import random
import threading
from queue import Queue

import numpy as np
import torch

from . import _lib as L

MEAN = (0.485, 0.456, 0.406)   # dataset.py:138-141, 152-154
STD = (0.229, 0.224, 0.225)


def str2img(b):
    """`Dataset_Base.str2img` (dataset.py:177-186): base64 (or raw bytes) JPEG -> uint8 RGB [H, W, 3]."""
    import cv2
    raw = base64.b64decode(b) if isinstance(b, str) else bytes(b)
    bgr = cv2.imdecode(np.frombuffer(raw, np.uint8), cv2.IMREAD_COLOR)
    if bgr is None:
        import io
        from PIL import Image
        return np.asarray(Image.open(io.BytesIO(raw)).convert("RGB"))
    return np.ascontiguousarray(bgr[:, :, ::-1])


def resized_size(h, w, size):
    """torchvision / torch_videovision Resize(int): the shorter side becomes `size`, the other int(size * long / short)."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


def crop_offsets(hr, wr, size, mode="center", rng=random):
    """CenterCrop (top = round((hr - size) / 2)) or RandomCrop offsets inside the resized frame."""
    if mode == "center":
        return int(round((hr - size) / 2.0)), int(round((wr - size) / 2.0))
    if mode == "random":
        return rng.randint(0, hr - size), rng.randint(0, wr - size)
    return mode   # explicit (top, left)


class GpuClipTransform:
    def __init__(self, size_img=224, mean=MEAN, std=STD, device=None):
        self.size = int(size_img)
        self._mean = (ctypes.c_float * 3)(*mean)
        self._std = (ctypes.c_float * 3)(*std)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def stage(self, frames, pinned=None):
        """Packs T decoded frames (uint8 [H, W, 3], one size) into a pinned buffer; returns (pinned, (T, H, W))."""
        T = len(frames)
        H, W, _ = frames[0].shape
        n = T * H * W * 3
        if pinned is None or pinned.numel() < n:
            pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
        view = pinned[:n].view(T, H, W, 3).numpy()
        for t, f in enumerate(frames):
            if f.shape != (H, W, 3):
                raise ValueError("all frames of a clip must have one size")
            view[t] = f
        return pinned, (T, H, W)

    def run(self, dev_u8, geom, out=None, crop="center", stream=None):
        """dev_u8: cuda uint8 holding [T, H, W, 3]; returns cuda fp32 [T, 3, S, S]."""
        T, H, W = geom
        S = self.size
        hr, wr = resized_size(H, W, S)
        top, left = crop_offsets(hr, wr, S, crop)
        if out is None:
            out = torch.empty(T, 3, S, S, dtype=torch.float32, device=dev_u8.device)
        st = torch.cuda.current_stream() if stream is None else stream
        rc = L.lib().lav_frames_resize_crop_norm_u8(ctypes.c_void_p(dev_u8.data_ptr()), T, H, W, H * W * 3, hr, wr, S, top, left,
                                                    self._mean, self._std, ctypes.c_void_p(out.data_ptr()),
                                                    ctypes.c_void_p(st.cuda_stream))
        L.check(rc, "lav_frames_resize_crop_norm_u8")
        return out

    def __call__(self, frames, crop="center"):
        pinned, geom = self.stage(frames)
        n = geom[0] * geom[1] * geom[2] * 3
        dev = pinned[:n].to(self.device, non_blocking=True)
        return self.run(dev, geom, crop=crop)


class GpuBatchLoader:
    """Double-buffered batches: a host thread decodes the JPEGs of batch i+1 (cv2 releases the GIL) into pinned memory while
    batch i trains; the H2D copies and the transform kernels run on a copy stream; iteration yields `[B, T, 3, S, S]`."""

    def __init__(self, transform, source, crop="center", depth=2):
        self.tf, self.source, self.crop = transform, source, crop
        self.q = Queue(maxsize=depth)
        self.stream = torch.cuda.Stream(device=transform.device)
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()

    def _work(self):
        for clips in self.source:
            staged = [self.tf.stage([str2img(b) for b in clip]) for clip in clips]
            with torch.cuda.stream(self.stream):
                outs = []
                for pinned, geom in staged:
                    n = geom[0] * geom[1] * geom[2] * 3
                    dev = pinned[:n].to(self.tf.device, non_blocking=True)
                    outs.append(self.tf.run(dev, geom, crop=self.crop, stream=self.stream))
                img = torch.stack(outs)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.q.put((img, ev, staged))
        self.q.put(None)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            img, ev, _staged = item
            torch.cuda.current_stream().wait_event(ev)
            img.record_stream(torch.cuda.current_stream())
            yield img
