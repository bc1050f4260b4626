#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_input_pipeline.py -m gpu -x -q -s > gpurun_out/r2o_input.log 2>&1
tail -n 25 gpurun_out/r2o_input.log
timeout 300 python - > gpurun_out/r2o_frames_bench.log 2>&1 <<'PY'
import time, numpy as np, torch
from lavender_b200.input_pipeline import GpuClipTransform
tf = GpuClipTransform(224)
rng = np.random.RandomState(0)
frames = [(rng.rand(360, 640, 3) * 255).astype(np.uint8) for _ in range(4)]
pinned, geom = tf.stage(frames)
n = geom[0] * geom[1] * geom[2] * 3
dev = pinned[:n].cuda()
out = tf.run(dev, geom)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    tf.run(dev, geom, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print(f"kernel: {ms*1e3:.1f} us per 4-frame 360x640 clip -> {4/ms*1e3:.0f} frames/s; bytes in {n} out {out.numel()*4}")
t0 = time.time()
for _ in range(50):
    o = tf(frames)
torch.cuda.synchronize()
print(f"stage+H2D+kernel: {(time.time()-t0)/50*1e3:.2f} ms per clip (host-bound, one thread)")
# PIL comparison on host
from PIL import Image
import torchvision.transforms as TT
c = TT.Compose([TT.Resize(224), TT.CenterCrop((224, 224)), TT.ToTensor(), TT.Normalize([0.485,0.456,0.406],[0.229,0.224,0.225])])
t0 = time.time()
for _ in range(20):
    r = torch.stack([c(Image.fromarray(f)) for f in frames])
print(f"PIL/torchvision on one host core: {(time.time()-t0)/20*1e3:.2f} ms per clip")
PY
cat gpurun_out/r2o_frames_bench.log
