"""the GPU frame transform (csrc/frames.cu) against the reference loader's own PIL / torchvision transform
(dataset.py:118-175: Resize(size_img) -> CenterCrop / RandomCrop -> ToTensor -> Normalize), bit-exact arithmetic."""
import numpy as np
import pytest
import torch


def _pil_reference(frames, size, top_left=None):
    from PIL import Image
    import torchvision.transforms as TT
    import torchvision.transforms.functional as TF
    outs = []
    for f in frames:
        im = TT.Resize(size)(Image.fromarray(f))
        if top_left is None:
            im = TT.CenterCrop((size, size))(im)
        else:
            im = TF.crop(im, top_left[0], top_left[1], size, size)
        t = TT.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(TT.ToTensor()(im))
        outs.append(t)
    return torch.stack(outs)


def test_resized_size_and_crop_offsets_match_torchvision():
    from PIL import Image
    import torchvision.transforms as TT
    from lavender_b200.input_pipeline import crop_offsets, resized_size
    for h, w in ((240, 426), (360, 640), (480, 360), (224, 224), (180, 320), (300, 225)):
        im = TT.Resize(224)(Image.fromarray(np.zeros((h, w, 3), np.uint8)))
        assert (im.size[1], im.size[0]) == resized_size(h, w, 224)
        hr, wr = resized_size(h, w, 224)
        top, left = crop_offsets(hr, wr, 224)
        assert 0 <= top <= hr - 224 and 0 <= left <= wr - 224


def test_str2img_decodes_like_the_reference():
    import base64
    import cv2
    from lavender_b200.input_pipeline import str2img
    rng = np.random.RandomState(0)
    img = cv2.GaussianBlur((rng.rand(64, 96, 3) * 255).astype(np.uint8), (7, 7), 0)
    ok, enc = cv2.imencode(".jpg", img[:, :, ::-1])
    assert ok
    a = str2img(base64.b64encode(enc.tobytes()).decode())
    b = cv2.imdecode(np.frombuffer(enc.tobytes(), np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
    assert a.shape == (64, 96, 3) and np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,size,crop", [(240, 426, 224, "center"), (360, 640, 224, (0, 117)), (480, 360, 224, "center"),
                                           (180, 320, 224, "center"), (224, 224, 224, "center"), (540, 960, 384, (0, 200))])
def test_gpu_frame_transform_is_bit_exact_vs_pil(h, w, size, crop):
    from lavender_b200.input_pipeline import GpuClipTransform
    rng = np.random.RandomState(h + w)
    base = rng.rand(5, h // 8 + 1, w // 8 + 1, 3)
    frames = [(np.kron(b, np.ones((8, 8, 1)))[:h, :w] * 200 + rng.rand(h, w, 3) * 55).astype(np.uint8) for b in base]
    tf = GpuClipTransform(size_img=size)
    out = tf(frames, crop=crop)
    torch.cuda.synchronize()
    ref = _pil_reference(frames, size, None if crop == "center" else crop)
    d = (out.cpu() - ref).abs()
    print(f"{h}x{w} -> {size}: max abs diff {d.max().item():.3e}, differing pixels {(d > 1e-6).float().mean().item():.2e}")
    assert d.max().item() < 1e-6   # same integer arithmetic as Pillow's resample: identical uint8 pixels


@pytest.mark.gpu
def test_gpu_batch_loader_overlaps_and_matches():
    import base64
    import cv2
    from lavender_b200.input_pipeline import GpuBatchLoader, GpuClipTransform, str2img
    rng = np.random.RandomState(3)

    def clip():
        out = []
        for _ in range(5):
            img = cv2.GaussianBlur((rng.rand(120, 160, 3) * 255).astype(np.uint8), (5, 5), 0)
            out.append(base64.b64encode(cv2.imencode(".jpg", img)[1].tobytes()).decode())
        return out
    batches = [[clip() for _ in range(3)] for _ in range(4)]
    tf = GpuClipTransform(size_img=96)
    got = [b.clone() for b in GpuBatchLoader(tf, iter(batches))]
    assert len(got) == 4 and got[0].shape == (3, 5, 3, 96, 96)
    ref = torch.stack([_pil_reference([str2img(b) for b in c], 96) for c in batches[2]])
    assert (got[2].cpu() - ref).abs().max().item() < 1e-6
