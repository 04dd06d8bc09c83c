"""clip `preprocess` on the device (rows f3 / f4).  CPU part: the host-side restatement of Pillow's coefficient computation
(divergen_b200.preprocess.resample_coeffs) drives two integer passes written here in numpy and must reproduce
`PIL.Image.resize(..., BICUBIC)` bit for bit -- Pillow itself is the checker.  GPU part: the kernels against PIL / torchvision-
style preprocessing."""
import numpy as np
import pytest
import torch
from PIL import Image

from divergen_b200.preprocess import CLIP_MEAN, CLIP_STD, resample_coeffs, resize_size

BITS = 22


def _numpy_resample(img, out_h, out_w):
    def one_pass(a, out_n, axis):
        bounds, kk, _ = resample_coeffs(a.shape[axis], out_n)
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((out_n,) + a.shape[1:], dtype=np.uint8)
        for o in range(out_n):
            lo, n = bounds[o]
            acc = (1 << (BITS - 1)) + np.tensordot(kk[o, :n].astype(np.int64), a[lo:lo + n], axes=(0, 0))
            out[o] = np.clip(acc >> BITS, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    a = img
    if out_w != a.shape[1]:
        a = one_pass(a, out_w, 1)
    if out_h != a.shape[0]:
        a = one_pass(a, out_h, 0)
    return a


@pytest.mark.parametrize("h,w,oh,ow", [(512, 512, 224, 224), (512, 768, 224, 336), (300, 200, 336, 224), (64, 64, 224, 224)])
def test_restated_pillow_coefficients_reproduce_pil(h, w, oh, ow):
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(_numpy_resample(img, oh, ow), want)


def test_resize_size_is_torchvision_resize():
    assert resize_size(512, 512, 224) == (224, 224)
    assert resize_size(512, 768, 224) == (224, 336)
    assert resize_size(300, 200, 224) == (336, 224)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(512, 512), (512, 768), (96, 160)])
def test_device_resize_is_bit_identical_to_pil(h, w):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from divergen_b200 import resize_u8
    imgs = np.random.default_rng(h + w).integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    oh, ow = resize_size(h, w, 224)
    got = resize_u8(torch.from_numpy(imgs).cuda(), oh, ow).cpu().numpy()
    for b in range(3):
        want = np.asarray(Image.fromarray(imgs[b]).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(got[b], want)


@pytest.mark.gpu
def test_device_clip_preprocess_matches_clip_transform():
    """Resize(224, BICUBIC) -> CenterCrop(224) -> ToTensor -> Normalize, as clip's `_transform` / get_clip_score.py:143-150."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from divergen_b200 import clip_preprocess
    imgs = np.random.default_rng(1).integers(0, 256, (2, 512, 640, 3), dtype=np.uint8)
    got = clip_preprocess(torch.from_numpy(imgs).cuda()).float().cpu()
    mean, std = torch.tensor(CLIP_MEAN).view(3, 1, 1), torch.tensor(CLIP_STD).view(3, 1, 1)
    for b in range(2):
        oh, ow = resize_size(512, 640, 224)
        r = np.asarray(Image.fromarray(imgs[b]).resize((ow, oh), Image.BICUBIC))
        top, left = int(round((oh - 224) / 2.0)), int(round((ow - 224) / 2.0))
        t = torch.from_numpy(r[top:top + 224, left:left + 224].copy()).permute(2, 0, 1).float() / 255.0
        want = (t - mean) / std
        assert (got[b] - want).abs().max().item() <= 2e-3          # fp16 rounding of values in [-1.8, 2.2]
