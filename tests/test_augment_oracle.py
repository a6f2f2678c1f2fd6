"""CPU: pins oracle/augment_oracle.py (the restatement of Pillow's 8-bit bilinear resample + the
crop / rotation / ToTensor / Normalize of /root/reference/dataset.py:16-22) against Pillow and torchvision
THEMSELVES, bit-exactly, for fixed random draws -- and the host-side parameter sampling of the drop-in
against torchvision's RandomResizedCrop.get_params under the same RNG state."""
import glob
import os

import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms
from torchvision.transforms import functional as TF

from oracle import augment_oracle as A

REF_IMGS = sorted(glob.glob("/root/reference/data/src/Tomato_healthy/*.JPG"))[:4]


def _images():
    rng = np.random.RandomState(0)
    imgs = [rng.randint(0, 256, (256, 256, 3), dtype=np.uint8), rng.randint(0, 256, (200, 320, 3), dtype=np.uint8)]
    # smooth content too (photographs are not noise): low-pass random field
    f = rng.rand(32, 32, 3)
    imgs.append(np.asarray(Image.fromarray((f * 255).astype(np.uint8)).resize((256, 256), Image.BICUBIC)))
    for p in REF_IMGS:                                   # the reference's own sample data, when present
        imgs.append(np.asarray(Image.open(p).convert("RGB")))
    return imgs


BOXES = [  # (top, left, height, width): full image, upscale, anisotropic, tiny (8% area), size-preserving axis
    (0, 0, None, None), (10, 20, 100, 140), (3, 0, 180, 75), (60, 70, 64, 80), (0, 0, 64, None), (17, 5, 131, 64),
]


@pytest.mark.parametrize("size", [64, 256])
def test_resize_crop_rotate_matches_pillow_bit_exact(size):
    for img in _images():
        H, W, _ = img.shape
        for bi, (t, l, h, w) in enumerate(BOXES):
            h = H if h is None else min(h, H - t)
            w = W if w is None else min(w, W - l)
            for k in range(4):
                pil = TF.resized_crop(Image.fromarray(img), t, l, h, w, [size, size])         # dataset.py:17
                pil = TF.rotate(pil, float(90 * k))                                           # dataset.py:18-19
                want_u8 = np.asarray(pil)
                got_u8 = A.augment_u8(img, t, l, h, w, k, size)
                assert got_u8.shape == want_u8.shape and np.array_equal(got_u8, want_u8), (img.shape, bi, k, size)
                want = TF.normalize(TF.to_tensor(pil), (0.5,) * 3, (0.5,) * 3).numpy()        # dataset.py:20-21
                got = A.augment(img, t, l, h, w, k, size)
                assert got.dtype == np.float32 and np.array_equal(got, want), (img.shape, bi, k, size)


def test_downscale_uses_the_widened_filter_support():
    """Pillow widens the triangle filter when shrinking (antialiasing): 5 and 7 taps here."""
    img = _images()[0]
    for size in (100, 40):
        want = np.asarray(Image.fromarray(img).resize((size, size), Image.BILINEAR))
        assert np.array_equal(A.resize_bilinear_u8(img, size, size), want)
    assert A.precompute_coeffs(256, 100)[2].shape[1] == 7 and A.precompute_coeffs(100, 256)[2].shape[1] == 3


def test_param_sampling_matches_torchvision_under_the_same_rng():
    """msig_b200.augment.sample_params consumes the torch RNG exactly like RandomResizedCrop.get_params +
    RandomChoice/RandomRotation of dataset.py:16-20 would for one image."""
    import msig_b200  # noqa: F401
    from msig_b200 import augment as G
    dummy = torch.zeros(3, 200, 320)
    for seed in range(20):
        torch.manual_seed(seed)
        want = transforms.RandomResizedCrop.get_params(dummy, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0))
        torch.manual_seed(seed)
        got = G.sample_crop(200, 320)
        assert tuple(got) == tuple(want), seed
    boxes, rots = G.sample_params(64, 256, 256, generator=torch.Generator().manual_seed(1))
    assert boxes.shape == (64, 4) and boxes.dtype == torch.int32 and rots.shape == (64,)
    assert int(rots.min()) >= 0 and int(rots.max()) <= 3 and len(set(rots.tolist())) == 4
    t, l, h, w = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    assert bool(((t >= 0) & (l >= 0) & (h > 0) & (w > 0) & (t + h <= 256) & (l + w <= 256)).all())
    area = (h * w).float() / (256 * 256)
    assert float(area.min()) >= 0.079 and float(area.max()) <= 1.0
