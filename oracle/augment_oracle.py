"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

CPU restatement (numpy, integer / double arithmetic) of the reference's training-time image augmentation,
/root/reference/dataset.py:16-22:

    transforms.RandomResizedCrop(image_size)                    # crop box, then PIL bilinear resize
    transforms.RandomChoice([RandomRotation([a, a]) for a in (0, 90, 180, 270)])
    transforms.ToTensor(); transforms.Normalize((0.5,)*3, (0.5,)*3)

for GIVEN random draws (crop box top/left/height/width, quarter turns), i.e. the deterministic part that the
GPU kernel msig_augment_u8 replaces. The arithmetic lives in third-party code that is not under
/root/reference: Pillow (PIL.Image.resize -> ImagingResample, src/libImaging/Resample.c, 8 bits per channel:
double-precision triangle-filter coefficients normalised and quantised to 22-bit fixed point, a horizontal pass
rounded to uint8, then a vertical pass rounded to uint8) and torchvision (F.resized_crop = crop + resize on PIL
images, F.rotate -> PIL transpose for multiples of 90 degrees on square images, to_tensor = /255, normalize).
Restated from the published algorithm; PINNED by tests/test_augment_oracle.py against Pillow 12.2 /
torchvision 0.26 themselves (bit-exact on uint8 and on the final fp32 tensor) on the reference's own sample
images when present and on random images otherwise.

Only tests/ may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2        # Resample.c


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter and
    box = the whole (cropped) axis. Returns (xmin[out], xcount[out], kk[out][ksize] int32)."""
    scale = float(in_size) / out_size               # (double)(in1 - in0) / outSize with in0 = 0
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """PIL Image.resize((out_w, out_h), BILINEAR) of a uint8 HWC image: horizontal pass, then vertical."""
    h, w, c = img.shape
    cur = img
    if out_w != w:                                   # ImagingResample: need_horizontal
        xmin, cnt, kk = precompute_coeffs(w, out_w)
        tmp = np.zeros((h, out_w, c), np.uint8)
        for xx in range(out_w):
            acc = np.full((h, c), 1 << (PRECISION_BITS - 1), np.int64)
            for x in range(cnt[xx]):
                acc += cur[:, xmin[xx] + x, :].astype(np.int64) * int(kk[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
    if out_h != h:                                   # need_vertical
        ymin, cnt, kk = precompute_coeffs(h, out_h)
        tmp = np.zeros((out_h, cur.shape[1], c), np.uint8)
        for yy in range(out_h):
            acc = np.full((cur.shape[1], c), 1 << (PRECISION_BITS - 1), np.int64)
            for y in range(cnt[yy]):
                acc += cur[ymin[yy] + y, :, :].astype(np.int64) * int(kk[yy, y])
            tmp[yy] = _clip8(acc)
        cur = tmp
    return cur


def augment_u8(img, top, left, height, width, quarter_turns, size):
    """dataset.py:16-22 for fixed draws: uint8 HWC image -> uint8 HWC [size, size, 3] after crop, bilinear
    resize and a counter-clockwise rotation by quarter_turns * 90 degrees (PIL Transpose.ROTATE_90/180/270)."""
    crop = img[top:top + height, left:left + width, :]
    out = resize_bilinear_u8(crop, size, size)
    return np.ascontiguousarray(np.rot90(out, quarter_turns % 4, axes=(0, 1)))


def to_normalized_chw(u8_hwc):
    """ToTensor (uint8 / 255 in fp32) + Normalize(0.5, 0.5): fp32 CHW in [-1, 1]."""
    x = u8_hwc.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def augment(img, top, left, height, width, quarter_turns, size):
    return to_normalized_chw(augment_u8(img, top, left, height, width, quarter_turns, size))
