"""GPU-side input pipeline for training batches: the deterministic part of the reference's
MultiDomainStyleTransferDataset transform (/root/reference/dataset.py:16-22) -- RandomResizedCrop's
crop + bilinear resize, the 0/90/180/270 rotation, ToTensor and Normalize(0.5) -- as one C-ABI call on a
whole uint8 batch (msig_augment_u8), bit-exact with PIL / torchvision. The random draws stay on the host and
consume the torch RNG like torchvision does.

The reference decodes and augments on the CPU in four DataLoader workers (trainer.py:287-290); at the
~3000 img/s an 8-GPU step consumes that is the bottleneck. JPEG decoding
itself is out of scope here: the kernel takes decoded uint8 HWC images (e.g. nvJPEG / a uint8 cache)."""
import ctypes
import math

import torch

from . import lib as L
from . import ops

SCALE = (0.08, 1.0)              # torchvision RandomResizedCrop defaults (dataset.py:17 passes none)
RATIO = (3.0 / 4.0, 4.0 / 3.0)


def sample_crop(height, width, scale=SCALE, ratio=RATIO):
    """RandomResizedCrop.get_params: (top, left, h, w); same draws, in the same order, from the global
    torch RNG as torchvision's implementation."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)          # fallback: central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def sample_params(batch, height, width, generator=None, scale=SCALE, ratio=RATIO):
    """Vectorised draws for a whole batch: boxes int32 [B,4] (top, left, h, w) and quarter turns int32 [B]
    (RandomChoice over the four RandomRotation([a, a]), dataset.py:18-19). Rejection sampling like
    get_params (10 attempts, then the central-crop fallback), one tensor op per attempt."""
    g = generator
    area = float(height * width)
    lo, hi = math.log(ratio[0]), math.log(ratio[1])
    boxes = torch.zeros((batch, 4), dtype=torch.int64)
    done = torch.zeros(batch, dtype=torch.bool)
    for _ in range(10):
        ta = area * torch.empty(batch).uniform_(scale[0], scale[1], generator=g)
        ar = torch.exp(torch.empty(batch).uniform_(lo, hi, generator=g))
        w = torch.round(torch.sqrt(ta * ar)).long()
        h = torch.round(torch.sqrt(ta / ar)).long()
        ok = (w > 0) & (w <= width) & (h > 0) & (h <= height) & ~done
        u = torch.rand(batch, 2, generator=g)
        top = (u[:, 0] * (height - h + 1).clamp_min(1)).long()
        left = (u[:, 1] * (width - w + 1).clamp_min(1)).long()
        boxes[ok] = torch.stack([top, left, h, w], dim=1)[ok]
        done |= ok
        if bool(done.all()):
            break
    if not bool(done.all()):
        in_ratio = float(width) / float(height)
        if in_ratio < min(ratio):
            w = width
            h = int(round(w / min(ratio)))
        elif in_ratio > max(ratio):
            h = height
            w = int(round(h * max(ratio)))
        else:
            w, h = width, height
        boxes[~done] = torch.tensor([(height - h) // 2, (width - w) // 2, h, w])
    rots = torch.randint(0, 4, (batch,), generator=g)
    return boxes.to(torch.int32), rots.to(torch.int32)


def augment(images_u8, boxes, quarter_turns, size, out=None):
    """images_u8: uint8 [B, H, W, 3] on the GPU; boxes int32 [B,4] (top, left, h, w); quarter_turns int32 [B]
    (counter-clockwise, PIL's direction). Returns fp32 [B, 3, size, size] in [-1, 1] -- what a DataLoader
    over the reference's dataset would have collated for the same draws."""
    if not images_u8.is_cuda:
        raise RuntimeError("augment: expected a CUDA uint8 batch; msig_b200 has no CPU path")
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
        raise RuntimeError("augment: expected uint8 [B, H, W, 3]")
    dev = images_u8.device
    with torch.cuda.device(dev):
        ops.ensure_init(dev)
        images_u8 = images_u8.contiguous()
        n, h, w, _ = images_u8.shape
        boxes = boxes.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        quarter_turns = quarter_turns.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        if boxes.shape != (n, 4) or quarter_turns.shape != (n,):
            raise RuntimeError("augment: boxes must be [B,4] and quarter_turns [B]")
        if out is None:
            out = torch.empty((n, 3, size, size), dtype=torch.float32, device=dev)
        ws = ops.workspace(L.load().msig_augment_workspace(n, h, w, size), dev)
        L.call("msig_augment_u8", ctypes.c_void_p(images_u8.data_ptr()), n, h, w, ctypes.c_void_p(boxes.data_ptr()),
               ctypes.c_void_p(quarter_turns.data_ptr()), size, ctypes.c_void_p(out.data_ptr()),
               ctypes.c_void_p(ws.data_ptr()), ws.numel(), ops._stream())
    return out
