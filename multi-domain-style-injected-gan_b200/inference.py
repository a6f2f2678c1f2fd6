"""Mirror of the hot-path part of the reference's inference.py: model loading (EMA preferred),
style-vector extraction, style sampling modes and the style-injection forward — here BATCHED
(the reference runs batch 1, inference.py:273-305; BASELINE.json config 2 is batch 16).

Reference: /root/reference/inference.py:19-77 (load_model), :80-129 (preload_style_vectors),
:132-169 (apply_style_mode), :273-305 (generation loop). File discovery, argparse and PNG writing
of the reference's main() are out of scope; `translate()` is the batched forward it would call.
"""
import os
import random

import torch

from . import ops
from .model import MultiDomainStyleEncoder, StyleCycleGANGenerator


def load_model(checkpoint_path, style_dim, num_domains, device):
    """Load generator + style encoder for inference; EMA weights when ema_checkpoint.pth exists."""
    print(f"Loading multi-domain model with {num_domains} domains...")
    generator = StyleCycleGANGenerator(in_channels=3, out_channels=3, style_dim=style_dim).to(device)
    style_encoder = MultiDomainStyleEncoder(style_dim=style_dim, num_domains=num_domains).to(device)
    checkpoint_dir = os.path.dirname(checkpoint_path)
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(f"Checkpoint not found: {checkpoint_path}")
    print(f"Loading checkpoint from: {checkpoint_path}")
    checkpoint = torch.load(checkpoint_path, map_location=device, weights_only=False)
    ema_checkpoint_path = os.path.join(checkpoint_dir, 'ema_checkpoint.pth')
    loaded = False
    if os.path.exists(ema_checkpoint_path):
        print("Loading EMA models from ema_checkpoint.pth...")
        ema_checkpoint = torch.load(ema_checkpoint_path, map_location=device, weights_only=False)
        try:
            generator.load_state_dict(ema_checkpoint['ema_G_A2B'])
            style_encoder.load_state_dict(ema_checkpoint['ema_SE_B'])
            loaded = True
            print("Successfully loaded EMA models")
        except KeyError as e:
            print(f"Error loading EMA models: {e}\nFalling back to regular models...")
    if not loaded:
        generator.load_state_dict(checkpoint['G_A2B'])
        style_encoder.load_state_dict(checkpoint['SE_B'])
        print("Successfully loaded regular models")
    generator.eval()
    style_encoder.eval()
    return generator, style_encoder


@torch.no_grad()
def extract_style_vectors(style_encoder, images, domain_idx, batch_size=16):
    """Style codes of a stack of reference images [N,3,S,S] (already normalised to [-1,1]) for one
    domain: the batched equivalent of the loop in preload_style_vectors (inference.py:106-123).
    Returns a list of [1, style_dim] tensors like the reference."""
    dev = next(style_encoder.parameters()).device
    out = []
    for i in range(0, images.shape[0], batch_size):
        chunk = images[i:i + batch_size].to(dev, non_blocking=True)
        y = torch.full((chunk.shape[0],), int(domain_idx), dtype=torch.int64, device=dev)
        codes = style_encoder(chunk, y)
        out.extend(codes[j:j + 1] for j in range(codes.shape[0]))
    return out


def preload_style_vectors(style_encoder, ref_domain_dir, domain_idx, image_size, device, max_styles=None):
    """Reference-compatible entry (inference.py:80-129): reads the images of one reference-domain
    directory (PIL + torchvision transforms on the host) and encodes them in batches."""
    import glob
    from PIL import Image
    from torchvision import transforms
    style_files = []
    for ext in ['*.jpg', '*.jpeg', '*.png', '*.JPG', '*.JPEG', '*.PNG']:
        style_files.extend(glob.glob(os.path.join(ref_domain_dir, ext)))
    if not style_files:
        raise ValueError(f"No images found in {ref_domain_dir}")
    if max_styles and len(style_files) > max_styles:
        style_files = random.sample(style_files, max_styles)
    print(f"Loading {len(style_files)} style vectors from {ref_domain_dir}")
    tf = transforms.Compose([transforms.Resize((image_size, image_size)), transforms.ToTensor(),
                             transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    imgs = []
    for path in style_files:
        try:
            imgs.append(tf(Image.open(path).convert('RGB')))
        except Exception as e:   # same tolerance as the reference loop
            print(f"Warning: Failed to process style image {path}: {e}")
    if not imgs:
        raise ValueError(f"No valid style vectors could be extracted from {ref_domain_dir}")
    vectors = extract_style_vectors(style_encoder, torch.stack(imgs).to(device), domain_idx)
    print(f"Loaded {len(vectors)} style vectors")
    return vectors


def apply_style_mode(style_vectors, mode, noise_level=0.1):
    """Style sampling strategies of the reference (inference.py:132-169)."""
    if not style_vectors:
        raise ValueError("No style vectors provided")
    if mode == 'average':
        style = torch.mean(torch.stack(style_vectors), dim=0)
    elif mode == 'random':
        style = random.choice(style_vectors)
    elif mode == 'interpolate':
        if len(style_vectors) < 2:
            style = style_vectors[0]
        else:
            s1, s2 = random.sample(style_vectors, 2)
            alpha = random.random()
            style = alpha * s1 + (1 - alpha) * s2
    elif mode == 'noise':
        style = random.choice(style_vectors)
        style = style + torch.randn_like(style) * noise_level
    elif mode == 'specific':
        style = style_vectors[0]
    else:
        raise ValueError(f"Unknown style mode: {mode}")
    return style


_translate_graphs = {}   # (generator, style encoder, shapes) -> captured forward, or "warm" after the first eager call


def _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev):
    src = src.to(dev, non_blocking=True)
    if style is None:
        y = None if ref_domain is None else ref_domain.to(dev, non_blocking=True)
        style = style_encoder(ref.to(dev, non_blocking=True), y)
    return generator(src, style.to(dev))


@torch.no_grad()
def translate(generator, style_encoder, src, ref=None, ref_domain=None, style=None, use_cuda_graph=None):
    """Style injection (inference.py:119,290), batched: src [B,3,S,S] in [-1,1]; either reference
    images `ref` [B,3,S,S] with their domain indices `ref_domain` [B], or a ready style code
    `style` [1|B, style_dim]. Returns the translated images [B,3,S,S].

    The forward has static shapes and no host synchronisation: from the second call with the same
    shapes it is replayed from a captured CUDA graph (the ~110 launches of one batch otherwise cost
    more host time than GPU time). The packed bf16 weights are refreshed outside the graph, in place,
    so a weight update does not invalidate it. The returned tensor is the graph's output buffer: it
    is overwritten by the next call with the same shapes (clone it to keep it).
    `use_cuda_graph=False` (or MSIG_CUDA_GRAPH=0) keeps every call eager."""
    dev = next(generator.parameters()).device
    if style is None and ref is None:
        raise ValueError("translate: give reference images or a style code")
    if use_cuda_graph is None:
        use_cuda_graph = os.environ.get("MSIG_CUDA_GRAPH", "1") != "0"
    if not use_cuda_graph or dev.type != "cuda":
        return _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev)
    other = style if style is not None else ref
    key = (id(generator), id(style_encoder), tuple(src.shape), style is None, tuple(other.shape), ref_domain is None)
    entry = _translate_graphs.get(key)
    if entry is None:                       # first call: eager (also builds the packed weights)
        _translate_graphs[key] = "warm"
        return _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev)
    generator._packed.get()                 # refresh packed weights (no-op unless the parameters changed)
    style_encoder._packed.get()
    if entry == "warm":
        st = {"src": torch.empty(src.shape, dtype=torch.float32, device=dev),
              "other": torch.empty(other.shape, dtype=torch.float32, device=dev),
              "dom": None if ref_domain is None else torch.empty(ref_domain.shape, dtype=torch.int64, device=dev)}
        st["src"].copy_(src, non_blocking=True)
        st["other"].copy_(other, non_blocking=True)
        if st["dom"] is not None:
            st["dom"].copy_(ref_domain, non_blocking=True)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        l0 = ops.kernel_launches()
        with torch.cuda.graph(graph):
            if style is None:
                out = _translate_eager(generator, style_encoder, st["src"], st["other"], st["dom"], None, dev)
            else:
                out = _translate_eager(generator, style_encoder, st["src"], None, None, st["other"], dev)
        launches = ops.kernel_launches() - l0
        ops.add_replayed_launches(-launches)      # recorded, not executed, during capture
        entry = {"graph": graph, "st": st, "out": out, "gen": generator, "se": style_encoder, "launches": launches}
        _translate_graphs[key] = entry
    st = entry["st"]
    st["src"].copy_(src, non_blocking=True)
    st["other"].copy_(other, non_blocking=True)
    if st["dom"] is not None:
        st["dom"].copy_(ref_domain, non_blocking=True)
    entry["graph"].replay()
    ops.add_replayed_launches(entry["launches"])
    return entry["out"]
