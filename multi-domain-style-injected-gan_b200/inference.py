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


@torch.no_grad()
def translate(generator, style_encoder, src, ref=None, ref_domain=None, style=None):
    """Style injection (inference.py:119,290), batched: src [B,3,S,S] in [-1,1]; either reference
    images `ref` [B,3,S,S] with their domain indices `ref_domain` [B], or a ready style code
    `style` [1|B, style_dim]. Returns the translated images [B,3,S,S]."""
    dev = next(generator.parameters()).device
    src = src.to(dev, non_blocking=True)
    if style is None:
        if ref is None:
            raise ValueError("translate: give reference images or a style code")
        y = None if ref_domain is None else ref_domain.to(dev, non_blocking=True)
        style = style_encoder(ref.to(dev, non_blocking=True), y)
    return generator(src, style.to(dev))
