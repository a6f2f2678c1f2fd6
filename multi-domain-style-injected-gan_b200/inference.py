"""Mirror of the hot-path part of the reference's inference.py: model loading (EMA preferred),
style-vector extraction, style sampling modes and the style-injection forward — here BATCHED
(the reference runs batch 1, inference.py:273-305; BASELINE.json config 2 is batch 16).

Reference: /root/reference/inference.py:19-77 (load_model), :80-129 (preload_style_vectors),
:132-169 (apply_style_mode), :273-305 (generation loop). File discovery, argparse and PNG writing
of the reference's main() are out of scope; `translate()` is the batched forward it would call.
"""
import os
import random
import weakref

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


def _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev):
    src = src.to(dev, non_blocking=True)
    if style is None:
        y = None if ref_domain is None else ref_domain.to(dev, non_blocking=True)
        style = style_encoder(ref.to(dev, non_blocking=True), y)
    return generator(src, style.to(dev))


class _CapturedTranslate:
    """One captured SE -> G forward: static device inputs, the CUDA graph, its output buffer."""

    def __init__(self, generator, style_encoder, src, other, ref_domain, with_ref, dev, pool=None):
        self.dev = dev
        self.se = None if style_encoder is None else weakref.ref(style_encoder)
        self.src = torch.empty(src.shape, dtype=torch.float32, device=dev)
        self.other = torch.empty(other.shape, dtype=torch.float32, device=dev)
        self.dom = None if ref_domain is None else torch.empty(ref_domain.shape, dtype=torch.int64, device=dev)
        self.load(src, other, ref_domain)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.kernel_launches()
        with torch.cuda.graph(self.graph, pool=pool):
            if with_ref:
                self.out = _translate_eager(generator, style_encoder, self.src, self.other, self.dom, None, dev)
            else:
                self.out = _translate_eager(generator, style_encoder, self.src, None, None, self.other, dev)
        self.launches = ops.kernel_launches() - l0
        ops.add_replayed_launches(-self.launches)      # recorded, not executed, during capture

    def load(self, src, other, ref_domain):
        self.src.copy_(src, non_blocking=True)
        self.other.copy_(other, non_blocking=True)
        if self.dom is not None:
            self.dom.copy_(ref_domain, non_blocking=True)

    def replay(self):
        self.graph.replay()
        ops.add_replayed_launches(self.launches)
        return self.out


def _graph_cache(generator):
    """Captured forwards live ON the generator instance (dropped with it; never keyed by a recyclable
    id()): {key: "warm" | _CapturedTranslate}. The style encoder is held weakly and re-validated."""
    c = generator.__dict__.get("_msig_translate_cache")
    if c is None:
        c = {}
        generator.__dict__["_msig_translate_cache"] = c
    return c


def _cache_lookup(generator, style_encoder, key):
    cache = _graph_cache(generator)
    entry = cache.get(key)
    if isinstance(entry, (_CapturedTranslate, _Pipeline)) and entry.se is not None and entry.se() is not style_encoder:
        del cache[key]                               # the encoder this graph was captured with is gone
        entry = None
    return cache, entry


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
    if dev.type != "cuda":
        return _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev)
    with torch.cuda.device(dev):
        if not use_cuda_graph:
            return _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev)
        other = style if style is not None else ref
        key = (id(style_encoder) if style is None else None, tuple(src.shape), style is None, tuple(other.shape),
               ref_domain is None)
        cache, entry = _cache_lookup(generator, style_encoder if style is None else None, key)
        if entry is None:                       # first call: eager (also builds the packed weights)
            cache[key] = "warm"
            return _translate_eager(generator, style_encoder, src, ref, ref_domain, style, dev)
        generator._packed.get()                 # refresh packed weights (no-op unless the parameters changed)
        if style is None:
            style_encoder._packed.get()
        if entry == "warm":
            entry = _CapturedTranslate(generator, style_encoder if style is None else None, src, other, ref_domain,
                                       style is None, dev)
            cache[key] = entry
        else:
            entry.load(src, other, ref_domain)
        return entry.replay()


def translate_batches(generator, style_encoder, batches, consume=None):
    """See _translate_pipeline. With `consume` the pipeline is run to completion here and
    `consume(index, images)` is called per batch; without it a generator of `(index, images)` is returned."""
    gen = _translate_pipeline(generator, style_encoder, batches)
    if consume is None:
        return gen
    for i, host in gen:
        consume(i, host)
    return None


@torch.no_grad()
def _translate_pipeline(generator, style_encoder, batches):
    """The batched inference DRIVER (the loop of inference.py:273-305, where the reference generates one
    image at a time and writes its PNG before starting the next): a software pipeline over an iterable of
    host batches `(src, ref, ref_domain)` (pinned CPU tensors for full overlap) that keeps three things in
    flight at once on three streams -- the host->device copy of batch i+1, the graph replay of batch i and
    the device->host copy of batch i-1 -- through two captured forwards with double-buffered staging.

    Yields `(index, images)` with `images` a pinned host tensor [B,3,S,S] (valid until two more batches
    have been yielded): whatever the caller does with a result (save_image, encoding) overlaps the GPU
    work of the next batches. All batches must share one shape (a ragged tail batch is run eagerly)."""
    dev = next(generator.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("translate_batches: expected CUDA models; msig_b200 has no CPU path")

    def emit(i, host):
        return ((i, host),)

    with torch.cuda.device(dev):
        pipe = None
        pending = None
        for i, (src, ref, dom) in enumerate(batches):
            if pipe is None:
                # streams, captured forwards and pinned output staging are built once per (models, shapes) and
                # kept on the generator, like translate()'s graph: a second call starts replaying at once
                key = ("pipe", id(style_encoder), tuple(src.shape), tuple(ref.shape), dom is None)
                cache, pipe = _cache_lookup(generator, style_encoder, key)
                if pipe is None:
                    _translate_eager(generator, style_encoder, src, ref, dom, None, dev)  # warm-up: packs weights
                    torch.cuda.synchronize(dev)
                    pipe = _Pipeline(style_encoder, dev, (tuple(src.shape), tuple(ref.shape)))
                    cache[key] = pipe
                s_in, s_run, s_out, slots, shape = pipe.s_in, pipe.s_run, pipe.s_out, pipe.slots, pipe.shape
                s_in.wait_stream(torch.cuda.current_stream(dev))
            if (tuple(src.shape), tuple(ref.shape)) != shape:          # ragged tail: plain call
                if pending is not None:
                    pending[2].synchronize()
                    yield from emit(pending[0], pending[1])
                    pending = None
                yield from emit(i, translate(generator, style_encoder, src, ref, dom, use_cuda_graph=False).cpu())
                continue
            generator._packed.get()
            style_encoder._packed.get()
            k = i % 2
            if len(slots) <= k:
                cur = torch.cuda.current_stream(dev)
                s_run.wait_stream(cur)
                with torch.cuda.stream(s_run):
                    cap = _CapturedTranslate(generator, style_encoder, src, ref, dom, True, dev, pool=pipe.pool)
                pipe.pool = cap.graph.pool()
                slots.append({"cap": cap, "host": torch.empty(cap.out.shape, dtype=cap.out.dtype).pin_memory(),
                              "in": torch.cuda.Event(), "run": torch.cuda.Event(), "out": torch.cuda.Event()})
                slots[k]["run"].record(s_run)
                slots[k]["out"].record(s_out)
            sl = slots[k]
            with torch.cuda.stream(s_in):              # H2D(i): the slot's inputs are free once replay(i-2) is done
                s_in.wait_event(sl["run"])
                sl["cap"].load(src, ref, dom)
                sl["in"].record(s_in)
            with torch.cuda.stream(s_run):             # replay(i): needs its inputs, and D2H(i-2) off its output
                s_run.wait_event(sl["in"])
                s_run.wait_event(sl["out"])
                sl["cap"].replay()
                sl["run"].record(s_run)
            with torch.cuda.stream(s_out):             # D2H(i)
                s_out.wait_event(sl["run"])
                sl["host"].copy_(sl["cap"].out, non_blocking=True)
                sl["out"].record(s_out)
            if pending is not None:                    # hand batch i-1 to the caller while i runs
                pending[2].synchronize()
                yield from emit(pending[0], pending[1])
            pending = (i, sl["host"], sl["out"])
        if pending is not None:
            pending[2].synchronize()
            yield from emit(pending[0], pending[1])
        if pipe is not None:
            torch.cuda.current_stream(dev).wait_stream(pipe.s_run)


class _Pipeline:
    """Three streams + up to two captured forwards (double-buffered staging) of translate_batches."""

    def __init__(self, style_encoder, dev, shape):
        self.se = weakref.ref(style_encoder)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slots, self.pool, self.shape = [], None, shape
