"""msig_b200 — B200-native hot path of chouyunming/Multi-Domain-Style-Injected-GAN.

The package directory is `multi-domain-style-injected-gan_b200/`; it is importable as `msig_b200`
through the loader shim `msig_b200.py` at the repository root.

Layout: csrc/ (CUDA kernels + the C ABI, built into libmsig.so), lib.py (ctypes binding),
ops.py (tensor-level wrappers), model.py / losses.py / trainer.py / inference.py (host-side mirror
of the reference's interface for this path).
"""
__version__ = "0.1.0"
