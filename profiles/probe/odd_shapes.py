"""Robustness: train_step + translate at batch / image sizes that leave the fast paths (odd tile counts ->
1-CTA kernel, widths < 128 -> generic kernels instead of the strip ring, ragged tiles)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import trainer as T, inference as I
from oracle import oracle as O
dev = torch.device("cuda", 0)
for (b, s, nd) in ((3, 96, 4), (5, 128, 10), (1, 256, 2), (2, 160, 3)):
    torch.manual_seed(0)
    tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=O.seeded_vgg_state())
    batch = O.synthetic_batch(b, s, nd)
    outs = [tr.train_step(batch, 0) for _ in range(3)]
    y = I.translate(tr.ema_G_A2B, tr.ema_SE_B, batch["source"], batch["target"], batch["target_domain"])
    torch.cuda.synchronize()
    vals = {k: round(float(v), 4) for k, v in outs[-1].items()}
    ok = all(torch.isfinite(v).all() for o in outs for v in o.values()) and bool(torch.isfinite(y).all())
    print((b, s, nd), "ok" if ok else "NON-FINITE", vals, flush=True)
    del tr
    torch.cuda.empty_cache()
