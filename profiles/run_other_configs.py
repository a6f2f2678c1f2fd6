import sys, time, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import trainer as T
from oracle import oracle as O
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), 10, vgg_state=O.seeded_vgg_state())
for (b, s) in ((8, 512), (64, 256)):
    batch = {k: v.to(dev) for k, v in O.synthetic_batch(b, s, 10).items()}
    for i in range(3):
        out = tr.train_step(batch, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        out = tr.train_step(batch, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(b, s, "ms/step", round(ms, 2), "img/s", round(b / ms * 1e3, 1), {k: round(float(v), 5) for k, v in out.items()},
          "mem GB", round(torch.cuda.max_memory_allocated() / 1e9, 1), flush=True)
    tr._graph = None; tr._graph_key = None
    del batch
    torch.cuda.empty_cache()
