#!/usr/bin/env python
"""Benchmark of the hot path: one full G+D training step (trainer.train_step) at 256x256.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c4|c5] [--batch B] [--impl ours|reference]

--config (BASELINE.json `configs`): c3 (default; the driver's line) = 256x256, batch 32 per GPU (weak scaling
when N > 1; at N = 8 this IS configs[3]); c4 = configs[3] as written: 256x256, GLOBAL batch 256, i.e.
256/N per GPU (strong scaling: N = 2 / 4 / 8 -> 128 / 64 / 32); c5 = configs[4]: 512x512, batch 8 per GPU.

Prints ONE JSON line (rank 0). Contract fields: metric/value/unit (BASELINE.json's metric: train
images/sec at 256^2), e2e (same metric through the public trainer API with pinned HOST batches and a
device->host loss read every step), roofline (dominant kernel: the 3x3 256->256 tcgen05 implicit
GEMM, timed live with CUDA events), cpu_baseline (the oracle port of the reference's CPU path timed
on the host cores, bounded sample), clocks, gpu_launches.

`--impl reference` times the reference's CPU implementation of the same step (the oracle port;
/root/reference is not present on the GPU box) with all host threads on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S = 256
ND = 10
CONFIGS = {
    "c3": dict(size=256, batch=lambda world: 32, scaling="weak",
               workload="full G+D training step 256x256 batch 32 per GPU (BASELINE.json configs[2]; configs[3] at 8 GPUs)"),
    "c4": dict(size=256, batch=lambda world: 256 // world, scaling="strong",
               workload="data-parallel training 256x256 global batch 256 (BASELINE.json configs[3]): 256/N per GPU"),
    "c5": dict(size=512, batch=lambda world: 8, scaling="weak",
               workload="512x512 multi-domain training batch 8 per GPU (BASELINE.json configs[4])"),
}
WORKLOAD = CONFIGS["c3"]["workload"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def step_flops(b, s=S):
    """Necessary algorithmic FLOPs of one train step: s*(1641.7*B + 21.47*B^2) GF."""
    sc = (s / 256.0) ** 2
    return sc * (1641.7 * b + 21.47 * b * b) * 1e9


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region with ONE long-running
    `nvidia-smi --query-gpu=... -lms 200` process (the recipe's clocks line, B200_PROFILING.md): spawning a new
    nvidia-smi every 200 ms re-initialises NVML each time and costs the step under test ~1 %."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.max_mhz, self.reasons = [], None, set()
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        super().start()

    def run(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            out = line.strip().split(",")
            try:
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(self.NAMES, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(sample_steps=5, threads=None, size=S):
    """The reference's CPU path (oracle port, fp32, all host threads) on a bounded sample: B=1 (SURVEY 8d:
    1 warm-up + 5 timed steps, median and min -- host timing is noisy)."""
    import torch
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    state = O.init_state(0, ND)
    tr = O.OracleTrainer(state, O.seeded_vgg_state(), ND)
    batch = O.synthetic_batch(1, size, ND)
    tr.train_step(batch, 0)                      # warm-up
    ts = []
    for _ in range(sample_steps):
        t0 = time.time()
        tr.train_step(batch, 0)
        ts.append(time.time() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": 1.0 / med, "unit": "images/sec", "cores": threads, "kind": "port",
            "median_s_per_step": med, "min_s_per_step": ts[0],
            "sample": f"oracle port of trainer.train_step, B=1 {size}x{size} nd={ND}, fp32, 1 warm-up + {sample_steps} timed "
                      f"steps (median {med:.2f} s/step, min {ts[0]:.2f})"}


def eager_cuda_baseline(dev, batch, size, steps=3):
    """The honest same-box bar: the reference's arithmetic as PyTorch-eager CUDA ops
    (cuDNN / cuBLAS / ATen, i.e. what the reference itself runs on a B200) on the SAME train step and batch
    -- the oracle restatement moved to the device -- (a) fp32 with the cuDNN TF32 default, (b) under bf16
    autocast with channels_last tensors. Reported beside our number; it is NOT the --impl reference arm."""
    import torch
    from oracle import oracle as O
    out = {}
    for name in ("fp32_tf32conv", "bf16_autocast_channels_last"):
        try:
            cl = name.startswith("bf16")
            state = {k: {n: (t.to(dev).contiguous(memory_format=torch.channels_last) if (cl and t.dim() == 4) else t.to(dev))
                         for n, t in sd.items()} for k, sd in O.init_state(0, ND).items()}
            vgg = {n: (t.to(dev).contiguous(memory_format=torch.channels_last) if (cl and t.dim() == 4) else t.to(dev))
                   for n, t in O.seeded_vgg_state().items()}
            tr = O.OracleTrainer(state, vgg, ND)
            b = {k: v.to(dev) for k, v in O.synthetic_batch(batch, size, ND).items()}
            if cl:
                b = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in b.items()}

            def step():
                if cl:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return tr.train_step(b, 0)
                return tr.train_step(b, 0)
            torch.cuda.reset_peak_memory_stats(dev)
            for _ in range(2):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms * 1e-3), "unit": "images/sec", "ms_per_step": ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
            del tr, state, vgg, b
        except Exception as e:          # e.g. out of memory at a large batch: reported, not fatal
            out[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    out["note"] = (f"oracle restatement of trainer.train_step as PyTorch-eager CUDA ops, B={batch} {size}x{size} nd={ND}, "
                   f"2 warm-up + {steps} timed steps, CUDA events; torch {torch.__version__}, cudnn.allow_tf32="
                   f"{torch.backends.cudnn.allow_tf32}")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = CONFIGS[args.config]
    size = cfg["size"]
    state = O.init_state(0, ND)
    tr = O.OracleTrainer(state, O.seeded_vgg_state(), ND)
    batch = O.synthetic_batch(1, size, ND)
    for _ in range(max(args.warmup, 0)):
        tr.train_step(batch, 0)
    t0 = time.time()
    for _ in range(args.steps):
        tr.train_step(batch, 0)
    dt = time.time() - t0
    v = args.steps * 1.0 / dt
    line = {"impl": "reference", "metric": f"train images/sec at {size}^2", "value": v, "unit": "images/sec",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": f"each step = B=1 {size}x{size} train_step on the host CPU",
                       "num_domains": ND},
            "cpu_baseline": {"value": v, "unit": "images/sec", "cores": threads, "kind": "port",
                             "sample": f"B=1 {size}x{size} train_step x {args.steps} (oracle port of the reference CPU path)"},
            "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _time_launch(torch, launch, sets, reps=4, iters=12):
    """Duration of one launch of a kernel timed alone, two ways, both with CUDA events on the launching stream after
    warm-up: (a) ONE launch between two events, a 256 MB flush kernel before every timed launch; (b) the average over
    `reps` back-to-back passes over `sets` DISTINCT operand sets whose footprint is several times the 126 MB L2 (cold
    operands, no flush kernel in between). (a) carries the event / launch latency of a single launch (~5 us: 20 % of a
    25 us kernel), (b) the gaps between launches; the kernel's own duration (ncu: profiles/ncu_*_r2.txt) is below
    both, so the smaller one is reported as kernel_ms and both are listed."""
    for k in range(sets):
        launch(k)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB)
    total = 0.0
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(i % sets)
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    single = total / iters
    del flush
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for k in range(sets):
            launch(k)
    e1.record()
    e1.synchronize()
    chain = e0.elapsed_time(e1) / (reps * sets)
    return min(single, chain), {"single_launch_l2_flushed_ms": single, "back_to_back_rotating_sets_ms": chain}


def time_dominant_kernel(torch, ops, L, batch, sets=6):
    """The dominant kernel: 3x3 s1 p1 256->256 implicit GEMM at 64x64 (80% of generator FLOPs); 6 rotating
    (input, output) pairs = 6 x 134 MB at batch 32."""
    dev = torch.device("cuda")
    xs = [torch.randn(batch, 64, 64, 256, device=dev).to(torch.bfloat16) for _ in range(sets)]
    ys = [torch.empty(batch, 64, 64, 256, device=dev, dtype=torch.bfloat16) for _ in range(sets)]
    w = torch.randn(256, 256, 3, 3, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_FWD, w, 256, 256, 3, 3)
    g = ops.conv_geom(batch, 64, 64, 256, 256, 3, 3, 1, 1, 1, 64, 64)
    ms, how = _time_launch(torch, lambda k: ops.conv2d_fwd(xs[k], wpk, g, out=ys[k]), sets)
    flops = 2.0 * batch * 4096 * 256 * 2304
    return ms, flops, how


def time_hbm_kernel(torch, ops, L, batch, sets=6):
    """The dominant HBM-bound kernel: fused norm-apply + ReLU (norm_act_fwd_kernel) on the residual-block
    activation [B,64,64,256] bf16: 1 read + 1 write per element; 6 rotating (input, output) pairs."""
    dev = torch.device("cuda")
    xs = [torch.randn(batch, 64, 64, 256, device=dev).to(torch.bfloat16) for _ in range(sets)]
    ys = [torch.empty_like(x) for x in xs]
    st = ops.in_stats(xs[0])
    ms, how = _time_launch(torch, lambda k: ops.norm_act_fwd(xs[k], st, L.ACT_RELU, out=ys[k]), sets)
    return ms, 2.0 * xs[0].numel() * 2, how


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/ncu_fprop_r2.txt, first line = plain forward launch)."""
    try:
        path = os.path.join(ROOT, "profiles", "ncu_fprop_r2.txt")
        if not os.path.exists(path):
            path = os.path.join(ROOT, "profiles", "ncu_fprop_r1.txt")
        line = open(path).readline()
        kv = dict(t.split("=") for t in line.split() if "=" in t)
        return (float(kv["dram_read_MB"]) + float(kv["dram_write_MB"])) * 1e6
    except Exception:
        return None


def time_inference(torch, T, O, dev, tr, batch=16, iters=10):
    """BASELINE.json configs[1]: style injection, src+ref 256x256 batch 16, style encoder + generator
    forward only (inference.py:119,290). Device-timed with resident inputs, and end to end through
    inference.translate with pinned host images in and the translated images copied back out."""
    from msig_b200 import inference as I
    gb = O.synthetic_batch(batch, S, ND)
    src_h, ref_h, dom_h = gb["source"].pin_memory(), gb["target"].pin_memory(), gb["target_domain"].pin_memory()
    src, ref, dom = src_h.to(dev), ref_h.to(dev), dom_h.to(dev)
    G, SE = tr.ema_G_A2B, tr.ema_SE_B
    for _ in range(3):
        out = I.translate(G, SE, src, ref, dom)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = I.translate(G, SE, src, ref, dom)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # end to end through the batched inference driver: pinned host images in, translated images out (pinned
    # host), H2D(i+1) / replay(i) / D2H(i-1) overlapped on three streams (inference.translate_batches)
    n_e2e = 4 * iters
    list(I.translate_batches(G, SE, ((src_h, ref_h, dom_h) for _ in range(4))))      # warm-up: captures both slots
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    checksum = 0.0
    for _, host in I.translate_batches(G, SE, ((src_h, ref_h, dom_h) for _ in range(n_e2e))):
        checksum += float(host[0, 0, 0, 0])         # the consumer touches every result on the host
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / n_e2e
    host_out = host
    flops = batch * 100.28e9                     # 96.96 (G) + 3.32 (SE) GF / image
    return {"workload": "inference.py style injection: src+ref 256x256 batch %d, SE + G forward (BASELINE.json configs[1])" % batch,
            "value": batch / (ms * 1e-3), "unit": "images/sec", "ms_per_batch": ms,
            "e2e": {"value": batch / (e2e_ms * 1e-3), "unit": "images/sec",
                    "h2d_bytes_per_step": 2 * src_h.numel() * 4 + dom_h.numel() * 8, "d2h_bytes_per_step": host_out.numel() * 4},
            "e2e_path": "inference.translate_batches (double-buffered pinned staging, 3 streams, 2 captured graphs)",
            "tflops": flops / (ms * 1e-3) / 1e12}


def run_ours(args):
    import torch
    import msig_b200  # noqa: F401
    from msig_b200 import lib as L
    from msig_b200 import ops
    from msig_b200 import trainer as T
    from oracle import oracle as O   # only for cpu_baseline and the seeded synthetic batch / VGG weights

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    S = args.size or cfg["size"]
    B = args.batch or cfg["batch"](world)
    if B < 1 or (args.config == "c4" and not args.batch and 256 % world):
        raise SystemExit(f"config {args.config}: world size {world} does not divide the global batch")
    pk = peaks()

    eager = None
    if world == 1 and rank == 0 and not args.no_eager_baseline:
        eager = eager_cuda_baseline(dev, B, S)      # first: it needs ~2.2 GB per sample of its own

    torch.manual_seed(0)
    tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), ND, vgg_state=O.seeded_vgg_state())
    gb = O.synthetic_batch(B * world, S, ND)
    host = {k: v[rank * B:(rank + 1) * B].contiguous().pin_memory() for k, v in gb.items()}
    devb = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    torch.cuda.reset_peak_memory_stats(dev)
    for _ in range(max(args.warmup, 3)):
        out = tr.train_step(devb, 0)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed region: inputs resident in HBM
    l0 = ops.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    th0 = time.perf_counter()
    for _ in range(args.steps):
        out = tr.train_step(devb, 0)
    host_ms = (time.perf_counter() - th0) * 1000.0 / args.steps      # time to ENQUEUE one step
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.kernel_launches() - l0
    # ---- end to end: pinned host batch in, loss scalar out, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = tr.train_step(host, 0)
        _ = float(out["G_loss"])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    value = B * world * args.steps / (ms / 1000.0)
    e2e = B * world * args.steps / (e2e_ms / 1000.0)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    kb = min(B, 32)                              # roofline kernels are timed at (up to) batch 32
    kms, kflops, khow = time_dominant_kernel(torch, ops, L, kb)
    ach = kflops / (kms * 1e-3) / 1e12
    hms, hbytes, hhow = time_hbm_kernel(torch, ops, L, kb)
    hach = hbytes / (hms * 1e-3) / 1e9
    step_tf = step_flops(B, S) / (ms / args.steps * 1e-3) / 1e12
    line = {
        "metric": f"train images/sec at {S}^2", "value": value, "unit": "images/sec", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "per_gpu_batch": B, "global_batch": B * world,
                   "image": S, "num_domains": ND, "peak_mem_gb": peak_mem,
                   "parallelism": f"dp{world}", "l2": "working set >> L2 (tens of GB of activations per step)",
                   "accumulate": "fp32", "loss_epoch": 0,
                   "cuda_graph": "step replayed from 4 captured graph segments (warm-up steps include the capture)"},
        "e2e": {"value": e2e, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": ach / pk["bf16_burst"],
                     "traffic": ncu_traffic(),
                     "kernel": "fprop2_kernel (tcgen05 cta_group::2 implicit GEMM) conv3x3 256->256 @64x64, batch %d" % kb,
                     "kernel_ms": kms, "algorithmic_flops_per_launch": kflops,
                     "algorithmic_bytes_per_launch": 2.0 * (2 * kb * 4096 * 256 + 256 * 2304),
                     "traffic_note": "dram__bytes_read+write of one launch (ncu --set full, profiles/ncu_fprop_r2.txt); "
                                     "below the algorithmic bytes because most of the output is still in L2 when the kernel ends",
                     "kernel_ms_by_method": khow,
                     "peak_source": pk["source"] + " (burst; kernel timed alone: the smaller of a single L2-flushed launch and back-to-back launches over 6 rotating operand sets)"},
        "roofline_hbm": {"bound": "hbm", "achieved": hach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hach / pk["hbm_gbs"],
                         "kernel": "norm_act_fwd_kernel (IN/AdaIN apply + ReLU) [%d,64,64,256] bf16, 1R+1W" % kb,
                         "kernel_ms": hms, "algorithmic_bytes_per_launch": hbytes,
                         "kernel_ms_by_method": hhow,
                         "peak_source": pk["source"] + " (kernel timed alone: the smaller of a single L2-flushed launch and back-to-back launches over 6 rotating operand sets)"},
        "step_tflops": {"achieved": step_tf, "peak_sustained": pk["bf16_sustained"], "frac": step_tf / pk["bf16_sustained"],
                        "flops_per_step": step_flops(B, S), "note": "necessary algorithmic FLOPs (SURVEY 8d) / step time"},
    }
    if eager is not None:
        line["eager_cuda_baseline"] = eager
        for k, v in eager.items():
            if isinstance(v, dict) and "value" in v:
                v["ours_over_this"] = value / v["value"]
    if world == 1 and args.config == "c3" and not args.no_inference:
        line["inference"] = time_inference(torch, T, O, dev, tr)
        line["inference"]["frac_of_sustained_peak"] = line["inference"]["tflops"] / pk["bf16_sustained"]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(size=S)
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json configs[2] / [3] / [4]")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--size", type=int, default=0, help="image side (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
