"""Mirror of the hot-path-adjacent parts of the reference's utils.py: EMA (utils.py:71-91) and
DynamicWeightScheduler (utils.py:94-134), plus the flat parameter / gradient buffers and the fused
clip + Adam + EMA optimizer the trainer uses (trainer.py:56-65,127-134,152-153).

Image-grid / plotting helpers of the reference (utils.py:9-68,136-155) are out of scope."""
import math

import torch

from . import ops

F32 = torch.float32
_ALIGN = 64   # floats; keeps every parameter 256-byte aligned inside the flat buffers


class FlatParams:
    """Re-homes the parameters of several modules into ONE fp32 buffer (and their gradients into
    another), in `module.parameters()` order, so that gradient all-reduce, global-norm clipping,
    Adam and EMA each touch one contiguous buffer. Parameter tensors stay ordinary views:
    state_dict() / load_state_dict() / checkpoints are unaffected."""

    def __init__(self, modules, device):
        self.modules = list(modules)
        self.params = [p for m in self.modules for p in m.parameters()]
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.data = torch.zeros(off, dtype=F32, device=device)
        self.grad = torch.zeros(off, dtype=F32, device=device)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.data[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.data[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)
        for m in self.modules:
            if hasattr(m, "mark_weights_dirty"):
                m.mark_weights_dirty()

    def views_of(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def rebind_grads(self):
        for p, o in zip(self.params, self.offsets):
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def mark_dirty(self):
        for m in self.modules:
            if hasattr(m, "mark_weights_dirty"):
                m.mark_weights_dirty()


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(betas, eps) semantics (no weight decay / amsgrad) over a FlatParams buffer,
    fused with clip_grad_norm_(max_norm) and the EMA update in one pass over memory
    (msig_sumsq + msig_adam_step). `param_groups[0]['lr']` is honoured, so torch LR schedulers
    (CosineAnnealingLR, trainer.py:64-65) work unchanged; state_dict() uses Adam's layout."""

    def __init__(self, flat, lr, betas=(0.5, 0.999), eps=1e-8, ema_flat=None, ema_beta=0.995):
        # every hyper-parameter key torch.optim.Adam keeps in a param_group (at its defaults): a
        # state_dict() written here then loads into the reference's torch.optim.Adam and can step
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                           foreach=None, capturable=False, differentiable=False, fused=None,
                                           decoupled_weight_decay=False))
        self.flat = flat
        self.ema_flat = ema_flat
        self.ema_beta = ema_beta
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.grad_sumsq = torch.zeros((), dtype=F32, device=flat.data.device)
        self.step_count = 0
        self._step_dev = None        # int32 device copy of step_count (CUDA-graph replay)
        m_views, v_views = flat.views_of(self.exp_avg), flat.views_of(self.exp_avg_sq)
        for p, m, v in zip(flat.params, m_views, v_views):
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": m, "exp_avg_sq": v}

    def zero_grad(self, set_to_none=False):
        self.flat.grad.zero_()
        self.flat.rebind_grads()

    def step_counter(self):
        """Device-resident step number, (re)synchronised with the host count."""
        if self._step_dev is None:
            self._step_dev = torch.zeros((), dtype=torch.int32, device=self.flat.data.device)
        self._step_dev.fill_(self.step_count)
        return self._step_dev

    def after_replay(self):
        """Host bookkeeping for one optimizer step executed inside a replayed CUDA graph."""
        self.step_count += 1
        self.flat.mark_dirty()
        if self.ema_flat is not None:
            self.ema_flat.mark_dirty()

    @torch.no_grad()
    def step(self, closure=None, max_norm=None, grad_scale=1.0, device_step=False):
        """device_step=True (CUDA-graph capture): the bias-correction step number is read from the
        device counter, which the captured kernels increment themselves; the host count is advanced
        by after_replay()."""
        g = self.param_groups[0]
        use_clip = max_norm is not None and max_norm > 0
        if use_clip:
            ops.sumsq(self.flat.grad, self.grad_sumsq, accumulate=False)
        args = (self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq,
                None if self.ema_flat is None else self.ema_flat.data,
                self.grad_sumsq if use_clip else None, max_norm if use_clip else 0.0, grad_scale,
                g["lr"], g["betas"][0], g["betas"][1], g["eps"])
        if device_step:
            ops.adam_step_dev(*args, self._step_dev, self.ema_beta)
            return
        self.step_count += 1
        ops.adam_step(*args, self.step_count, self.ema_beta)
        self.flat.mark_dirty()
        if self.ema_flat is not None:
            self.ema_flat.mark_dirty()

    def state_dict(self):
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.step_count))
        return super().state_dict()

    def grad_norm(self):
        """Pre-clip global gradient norm of the last step (device scalar)."""
        return torch.sqrt(self.grad_sumsq)

    def load_state_dict(self, state_dict):
        # copy into the flat-backed state instead of replacing the tensors
        groups = state_dict["param_groups"]
        for g, sg in zip(self.param_groups, groups):
            if sg.get("weight_decay", 0) or sg.get("amsgrad", False) or sg.get("maximize", False):
                raise RuntimeError("FusedAdam: weight_decay / amsgrad / maximize are not supported (the reference "
                                   "uses none of them, trainer.py:58-61)")
            for k in ("lr", "betas", "eps", "initial_lr"):
                if k in sg:
                    g[k] = tuple(sg[k]) if k == "betas" else sg[k]
        ids = [i for sg in groups for i in sg["params"]]
        steps = []
        for i, p in zip(ids, self.flat.params):
            st = state_dict["state"].get(i)
            if st is None:
                continue
            self.state[p]["exp_avg"].copy_(st["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
            steps.append(int(float(st["step"])))
        if steps:
            self.step_count = max(steps)


class EMA:
    """Exponential moving average of model parameters (reference utils.py:71-91). The trainer folds
    this update into the fused optimizer pass; this class keeps the reference's stand-alone API."""

    def __init__(self, beta):
        self.beta = beta

    @torch.no_grad()
    def update_model_average(self, ma_model, current_model):
        for cur, ma in zip(current_model.parameters(), ma_model.parameters()):
            ma.data.mul_(self.beta).add_(cur.data, alpha=1 - self.beta)
        if hasattr(ma_model, "mark_weights_dirty"):
            ma_model.mark_weights_dirty()

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new


class DynamicWeightScheduler:
    """Warm-up x cosine-decay loss weights (reference utils.py:94-134). The weights depend on the epoch
    only; the reference's per-step `.item()` host syncs (utils.py:114) are replaced by a preallocated
    device ring: each step's loss scalars are copied into one ring row (stream-ordered, no sync) and the
    ring is flushed to host floats with ONE device->host copy when it fills or when `loss_history` is
    read. `loss_history` therefore holds Python floats like the reference's, and no device memory is
    retained per step."""

    RING_ROWS = 1024

    def __init__(self, init_weights, warmup_epochs=10, decay_epochs=100, total_epochs=200):
        self.init_weights = init_weights
        self.current_weights = init_weights.copy()
        self.warmup_epochs = warmup_epochs
        self.decay_end_epoch = warmup_epochs + decay_epochs
        self.total_epochs = total_epochs
        self._keys = list(init_weights.keys())
        self._loss_history = {k: [] for k in self._keys}
        self.weight_history = {k: [] for k in self._keys}
        self._ring = None          # [RING_ROWS, len(keys)] fp32 on the losses' device
        self._pending = 0

    # -- loss history: floats, flushed lazily -------------------------------------------------
    @property
    def loss_history(self):
        self._flush()
        return self._loss_history

    @loss_history.setter
    def loss_history(self, value):
        self._pending = 0
        self._loss_history = value

    def _flush(self):
        if self._pending:
            rows = self._ring[:self._pending].cpu().tolist()       # one D2H copy (synchronises once)
            self._pending = 0
            for row in rows:
                for k, v in zip(self._keys, row):
                    self._loss_history[k].append(v)

    def record_device(self, row):
        """Append one step's losses given as a device vector ordered like the weight keys."""
        if self._ring is None or self._ring.device != row.device:
            self._flush()
            self._ring = torch.empty((self.RING_ROWS, len(self._keys)), dtype=F32, device=row.device)
        if self._pending == self.RING_ROWS:
            self._flush()
        self._ring[self._pending].copy_(row, non_blocking=True)
        self._pending += 1

    def _record(self, current_losses):
        vals = [current_losses.get(k) for k in self._keys]
        if any(v is None for v in vals):            # partial dict (e.g. {}): reference semantics, per key
            for k, v in current_losses.items():
                if k in self._loss_history:
                    self._flush()
                    self._loss_history[k].append(float(v))
            return
        if any(torch.is_tensor(v) and v.is_cuda for v in vals):
            dev = next(v.device for v in vals if torch.is_tensor(v) and v.is_cuda)
            self.record_device(torch.stack([v.detach().float().reshape(()) if torch.is_tensor(v)
                                            else torch.tensor(float(v), device=dev) for v in vals]))
        else:
            self._flush()
            for k, v in zip(self._keys, vals):
                self._loss_history[k].append(float(v))

    def get_current_weights(self, epoch, current_losses, record=True):
        """record=False (CUDA-graph capture): compute the weights only; the histories are appended
        after each replay, from that replay's loss values."""
        if not record:
            saved = {k: list(v) for k, v in self.weight_history.items()}
            w = self.get_current_weights(epoch, {})
            self.weight_history = saved
            return w
        if current_losses:
            self._record(current_losses)
        warmup_factor = min(1.0, (epoch + 1) / self.warmup_epochs)
        decay_factor = 1.0
        if epoch >= self.warmup_epochs:
            progress = min(1.0, (epoch - self.warmup_epochs) / (self.decay_end_epoch - self.warmup_epochs))
            decay_factor = 0.1 + 0.9 * 0.5 * (1 + math.cos(math.pi * progress))
        for k in self.current_weights.keys():
            self.current_weights[k] = self.init_weights[k] * warmup_factor * decay_factor
            self.weight_history[k].append(self.current_weights[k])
        return self.current_weights

    def loss_history_values(self):
        return {k: list(v) for k, v in self.loss_history.items()}
