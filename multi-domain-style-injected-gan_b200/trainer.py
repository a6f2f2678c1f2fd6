"""Drop-in mirror of the reference's trainer.py::MultiDomainStyleCycleGAN (constructor signature,
attributes, train_step / save_models / load_models / generate_cyclegan_style_grid), with the step
executed by the C-ABI CUDA kernels and batch-sharded data parallelism over NCCL.

Reference: /root/reference/trainer.py:19-239. Result-preserving differences:
  * the shape-probe D forward (trainer.py:85) is dropped: the LSGAN targets are the constants 1 / 0;
  * D parameter gradients produced during the generator phase are not computed (the reference
    zeroes them at trainer.py:139 before using them);
  * the loss-weight scheduler keeps device scalars instead of calling .item() (utils.py:114);
  * clip_grad_norm_ + Adam + EMA run as one fused pass over flat buffers;
  * with world_size > 1 each rank takes an equal shard of the batch and gradients are averaged with
    NCCL all-reduces on a communication stream: the generator-side one (135.6 MB) overlaps the whole
    discriminator phase (which does not depend on the generator update), the discriminator-side one
    (22.7 MB) overlaps the generator's clip + Adam + EMA pass;
  * the step has static shapes and no host synchronisation, so from the second call with the same
    (batch shape, epoch, learning rates) it is replayed from ONE captured CUDA graph (~1900 kernel
    launches per step otherwise cost as much host time as the GPU needs to run them). The first call
    runs eagerly. `use_cuda_graph=False` (or MSIG_CUDA_GRAPH=0) keeps every step eager.
"""
import copy
import os

import torch

from . import ops
from .losses import L1Loss, MSELoss, VGGStyleContentLoss
from .model import MultiDomainDiscriminator, MultiDomainStyleEncoder, StyleCycleGANGenerator, param_grad_delivery
from .parallel import FlatAllReduce
from .utils import EMA, DynamicWeightScheduler, FlatParams, FusedAdam


def checkpoint_payload(nets, g_optimizer, d_optimizer, g_scheduler, d_scheduler, loss_history, num_domains):
    """The two dicts the reference writes to checkpoint.pth / ema_checkpoint.pth (trainer.py:157-174), from
    `nets` = {name: module} holding the six networks and the four EMA copies. Plain tensors / floats /
    ints only, so a reference build (torch.load with weights_only=True) reads them; the optimizer entries
    use torch.optim.Adam's state_dict layout (FusedAdam.state_dict)."""
    main = {k: nets[k].state_dict() for k in ('G_A2B', 'G_B2A', 'SE_A', 'SE_B', 'D_A', 'D_B')}
    main.update({'g_optimizer': g_optimizer.state_dict(), 'd_optimizer': d_optimizer.state_dict(),
                 'g_scheduler': g_scheduler.state_dict(), 'd_scheduler': d_scheduler.state_dict(),
                 'loss_history': loss_history, 'num_domains': num_domains})
    ema = {k: nets[k].state_dict() for k in ('ema_G_A2B', 'ema_G_B2A', 'ema_SE_A', 'ema_SE_B')}
    return main, ema


class MultiDomainStyleCycleGAN:
    """Multi-domain StyleCycleGAN trainer (reference trainer.py:19-72)."""

    def __init__(self, device, total_epochs, lr_g, lr_d, loss_weights, num_domains, process_group=None,
                 vgg_state=None, use_cuda_graph=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("msig_b200 trainer needs a CUDA device (sm_100a); there is no CPU path")
        self.num_domains = num_domains
        self.process_group = process_group
        self.world_size = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world_size = torch.distributed.get_world_size(process_group)

        # --- models, constructed in the reference's order so that a seeded init is identical
        self.G_A2B = StyleCycleGANGenerator().to(self.device)
        self.G_B2A = StyleCycleGANGenerator().to(self.device)
        self.SE_A = MultiDomainStyleEncoder(num_domains=num_domains).to(self.device)
        self.SE_B = MultiDomainStyleEncoder(num_domains=num_domains).to(self.device)
        self.D_A = MultiDomainDiscriminator(num_domains=num_domains).to(self.device)
        self.D_B = MultiDomainDiscriminator(num_domains=num_domains).to(self.device)

        # --- EMA copies
        self.ema = EMA(beta=0.995)
        self.ema_G_A2B = copy.deepcopy(self.G_A2B).eval()
        self.ema_G_B2A = copy.deepcopy(self.G_B2A).eval()
        self.ema_SE_A = copy.deepcopy(self.SE_A).eval()
        self.ema_SE_B = copy.deepcopy(self.SE_B).eval()

        # --- criteria
        self.criterion_gan = MSELoss()
        self.criterion_cycle = L1Loss()
        self.criterion_identity = L1Loss()
        self.criterion_style_content = VGGStyleContentLoss(self.device, vgg_state=vgg_state)

        # --- flat parameter / gradient buffers + fused optimizers (same parameter order as
        #     trainer.py:56-61)
        g_nets = [self.G_A2B, self.G_B2A, self.SE_A, self.SE_B]
        self._g_flat = FlatParams(g_nets, self.device)
        self._ema_flat = FlatParams([self.ema_G_A2B, self.ema_G_B2A, self.ema_SE_A, self.ema_SE_B], self.device)
        for p in self._ema_flat.params:
            p.requires_grad_(False)
            p.grad = None
        self._d_flat = FlatParams([self.D_A, self.D_B], self.device)
        self.g_optimizer = FusedAdam(self._g_flat, lr=lr_g, betas=(0.5, 0.999), ema_flat=self._ema_flat,
                                     ema_beta=self.ema.beta)
        self.d_optimizer = FusedAdam(self._d_flat, lr=lr_d, betas=(0.5, 0.999))

        self.g_scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.g_optimizer, T_max=total_epochs, eta_min=1e-6)
        self.d_scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.d_optimizer, T_max=total_epochs, eta_min=1e-6)
        self.weight_scheduler = DynamicWeightScheduler(loss_weights, warmup_epochs=10, decay_epochs=100,
                                                       total_epochs=total_epochs)
        self.loss_history = {k: [] for k in (list(loss_weights.keys()) + ['D_loss', 'G_loss'])}
        self.current_epoch_losses = {k: [] for k in self.loss_history.keys()}
        self._comm = FlatAllReduce(process_group, self.device)
        if use_cuda_graph is None:
            use_cuda_graph = os.environ.get("MSIG_CUDA_GRAPH", "1") != "0"
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph_key = None     # configuration of the last call
        self._graph = None         # captured step for that configuration (None until the second call)
        self._graph_stream = None

    def train_step(self, batch, epoch):
        """One G+D optimisation step (reference trainer.py:74-155). Returns the same dict of loss
        tensors: D_loss, G_loss, gan, cycle, identity, style, content."""
        # streams / workspaces of the trainer's device, whatever is current; the flat gradient buffers and the fused
        # optimizer need the wgrad kernels to write param.grad in place ("direct" delivery)
        with torch.cuda.device(self.device), param_grad_delivery("direct"):
            if self.use_cuda_graph:
                return self._train_step_graphed(batch, epoch)
            return self._train_step_eager(batch, epoch)

    def _train_step_eager(self, batch, epoch):
        ops.step_cache_begin()
        try:
            st = {}
            for seg, comm in self._segments(batch, epoch, st, device_step=False):
                torch.cuda.nvtx.range_push("msig." + seg.__name__)
                seg()
                torch.cuda.nvtx.range_pop()
                comm()
            return st["out"]
        finally:
            ops.step_cache_end()

    # ------------------------------------------------------------------ the step, in four segments
    def _segments(self, batch, epoch, st, device_step):
        """The step as (compute segment, communication hook) pairs. A segment only launches kernels
        of this library on the current stream (capturable into a CUDA graph); the hooks hold the
        gradient all-reduces (NCCL, own stream), which stay outside the graphs. `st` carries tensors
        from one segment to the next."""
        dev = self.device
        valid, fake = 1.0, 0.0     # LSGAN targets (all-ones / all-zeros, trainer.py:85-86)
        dp = self.world_size > 1
        scale = self._comm.grad_scale

        def inputs():
            return (batch['source'].to(dev, non_blocking=True), batch['target'].to(dev, non_blocking=True),
                    batch['source_domain'].to(dev, non_blocking=True), batch['target_domain'].to(dev, non_blocking=True))

        def seg_generators():        # trainer.py:88-125
            real_A, real_B, y_org, y_trg = inputs()
            self.g_optimizer.zero_grad()
            self.D_A.skip_param_grads = self.D_B.skip_param_grads = True
            style_A = self.SE_A(real_A, y_org)
            style_B = self.SE_B(real_B, y_trg)
            loss_identity = self.criterion_identity(self.G_A2B(real_B, style_B), real_B)

            fake_B = self.G_A2B(real_A, style_B)
            loss_gan_A2B = self.criterion_gan(self.D_B(fake_B, y_trg), valid)
            content_loss_B, style_loss_B = self.criterion_style_content(fake_B, real_B, real_A)

            fake_A = self.G_B2A(real_B, style_A)
            loss_gan_B2A = self.criterion_gan(self.D_A(fake_A, y_org), valid)
            content_loss_A, style_loss_A = self.criterion_style_content(fake_A, real_A, real_B)

            loss_gan = (loss_gan_A2B + loss_gan_B2A) / 2
            loss_style = (style_loss_A + style_loss_B) / 2
            loss_content = (content_loss_A + content_loss_B) / 2
            loss_cycle = (self.criterion_cycle(self.G_B2A(fake_B, style_A), real_A) +
                          self.criterion_cycle(self.G_A2B(fake_A, style_B), real_B)) / 2

            individual_losses = {'gan': loss_gan, 'cycle': loss_cycle, 'identity': loss_identity,
                                 'style': loss_style, 'content': loss_content}
            weights = self.weight_scheduler.get_current_weights(epoch, individual_losses, record=not device_step)
            g_loss = sum(loss * weights[name] for name, loss in individual_losses.items())
            g_loss.backward()
            self.D_A.skip_param_grads = self.D_B.skip_param_grads = False
            st["fakes"] = (fake_A.detach(), fake_B.detach())
            st["out"] = {'G_loss': g_loss, **individual_losses}

        def comm_g_start():          # overlaps the whole discriminator phase
            st["g_ev"] = self._comm.start(self._g_flat.grad) if dp else None

        def seg_discriminators():    # trainer.py:136-151; independent of the generator update
            real_A, real_B, y_org, y_trg = inputs()
            fake_A_d, fake_B_d = st.pop("fakes")
            self.d_optimizer.zero_grad()
            loss_real_A = self.criterion_gan(self.D_A(real_A, y_org), valid)
            loss_real_B = self.criterion_gan(self.D_B(real_B, y_trg), valid)
            loss_fake_A = self.criterion_gan(self.D_A(fake_A_d, y_org), fake)
            loss_fake_B = self.criterion_gan(self.D_B(fake_B_d, y_trg), fake)
            d_loss = (loss_real_A + loss_fake_A + loss_real_B + loss_fake_B) / 2
            d_loss.backward()
            st["out"] = {'D_loss': d_loss, **st["out"]}

        def comm_g_wait_d_start():   # G grads are needed now; the D-side all-reduce hides behind the G update
            if dp:
                d_ev = self._comm.start(self._d_flat.grad)
                self._comm.wait(st.pop("g_ev"), dev)
                st["d_ev"] = d_ev

        def seg_g_update():          # clip_grad_norm_(1.0) + Adam + EMA (trainer.py:127-134), fused
            self.g_optimizer.step(max_norm=1.0, grad_scale=scale, device_step=device_step)

        def comm_d_wait():
            if dp:
                self._comm.wait(st.pop("d_ev"), dev)

        def seg_d_update():          # trainer.py:152-153
            self.d_optimizer.step(max_norm=1.0, grad_scale=scale, device_step=device_step)

        return [(seg_generators, comm_g_start), (seg_discriminators, comm_g_wait_d_start), (seg_g_update, comm_d_wait),
                (seg_d_update, lambda: None)]

    # ------------------------------------------------------------------ CUDA-graph replay
    _BATCH_KEYS = ('source', 'target', 'source_domain', 'target_domain')

    def _train_step_graphed(self, batch, epoch):
        key = (tuple(batch['source'].shape), tuple(batch['target'].shape), int(epoch),
               float(self.g_optimizer.param_groups[0]['lr']), float(self.d_optimizer.param_groups[0]['lr']))
        if self._graph_stream is None:
            self._graph_stream = torch.cuda.Stream(device=self.device)
        if key != self._graph_key:          # new configuration: one eager step, capture on the next call
            self._graph_key, self._graph = key, None
            # ... on the capture stream: autograd's gradient-accumulator nodes remember the stream they
            # were created on, and a capture may only depend on work of its own stream
            cur = torch.cuda.current_stream(self.device)
            self._graph_stream.wait_stream(cur)
            with torch.cuda.stream(self._graph_stream):
                out = self._train_step_eager(batch, epoch)
            cur.wait_stream(self._graph_stream)
            return out
        if self._graph is None:
            self._graph = self._capture(batch, epoch)
        gs = self._graph
        for k in self._BATCH_KEYS:
            gs["static"][k].copy_(batch[k], non_blocking=True)
        for (graph, comm), name in zip(gs["segments"], gs["seg_names"]):
            torch.cuda.nvtx.range_push(name)
            graph.replay()
            torch.cuda.nvtx.range_pop()
            comm()
        self.g_optimizer.after_replay()
        self.d_optimizer.after_replay()
        ops.add_replayed_launches(gs["launches"])
        vals = gs["out"].clone()            # the graphs' output buffer is overwritten by the next replay
        out = {name: vals[i] for i, name in enumerate(gs["names"])}
        # loss history: one gathered row into the scheduler's device ring (flushed to floats lazily)
        self.weight_scheduler.record_device(vals.index_select(0, gs["hist_idx"]))
        self.weight_scheduler.get_current_weights(epoch, {})
        return out

    def _capture(self, batch, epoch):
        """Captures the four compute segments as four CUDA graphs sharing one memory pool (replayed
        in capture order); the all-reduce hooks run between the replays."""
        dev = self.device
        static = {k: torch.empty(batch[k].shape, dtype=batch[k].dtype, device=dev) for k in self._BATCH_KEYS}
        for k in self._BATCH_KEYS:
            static[k].copy_(batch[k], non_blocking=True)
        for flat in (self._g_flat, self._d_flat):
            flat.mark_dirty()               # the captured step re-packs the bf16 weight copies it uses
        self.g_optimizer.step_counter()
        self.d_optimizer.step_counter()
        torch.cuda.synchronize(dev)
        # The graphs get a private memory pool holding the whole step's working set; hand the eager step's
        # cached blocks back first, or a large batch needs that working set twice (B=128: ~2 x 100 GB).
        torch.cuda.empty_cache()
        l0 = ops.kernel_launches()
        st, segments, pool = {}, [], None
        ops.step_cache_begin()
        try:
            segs = self._segments(static, epoch, st, device_step=True)
            for i, (seg, comm) in enumerate(segs):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, pool=pool, stream=self._graph_stream, capture_error_mode="thread_local"):
                    seg()
                    if i == len(segs) - 1:
                        names = list(st["out"].keys())
                        stacked = torch.stack([st["out"][k].detach().float().reshape(()) for k in names])
                pool = graph.pool()
                segments.append((graph, comm))
        finally:
            ops.step_cache_end()
        launches = ops.kernel_launches() - l0
        ops.add_replayed_launches(-launches)   # recorded, not executed, during capture
        hist_idx = torch.tensor([names.index(k) for k in self.weight_scheduler._keys], dtype=torch.int64, device=dev)
        return {"segments": segments, "static": static, "out": stacked, "names": names, "launches": launches,
                "keep": st, "hist_idx": hist_idx, "seg_names": ["msig." + seg.__name__ for seg, _ in segs]}

    # ------------------------------------------------------------------ checkpoints (trainer.py:157-207)
    def save_models(self, save_dir):
        os.makedirs(save_dir, exist_ok=True)
        nets = {k: getattr(self, k) for k in ('G_A2B', 'G_B2A', 'SE_A', 'SE_B', 'D_A', 'D_B',
                                              'ema_G_A2B', 'ema_G_B2A', 'ema_SE_A', 'ema_SE_B')}
        main, ema = checkpoint_payload(nets, self.g_optimizer, self.d_optimizer, self.g_scheduler, self.d_scheduler,
                                       self.loss_history, self.num_domains)
        torch.save(main, os.path.join(save_dir, 'checkpoint.pth'))
        torch.save(ema, os.path.join(save_dir, 'ema_checkpoint.pth'))
        print(f"Models successfully saved to {save_dir}")

    def load_models(self, checkpoint_dir):
        ckpt_path = os.path.join(checkpoint_dir, 'checkpoint.pth')
        if not os.path.exists(ckpt_path):
            print(f"Checkpoint not found at {ckpt_path}. Starting from scratch.")
            return 0
        print(f"Loading checkpoint from {ckpt_path}...")
        ckpt = torch.load(ckpt_path, map_location=self.device, weights_only=False)
        saved_num_domains = ckpt.get('num_domains', 2)
        if saved_num_domains != self.num_domains:
            print(f"Warning: Saved model has {saved_num_domains} domains, but current model expects {self.num_domains}")
            return 0
        self.G_A2B.load_state_dict(ckpt['G_A2B']); self.G_B2A.load_state_dict(ckpt['G_B2A'])
        self.SE_A.load_state_dict(ckpt['SE_A']); self.SE_B.load_state_dict(ckpt['SE_B'])
        self.D_A.load_state_dict(ckpt['D_A']); self.D_B.load_state_dict(ckpt['D_B'])
        self.g_optimizer.load_state_dict(ckpt['g_optimizer']); self.d_optimizer.load_state_dict(ckpt['d_optimizer'])
        self.g_scheduler.load_state_dict(ckpt['g_scheduler']); self.d_scheduler.load_state_dict(ckpt['d_scheduler'])
        self.loss_history = ckpt.get('loss_history', self.loss_history)
        ema_ckpt_path = os.path.join(checkpoint_dir, 'ema_checkpoint.pth')
        if os.path.exists(ema_ckpt_path):
            ema_ckpt = torch.load(ema_ckpt_path, map_location=self.device, weights_only=False)
            self.ema_G_A2B.load_state_dict(ema_ckpt['ema_G_A2B']); self.ema_G_B2A.load_state_dict(ema_ckpt['ema_G_B2A'])
            self.ema_SE_A.load_state_dict(ema_ckpt['ema_SE_A']); self.ema_SE_B.load_state_dict(ema_ckpt['ema_SE_B'])
        for flat in (self._g_flat, self._d_flat, self._ema_flat):
            flat.mark_dirty()
        self._graph_key = self._graph = None    # optimizer step counts changed: re-capture
        print(f"Models successfully loaded from {checkpoint_dir}")
        return len(self.loss_history.get('G_loss', []))

    def plot_losses(self, save_path):
        """Loss curves (reference trainer.py:209-217); needs matplotlib, which is optional here."""
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            print("matplotlib is not installed; skipping loss plot")
            return
        if not self.loss_history or not any(v for k, v in self.loss_history.items() if k in ['G_loss', 'D_loss']):
            return
        plt.figure(figsize=(12, 8))
        epochs = range(1, len(self.loss_history['G_loss']) + 1)
        for loss_type, values in self.loss_history.items():
            if values:
                plt.plot(epochs, values, label=loss_type)
        plt.legend(); plt.xlabel('Epochs'); plt.ylabel('Loss'); plt.title('Training Losses Over Epochs')
        plt.grid(True, linestyle='--', alpha=0.6)
        plt.savefig(save_path, dpi=300); plt.close()

    def generate_cyclegan_style_grid(self, batch):
        """2x2 grid Real A, Fake B, Real B, Fake A from the EMA networks (trainer.py:219-239)."""
        with torch.no_grad():
            real_A = batch['source'][0:1].to(self.device)
            real_B = batch['target'][0:1].to(self.device)
            y_org = batch['source_domain'][0:1].to(self.device)
            y_trg = batch['target_domain'][0:1].to(self.device)
            style_A = self.ema_SE_A(real_A, y_org)
            style_B = self.ema_SE_B(real_B, y_trg)
            fake_B = self.ema_G_A2B(real_A, style_B)
            fake_A = self.ema_G_B2A(real_B, style_A)
            grid = torch.cat([real_A, fake_B, real_B, fake_A], dim=0)
            return grid, y_trg[0].item()
