"""Drop-in mirror of the reference's model.py (same class names, constructor signatures,
parameter names / state_dict layout and seeded-init behaviour) whose forward AND backward are
sequenced explicitly over the C-ABI CUDA kernels (libmsig.so): tcgen05 implicit-GEMM convs,
fused InstanceNorm/AdaIN kernels, gathered-patch GEMMs for the 3-channel layers.

Reference: /root/reference/model.py:9-214. Tensors at the module surface are fp32 NCHW like the
reference's; internally activations are bf16 NHWC and accumulation / statistics are fp32.

Autograd contract: each network is ONE torch.autograd.Function. Its backward returns the gradients
of the image and the style code; by default ("direct" delivery) parameter gradients are accumulated by
the wgrad kernels DIRECTLY into `param.grad` (allocated on demand), so flat gradient buffers (trainer /
data parallel) are written in place and `torch.autograd.grad(out, params)` reports None for parameters.
`set_param_grad_delivery("autograd")` switches to returning them from backward like any PyTorch op
(`torch.autograd.grad`, tensor hooks, `GradScaler`, a DDP wrapper then see them; costs one extra
accumulation pass); the drop-in trainer always runs "direct".
"""
import contextlib
import functools

import torch
import torch.nn as nn

from . import ops
from .lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, AUX_ADD, AUX_RELU_MASK, AUX_LRELU_MASK,
                  OUT_F32_NCHW, OUT_F32_NHWC, WPACK_CONVT_DGRAD, WPACK_CONVT_FWD, WPACK_DGRAD_S1,
                  WPACK_DGRAD_S2, WPACK_FWD, WPACK_IM2COL, WPACK_IM2COL_DGRAD, WPACK_IM2COL_FLIP,
                  WPACK_ROWFOLD, WPACK_ROWFOLD_DGRAD, WPACK_ROWPATCH, WPACK_ROWPATCH_FLIP)

N_RESIDUAL_BLOCKS = 8   # reference config.py:19
F32 = torch.float32
BF16 = torch.bfloat16


_DELIVERY = "direct"     # "direct": wgrad kernels accumulate into param.grad; "autograd": backward returns the gradients
_sink = None             # id(param) -> gradient tensor of the backward call in flight ("autograd" delivery)


def set_param_grad_delivery(mode):
    """How the network Functions hand out PARAMETER gradients: "direct" (default) or "autograd" (see the module
    docstring). Process-wide; returns the previous mode."""
    global _DELIVERY
    if mode not in ("direct", "autograd"):
        raise ValueError("param grad delivery must be 'direct' or 'autograd'")
    prev, _DELIVERY = _DELIVERY, mode
    ops.PTR_TABLE_CACHE = mode == "direct"      # the cached device pointer tables assume gradient buffers that stay put
    return prev


@contextlib.contextmanager
def param_grad_delivery(mode):
    prev = set_param_grad_delivery(mode)
    try:
        yield
    finally:
        set_param_grad_delivery(prev)


def _grad_buf(p):
    """The accumulation target of a parameter's gradient: param.grad (zero-filled on first use), or in "autograd"
    delivery a fresh zero tensor that the running backward returns for that parameter."""
    if _sink is not None:
        g = _sink.get(id(p))
        if g is None:
            g = _sink[id(p)] = torch.zeros_like(p)
        return g
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


def _delivers_param_grads(backward):
    """Decorator of a network Function's backward (inputs: module, two tensors, then *module.parameters()): in
    "autograd" delivery the trailing Nones become the gradients the wgrad kernels wrote during this call."""
    @functools.wraps(backward)
    def wrapped(ctx, *grads):
        global _sink
        if _DELIVERY != "autograd":
            return backward(ctx, *grads)
        params = list(ctx.mod.parameters())
        outer, _sink = _sink, {}
        try:
            out = backward(ctx, *grads)
            sink = _sink
        finally:
            _sink = outer
        return tuple(out[:len(out) - len(params)]) + tuple(sink.get(id(p)) for p in params)
    return wrapped


def _trace(mod, name, t):
    """Test hook (tests/layerwise_cases.py): when a module carries a `_msig_trace` dict, the network-level
    Functions record their saved activations and the intermediate gradient tensors of their backward in it,
    so that every layer can be checked on its own against the oracle. A no-op otherwise."""
    tr = mod.__dict__.get("_msig_trace")
    if tr is not None:
        tr[name] = t


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor; msig_b200 has no CPU path")


def _conv_stats(x, wpk, g, gamma=None, beta=None, gstride=0, transposed=False):
    """conv (or k4 s2 transposed conv) whose epilogue also emits the InstanceNorm / AdaIN statistics of
    its output: returns (z, NormStats). Replaces conv + a separate statistics pass (model.py:16,28-36)."""
    if not ops.epi_fusable(4 if transposed else g.r * g.s, g.c):   # short K: the epilogue would dominate
        es = ops.ring_stats(ops.RING_CONVT if transposed else ops.RING_CONV, g, x.device)
        if es is not None:      # strip-ring kernel: per-item statistics out of the lean epilogue
            e = ops.epilogue(stats=es)
            z = ops.convT2d_fwd(x, wpk, g, e) if transposed else ops.conv2d_fwd(x, wpk, g, e)
            return z, ops.in_stats_from(es, g.oh * g.ow, g.k, gamma, beta, gstride)
        z = ops.convT2d_fwd(x, wpk, g) if transposed else ops.conv2d_fwd(x, wpk, g)
        return z, ops.in_stats(z, gamma, beta, gstride)
    if transposed:
        es = ops.epi_stats(g.n, g.h, g.w, g.k, x.device, phases=4)
        z = ops.convT2d_fwd(x, wpk, g, ops.epilogue(stats=es))
    else:
        es = ops.epi_stats(g.n, g.oh, g.ow, g.k, x.device)
        z = ops.conv2d_fwd(x, wpk, g, ops.epilogue(stats=es))
    return z, ops.in_stats_from(es, g.oh * g.ow, g.k, gamma, beta, gstride)


class _PackedWeights:
    """bf16 packed copies of a module's fp32 master weights, rebuilt when the parameters change
    (detected through tensor versions plus an explicit dirty counter bumped by the fused optimizer,
    which updates parameters through raw pointers)."""

    def __init__(self, module):
        self._module = module
        self._key = None
        self._table = None      # ops.PackTable: all packs of the network in one launch
        self._ptrs = None
        self.t = {}

    def get(self):
        m = self._module
        key = (m._dirty,) + tuple((p._version, p.data_ptr()) for p in m.parameters())
        if key != self._key:
            ptrs = tuple(p.data_ptr() for p in m.parameters())
            with torch.no_grad():
                if self._table is None or ptrs != self._ptrs:
                    # first use (or the parameters moved): run the packs one by one and record them
                    with ops.record_packs() as jobs:
                        m._pack(self.t)
                    self._table = ops.PackTable(jobs, next(m.parameters()).device)
                    self._ptrs = ptrs
                else:
                    self._table.run()
            self._key = key
        return self.t


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self._dirty = 0
        self._packed = _PackedWeights(self)
        self.skip_param_grads = False   # trainer sets this on D during the generator phase

    def mark_weights_dirty(self):
        self._dirty += 1
        for m in self._child_nets():       # nested stand-alone-capable blocks keep their own packed copies
            m._dirty += 1

    def _child_nets(self):
        c = self.__dict__.get("_child_nets_cache")
        if c is None:
            c = [m for m in self.modules() if m is not self and isinstance(m, _Net)]
            self.__dict__["_child_nets_cache"] = c
        return c

    def _apply(self, fn, *a, **k):   # .to()/.cuda() move parameters: invalidate packed copies
        r = super()._apply(fn, *a, **k)
        self._dirty += 1
        return r

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_packed", "_child_nets_cache") or k.startswith("_msig_"):   # caches are per instance
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._packed = _PackedWeights(new)
        return new


# ######################################################################
# Building blocks (parameter containers; the arithmetic lives in the network-level Functions)
# ######################################################################
class AdaIN(nn.Module):
    """Adaptive Instance Normalization (reference model.py:9-36). Stand-alone use runs the fused
    statistics + modulation kernels; inside the generator it is part of the fused network pass."""

    def __init__(self, content_channels, style_dim):
        super().__init__()
        self.instance_norm = nn.InstanceNorm2d(content_channels, affine=False)
        self.style_modulation = nn.Linear(style_dim, content_channels * 2)

    def forward(self, content_features, style_code):
        return _AdaINFn.apply(content_features, style_code, self.style_modulation.weight,
                              self.style_modulation.bias)


class ResidualBlockWithAdaIN(_Net):
    """Residual block with two AdaIN layers (reference model.py:38-55). Inside the generator the block
    is part of the fused network pass (_GeneratorFn); called on its own it runs the same kernels through
    _ResBlockFn: conv + epilogue statistics, fused AdaIN apply (+ReLU / +residual), dgrads with fused
    mask + norm-backward reductions."""

    def __init__(self, channels, style_dim):
        super().__init__()
        self.channels, self.style_dim = channels, style_dim
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.adain1 = AdaIN(channels, style_dim)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.adain2 = AdaIN(channels, style_dim)

    def _pack(self, t):
        c, sd = self.channels, self.style_dim
        dev = self.conv1.weight.device
        if "lin" not in t:
            t["lin"] = torch.zeros(2 * 2 * c * sd, dtype=BF16, device=dev)
            t["lin_d"] = torch.zeros(ops.pad_rows(sd) * 2 * 2 * c, dtype=BF16, device=dev)
            t["lin_b"] = torch.zeros(2 * 2 * c, dtype=F32, device=dev)
        for j, (conv, ada) in enumerate(((self.conv1, self.adain1), (self.conv2, self.adain2))):
            t[f"w{j}"] = ops.wpack(WPACK_FWD, conv.weight, c, c, 3, 3, out=t.get(f"w{j}"))
            t[f"w{j}_d"] = ops.wpack(WPACK_DGRAD_S1, conv.weight, c, c, 3, 3, out=t.get(f"w{j}_d"))
            ops.wpack(WPACK_FWD, ada.style_modulation.weight, 2 * c, sd, 1, 1, out=t["lin"], oc=4 * c, o_off=j * 2 * c)
            ops.wpack(WPACK_DGRAD_S1, ada.style_modulation.weight, 2 * c, sd, 1, 1, out=t["lin_d"], oc=4 * c,
                      o_off=j * 2 * c)
            ops.copy_f32(t["lin_b"][j * 2 * c:(j + 1) * 2 * c], ada.style_modulation.bias)

    def _dead_biases(self):
        return [self.conv1.bias, self.conv2.bias]

    def forward(self, x, style_code):
        return _ResBlockFn.apply(self, x, style_code, *self.parameters())


def _style_linears_backward(dgb, B, bs, nl, c2, sd, sb, lin_d, adains, wg, want_dstyle, style_shape):
    """Backward of the batched style Linears (model.py:28): dgb [B, nl*c2] fp32 holds (dgamma | dbeta) of
    every AdaIN site. Returns dstyle (or None) and accumulates dW / db into the Linear parameters."""
    if bs == 1 and B > 1:                      # one style code broadcast over the batch
        red = torch.empty((1, nl * c2), dtype=F32, device=dgb.device)
        ops.colsum_f32(dgb, B, nl * c2, red, accumulate=False)
        dgb = red
    rows = dgb.shape[0]
    dgbb = ops.to_bf16(dgb)
    dstyle = None
    if want_dstyle:
        dstyle = ops.conv2d_fwd(dgbb.view(1, 1, rows, nl * c2), lin_d, ops.gemm_geom(rows, nl * c2, sd),
                                ops.epilogue(out_layout=OUT_F32_NHWC)).view(rows, sd).view(style_shape)
    if wg:
        ws, splits = ops.gemm_tn_partial(rows, dgbb, nl * c2, sb, sd)
        ops.multi_linear_grads(ws, splits, nl * c2 * sd, dgb, rows, nl * c2, nl, c2, sd,
                               [_grad_buf(a.style_modulation.weight) for a in adains],
                               [_grad_buf(a.style_modulation.bias) for a in adains])
    return dstyle


class _ResBlockFn(torch.autograd.Function):
    """x + AdaIN2(conv2(ReLU(AdaIN1(conv1(x))))) on fp32 NCHW features (model.py:51-55)."""

    @staticmethod
    @ops.dev_guard
    def forward(ctx, mod, x, style, *params):
        _require_cuda(x, "ResidualBlockWithAdaIN")
        ops.ensure_init(x.device)
        P = mod._packed.get()
        c, sd = mod.channels, mod.style_dim
        B, C, H, W = x.shape
        if C != c or c % 64:
            raise RuntimeError(f"ResidualBlockWithAdaIN: expected {c} channels (a multiple of 64), got {C}")
        s2 = _style_2d(style).contiguous().float()
        bs = s2.shape[0]
        if bs not in (1, B):
            raise RuntimeError(f"style code batch {bs} does not match feature batch {B}")
        xh = ops.to_bf16(x.float().permute(0, 2, 3, 1).contiguous())
        sb = ops.to_bf16(s2)
        c2 = 2 * c
        gb = ops.conv2d_fwd(sb.view(1, 1, bs, sd), P["lin"], ops.gemm_geom(bs, sd, 2 * c2),
                            ops.epilogue(bias=P["lin_b"], out_layout=OUT_F32_NHWC)).view(bs, 2 * c2)
        gstride = 0 if bs == 1 else 2 * c2
        g3 = ops.conv_geom(B, H, W, c, c, 3, 3, 1, 1, 1, H, W)
        za, sta = _conv_stats(xh, P["w0"], g3, gb[:, 0:], gb[:, c:], gstride)
        ha = ops.norm_act_fwd(za, sta, ACT_RELU)
        zb, stb = _conv_stats(ha, P["w1"], g3, gb[:, c2:], gb[:, c2 + c:], gstride)
        out = ops.norm_act_fwd(zb, stb, ACT_NONE, residual=xh)
        if any(ctx.needs_input_grad):
            ctx.mod = mod
            ctx.saved = dict(xh=xh, za=za, sta=sta, ha=ha, zb=zb, stb=stb, sb=sb, bs=bs, g3=g3,
                             style_shape=style.shape, x_grad=ctx.needs_input_grad[1],
                             style_grad=ctx.needs_input_grad[2])
        return ops.to_f32(out).permute(0, 3, 1, 2)

    @staticmethod
    @_delivers_param_grads
    @ops.dev_guard
    def backward(ctx, dout):
        mod, S = ctx.mod, ctx.saved
        P = mod._packed.get()
        c, sd = mod.channels, mod.style_dim
        wg = not mod.skip_param_grads
        g3 = S["g3"]
        B = g3.n
        c2 = 2 * c
        dy = ops.to_bf16(dout.float().permute(0, 2, 3, 1).contiguous())
        dgb = torch.zeros((B, 2 * c2), dtype=F32, device=dout.device)
        # second AdaIN (no activation): the upstream gradient comes from outside, so its two reductions
        # are a separate pass; the dgrads below emit the reductions of the first AdaIN in their epilogue
        dzb = ops.norm_act_bwd(dy, S["zb"], S["stb"], ACT_NONE, dgamma=dgb[:, c2:], dbeta=dgb[:, c2 + c:],
                               dgb_stride=2 * c2)
        if wg:
            ops.conv2d_wgrad(S["ha"], dzb, g3, _grad_buf(mod.conv2.weight))
        if ops.epi_fusable(9, c):
            es = ops.epi_stats(B, g3.h, g3.w, c, dout.device)
            dh = ops.conv2d_dgrad(dzb, P["w1_d"], g3,
                                  ops.epilogue(aux=S["ha"], aux_mode=AUX_RELU_MASK, stats=es, stats_z=S["za"],
                                               mask_norm=S["sta"]))
            dza = ops.norm_bwd_from(es, dh, S["za"], S["sta"], dgamma=dgb[:, 0:], dbeta=dgb[:, c:], dgb_stride=2 * c2)
        else:
            dh = ops.conv2d_dgrad(dzb, P["w1_d"], g3)
            dza = ops.norm_act_bwd(dh, S["za"], S["sta"], ACT_RELU, dgamma=dgb[:, 0:], dbeta=dgb[:, c:],
                                   dgb_stride=2 * c2)
        if wg:
            ops.conv2d_wgrad(S["xh"], dza, g3, _grad_buf(mod.conv1.weight))
            for p in mod._dead_biases():
                _grad_buf(p)
        dx = None
        if S["x_grad"]:
            dxh = ops.conv2d_dgrad(dza, P["w0_d"], g3, ops.epilogue(aux=dy, aux_mode=AUX_ADD))
            dx = ops.to_f32(dxh).permute(0, 3, 1, 2)
        dstyle = _style_linears_backward(dgb, B, S["bs"], 2, c2, sd, S["sb"], P["lin_d"], (mod.adain1, mod.adain2),
                                         wg, S["style_grad"], S["style_shape"])
        ctx.saved = None
        return (None, dx, dstyle) + (None,) * (len(ctx.needs_input_grad) - 3)


def _style_2d(style_code):
    if style_code.dim() == 4:
        style_code = style_code.squeeze(-1).squeeze(-1)
    return style_code


class _AdaINFn(torch.autograd.Function):
    """Stand-alone AdaIN (model.py:20-36) on fp32 NCHW features: Linear -> gamma/beta, fused
    statistics + modulation. Used by the unit tests of the building block."""

    @staticmethod
    @ops.dev_guard
    def forward(ctx, x, style, w, b):
        _require_cuda(x, "AdaIN")
        ops.ensure_init(x.device)
        n, c, h, wd = x.shape
        s2 = _style_2d(style).contiguous().float()
        xh = ops.to_bf16(x.permute(0, 2, 3, 1).contiguous())
        wpk = ops.wpack(WPACK_FWD, w.detach().contiguous(), 2 * c, w.shape[1], 1, 1)
        sb = ops.to_bf16(s2)
        gb = ops.conv2d_fwd(sb.view(1, 1, s2.shape[0], -1), wpk, ops.gemm_geom(s2.shape[0], w.shape[1], 2 * c),
                            ops.epilogue(bias=b.detach().contiguous(), out_layout=OUT_F32_NHWC)).view(s2.shape[0], 2 * c)
        stride = 0 if s2.shape[0] == 1 else 2 * c
        st = ops.in_stats(xh, gb[:, :c], gb[:, c:], stride)
        y = ops.norm_act_fwd(xh, st, ACT_NONE)
        ctx.saved = (xh, st, sb, w, style.shape, s2.shape[0])
        return ops.to_f32(y).permute(0, 3, 1, 2)

    @staticmethod
    @ops.dev_guard
    def backward(ctx, dy):
        xh, st, sb, w, style_shape, bs = ctx.saved
        n, h, wd, c = xh.shape
        dyh = ops.to_bf16(dy.permute(0, 2, 3, 1).contiguous())
        dgb = torch.zeros((n, 2 * c), dtype=F32, device=dy.device)
        dx = ops.norm_act_bwd(dyh, xh, st, ACT_NONE, dgamma=dgb[:, :c], dbeta=dgb[:, c:], dgb_stride=2 * c)
        if bs == 1 and n > 1:
            red = torch.zeros((1, 2 * c), dtype=F32, device=dy.device)
            ops.colsum_f32(dgb, n, 2 * c, red, accumulate=False)
            dgb = red
        db = torch.zeros(2 * c, dtype=F32, device=dy.device)
        ops.colsum_f32(dgb, dgb.shape[0], 2 * c, db, accumulate=False)
        dgbb = ops.to_bf16(dgb)
        wd_pk = ops.wpack(WPACK_DGRAD_S1, w.detach().contiguous(), 2 * c, w.shape[1], 1, 1)
        ds = ops.conv2d_fwd(dgbb.view(1, 1, dgb.shape[0], 2 * c), wd_pk,
                            ops.gemm_geom(dgb.shape[0], 2 * c, w.shape[1]),
                            ops.epilogue(out_layout=OUT_F32_NHWC)).view(dgb.shape[0], -1)
        dw = torch.zeros_like(w)
        ops.patch_wgrad(WPACK_FWD, 2 * c, w.shape[1], 1, 1, dgb.shape[0], dgbb, 2 * c, sb, w.shape[1], dw,
                        accumulate=False)
        return ops.to_f32(dx).permute(0, 3, 1, 2), ds.view(style_shape), dw, db


# ######################################################################
# Generator
# ######################################################################
class StyleCycleGANGenerator(_Net):
    """Content encoder + 8 AdaIN residual blocks + decoder (reference model.py:121-151)."""

    def __init__(self, in_channels=3, out_channels=3, style_dim=256, n_residual_blocks=N_RESIDUAL_BLOCKS):
        super().__init__()
        if in_channels != 3 or out_channels != 3:
            raise RuntimeError("msig_b200 generator supports 3-channel images only (the reference's configuration)")
        self.style_dim = style_dim
        self.n_res = n_residual_blocks
        self.content_encoder = nn.Sequential(
            nn.Conv2d(in_channels, 64, 7, 1, 3, padding_mode='reflect'), nn.InstanceNorm2d(64), nn.ReLU(inplace=True),
            nn.Conv2d(64, 128, 4, 2, 1), nn.InstanceNorm2d(128), nn.ReLU(inplace=True),
            nn.Conv2d(128, 256, 4, 2, 1), nn.InstanceNorm2d(256), nn.ReLU(inplace=True)
        )
        decoder_blocks = [ResidualBlockWithAdaIN(256, style_dim) for _ in range(n_residual_blocks)]
        decoder_blocks.extend([
            nn.ConvTranspose2d(256, 128, 4, 2, 1), nn.InstanceNorm2d(128), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(128, 64, 4, 2, 1), nn.InstanceNorm2d(64), nn.ReLU(inplace=True),
            nn.Conv2d(64, out_channels, 7, 1, 3, padding_mode='reflect'), nn.Tanh()
        ])
        self.decoder = nn.ModuleList(decoder_blocks)

    # -- packed weights ------------------------------------------------------------------
    def _pack(self, t):
        enc, dec, k, sd = self.content_encoder, self.decoder, self.n_res, self.style_dim
        t["e0"] = ops.wpack(WPACK_ROWPATCH, enc[0].weight, 64, 3, 7, 7, out=t.get("e0"))
        t["e0_rf"] = ops.wpack(WPACK_ROWFOLD_DGRAD, enc[0].weight, 64, 3, 7, 7, out=t.get("e0_rf"))
        t["e1"] = ops.wpack(WPACK_FWD, enc[3].weight, 128, 64, 4, 4, out=t.get("e1"))
        t["e1_d"] = ops.wpack(WPACK_DGRAD_S2, enc[3].weight, 128, 64, 4, 4, out=t.get("e1_d"))
        t["e2"] = ops.wpack(WPACK_FWD, enc[6].weight, 256, 128, 4, 4, out=t.get("e2"))
        t["e2_d"] = ops.wpack(WPACK_DGRAD_S2, enc[6].weight, 256, 128, 4, 4, out=t.get("e2_d"))
        nl = 2 * k
        if "lin" not in t:
            t["lin"] = torch.zeros(nl * 512 * sd, dtype=BF16, device=enc[0].weight.device)
            t["lin_d"] = torch.zeros(ops.pad_rows(sd) * nl * 512, dtype=BF16, device=enc[0].weight.device)
            t["lin_b"] = torch.zeros(nl * 512, dtype=F32, device=enc[0].weight.device)
        for i in range(k):
            blk = dec[i]
            for j, (conv, ada) in enumerate(((blk.conv1, blk.adain1), (blk.conv2, blk.adain2))):
                t[f"r{i}{j}"] = ops.wpack(WPACK_FWD, conv.weight, 256, 256, 3, 3, out=t.get(f"r{i}{j}"))
                t[f"r{i}{j}_d"] = ops.wpack(WPACK_DGRAD_S1, conv.weight, 256, 256, 3, 3, out=t.get(f"r{i}{j}_d"))
                l = 2 * i + j
                ops.wpack(WPACK_FWD, ada.style_modulation.weight, 512, sd, 1, 1, out=t["lin"], oc=nl * 512, o_off=l * 512)
                ops.wpack(WPACK_DGRAD_S1, ada.style_modulation.weight, 512, sd, 1, 1, out=t["lin_d"], oc=nl * 512,
                          o_off=l * 512)
                ops.copy_f32(t["lin_b"][l * 512:(l + 1) * 512], ada.style_modulation.bias)
        t["u1"] = ops.wpack(WPACK_CONVT_FWD, dec[k].weight, 128, 256, 4, 4, out=t.get("u1"))
        t["u1_d"] = ops.wpack(WPACK_CONVT_DGRAD, dec[k].weight, 128, 256, 4, 4, out=t.get("u1_d"))
        t["u2"] = ops.wpack(WPACK_CONVT_FWD, dec[k + 3].weight, 64, 128, 4, 4, out=t.get("u2"))
        t["u2_d"] = ops.wpack(WPACK_CONVT_DGRAD, dec[k + 3].weight, 64, 128, 4, 4, out=t.get("u2_d"))
        t["f"] = ops.wpack(WPACK_ROWFOLD, dec[k + 6].weight, 3, 64, 7, 7, out=t.get("f"))
        t["f_d"] = ops.wpack(WPACK_ROWPATCH_FLIP, dec[k + 6].weight, 3, 64, 7, 7, out=t.get("f_d"))

    def forward(self, content_image, style_code):
        return _GeneratorFn.apply(self, content_image, style_code, *self.parameters())


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, mod, img, style, *params):
        _require_cuda(img, "StyleCycleGANGenerator")
        ops.ensure_init(img.device)
        P = mod._packed.get()
        k, sd = mod.n_res, mod.style_dim
        img = img.contiguous().float()
        B, _, H, W = img.shape
        if H % 4 or W % 4:
            raise RuntimeError("generator input height/width must be multiples of 4")
        s2 = _style_2d(style).contiguous().float()
        bs = s2.shape[0]
        if bs not in (1, B):
            raise RuntimeError(f"style code batch {bs} does not match image batch {B}")
        need_grad = any(ctx.needs_input_grad)   # (forward itself always runs with grad mode off)
        S = {}   # saved activations
        # ---- content encoder (model.py:130-134)
        # 7x7 reflect conv on the 3-channel image: row-patch implicit GEMM over the padded bf16 copy
        g0 = ops.conv_geom(B, H, W, 3, 64, 7, 7, 1, 3, 3, H, W)
        es0 = ops.ring_stats(ops.RING_ROWPATCH, g0, img.device)   # (None on planes narrower than 128 pixels)
        z0 = ops.conv_rowpatch_fwd(ops.img_pad8_cached(img, 3, True), P["e0"], g0,
                                   None if es0 is None else ops.epilogue(stats=es0))
        st0 = ops.in_stats(z0) if es0 is None else ops.in_stats_from(es0, H * W, 64)
        y0 = ops.norm_act_fwd(z0, st0, ACT_RELU)
        g1 = ops.conv_geom(B, H, W, 64, 128, 4, 4, 2, 1, 1, H // 2, W // 2)
        z1, st1 = _conv_stats(y0, P["e1"], g1)
        y1 = ops.norm_act_fwd(z1, st1, ACT_RELU)
        g2 = ops.conv_geom(B, H // 2, W // 2, 128, 256, 4, 4, 2, 1, 1, H // 4, W // 4)
        z2, st2 = _conv_stats(y1, P["e2"], g2)
        x = ops.norm_act_fwd(z2, st2, ACT_RELU)
        # ---- all 2k style Linears in one GEMM (model.py:28)
        nl = 2 * k
        sb = ops.to_bf16(s2)
        gb = ops.conv2d_fwd(sb.view(1, 1, bs, sd), P["lin"], ops.gemm_geom(bs, sd, nl * 512),
                            ops.epilogue(bias=P["lin_b"], out_layout=OUT_F32_NHWC)).view(bs, nl * 512)
        gstride = 0 if bs == 1 else nl * 512
        # ---- residual AdaIN blocks (model.py:51-55)
        h4, w4 = H // 4, W // 4
        g3 = ops.conv_geom(B, h4, w4, 256, 256, 3, 3, 1, 1, 1, h4, w4)
        res = []
        for i in range(k):
            l = 2 * i
            za, sta = _conv_stats(x, P[f"r{i}0"], g3, gb[:, l * 512:], gb[:, l * 512 + 256:], gstride)
            ha = ops.norm_act_fwd(za, sta, ACT_RELU)
            zb, stb = _conv_stats(ha, P[f"r{i}1"], g3, gb[:, (l + 1) * 512:], gb[:, (l + 1) * 512 + 256:], gstride)
            xn = ops.norm_act_fwd(zb, stb, ACT_NONE, residual=x)
            res.append((x, za, sta, ha, zb, stb))
            x = xn
        # ---- decoder (model.py:139-141)
        gu1 = ops.conv_geom(B, h4, w4, 256, 128, 4, 4, 2, 1, 1, H // 2, W // 2)
        zu1, stu1 = _conv_stats(x, P["u1"], gu1, transposed=True)
        yu1 = ops.norm_act_fwd(zu1, stu1, ACT_RELU)
        gu2 = ops.conv_geom(B, H // 2, W // 2, 128, 64, 4, 4, 2, 1, 1, H, W)
        zu2, stu2 = _conv_stats(yu1, P["u2"], gu2, transposed=True)
        xp = ops.norm_act_fwd_pad(zu2, stu2, ACT_RELU, 3)   # IN + ReLU written straight into the reflect-padded buffer
        gf = ops.conv_geom(B, H + 6, W + 6, 64, 3, 7, 7, 1, 0, 0, H, W)
        out = ops.conv_narrow_fwd(xp, P["f"], gf, ops.epilogue(bias=mod.decoder[k + 6].bias.detach(), act=ACT_TANH,
                                                               out_layout=OUT_F32_NCHW))
        if need_grad:
            ctx.mod = mod
            ctx.saved = dict(img=img, g0=g0, z0=z0, st0=st0, y0=y0, g1=g1, z1=z1, st1=st1, y1=y1, g2=g2, z2=z2,
                             st2=st2, x2=res[0][0] if k else x, sb=sb, bs=bs, res=res, g3=g3, x_res=x, gu1=gu1,
                             zu1=zu1, stu1=stu1, yu1=yu1, gu2=gu2, zu2=zu2, stu2=stu2, xp=xp, out=out,
                             style_shape=style.shape, img_grad=ctx.needs_input_grad[1], style_grad=ctx.needs_input_grad[2])
            _trace(mod, "fwd", ctx.saved)
            _trace(mod, "gb", gb)
        return out

    @staticmethod
    @_delivers_param_grads
    @ops.dev_guard
    def backward(ctx, dout):
        mod, S = ctx.mod, ctx.saved
        P = mod._packed.get()
        k, sd = mod.n_res, mod.style_dim
        enc, dec = mod.content_encoder, mod.decoder
        wg = not mod.skip_param_grads
        out = S["out"]
        B, _, H, W = out.shape
        h4, w4 = H // 4, W // 4
        dout = dout.contiguous().float()
        # ---- final 7x7 reflect conv + tanh (model.py:141): gather dy patches ("full" correlation)
        dz = ops.tanh_bwd(dout, out)
        _trace(mod, "dz_f", dz)
        if wg:
            ops.nchw_chansum(dz, _grad_buf(dec[k + 6].bias))
        gfd = ops.conv_geom(B, H, W, 3, 64, 7, 7, 1, 6, 6, H + 6, W + 6)
        dz8 = ops.img_pad8(dz, 6, False)            # zero-extended bf16 copy of the 3-channel gradient
        dxp = ops.conv_rowpatch_fwd(dz8, P["f_d"], gfd)
        _trace(mod, "dxp", dxp)
        if wg:
            ops.conv_rowpatch_wgrad(dz8, S["xp"], gfd, _grad_buf(dec[k + 6].weight), flip=True)
        del dz8
        # ---- up 2 (ConvTranspose 128->64 + IN + ReLU); the reflect fold of dxp happens inside the norm backward
        dzu2 = ops.norm_act_bwd_pad(dxp, S["zu2"], S["stu2"], ACT_RELU, 3)
        _trace(mod, "dzu2", dzu2)
        del dxp
        if wg:
            ops.convT2d_wgrad(S["yu1"], dzu2, S["gu2"], _grad_buf(dec[k + 3].weight))
        # Every dgrad below also applies the previous layer's ReLU mask and reduces (sum g, sum g*z)
        # over pixels in its epilogue, so the norm backward that follows needs no reduction pass.
        dev = dout.device
        gu2, gu1 = S["gu2"], S["gu1"]
        es = ops.epi_stats(B, gu2.h, gu2.w, gu2.c, dev) if ops.FUSE_N128_REDUCTIONS else None
        dy = ops.convT2d_dgrad(dzu2, P["u2_d"], gu2,
                               ops.epilogue(aux=S["yu1"], aux_mode=AUX_RELU_MASK, stats=es, stats_z=S["zu1"],
                                            mask_norm=S["stu1"]))
        del dzu2
        # ---- up 1 (ConvTranspose 256->128 + IN + ReLU)
        _trace(mod, "dyu1", dy)
        dzu1 = (ops.norm_bwd_from(es, dy, S["zu1"], S["stu1"]) if es is not None
                else ops.norm_act_bwd(dy, S["zu1"], S["stu1"], ACT_NONE))
        _trace(mod, "dzu1", dzu1)
        if wg:
            ops.convT2d_wgrad(S["x_res"], dzu1, gu1, _grad_buf(dec[k].weight))
        es = ops.epi_stats(B, gu1.h, gu1.w, gu1.c, dev) if k else None
        dy = ops.convT2d_dgrad(dzu1, P["u1_d"], gu1,
                               ops.epilogue(stats=es, stats_z=S["res"][k - 1][4]) if k else None)
        del dzu1
        # ---- residual blocks, reversed
        nl = 2 * k
        dgb = torch.zeros((B, nl * 512), dtype=F32, device=dout.device)
        _trace(mod, "dgb", dgb)
        _trace(mod, "dx_res", dy)
        g3 = S["g3"]
        for i in reversed(range(k)):
            x_in, za, sta, ha, zb, stb = S["res"][i]
            blk = dec[i]
            l = 2 * i
            dzb = ops.norm_bwd_from(es, dy, zb, stb, dgamma=dgb[:, (l + 1) * 512:], dbeta=dgb[:, (l + 1) * 512 + 256:],
                                    dgb_stride=nl * 512)
            if wg:
                ops.conv2d_wgrad(ha, dzb, g3, _grad_buf(blk.conv2.weight))
            es = ops.epi_stats(B, h4, w4, 256, dev)
            dh = ops.conv2d_dgrad(dzb, P[f"r{i}1_d"], g3,
                                  ops.epilogue(aux=ha, aux_mode=AUX_RELU_MASK, stats=es, stats_z=za, mask_norm=sta))
            dza = ops.norm_bwd_from(es, dh, za, sta, dgamma=dgb[:, l * 512:], dbeta=dgb[:, l * 512 + 256:],
                                    dgb_stride=nl * 512)
            _trace(mod, f"res{i}", (dy, dzb, dh, dza))       # (dy into the block, dz conv2, dh, dz conv1)
            if wg:
                ops.conv2d_wgrad(x_in, dza, g3, _grad_buf(blk.conv1.weight))
            # + the skip connection's gradient; for i > 0 also the reductions of block i-1's second AdaIN
            es = ops.epi_stats(B, h4, w4, 256, dev) if i > 0 else None
            dy = ops.conv2d_dgrad(dza, P[f"r{i}0_d"], g3,
                                  ops.epilogue(aux=dy, aux_mode=AUX_ADD, stats=es,
                                               stats_z=S["res"][i - 1][4] if i > 0 else None))
        # ---- style Linears (model.py:28): dstyle, dW, db for all 2k layers
        adains = [a for i in range(k) for a in (dec[i].adain1, dec[i].adain2)]
        dstyle = _style_linears_backward(dgb, B, S["bs"], nl, 512, sd, S["sb"], P["lin_d"], adains, wg,
                                         S["style_grad"], S["style_shape"])
        # ---- encoder, reversed
        _trace(mod, "dx2", dy)
        dz2 = ops.norm_act_bwd(dy, S["z2"], S["st2"], ACT_RELU)
        _trace(mod, "dz2", dz2)
        if wg:
            ops.conv2d_wgrad(S["y1"], dz2, S["g2"], _grad_buf(enc[6].weight))
        g2, g1 = S["g2"], S["g1"]
        es = ops.epi_stats(B, g2.oh, g2.ow, g2.c, dev, phases=4) if ops.FUSE_N128_REDUCTIONS else None
        dy = ops.conv2d_dgrad(dz2, P["e2_d"], g2,
                              ops.epilogue(aux=S["y1"], aux_mode=AUX_RELU_MASK, stats=es, stats_z=S["z1"],
                                           mask_norm=S["st1"]))
        del dz2
        _trace(mod, "dy1", dy)
        dz1 = (ops.norm_bwd_from(es, dy, S["z1"], S["st1"]) if es is not None
               else ops.norm_act_bwd(dy, S["z1"], S["st1"], ACT_NONE))
        _trace(mod, "dz1", dz1)
        if wg:
            ops.conv2d_wgrad(S["y0"], dz1, g1, _grad_buf(enc[3].weight))
        es0 = ops.ring_stats(ops.RING_DGRAD_S2, g1, dev) if ops.MASK_FROM_Z else None
        if es0 is not None:
            # strip-ring kernel: ReLU mask from z0 and the reductions of the norm backward in the lean epilogue
            # (per work item, in registers) -- no separate reduction pass over dy and z0
            dy = ops.conv2d_dgrad(dz1, P["e1_d"], g1,
                                  ops.epilogue(aux=S["y0"], aux_mode=AUX_RELU_MASK, stats=es0, stats_z=S["z0"],
                                               mask_norm=S["st0"]))
            del dz1
            _trace(mod, "dy0_masked", dy)
            dz0 = ops.norm_bwd_from(es0, dy, S["z0"], S["st0"])
        else:
            dy = ops.conv2d_dgrad(dz1, P["e1_d"], g1)      # K = 4*128: too short to hide per-tile reductions
            del dz1
            _trace(mod, "dy0", dy)
            dz0 = ops.norm_act_bwd(dy, S["z0"], S["st0"], ACT_RELU)
        _trace(mod, "dz0", dz0)
        if wg:
            # the padded bf16 image copy is re-made (or step-cached), not saved
            ops.conv_rowpatch_wgrad(ops.img_pad8_cached(S["img"], 3, True), dz0, S["g0"], _grad_buf(enc[0].weight))
            # conv biases that feed an InstanceNorm have an exactly-zero true gradient (the mean
            # subtraction removes them); the reference produces float noise ~1e-9 there.
            for p in mod._dead_biases():
                _grad_buf(p)
        dimg = None
        if S["img_grad"]:
            # 7x7 conv of dz0 (64 -> 3, flipped filter) over the reflect-padded domain, then the fold
            gd = ops.conv_geom(B, H, W, 64, 3, 7, 7, 1, 6, 6, H + 6, W + 6)
            dimg = ops.reflect_fold_nchw(ops.conv_narrow_fwd(dz0, P["e0_rf"], gd), 3)
        ctx.saved = None
        return (None, dimg, dstyle) + (None,) * (len(ctx.needs_input_grad) - 3)


def _gen_dead_biases(self):
    enc, dec, k = self.content_encoder, self.decoder, self.n_res
    ps = [enc[0].bias, enc[3].bias, enc[6].bias, dec[k].bias, dec[k + 3].bias]
    for i in range(k):
        ps += [dec[i].conv1.bias, dec[i].conv2.bias]
    return ps


StyleCycleGANGenerator._dead_biases = _gen_dead_biases


# ######################################################################
# Shared trunk helpers for the 4x4 stride-2 stacks of the style encoder / discriminator
# ######################################################################
_TRUNK = ((3, 64), (64, 128), (128, 256), (256, 512))


def _pack_trunk(t, convs):
    t["c0"] = ops.wpack(WPACK_IM2COL, convs[0].weight, 64, 3, 4, 4, out=t.get("c0"))
    t["c0_d"] = ops.wpack(WPACK_IM2COL_DGRAD, convs[0].weight, 64, 3, 4, 4, out=t.get("c0_d"))
    for j in (1, 2, 3):
        ci, co = _TRUNK[j]
        t[f"c{j}"] = ops.wpack(WPACK_FWD, convs[j].weight, co, ci, 4, 4, out=t.get(f"c{j}"))
        t[f"c{j}_d"] = ops.wpack(WPACK_DGRAD_S2, convs[j].weight, co, ci, 4, 4, out=t.get(f"c{j}_d"))


# ######################################################################
# Style encoder
# ######################################################################
class MultiDomainStyleEncoder(_Net):
    """Shared conv trunk + one 1x1-conv head per domain (reference model.py:61-118)."""

    def __init__(self, style_dim=256, num_domains=2):
        super().__init__()
        self.num_domains = num_domains
        self.style_dim = style_dim
        self.shared_layers = nn.Sequential(
            nn.Conv2d(3, 64, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(64, 128, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(128, 256, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(256, 512, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.AdaptiveAvgPool2d(1)
        )
        self.domain_branches = nn.ModuleList()
        for _ in range(num_domains):
            self.domain_branches.append(nn.Sequential(nn.Conv2d(512, style_dim, kernel_size=1), nn.Flatten()))

    def _convs(self):
        sl = self.shared_layers
        return [sl[0], sl[2], sl[4], sl[6]]

    def _pack(self, t):
        _pack_trunk(t, self._convs())
        nd, sd = self.num_domains, self.style_dim
        dev = self.shared_layers[0].weight.device
        if "h" not in t:
            t["h"] = torch.zeros(ops.pad_rows(nd * sd) * 512, dtype=BF16, device=dev)
            t["h_d"] = torch.zeros(512 * nd * sd, dtype=BF16, device=dev)
            t["h_b"] = torch.zeros(nd * sd, dtype=F32, device=dev)
        for kx, br in enumerate(self.domain_branches):
            ops.wpack(WPACK_FWD, br[0].weight, sd, 512, 1, 1, out=t["h"], oc=nd * sd, o_off=kx * sd)
            ops.wpack(WPACK_DGRAD_S1, br[0].weight, sd, 512, 1, 1, out=t["h_d"], oc=nd * sd, o_off=kx * sd)
            ops.copy_f32(t["h_b"][kx * sd:(kx + 1) * sd], br[0].bias)

    def forward(self, img, domain_idx=None):
        return _StyleEncoderFn.apply(self, img, domain_idx, *self.parameters())


class _StyleEncoderFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, mod, img, domain_idx, *params):
        _require_cuda(img, "MultiDomainStyleEncoder")
        ops.ensure_init(img.device)
        P = mod._packed.get()
        nd, sd = mod.num_domains, mod.style_dim
        convs = mod._convs()
        img = img.contiguous().float()
        B, _, H, W = img.shape
        if H % 16 or W % 16:
            raise RuntimeError("style encoder input height/width must be multiples of 16")
        idx = ops.domain_index(domain_idx, nd, B, img.device)
        pg = ops.patch_geom(B, 3, H, W, 4, 4, 2, 1, 1, H // 2, W // 2, False)
        a = ops.patch_gather_cached(img, pg)
        m0 = B * (H // 2) * (W // 2)
        y = ops.conv2d_fwd(a.view(1, 1, m0, pg.kpad), P["c0"], ops.gemm_geom(m0, pg.kpad, 64),
                           ops.epilogue(bias=convs[0].bias.detach(), act=ACT_RELU)).view(B, H // 2, W // 2, 64)
        del a
        ys, gs = [y], []
        h, w = H // 2, W // 2
        for j in (1, 2, 3):
            ci, co = _TRUNK[j]
            g = ops.conv_geom(B, h, w, ci, co, 4, 4, 2, 1, 1, h // 2, w // 2)
            y = ops.conv2d_fwd(y, P[f"c{j}"], g, ops.epilogue(bias=convs[j].bias.detach(), act=ACT_RELU))
            ys.append(y)
            gs.append(g)
            h, w = h // 2, w // 2
        pooled = ops.avgpool_fwd(y)
        allh = ops.conv2d_fwd(pooled.view(1, 1, B, 512), P["h"], ops.gemm_geom(B, 512, nd * sd),
                              ops.epilogue(bias=P["h_b"], out_layout=OUT_F32_NHWC)).view(B, nd * sd)
        out = ops.head_gather(allh, idx, B, 1, nd, sd, True).view(B, sd)
        if any(ctx.needs_input_grad):
            ctx.mod = mod
            ctx.saved = dict(img=img, pg=pg, ys=ys, gs=gs, pooled=pooled, idx=idx, hw=(h, w),
                             img_grad=ctx.needs_input_grad[1])
        return out

    @staticmethod
    @_delivers_param_grads
    @ops.dev_guard
    def backward(ctx, dout):
        mod, S = ctx.mod, ctx.saved
        P = mod._packed.get()
        nd, sd = mod.num_domains, mod.style_dim
        convs = mod._convs()
        B = dout.shape[0]
        dout = dout.contiguous().float()
        dall = ops.head_scatter(dout, S["idx"], B, 1, nd, sd, True).view(B, nd * sd)
        dallb = ops.to_bf16(dall)
        # heads: dW_k = dall[:, k]^T pooled, db_k = colsum(dall[:, k])
        ws, splits = ops.gemm_tn_partial(B, dallb, nd * sd, S["pooled"], 512)
        ops.multi_linear_grads(ws, splits, nd * sd * 512, dall, B, nd * sd, nd, sd, 512,
                               [_grad_buf(br[0].weight) for br in mod.domain_branches],
                               [_grad_buf(br[0].bias) for br in mod.domain_branches])
        dpool = ops.conv2d_fwd(dallb.view(1, 1, B, nd * sd), P["h_d"], ops.gemm_geom(B, nd * sd, 512)).view(B, 512)
        h, w = S["hw"]
        dy = ops.avgpool_bwd(dpool, h, w)
        dz = ops.act_bwd(dy, S["ys"][3], ACT_RELU)
        for j in (3, 2, 1):
            ci, co = _TRUNK[j]
            g = S["gs"][j - 1]
            ops.colsum(dz, co, _grad_buf(convs[j].bias))
            ops.conv2d_wgrad(S["ys"][j - 1], dz, g, _grad_buf(convs[j].weight))
            # dgrad fused with the ReLU mask of the previous layer's output
            dz = ops.conv2d_dgrad(dz, P[f"c{j}_d"], g, ops.epilogue(aux=S["ys"][j - 1], aux_mode=AUX_RELU_MASK))
        pg = S["pg"]
        m0 = dz.shape[0] * dz.shape[1] * dz.shape[2]
        ops.colsum(dz, 64, _grad_buf(convs[0].bias))
        a = ops.patch_gather_cached(S["img"], pg)
        ops.patch_wgrad(WPACK_IM2COL, 64, 3, 4, 4, m0, dz, 64, a, pg.kpad, _grad_buf(convs[0].weight))
        del a
        dimg = None
        if S["img_grad"]:      # (train_step only feeds leaf images, trainer.py:94-95; model.py:89-118 is differentiable)
            da = ops.conv2d_fwd(dz.view(1, 1, m0, 64), P["c0_d"], ops.gemm_geom(m0, 64, pg.kpad))
            dimg = ops.patch_scatter(da.view(m0, pg.kpad), pg)
        ctx.saved = None
        return (None, dimg, None) + (None,) * (len(ctx.needs_input_grad) - 3)


# ######################################################################
# Discriminator
# ######################################################################
class MultiDomainDiscriminator(_Net):
    """PatchGAN trunk + one 4x4 head per domain (reference model.py:154-214)."""

    def __init__(self, in_channels=3, num_domains=2):
        super().__init__()
        if in_channels != 3:
            raise RuntimeError("msig_b200 discriminator supports 3-channel images only")
        self.num_domains = num_domains

        def discriminator_block(in_feat, out_feat, normalize=True):
            layers = [nn.Conv2d(in_feat, out_feat, 4, 2, 1)]
            if normalize:
                layers.append(nn.InstanceNorm2d(out_feat))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        self.shared_layers = nn.Sequential(
            *discriminator_block(in_channels, 64, normalize=False),
            *discriminator_block(64, 128),
            *discriminator_block(128, 256),
            *discriminator_block(256, 512),
        )
        self.domain_branches = nn.ModuleList()
        for _ in range(num_domains):
            self.domain_branches.append(nn.Sequential(nn.ZeroPad2d((1, 0, 1, 0)), nn.Conv2d(512, 1, 4, padding=1)))

    def _convs(self):
        sl = self.shared_layers
        return [sl[0], sl[2], sl[5], sl[8]]

    def _pack(self, t):
        _pack_trunk(t, self._convs())
        nd = self.num_domains
        dev = self.shared_layers[0].weight.device
        kflip = (16 * nd + 63) // 64 * 64
        if "h" not in t:
            t["h"] = torch.zeros(ops.pad_rows(nd) * 16 * 512, dtype=BF16, device=dev)
            t["h_d"] = torch.zeros(512 * kflip, dtype=BF16, device=dev)
            t["h_b"] = torch.zeros(ops.pad_rows(nd), dtype=F32, device=dev)
        for kx, br in enumerate(self.domain_branches):
            ops.wpack(WPACK_FWD, br[1].weight, 1, 512, 4, 4, out=t["h"], oc=nd, o_off=kx)
            ops.wpack(WPACK_IM2COL_FLIP, br[1].weight, 1, 512, 4, 4, out=t["h_d"], oc=nd, o_off=kx)
            ops.copy_f32(t["h_b"][kx:kx + 1], br[1].bias)

    def _dead_biases(self):
        c = self._convs()
        return [c[1].bias, c[2].bias, c[3].bias]

    def forward(self, img, domain_idx=None):
        return _DiscriminatorFn.apply(self, img, domain_idx, *self.parameters())


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, mod, img, domain_idx, *params):
        _require_cuda(img, "MultiDomainDiscriminator")
        ops.ensure_init(img.device)
        P = mod._packed.get()
        nd = mod.num_domains
        convs = mod._convs()
        img = img.contiguous().float()
        B, _, H, W = img.shape
        if H % 16 or W % 16:
            raise RuntimeError("discriminator input height/width must be multiples of 16")
        idx = ops.domain_index(domain_idx, nd, B, img.device)
        pg = ops.patch_geom(B, 3, H, W, 4, 4, 2, 1, 1, H // 2, W // 2, False)
        a = ops.patch_gather_cached(img, pg)
        m0 = B * (H // 2) * (W // 2)
        y0 = ops.conv2d_fwd(a.view(1, 1, m0, pg.kpad), P["c0"], ops.gemm_geom(m0, pg.kpad, 64),
                            ops.epilogue(bias=convs[0].bias.detach(), act=ACT_LRELU)).view(B, H // 2, W // 2, 64)
        del a
        y = y0
        layers = []
        h, w = H // 2, W // 2
        for j in (1, 2, 3):
            ci, co = _TRUNK[j]
            g = ops.conv_geom(B, h, w, ci, co, 4, 4, 2, 1, 1, h // 2, w // 2)
            z, st = _conv_stats(y, P[f"c{j}"], g)         # bias feeds an InstanceNorm: a no-op on the output
            yn = ops.norm_act_fwd(z, st, ACT_LRELU)
            layers.append((y, g, z, st))
            y = yn
            h, w = h // 2, w // 2
        # all heads in one implicit GEMM: ZeroPad2d((1,0,1,0)) + padding=1 == top/left pad 2 (model.py:182-183)
        gh = ops.conv_geom(B, h, w, 512, nd, 4, 4, 1, 2, 2, h, w)
        allh = ops.conv2d_fwd(y, P["h"], gh, ops.epilogue(bias=P["h_b"], out_layout=OUT_F32_NHWC))
        out = ops.head_gather(allh, idx, B, h * w, nd, 1, False).view(B, 1, h, w)
        if any(ctx.needs_input_grad):
            ctx.mod = mod
            ctx.saved = dict(img=img, pg=pg, y0=y0, layers=layers, y3=y, idx=idx, hw=(h, w),
                             img_grad=ctx.needs_input_grad[1])
            _trace(mod, "fwd", ctx.saved)
        return out

    @staticmethod
    @_delivers_param_grads
    @ops.dev_guard
    def backward(ctx, dout):
        mod, S = ctx.mod, ctx.saved
        P = mod._packed.get()
        nd = mod.num_domains
        convs = mod._convs()
        wg = not mod.skip_param_grads
        h, w = S["hw"]
        B = dout.shape[0]
        dout = dout.contiguous().float()
        # zeros for the unselected heads (they still receive zero gradients, model.py:208-212)
        dall = ops.head_scatter(dout, S["idx"], B, 1, nd, h * w, True)          # NCHW [B, nd, h, w]
        pgh = ops.patch_geom(B, nd, h, w, 4, 4, 1, 1, 1, h, w, False)
        ad = ops.patch_gather(dall, pgh)
        mh = B * h * w
        # the dgrads apply LeakyReLU' of the layer below and reduce (sum g, sum g*z) for its norm backward
        dev = dout.device
        es = None                                      # (the head GEMM's K is too short to fuse them)
        dy = ops.conv2d_fwd(ad.view(1, 1, mh, pgh.kpad), P["h_d"], ops.gemm_geom(mh, pgh.kpad, 512)).view(B, h, w, 512)
        _trace(mod, "dy3", dy)
        if wg:
            ws, splits = ops.gemm_tn_partial(mh, ad, pgh.kpad, S["y3"], 512)
            for kx, br in enumerate(mod.domain_branches):
                ops.wgrad_unpack(WPACK_IM2COL_FLIP, 1, 512, 4, 4, ws, splits, pgh.kpad * 512, _grad_buf(br[1].weight),
                                 oc=nd, o_off=kx)
            # (after the last unpack: nchw_chansum takes its scratch from the same stream workspace `ws` lives in)
            for kx, br in enumerate(mod.domain_branches):
                ops.nchw_chansum(dall, _grad_buf(br[1].bias), n=B, c=1, hw=h * w, img_stride=nd * h * w,
                                 offset=kx * h * w)
        for j in (3, 2, 1):
            ci, co = _TRUNK[j]
            y_in, g, z, st = S["layers"][j - 1]
            dz = ops.norm_bwd_from(es, dy, z, st) if es is not None else ops.norm_act_bwd(dy, z, st, ACT_LRELU)
            _trace(mod, f"dz{j}", dz)
            if wg:
                ops.conv2d_wgrad(y_in, dz, g, _grad_buf(convs[j].weight))
            if j > 1:
                es = ops.epi_stats(B, g.oh, g.ow, g.c, dev, phases=4)
                dy = ops.conv2d_dgrad(dz, P[f"c{j}_d"], g,
                                      ops.epilogue(aux=y_in, aux_mode=AUX_LRELU_MASK, stats=es,
                                                   stats_z=S["layers"][j - 2][2], mask_norm=S["layers"][j - 2][3]))
            else:   # dgrad fused with LeakyReLU' of the first layer's output
                dy = ops.conv2d_dgrad(dz, P[f"c{j}_d"], g, ops.epilogue(aux=S["y0"], aux_mode=AUX_LRELU_MASK))
            _trace(mod, f"dy{j - 1}", dy)
        dz0 = dy
        pg = S["pg"]
        m0 = dz0.shape[0] * dz0.shape[1] * dz0.shape[2]
        if wg:
            ops.colsum(dz0, 64, _grad_buf(convs[0].bias))
            a = ops.patch_gather_cached(S["img"], pg)
            ops.patch_wgrad(WPACK_IM2COL, 64, 3, 4, 4, m0, dz0, 64, a, pg.kpad, _grad_buf(convs[0].weight))
            del a
            for p in mod._dead_biases():
                _grad_buf(p)
        dimg = None
        if S["img_grad"]:
            da = ops.conv2d_fwd(dz0.view(1, 1, m0, 64), P["c0_d"], ops.gemm_geom(m0, 64, pg.kpad))
            dimg = ops.patch_scatter(da.view(m0, pg.kpad), pg)
        ctx.saved = None
        return (None, dimg, None) + (None,) * (len(ctx.needs_input_grad) - 3)
