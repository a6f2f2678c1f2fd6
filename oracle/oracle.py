"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

CPU fp32 restatement (plain torch.nn.functional, no custom kernels) of the reference's hot path:
the style-injected Generator, the multi-domain Style Encoder and Discriminator (model.py), the
VGG style/content loss (losses.py) and one G+D optimisation step (trainer.py::train_step,
utils.py EMA / DynamicWeightScheduler). Every function cites the reference lines it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module, and only as the checker / the CPU baseline.

Pinning: the reference ships no tests or golden vectors, so this restatement
is pinned against the reference ITSELF: oracle/make_golden.py imports /root/reference in the build
container, runs it on seeded inputs and writes small fixtures to tests/golden/; tests/test_oracle.py
checks this module against them (and, when /root/reference is present, against the live reference).

Weights are plain dicts of tensors keyed exactly like the reference's state_dict().
"""
import math

import torch
import torch.nn.functional as F

LRELU = 0.2
IN_EPS = 1e-5


# ----------------------------------------------------------------------------- bf16-storage emulation
# The CUDA path keeps operands and activations in bf16 and accumulates / normalises / reduces in fp32
# (DESIGN.md "Precision contract"). `with emulate_bf16():` makes this restatement round to bf16 at exactly
# the points where that path STORES a tensor -- conv operands (images, packed weights), conv outputs z,
# post-norm activations y, the bf16 gradient tensors of the backward pass (dy / dz / the gathered-patch
# gradients), pooled features -- while everything in between stays fp32, like the kernels' registers.
# With it, CUDA-vs-oracle gradient parity can be held to bounds that are ~50x tighter than against the
# plain fp32 oracle, because the two sides then differ only where an fp32 summation-order difference
# flips a bf16 rounding. Outside the context manager every helper below is the identity and the
# restatement is the plain fp32 reference arithmetic (pinned by tests/golden/ref_small.pt).
_EMU = False


class emulate_bf16:
    def __enter__(self):
        global _EMU
        self._prev, _EMU = _EMU, True
        return self

    def __exit__(self, *exc):
        global _EMU
        _EMU = self._prev
        return False


def _r(x):
    return x.to(torch.bfloat16).to(x.dtype)


class _Store(torch.autograd.Function):
    """A tensor stored in bf16 whose gradient is stored in bf16 too (activations z / y, pooled features)."""
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _RoundFwd(torch.autograd.Function):
    """Rounded operand with an fp32 gradient (packed weights; images and style codes at a network input)."""
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """fp32 tensor whose GRADIENT is consumed in bf16 (GEMM outputs in front of an fp32 bias / tanh, the
    gathered patch matrices, the reflect-padded activation)."""
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _ActStore(torch.autograd.Function):
    """y = bf16(act(u)); backward g = bf16(dy * act'(u)) (mask fused into the producing dgrad epilogue,
    mask_first) or bf16(dy) * act'(u) (dy stored first, mask applied by the norm backward)."""
    @staticmethod
    def forward(ctx, u, slope, mask_first):
        ctx.save_for_backward(u)
        ctx.slope, ctx.mask_first = slope, mask_first
        return _r(torch.where(u > 0, u, u * slope))

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        d = torch.where(u > 0, torch.ones_like(u), torch.full_like(u, ctx.slope))
        return (_r(g * d) if ctx.mask_first else _r(g) * d), None, None


class _Tap(torch.autograd.Function):
    """A stored feature map with a main consumer and 1-2 side consumers whose gradients are added one
    after the other into the bf16 gradient tensor (losses.py VGG taps: conv/pool path first, then the
    content L1, then the Gram term; each addition re-rounds, like the aux operand of the kernels)."""
    @staticmethod
    def forward(ctx, f, n_side):
        ctx.n_side = n_side
        return tuple(f.clone() for _ in range(1 + n_side))

    @staticmethod
    def backward(ctx, *gs):
        acc = None
        for g in gs:
            if g is None:
                continue
            acc = _r(g) if acc is None else _r(acc + g)
        return acc, None


def st(x):
    return _Store.apply(x) if _EMU else x


def wq(x):
    return _RoundFwd.apply(x) if (_EMU and x is not None) else x


def gq(x):
    return _RoundBwd.apply(x) if _EMU else x


def act_store(u, slope=0.0, mask_first=True, round_grad=True):
    """act = ReLU (slope 0) / LeakyReLU(slope) followed by the bf16 store of the emulated path."""
    if not _EMU:
        return F.relu(u) if slope == 0.0 else F.leaky_relu(u, slope)
    if not round_grad:
        return _RoundFwd.apply(F.relu(u) if slope == 0.0 else F.leaky_relu(u, slope))
    return _ActStore.apply(u, slope, mask_first)


def _dead(b):
    """A conv bias in front of an InstanceNorm: removed by the mean subtraction, so the CUDA path neither
    adds it nor gives it a gradient; the emulation drops it to round the same values."""
    return None if _EMU else b


# ----------------------------------------------------------------------------- building blocks
def instance_norm(x, gamma=None, beta=None):
    """nn.InstanceNorm2d(affine=False): per-(n,c) biased variance, eps inside the sqrt
    (model.py:16,131-133,139-140,167); with gamma / beta the AdaIN modulation of model.py:28-36.
    Emulation: the kernels apply it as x*scale + shift with scale = gamma*rstd, shift = beta - mean*scale."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    if _EMU:
        rstd = 1.0 / torch.sqrt(var + IN_EPS)
        scale = rstd if gamma is None else gamma * rstd
        shift = -mu * scale if beta is None else beta - mu * scale
        return x * scale + shift
    xh = (x - mu) / torch.sqrt(var + IN_EPS)
    return xh if gamma is None else gamma * xh + beta


def style_affine(style, w, b):
    """[gamma | beta] = Linear(style) (model.py:18,28). Emulation: bf16 style code and weight, fp32 output
    and bias; the output's gradient is rounded before the two backward GEMMs (the bias sums the fp32 one)."""
    if style.dim() == 4:
        style = style.squeeze(-1).squeeze(-1)
    if _EMU:
        return gq(F.linear(wq(style), wq(w))) + b
    return F.linear(style, w, b)


def adain(x, style, w, b):
    """AdaIN.forward (model.py:20-36): gamma = first C outputs of the Linear, beta = last C."""
    gb = style_affine(style, w, b)
    c = x.shape[1]
    gamma = gb[:, :c].reshape(-1, c, 1, 1)
    beta = gb[:, c:].reshape(-1, c, 1, 1)
    return instance_norm(x, gamma, beta)


def reflect_conv7(x, w, b):
    """nn.Conv2d(k=7, s=1, p=3, padding_mode='reflect') (model.py:131,141)."""
    return F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w, b)


def residual_block_forward(sd, p, x, style):
    """ResidualBlockWithAdaIN.forward (model.py:51-55); `p` = key prefix of the block in `sd` ("" for a
    stand-alone block's own state_dict). Storage points of the emulation: both conv outputs, the
    activation between them, and the block output (after the residual add)."""
    r = x
    h = st(F.conv2d(x, wq(sd[p + "conv1.weight"]), _dead(sd[p + "conv1.bias"]), padding=1))
    h = act_store(adain(h, style, sd[p + "adain1.style_modulation.weight"], sd[p + "adain1.style_modulation.bias"]))
    h = st(F.conv2d(h, wq(sd[p + "conv2.weight"]), _dead(sd[p + "conv2.bias"]), padding=1))
    h = adain(h, style, sd[p + "adain2.style_modulation.weight"], sd[p + "adain2.style_modulation.bias"])
    return st(h + r)


def generator_forward(sd, img, style, n_res=8):
    """StyleCycleGANGenerator.forward (model.py:145-151; layers :130-143)."""
    x = wq(img)
    x = act_store(instance_norm(st(reflect_conv7(x, wq(sd["content_encoder.0.weight"]), _dead(sd["content_encoder.0.bias"])))))
    x = act_store(instance_norm(st(F.conv2d(x, wq(sd["content_encoder.3.weight"]), _dead(sd["content_encoder.3.bias"]), stride=2, padding=1))))
    x = act_store(instance_norm(st(F.conv2d(x, wq(sd["content_encoder.6.weight"]), _dead(sd["content_encoder.6.bias"]), stride=2, padding=1))))
    for i in range(n_res):
        x = residual_block_forward(sd, f"decoder.{i}.", x, style)
    k = n_res
    x = act_store(instance_norm(st(F.conv_transpose2d(x, wq(sd[f"decoder.{k}.weight"]), _dead(sd[f"decoder.{k}.bias"]), stride=2, padding=1))))
    # the IN + ReLU in front of the final conv writes the reflect-padded buffer; its gradient arrives as a
    # bf16 padded tensor and is folded + masked in fp32 (no second rounding)
    x = act_store(instance_norm(st(F.conv_transpose2d(x, wq(sd[f"decoder.{k + 3}.weight"]), _dead(sd[f"decoder.{k + 3}.bias"]), stride=2, padding=1))),
                  round_grad=False)
    if not _EMU:
        return torch.tanh(reflect_conv7(x, sd[f"decoder.{k + 6}.weight"], sd[f"decoder.{k + 6}.bias"]))
    xp = gq(F.pad(x, (3, 3, 3, 3), mode="reflect"))
    # fp32 output; tanh' is applied in fp32 and the product is rounded for the dgrad / wgrad (the bias sums fp32)
    return torch.tanh(gq(F.conv2d(xp, wq(sd[f"decoder.{k + 6}.weight"]))) + sd[f"decoder.{k + 6}.bias"].view(1, -1, 1, 1))


def _first_conv(img, w, b, k, stride, pad, slope, pre_scale=None, pre_shift=None):
    """The 3-channel first conv of SE / D / VGG + bias + (Leaky)ReLU. Emulation: SE / D run it as a GEMM over
    a gathered bf16 patch matrix, and the image gradient is the scatter-add (fp32) of the bf16 patch-matrix
    gradient -- restated with unfold so that the per-tap products are rounded the same way. The VGG conv
    (pre_scale / pre_shift = the input renormalisation x*scale + shift, applied before the rounding) runs
    on the row-patch / row-fold kernels: its image gradient is accumulated in fp32 and scaled in fp32."""
    if not _EMU:
        x = img if pre_scale is None else img * pre_scale + pre_shift
        y = F.conv2d(x, w, b, stride=stride, padding=pad)
        return F.relu(y) if slope == 0.0 else F.leaky_relu(y, slope)
    x = wq(img if pre_scale is None else img * pre_scale + pre_shift)
    if pre_scale is not None:
        return act_store(F.conv2d(x, wq(w), b, stride=stride, padding=pad), slope)
    n, c, h, wd = x.shape
    oh, ow = (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1
    a = gq(F.unfold(x, k, padding=pad, stride=stride))                       # [n, c*k*k, oh*ow]
    y = (wq(w).reshape(w.shape[0], -1) @ a).view(n, w.shape[0], oh, ow) + b.view(1, -1, 1, 1)
    return act_store(y, slope)


def _select_heads(all_out, domain_idx):
    """torch.stack(..., dim=1)[arange(B), domain_idx] (model.py:112-116, 208-212)."""
    b = all_out.shape[0]
    return all_out[torch.arange(b, device=all_out.device), domain_idx]


def style_encoder_forward(sd, img, domain_idx, num_domains):
    """MultiDomainStyleEncoder.forward (model.py:89-118): 4x (conv4x4 s2 + ReLU), global average
    pool, one 1x1-conv head per domain, per-sample head selection."""
    x = _first_conv(img, sd["shared_layers.0.weight"], sd["shared_layers.0.bias"], 4, 2, 1, 0.0)
    for i in (2, 4, 6):
        x = act_store(F.conv2d(x, wq(sd[f"shared_layers.{i}.weight"]), sd[f"shared_layers.{i}.bias"], stride=2, padding=1))
    x = st(x.mean(dim=(2, 3), keepdim=True))

    def head(k):   # fp32 output and bias; the output's gradient is rounded for the two backward GEMMs
        return (gq(F.conv2d(x, wq(sd[f"domain_branches.{k}.0.weight"]))) +
                sd[f"domain_branches.{k}.0.bias"].view(1, -1, 1, 1)).flatten(1) if _EMU else \
            F.conv2d(x, sd[f"domain_branches.{k}.0.weight"], sd[f"domain_branches.{k}.0.bias"]).flatten(1)
    if domain_idx is None:
        return head(0)
    return _select_heads(torch.stack([head(k) for k in range(num_domains)], dim=1), domain_idx)


def discriminator_forward(sd, img, domain_idx, num_domains):
    """MultiDomainDiscriminator.forward (model.py:186-214): conv+LeakyReLU, 3x (conv + IN +
    LeakyReLU), per-domain ZeroPad2d((1,0,1,0)) + conv4x4 p1 head, per-sample selection."""
    x = _first_conv(img, sd["shared_layers.0.weight"], sd["shared_layers.0.bias"], 4, 2, 1, LRELU)
    for i in (2, 5, 8):
        x = st(F.conv2d(x, wq(sd[f"shared_layers.{i}.weight"]), _dead(sd[f"shared_layers.{i}.bias"]), stride=2, padding=1))
        # the gradient of the last trunk activation is stored by the head GEMM before LeakyReLU' is applied
        x = act_store(instance_norm(x), LRELU, mask_first=(i != 8))
    xp = F.pad(x, (1, 0, 1, 0))

    def head(k):
        return gq(F.conv2d(xp, wq(sd[f"domain_branches.{k}.1.weight"]), padding=1)) + \
            sd[f"domain_branches.{k}.1.bias"].view(1, -1, 1, 1) if _EMU else \
            F.conv2d(xp, sd[f"domain_branches.{k}.1.weight"], sd[f"domain_branches.{k}.1.bias"], padding=1)
    if domain_idx is None:
        return head(0)
    return _select_heads(torch.stack([head(k) for k in range(num_domains)], dim=1), domain_idx)


# ----------------------------------------------------------------------------- VGG loss
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)
# torchvision vgg19().features indices of the first five convs; max-pools sit after conv 2 and 4.
VGG_CONV_IDX = (0, 2, 5, 7, 10)


def vgg_features(vgg_sd, img):
    """The five taps the reference actually uses (losses.py:23-35 names the ReLUs after the FIRST
    FIVE convs 'relu_1_1'..'relu_5_1'; everything after conv 5 is dead).
    Input renormalisation: losses.py:49-56."""
    mean = torch.tensor(VGG_MEAN, device=img.device).view(1, 3, 1, 1)
    std = torch.tensor(VGG_STD, device=img.device).view(1, 3, 1, 1)
    if _EMU:
        return _vgg_features_emulated(vgg_sd, img, 0.5 / std, (0.5 - mean) / std)
    x = ((img + 1) / 2 - mean) / std
    feats = []
    for j, idx in enumerate(VGG_CONV_IDX):
        x = F.relu(F.conv2d(x, vgg_sd[f"{idx}.weight"], vgg_sd[f"{idx}.bias"], padding=1))
        feats.append(x)
        if j in (1, 3):
            x = F.max_pool2d(x, 2)
    return feats


def _vgg_features_emulated(vgg_sd, img, scale, shift):
    """Storage points of the CUDA VGG pass: renormalised bf16 patch matrix, every tap (bf16), the pooled
    maps' gradients. Each tap is returned as (feature for the Gram term, feature for the content term):
    the gradient tensor of a tap receives the conv / pool path first, then the content L1 (tap 4 only),
    then the Gram term, re-rounded after each addition (aux operands of l1_loss_bf16_bwd / gram_bwd)."""
    feats = []
    x = None
    for j, idx in enumerate(VGG_CONV_IDX):
        w, b = vgg_sd[f"{idx}.weight"], vgg_sd[f"{idx}.bias"]
        if j == 0:
            f = _first_conv(img, w, b, 3, 1, 1, 0.0, scale, shift)
        else:
            f = act_store(F.conv2d(x, wq(w), b, padding=1))
        if j == 4:
            main, gram_side, content_side = None, f, None
        elif j == 3:
            main, content_side, gram_side = _Tap.apply(f, 2)
        else:
            main, gram_side = _Tap.apply(f, 1)
            content_side = None
        feats.append((gram_side, content_side))
        if j in (1, 3):
            x = gq(F.max_pool2d(main, 2))
        else:
            x = main
    return feats


def gram(x):
    """compute_gram_matrix (losses.py:70-78): batch folded into rows."""
    a, b, c, d = x.shape
    f = x.reshape(a * b, c * d)
    return (f @ f.t()) / (a * b * c * d)


def vgg_loss(vgg_sd, generated, real_style, real_content):
    """VGGStyleContentLoss.forward (losses.py:100-115) -> (content_loss, style_loss)."""
    fg = vgg_features(vgg_sd, generated)
    fs = vgg_features(vgg_sd, real_style)
    fc = vgg_features(vgg_sd, real_content)
    if _EMU:   # (Gram-term view, content-term view) per tap, see _vgg_features_emulated
        style = sum(F.l1_loss(gram(a[0]), gram(b[0])) for a, b in zip(fg, fs))
        return F.l1_loss(fg[3][1], fc[3][1]), style
    style = sum(F.l1_loss(gram(a), gram(b)) for a, b in zip(fg, fs))
    content = F.l1_loss(fg[3], fc[3])   # 'relu_4_1' = ReLU after the 4th conv
    return content, style


# ----------------------------------------------------------------------------- schedules
def loss_weights(init_weights, epoch, warmup_epochs=10, decay_epochs=100):
    """DynamicWeightScheduler.get_current_weights (utils.py:117-132): depends on epoch only."""
    warm = min(1.0, (epoch + 1) / warmup_epochs)
    decay = 1.0
    if epoch >= warmup_epochs:
        prog = min(1.0, (epoch - warmup_epochs) / decay_epochs)
        decay = 0.1 + 0.9 * 0.5 * (1 + math.cos(math.pi * prog))
    return {k: v * warm * decay for k, v in init_weights.items()}


DEFAULT_LOSS_WEIGHTS = {"gan": 1.0, "cycle": 10.0, "identity": 5.0, "content": 1.0, "style": 1.0}  # config.py:27-33


class AdamState:
    """torch.optim.Adam defaults used by the reference (trainer.py:58,61): betas (0.5, 0.999),
    eps 1e-8, no weight decay; preceded by clip_grad_norm_(params, 1.0) (trainer.py:127,152)."""

    def __init__(self, params, lr, betas=(0.5, 0.999), eps=1e-8):
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    def step(self, params, grads, max_norm=1.0):
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        for p, g, m, v in zip(params, grads, self.m, self.v):
            g = g * coef
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = v.sqrt() / math.sqrt(bc2) + self.eps
            p.sub_((self.lr / bc1) * m / denom)
        return total


class OracleTrainer:
    """One G+D optimisation step, trainer.py:74-155, on dict-of-tensor weights."""

    NETS = ("G_A2B", "G_B2A", "SE_A", "SE_B", "D_A", "D_B")

    def __init__(self, state, vgg_sd, num_domains, lr_g=2e-4, lr_d=1e-4, loss_w=None, ema_beta=0.995):
        self.sd = {k: {n: t.detach().clone().float() for n, t in state[k].items()} for k in self.NETS}
        self.ema = {k: {n: t.clone() for n, t in self.sd[k].items()} for k in ("G_A2B", "G_B2A", "SE_A", "SE_B")}
        self.vgg = {n: t.detach().clone().float() for n, t in vgg_sd.items()}
        self.nd = num_domains
        self.loss_w = dict(loss_w or DEFAULT_LOSS_WEIGHTS)
        self.ema_beta = ema_beta
        self.g_names = [(k, n) for k in ("G_A2B", "G_B2A", "SE_A", "SE_B") for n in self.sd[k]]
        self.d_names = [(k, n) for k in ("D_A", "D_B") for n in self.sd[k]]
        self.g_opt = AdamState([self.sd[k][n] for k, n in self.g_names], lr_g)
        self.d_opt = AdamState([self.sd[k][n] for k, n in self.d_names], lr_d)

    def _leaf(self, names):
        for k, n in names:
            self.sd[k][n] = self.sd[k][n].detach().requires_grad_(True)

    def train_step(self, batch, epoch):
        sd, nd = self.sd, self.nd
        real_A, real_B = batch["source"].float(), batch["target"].float()
        y_org, y_trg = batch["source_domain"], batch["target_domain"]
        self._leaf(self.g_names + self.d_names)
        # ---- generators (trainer.py:91-128)
        style_A = style_encoder_forward(sd["SE_A"], real_A, y_org, nd)
        style_B = style_encoder_forward(sd["SE_B"], real_B, y_trg, nd)
        l_id = F.l1_loss(generator_forward(sd["G_A2B"], real_B, style_B), real_B)
        fake_B = generator_forward(sd["G_A2B"], real_A, style_B)
        d_fb = discriminator_forward(sd["D_B"], fake_B, y_trg, nd)
        l_gan_ab = F.mse_loss(d_fb, torch.ones_like(d_fb))
        c_b, s_b = vgg_loss(self.vgg, fake_B, real_B, real_A)
        fake_A = generator_forward(sd["G_B2A"], real_B, style_A)
        d_fa = discriminator_forward(sd["D_A"], fake_A, y_org, nd)
        l_gan_ba = F.mse_loss(d_fa, torch.ones_like(d_fa))
        c_a, s_a = vgg_loss(self.vgg, fake_A, real_A, real_B)
        l_gan = (l_gan_ab + l_gan_ba) / 2
        l_style = (s_a + s_b) / 2
        l_content = (c_a + c_b) / 2
        l_cycle = (F.l1_loss(generator_forward(sd["G_B2A"], fake_B, style_A), real_A) +
                   F.l1_loss(generator_forward(sd["G_A2B"], fake_A, style_B), real_B)) / 2
        indiv = {"gan": l_gan, "cycle": l_cycle, "identity": l_id, "style": l_style, "content": l_content}
        w = loss_weights(self.loss_w, epoch)
        g_loss = sum(indiv[k] * w[k] for k in indiv)
        g_params = [sd[k][n] for k, n in self.g_names]
        g_grads = torch.autograd.grad(g_loss, g_params, allow_unused=True)
        g_grads = [torch.zeros_like(p) if g is None else g for p, g in zip(g_params, g_grads)]
        fake_A_d, fake_B_d = fake_A.detach(), fake_B.detach()
        with torch.no_grad():
            g_norm = self.g_opt.step(g_params, g_grads)
            for (k, n), p in zip(self.g_names, g_params):   # EMA, utils.py:80-91
                self.ema[k][n].mul_(self.ema_beta).add_(p, alpha=1 - self.ema_beta)
        # ---- discriminators (trainer.py:139-153)
        def dl(net, x, y, target):
            o = discriminator_forward(sd[net], x, y, nd)
            return F.mse_loss(o, torch.full_like(o, target))
        d_loss = (dl("D_A", real_A, y_org, 1.0) + dl("D_A", fake_A_d, y_org, 0.0) +
                  dl("D_B", real_B, y_trg, 1.0) + dl("D_B", fake_B_d, y_trg, 0.0)) / 2
        d_params = [sd[k][n] for k, n in self.d_names]
        d_grads = torch.autograd.grad(d_loss, d_params, allow_unused=True)
        d_grads = [torch.zeros_like(p) if g is None else g for p, g in zip(d_params, d_grads)]
        with torch.no_grad():
            d_norm = self.d_opt.step(d_params, d_grads)
        losses = {"D_loss": d_loss.detach(), "G_loss": g_loss.detach(), **{k: v.detach() for k, v in indiv.items()}}
        grads = {f"{k}.{n}": g.detach() for (k, n), g in zip(self.g_names + self.d_names, g_grads + d_grads)}
        return {"losses": losses, "grads": grads, "g_grad_norm": g_norm, "d_grad_norm": d_norm,
                "fake_A": fake_A_d, "fake_B": fake_B_d}


def synthetic_batch(b, s, num_domains, seed=42):
    """The seeded synthetic src/ref batch of SURVEY §8(d) / BASELINE.md §4."""
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    tgt = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    return {"source": src, "target": tgt, "source_domain": torch.zeros(b, dtype=torch.int64),
            "target_domain": 1 + (torch.arange(b) % max(num_domains - 1, 1))}


def seeded_vgg_state(seed=1234):
    """Random-init VGG19 feature weights under a fixed seed (the pretrained file cannot be
    downloaded offline); RNG state is restored afterwards."""
    import torchvision
    st = torch.get_rng_state()
    torch.manual_seed(seed)
    sd = {k: v.clone() for k, v in torchvision.models.vgg19(weights=None).features.state_dict().items()}
    torch.set_rng_state(st)
    return sd


# ----------------------------------------------------------------------------- seeded init
def _conv(sd, key, cin, cout, k, transposed=False):
    m = (torch.nn.ConvTranspose2d if transposed else torch.nn.Conv2d)(cin, cout, k)
    sd[key + ".weight"], sd[key + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()


def init_generator(n_res=8, style_dim=256):
    """Consumes the global RNG exactly like StyleCycleGANGenerator.__init__ (model.py:127-143):
    PyTorch default inits in module-construction order."""
    sd = {}
    _conv(sd, "content_encoder.0", 3, 64, 7)
    _conv(sd, "content_encoder.3", 64, 128, 4)
    _conv(sd, "content_encoder.6", 128, 256, 4)
    for i in range(n_res):
        for j in (1, 2):
            _conv(sd, f"decoder.{i}.conv{j}", 256, 256, 3)
            lin = torch.nn.Linear(style_dim, 512)
            sd[f"decoder.{i}.adain{j}.style_modulation.weight"] = lin.weight.detach().clone()
            sd[f"decoder.{i}.adain{j}.style_modulation.bias"] = lin.bias.detach().clone()
    _conv(sd, f"decoder.{n_res}", 256, 128, 4, transposed=True)
    _conv(sd, f"decoder.{n_res + 3}", 128, 64, 4, transposed=True)
    _conv(sd, f"decoder.{n_res + 6}", 64, 3, 7)
    # state_dict order of the reference: per block conv1, adain1, conv2, adain2
    return sd


def init_style_encoder(num_domains, style_dim=256):
    sd = {}
    for i, (a, b) in zip((0, 2, 4, 6), ((3, 64), (64, 128), (128, 256), (256, 512))):
        _conv(sd, f"shared_layers.{i}", a, b, 4)
    for k in range(num_domains):
        _conv(sd, f"domain_branches.{k}.0", 512, style_dim, 1)
    return sd


def init_discriminator(num_domains):
    sd = {}
    for i, (a, b) in zip((0, 2, 5, 8), ((3, 64), (64, 128), (128, 256), (256, 512))):
        _conv(sd, f"shared_layers.{i}", a, b, 4)
    for k in range(num_domains):
        _conv(sd, f"domain_branches.{k}.1", 512, 1, 4)
    return sd


def init_state(seed, num_domains):
    """All six networks in the construction order of trainer.py:31-40 after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    st = {"G_A2B": init_generator(), "G_B2A": init_generator()}
    st["SE_A"] = init_style_encoder(num_domains)
    st["SE_B"] = init_style_encoder(num_domains)
    st["D_A"] = init_discriminator(num_domains)
    st["D_B"] = init_discriminator(num_domains)
    return st
