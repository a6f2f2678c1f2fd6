"""Drop-in mirror of the reference's losses.py plus the adversarial / reconstruction criteria of
trainer.py:50-52, executed by the C-ABI CUDA kernels.

VGGStyleContentLoss(device).forward(generated, real_style, real_content) -> (content, style)
follows /root/reference/losses.py:100-115 with two result-preserving savings:
only the first five VGG19 convs are evaluated (everything after 'relu_5_1' is dead code in the
reference, losses.py:64-68), and the two reference images are run forward-only.
"""
import warnings

import torch
import torch.nn as nn

from . import ops
from .lib import ACT_RELU, OUT_F32_NCHW, WPACK_DGRAD_S1, WPACK_FWD, WPACK_ROWFOLD_DGRAD, WPACK_ROWPATCH

F32 = torch.float32
VGG_CONV_IDX = (0, 2, 5, 7, 10)          # torchvision vgg19().features indices of convs 1..5
VGG_CH = ((3, 64), (64, 64), (64, 128), (128, 128), (128, 256))
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


def _vgg19_feature_state(seed=1234):
    """Pretrained VGG19 weights when torchvision can provide them (reference losses.py:15);
    offline, a random-init VGG19 under a fixed seed (global RNG state preserved)."""
    import torchvision
    try:
        m = torchvision.models.vgg19(weights=torchvision.models.VGG19_Weights.DEFAULT)
        return m.features.state_dict(), True
    except Exception:
        st = torch.get_rng_state()
        torch.manual_seed(seed)
        m = torchvision.models.vgg19(weights=None)
        torch.set_rng_state(st)
        warnings.warn("VGG19 pretrained weights unavailable (offline): using seeded random-init weights")
        return m.features.state_dict(), False


class VGGStyleContentLoss(nn.Module):
    """Perceptual style (L1 of Gram matrices at five taps) and content (L1 at 'relu_4_1') losses."""

    def __init__(self, device, vgg_state=None):
        super().__init__()
        self.device = torch.device(device)
        self.content_layers_default = ['relu_4_1']
        self.style_layers_default = ['relu_1_1', 'relu_2_1', 'relu_3_1', 'relu_4_1', 'relu_5_1']
        if vgg_state is None:
            vgg_state, self.pretrained = _vgg19_feature_state()
        else:
            self.pretrained = None
        # frozen parameters, named like the reference's ModuleDict (conv_{i}_1.weight / .bias)
        self.vgg_layers = nn.ModuleDict()
        for i, (idx, (ci, co)) in enumerate(zip(VGG_CONV_IDX, VGG_CH), start=1):
            conv = nn.Conv2d(ci, co, 3, padding=1)
            conv.weight.data.copy_(vgg_state[f"{idx}.weight"])
            conv.bias.data.copy_(vgg_state[f"{idx}.bias"])
            self.vgg_layers[f"conv_{i}_1"] = conv
        for p in self.parameters():
            p.requires_grad = False
        self.mean = torch.tensor(VGG_MEAN, device=self.device).view(1, 3, 1, 1)
        self.std = torch.tensor(VGG_STD, device=self.device).view(1, 3, 1, 1)
        self.to(self.device)
        self._pk = None

    def _packed(self):
        if self._pk is None:
            t = {}
            convs = [self.vgg_layers[f"conv_{i}_1"] for i in range(1, 6)]
            with torch.no_grad():
                # conv 1_1 on the 3-channel image: row-patch forward, row-fold image gradient (no patch matrix)
                t["w0"] = ops.wpack(WPACK_ROWPATCH, convs[0].weight, 64, 3, 3, 3)
                t["w0_rf"] = ops.wpack(WPACK_ROWFOLD_DGRAD, convs[0].weight, 64, 3, 3, 3)
                for j in range(1, 5):
                    ci, co = VGG_CH[j]
                    t[f"w{j}"] = ops.wpack(WPACK_FWD, convs[j].weight, co, ci, 3, 3)
                    t[f"w{j}_d"] = ops.wpack(WPACK_DGRAD_S1, convs[j].weight, co, ci, 3, 3)
                t["b"] = [c.bias.detach().contiguous() for c in convs]
                # ((x+1)/2 - mean)/std  ==  x*scale + shift   (losses.py:49-56)
                std = torch.tensor(VGG_STD, device=self.device)
                mean = torch.tensor(VGG_MEAN, device=self.device)
                t["scale"] = (0.5 / std).contiguous()
                t["shift"] = ((0.5 - mean) / std).contiguous()
            self._pk = t
        return self._pk

    def features(self, img, upto=5):
        """relu_1_1 .. relu_{upto}_1 (bf16 NHWC) of an fp32 NCHW image in [-1, 1]."""
        P = self._packed()
        img = img.contiguous().float()
        B, _, H, W = img.shape
        # input renormalisation (losses.py:49-56) fused into the padded bf16 copy; zero padding = padding=1
        g0 = ops.conv_geom(B, H, W, 3, 64, 3, 3, 1, 1, 1, H, W)
        f1 = ops.conv_rowpatch_fwd(ops.img_pad8_cached(img, 1, False, P["scale"], P["shift"]), P["w0"], g0,
                                   ops.epilogue(bias=P["b"][0], act=ACT_RELU))
        g = [None] * 5
        g[0] = g0
        g[1] = ops.conv_geom(B, H, W, 64, 64, 3, 3, 1, 1, 1, H, W)
        f2 = ops.conv2d_fwd(f1, P["w1"], g[1], ops.epilogue(bias=P["b"][1], act=ACT_RELU))
        feats = [f1, f2]
        if upto > 2:
            p2 = ops.maxpool2_fwd(f2)
            g[2] = ops.conv_geom(B, H // 2, W // 2, 64, 128, 3, 3, 1, 1, 1, H // 2, W // 2)
            f3 = ops.conv2d_fwd(p2, P["w2"], g[2], ops.epilogue(bias=P["b"][2], act=ACT_RELU))
            g[3] = ops.conv_geom(B, H // 2, W // 2, 128, 128, 3, 3, 1, 1, 1, H // 2, W // 2)
            f4 = ops.conv2d_fwd(f3, P["w3"], g[3], ops.epilogue(bias=P["b"][3], act=ACT_RELU))
            feats += [f3, f4]
            if upto > 4:
                p4 = ops.maxpool2_fwd(f4)
                g[4] = ops.conv_geom(B, H // 4, W // 4, 128, 256, 3, 3, 1, 1, 1, H // 4, W // 4)
                f5 = ops.conv2d_fwd(p4, P["w4"], g[4], ops.epilogue(bias=P["b"][4], act=ACT_RELU))
                feats.append(f5)
        return feats, g

    def features_of_real(self, img, upto):
        """Features of a data image; inside a train_step they are computed once (all five levels)
        and shared by the two perceptual-loss calls (trainer.py:104,109)."""
        cache = ops.step_cache()
        if cache is None:
            return self.features(img, upto)[0]
        key = ("vgg", img.data_ptr(), img._version, tuple(img.shape))
        hit = cache.get(key)
        if hit is None:
            hit = (img, self.features(img, 5)[0])
            cache[key] = hit
        return hit[1]

    def forward(self, generated, real_style, real_content):
        return _VGGLossFn.apply(self, generated, real_style, real_content)


class _VGGLossFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, mod, gen, real_style, real_content):
        if not gen.is_cuda:
            raise RuntimeError("VGGStyleContentLoss: expected CUDA tensors; msig_b200 has no CPU path")
        ops.ensure_init(gen.device)
        if gen.shape[2] % 4 or gen.shape[3] % 4:
            raise RuntimeError("VGG loss input height/width must be multiples of 4")
        fg, geoms = mod.features(gen, 5)
        fs = mod.features_of_real(real_style, 5)
        fc = mod.features_of_real(real_content, 4)         # only relu_4_1 is used (losses.py:110)
        style = torch.zeros((), dtype=F32, device=gen.device)
        ssyms = []
        for a, b in zip(fg, fs):                            # losses.py:80-89
            ga = ops.gram_fwd(a)
            gb_ = ops.gram_fwd(b)
            _, ss = ops.gram_l1(ga, gb_, loss=style)
            ssyms.append(ss)
        content = ops.l1_loss_bf16_fwd(fg[3], fc[3])        # losses.py:91-98
        if ctx.needs_input_grad[1]:
            ctx.mod = mod
            ctx.saved = (fg, fc[3], ssyms, geoms)
        return content, style

    @staticmethod
    @ops.dev_guard
    def backward(ctx, g_content, g_style):
        mod = ctx.mod
        fg, fc4, ssyms, g = ctx.saved
        P = mod._packed()
        g_content = g_content.contiguous().float()
        g_style = g_style.contiguous().float()

        def alpha(f):
            n, h, w, c = f.shape
            dim = n * c
            return 1.0 / (float(dim) * dim * n * c * h * w)

        f1, f2, f3, f4, f5 = fg
        # each Gram backward adds the gradient that arrived through the conv / pool path (aux) and applies the
        # ReLU backward of its tap in the same epilogue (relu_mask): no separate act_bwd pass
        dz5 = ops.gram_bwd(f5, ssyms[4], alpha(f5), g_style, relu_mask=True)
        dp4 = ops.conv2d_dgrad(dz5, P["w4_d"], g[4])
        d4 = ops.maxpool2_bwd(dp4, f4)
        d4 = ops.l1_loss_bf16_bwd(f4, fc4, g_content, aux=d4)
        dz4 = ops.gram_bwd(f4, ssyms[3], alpha(f4), g_style, aux=d4, relu_mask=True)
        d3 = ops.conv2d_dgrad(dz4, P["w3_d"], g[3])
        dz3 = ops.gram_bwd(f3, ssyms[2], alpha(f3), g_style, aux=d3, relu_mask=True)
        dp2 = ops.conv2d_dgrad(dz3, P["w2_d"], g[2])
        d2 = ops.maxpool2_bwd(dp2, f2)
        dz2 = ops.gram_bwd(f2, ssyms[1], alpha(f2), g_style, aux=d2, relu_mask=True)
        d1 = ops.conv2d_dgrad(dz2, P["w1_d"], g[1])
        dz1 = ops.gram_bwd(f1, ssyms[0], alpha(f1), g_style, aux=d1, relu_mask=True)
        # image gradient of conv 1_1: 3x3 conv of dz1 with the flipped filter (row-fold kernel, fp32 NCHW out),
        # times the renormalisation's per-channel scale (epilogue ch_scale)
        B, H, W, _ = f1.shape
        gd = ops.conv_geom(B, H, W, 64, 3, 3, 3, 1, 1, 1, H, W)          # pad' = R - 1 - pad = 1
        dgen = ops.conv_narrow_fwd(dz1, P["w0_rf"], gd, ops.epilogue(out_layout=OUT_F32_NCHW, ch_scale=P["scale"]))
        ctx.saved = None
        return None, dgen, None, None


# ---------------------------------------------------------------------------- simple criteria
class _L1Fn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, a, b):
        a = a.contiguous().float()
        b = b.contiguous().float()
        ops.ensure_init(a.device)
        ctx.save_for_backward(a, b)
        return ops.l1_loss_f32_fwd(a, b)

    @staticmethod
    @ops.dev_guard
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return ops.l1_loss_f32_bwd(a, b, g.contiguous().float()), None


class _MSEConstFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, a, target):
        a = a.contiguous().float()
        ops.ensure_init(a.device)
        ctx.save_for_backward(a)
        ctx.target = target
        return ops.mse_const_fwd(a, target)

    @staticmethod
    @ops.dev_guard
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        return ops.mse_const_bwd(a, ctx.target, g.contiguous().float()), None


class _MSEFn(torch.autograd.Function):
    @staticmethod
    @ops.dev_guard
    def forward(ctx, a, target):
        a = a.contiguous().float()
        ops.ensure_init(a.device)
        target = target.to(device=a.device, dtype=F32).expand_as(a).contiguous()
        ctx.save_for_backward(a, target)
        return ops.mse_loss_fwd(a, target)

    @staticmethod
    @ops.dev_guard
    def backward(ctx, g):
        a, target = ctx.saved_tensors
        return ops.mse_loss_bwd(a, target, g.contiguous().float()), None


class L1Loss(nn.Module):
    """nn.L1Loss() (mean) on images; the target does not receive a gradient (the reference's
    targets are data, trainer.py:99,116-117)."""

    def forward(self, input, target):
        return _L1Fn.apply(input, target)


class MSELoss(nn.Module):
    """nn.MSELoss() (mean). `target` is either a float (the LSGAN constants 1.0 / 0.0: no target tensor is
    read at all) or a tensor like the reference's `valid` / `fake` (trainer.py:85-86,103): both forms run
    on the device without a host synchronisation."""

    def forward(self, input, target):
        if torch.is_tensor(target):
            return _MSEFn.apply(input, target)
        return _MSEConstFn.apply(input, float(target))
