"""Drop-in for the generator half of the reference's `models/stylegan2.py` (lines 1-576) on sm_100a kernels.

Same class names, constructor signatures, parameter/buffer names (=> the reference's `ckpt["g_ema"]` state_dict
loads unchanged) and the same `Generator.forward(styles, ..., noise, truncation,
transform_dict_list, input_is_latent, ...)` call signature, so `render.render` / `generate_audiovisual.generate`
can use it as-is.  What runs underneath is different: `Generator.forward` does not call the sub-modules one by
one; it drives libmaua_b200.so through the C ABI (include/maua_b200.h):

  style prologue (2 launches: truncation + 26 affines, 17 demod vectors)  ->  per layer:
     impl="tc"   tcgen05 implicit-GEMM conv on NHWC split-bf16 activations with the noise/bias/lrelu/next-style
                 epilogue fused (up layers: 4-phase transposed conv + NHWC blur/activation kernel)
     impl="simt" fp32 SIMT conv -> upfirdn2d blur -> noise+bias+lrelu (reference op order, exact fp32)
  ToRGB (+ in-place polyphase skip upsample) -> optional uint8 NHWC pack.

The sub-modules keep working `forward`s of their own (SIMT path) for users that call them directly.
Inference only; CUDA only (no CPU fallback).
"""
import ctypes as C
import math
import os
import random

import torch
from torch import nn

from . import _lib as L
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

SQRT2 = 2 ** 0.5


def _impl_default():
    return os.environ.get("MAUA_CONV_IMPL", "tc")


# ----------------------------------------------------------------------------------------------------------------
# small modules (parameter containers with reference-compatible forwards)
# ----------------------------------------------------------------------------------------------------------------

class PixelNorm(nn.Module):
    """models/stylegan2.py:15-20 — fused into the first mapping EqualLinear by Generator.get_latent."""

    def forward(self, inputs):
        return inputs * torch.rsqrt(torch.mean(inputs ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """models/stylegan2.py:23-31"""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """models/stylegan2.py:34-52"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, inputs):
        return upfirdn2d(inputs, self.kernel, up=self.factor, down=1, pad=self.pad)


class Blur(nn.Module):
    """models/stylegan2.py:76-92"""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, inputs):
        return upfirdn2d(inputs, self.kernel, pad=self.pad)


def _linear(x, weight, bias, w_scale, bias_scale, act, pixel_norm=False):
    x = x.contiguous()
    y = torch.empty((x.shape[0], weight.shape[0]), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.call("maua_linear_f32", x.data_ptr(), weight.data_ptr(), L.ptr(bias), y.data_ptr(), x.shape[0],
               weight.shape[1], weight.shape[0], float(w_scale), float(bias_scale), int(act), int(pixel_norm),
               L.stream_ptr(x.device))
    return y


class EqualLinear(nn.Module):
    """models/stylegan2.py:123-149"""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, inputs, pixel_norm=False, bias0=False):
        lead = inputs.shape[:-1]
        out = _linear(inputs.reshape(-1, inputs.shape[-1]), self.weight, self.bias, self.scale, self.lr_mul,
                      (2 if bias0 else 1) if self.activation else 0, pixel_norm)
        return out.view(*lead, -1)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ModulatedConv2d(nn.Module):
    """models/stylegan2.py:164-254 (generator cases: same-resolution and upsample)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if downsample:
            raise NotImplementedError("downsample=True is discriminator-only (out of scope, SURVEY.md §2 #8)")
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def forward(self, inputs, style):
        """Stand-alone forward (SIMT fp32 path)."""
        from .plan import style_single

        wsq = _weight_sq(self.weight, self.scale) if self.demodulate else None
        s, d = style_single(self.modulation, style, wsq, self.out_channel)
        out = _modconv_simt(inputs, self.weight, s, d, self.scale, self.kernel_size, self.upsample)
        if self.upsample:
            out = self.blur(out)
        return out


class NoiseInjection(nn.Module):
    """models/stylegan2.py:257-266"""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return _noise_bias_act(image, noise, self.weight, None, slope=1.0, scale=1.0)


class ConstantInput(nn.Module):
    """models/stylegan2.py:269-278"""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, inputs):
        return self.input.repeat(inputs.shape[0], 1, 1, 1)


class LatentInput(nn.Module):
    """models/stylegan2.py:281-294"""

    def __init__(self, latent_dim, channel, size=4):
        super().__init__()
        self.channel = channel
        self.size = size
        self.linear = EqualLinear(latent_dim, channel * size * size, activation="fused_lrelu")
        self.activate = FusedLeakyReLU(channel * size * size)
        self.input = nn.Parameter(torch.randn(1))

    def forward(self, inputs):
        batch = inputs.shape[0]
        out = self.linear(inputs[:, 0])
        out = self.activate(out)
        return out.reshape((batch, self.channel, self.size, self.size))


class ManipulationLayer(nn.Module):
    """models/stylegan2.py:297-307"""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer

    def forward(self, input, tranforms_dict_list):
        out = input
        for transform_dict in tranforms_dict_list:
            if transform_dict["layer"] == self.layer:
                out = transform_dict["transform"].to(out.device)(out)
        return out


class StyledConv(nn.Module):
    """models/stylegan2.py:310-343"""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True, layerID=-1):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)
        self.manipulation = ManipulationLayer(layerID)

    def forward(self, inputs, style, noise=None, transform_dict_list=[]):
        out = self.conv(inputs, style)
        if noise is None:
            b, _, h, w = out.shape
            noise = out.new_empty(b, 1, h, w).normal_()
        out = _noise_bias_act(out, noise, self.noise.weight, self.activate.bias, 0.2, SQRT2)
        return self.manipulation(out, transform_dict_list)


class ToRGB(nn.Module):
    """models/stylegan2.py:346-365"""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, inputs, style, skip=None):
        from .plan import style_single

        s, _ = style_single(self.conv.modulation, style)
        return _torgb(inputs, self.conv.weight, s, self.bias, skip,
                      self.upsample.kernel if skip is not None else None, self.conv.scale)


# ----------------------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ----------------------------------------------------------------------------------------------------------------

def _weight_sq(weight, scale):
    cout, cin, k = weight.shape[1], weight.shape[2], weight.shape[3]
    wsq = torch.empty((cout, cin), device=weight.device, dtype=torch.float32)
    with torch.cuda.device(weight.device):
        L.call("maua_weight_sq_f32", weight.data_ptr(), wsq.data_ptr(), cout, cin, k, float(scale),
               L.stream_ptr(weight.device))
    return wsq


def _modconv_simt(x, weight, s, d, scale, ksize, up):
    x = x.contiguous()
    b, cin, h, w = x.shape
    cout = weight.shape[1]
    oh, ow = (2 * h + 1, 2 * w + 1) if up else (h, w)
    y = torch.empty((b, cout, oh, ow), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.call("maua_modconv_simt_f32", x.data_ptr(), weight.data_ptr(), L.ptr(s), L.ptr(d), y.data_ptr(), b, cin,
               cout, h, w, ksize, 1 if up else 0, float(scale), L.stream_ptr(x.device))
    return y


def _prep_noise(noise, device, batch, hw):
    """-> (contiguous fp32 tensor on `device`, batch stride in elements: 0 = broadcast [1,1,H,W])."""
    n = noise.to(device=device, dtype=torch.float32).contiguous()
    if n.numel() == hw:
        return n, 0
    if n.numel() == batch * hw:
        return n, hw
    raise L.MauaError(f"noise of shape {tuple(noise.shape)} does not match a batch of {batch} maps of {hw} pixels")


def _noise_bias_act(x, noise, noise_weight, bias, slope, scale):
    x = x.contiguous()
    b, c, h, w = x.shape
    y = torch.empty_like(x)
    n, bstride = (None, 0)
    if noise is not None:
        n, bstride = _prep_noise(noise, x.device, b, h * w)
    with torch.cuda.device(x.device):
        L.call("maua_noise_bias_act_f32", x.data_ptr(), L.ptr(n), L.ptr(noise_weight), L.ptr(bias), y.data_ptr(), b,
               c, h, w, bstride, float(slope), float(scale), L.stream_ptr(x.device))
    return y


def _torgb(x, weight, s, bias, skip, k4, scale):
    x = x.contiguous()
    b, cin, h, w = x.shape
    y = torch.empty((b, 3, h, w), device=x.device, dtype=torch.float32)
    if skip is not None:
        skip = skip.contiguous()
        if skip.shape[-2] * 2 != h or skip.shape[-1] * 2 != w:
            raise L.MauaError(f"ToRGB skip {tuple(skip.shape)} is not half of {tuple(x.shape)}")
    with torch.cuda.device(x.device):
        L.call("maua_torgb_f32", x.data_ptr(), weight.data_ptr(), L.ptr(s), L.ptr(bias), L.ptr(skip), L.ptr(k4),
               y.data_ptr(), b, cin, h, w, float(scale), L.stream_ptr(x.device))
    return y


def frames_to_u8(image):
    """render.py:40-43 on device: [B,3,H,W] fp32 -> [B,H,W,3] uint8 (truncating)."""
    image = image.contiguous()
    b, _, h, w = image.shape
    out = torch.empty((b, h, w, 3), device=image.device, dtype=torch.uint8)
    with torch.cuda.device(image.device):
        L.call("maua_rgb_to_u8_nhwc", image.data_ptr(), out.data_ptr(), b, h, w, L.stream_ptr(image.device))
    return out


# ----------------------------------------------------------------------------------------------------------------
# Generator
# ----------------------------------------------------------------------------------------------------------------

class _ConvSpec:
    __slots__ = ("mod", "latent_index", "noise_index", "layer_id", "up", "cin", "cout", "rgb", "rgb_latent_index")

    def __init__(self, mod, latent_index, noise_index, layer_id, rgb=None, rgb_latent_index=-1):
        self.mod = mod
        self.latent_index = latent_index
        self.noise_index = noise_index
        self.layer_id = layer_id
        self.up = mod.conv.upsample
        self.cin = mod.conv.in_channel
        self.cout = mod.conv.out_channel
        self.rgb = rgb
        self.rgb_latent_index = rgb_latent_index


class Generator(nn.Module):
    """models/stylegan2.py:368-576"""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 constant_input=False, checkpoint=None, output_size=None, min_rgb_size=4, base_res_factor=1,
                 impl=None, precision=None):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        self.impl = impl or _impl_default()
        # "bf16x3": hi*hi + hi*lo + lo*hi on every layer (~2^-16 rel per product; network error ~5e-5)
        # "mixed"  : as bf16x3 below 512^2; the >= 512^2 layers take ONE fp16 activation plane and the weights as an fp16
        #            (hi, lo) pair — one tensor-core pass instead of three (network error ~5e-4, inside the 1e-3 parity bar)
        # "bf16"   : single bf16 product (fast preview, ~1e-2)
        # Default "mixed": measured on the B200 against the CPU oracle / the reference's own 1024^2 fixture the activation
        # maps stay within 4.9e-4 and the image within 6.2e-4 of the tensor max (bar 1e-3; the reference's own default GPU
        # path, TF32 cuDNN convs, is at 8e-4 already at 256^2) for +12 % frames/s; "bf16x3" is the 5e-5 option.
        self.precision = precision or os.environ.get("MAUA_TC_PRECISION", "mixed")
        if self.precision not in ("bf16x3", "mixed", "bf16"):
            raise ValueError(f"precision must be 'bf16x3', 'mixed' or 'bf16', got {self.precision!r}")

        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)

        cm = channel_multiplier
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm,
                         512: 32 * cm, 1024: 16 * cm}
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2
        self.min_rgb_size = min_rgb_size

        if constant_input:
            self.input = ConstantInput(self.channels[4])
        else:
            self.input = LatentInput(style_dim, self.channels[4])
        self.const_manipulation = ManipulationLayer(0)

        layerID = 1
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel,
                                layerID=layerID)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()

        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            layerID += 1
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel, layerID=layerID))
            layerID += 1
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel,
                                         layerID=layerID))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel

        self.truncation_latent = None

        if checkpoint is not None:
            self.load_state_dict(torch.load(checkpoint)["g_ema"])

        if size != output_size or base_res_factor != 1:  # models/stylegan2.py:461-470
            for layer_idx in range(self.num_layers):
                res = (layer_idx + 5) // 2
                shape = [1, 1,
                         int(base_res_factor * 2 ** res * (2 if output_size == 1080 else 1)),
                         int(base_res_factor * 2 ** res * (2 if output_size == 1920 else 1))]
                setattr(self.noises, f"noise_{layer_idx}", torch.randn(*shape))

        # execution order (models/stylegan2.py:549-569; latent indexing SURVEY.md Appendix A)
        self._specs = [_ConvSpec(self.conv1, 0, 0, 1, rgb=self.to_rgb1, rgb_latent_index=1)]
        i = 1
        for j in range(self.log_size - 2):
            self._specs.append(_ConvSpec(self.convs[2 * j], i, 1 + 2 * j, 2 * j + 2))
            self._specs.append(_ConvSpec(self.convs[2 * j + 1], i + 1, 2 + 2 * j, 2 * j + 3, rgb=self.to_rgbs[j],
                                         rgb_latent_index=i + 2))
            i += 2
        self._plan = None

    # ---- reference helpers --------------------------------------------------------------------------------------
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.get_latent(latent_in).mean(0, keepdim=True)

    def get_latent(self, inputs, reference_3d_path=False):
        """Mapping network on [N, style_dim] (PixelNorm fused into the first EqualLinear launch).
        reference_3d_path: evaluate what the reference's CUDA path computes when each z is fed as a [1,1,512] tensor
        (`map_latents=True`, models/stylegan2.py:506-509): PixelNorm over the singleton axis (every element normalised
        by itself) and bias[0] broadcast by the fused bias-act op (SURVEY.md §8(b) "Op API", §8(f) row 3)."""
        x = inputs.reshape(-1, inputs.shape[-1]).float()
        first = True
        for layer in self.style:
            if isinstance(layer, PixelNorm):
                continue
            x = layer(x, pixel_norm=(2 if reference_3d_path else 1) if first else 0, bias0=reference_3d_path)
            first = False
        return x.view(*inputs.shape[:-1], -1)

    # ---- weight-derived plan ------------------------------------------------------------------------------------
    def _plan_key(self):
        key = [self.impl, self.precision]
        for sp in self._specs:
            for p in (sp.mod.conv.weight, sp.mod.conv.modulation.weight, sp.mod.conv.modulation.bias):
                key.append((p.data_ptr(), p._version))
            if sp.rgb is not None:
                for p in (sp.rgb.conv.weight, sp.rgb.conv.modulation.weight, sp.rgb.conv.modulation.bias):
                    key.append((p.data_ptr(), p._version))
        return tuple(key)

    def invalidate_plan(self):
        """Drop the cached weight-derived plan (packed tensor-core weights, Wsq, style job tables).  The cache key follows
        `(data_ptr, _version)` of every conv / modulation parameter, which in-place ops and `nn.Parameter` replacement
        (render's `rewrites`) bump — edits made through `param.data` (`w.data.mul_()`) do NOT: call this after them."""
        self._plan = None
        self._synth_handle = None

    def _get_plan(self):
        key = self._plan_key()
        if self._plan is None or self._plan["key"] != key:
            from .plan import build_plan

            self._plan = build_plan(self, key)
        return self._plan

    # ---- forward ------------------------------------------------------------------------------------------------
    def forward(self, styles, return_latents=False, return_activation_maps=False, inject_index=None, truncation=1.0,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True,
                transform_dict_list=[], map_latents=False, return_u8=False):
        if map_latents:
            # Reference: th.cat([self.style(s[None, None, :]) for s in styles]).repeat(1, n_latent, 1)
            # (models/stylegan2.py:506-509).  Default = bit-for-bit what the reference's CUDA path returns for that 3-D
            # input (see get_latent); MAUA_MAP_LATENTS=2d selects the ordinary 2-D mapping network instead.
            quirk = os.environ.get("MAUA_MAP_LATENTS", "reference") != "2d"
            latent = self.get_latent(styles, reference_3d_path=quirk)[:, None, :]
            return latent.repeat(1, self.n_latent, 1)

        if not input_is_latent:
            if torch.is_tensor(styles):
                styles = [styles]
            styles = [self.get_latent(s) for s in styles]
            if len(styles) < 2:
                inject_index = self.n_latent
                if styles[0].ndim < 3:
                    latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
                else:
                    latent = styles[0]
            else:
                if inject_index is None:
                    inject_index = random.randint(1, self.n_latent - 1)
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
                latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
                latent = torch.cat([latent, latent2], 1)
        else:
            latent = styles
            if latent.dim() == 2:
                latent = latent[:, None, :].repeat(1, self.n_latent, 1)

        device = self.input.input.device
        if not latent.is_cuda:
            latent = latent.to(device)
        latent = latent.float().contiguous()

        noise = list(noise) if noise is not None else [None] * self.num_layers
        for ns, noise_scale in enumerate(noise):
            if not randomize_noise and noise_scale is None:
                noise[ns] = getattr(self.noises, f"noise_{ns}")

        if self.truncation_latent is None:
            self.truncation_latent = truncation_latent if truncation_latent is not None else self.mean_latent(2 ** 14)

        from .synthesis import synthesize

        image, latent_t, acts = synthesize(self, latent, noise, truncation, transform_dict_list or [],
                                           want_acts=return_activation_maps, want_u8=return_u8)
        if return_activation_maps:
            return image, acts
        elif return_latents:
            return image, (latent_t.get() if hasattr(latent_t, "get") else latent_t)
        else:
            return image, None
