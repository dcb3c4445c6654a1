"""CPU oracle for the StyleGAN2 synthesis hot path (`models/stylegan2.py:492-576` of the reference).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  The product path (maua_stylegan2_b200/) never imports this module.

It is a functional restatement over a plain `state_dict` (no nn.Module), written from the index-level
formulas of SURVEY.md Appendix B, each function citing the reference lines it follows.  It is PINNED:
tests/golden/generator_*.npz were produced by importing the real reference from /root/reference
(tests/golden/make_golden.py) and tests/test_oracle_golden.py checks this file against them.

All arithmetic is torch CPU fp32 (or fp64 with `dtype=torch.float64` for an order-free reference value).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# configuration helpers
# ----------------------------------------------------------------------------------------------------------------

def channels_for(channel_multiplier=2):
    """`models/stylegan2.py:395-405`."""
    cm = channel_multiplier
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm, 512: 32 * cm, 1024: 16 * cm}


def layout(size):
    """(log_size, num_layers, n_latent) — `models/stylegan2.py:407-409`."""
    log_size = int(math.log(size, 2))
    return log_size, (log_size - 2) * 2 + 1, log_size * 2 - 2


def make_kernel(k):
    """`models/stylegan2.py:23-31`."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def synth_state_dict(size, style_dim=512, n_mlp=8, channel_multiplier=2, seed=0, perturb=0.1, lr_mlp=0.01,
                     noconst=False):
    """Deterministic random-init parameters with the reference's key layout and init
    distributions (`models/stylegan2.py:127,205,260,273,354`, `op/fused_act.py:78`), drawn from numpy's PCG64
    (stable across torch versions).  Zero-initialised parameters (noise.weight, activate.bias, to_rgb bias)
    are perturbed with N(0, perturb^2) so that the noise / bias paths are actually exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ch = channels_for(channel_multiplier)
    log_size, num_layers, _ = layout(size)

    def randn(*shape):
        return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))

    sd = {}
    for i in range(n_mlp):
        sd[f"style.{i + 1}.weight"] = randn(style_dim, style_dim) / lr_mlp
        sd[f"style.{i + 1}.bias"] = randn(style_dim) * perturb
    sd["input.input"] = randn(1, ch[4], 4, 4)

    def styled(prefix, cin, cout, upsample):
        sd[f"{prefix}.conv.weight"] = randn(1, cout, cin, 3, 3)
        sd[f"{prefix}.conv.modulation.weight"] = randn(cin, style_dim)
        sd[f"{prefix}.conv.modulation.bias"] = torch.ones(cin) + randn(cin) * perturb
        sd[f"{prefix}.noise.weight"] = randn(1) * perturb * 5
        sd[f"{prefix}.activate.bias"] = randn(cout) * perturb
        if upsample:
            sd[f"{prefix}.conv.blur.kernel"] = make_kernel([1, 3, 3, 1]) * 4

    def rgb(prefix, cin, upsample):
        sd[f"{prefix}.conv.weight"] = randn(1, 3, cin, 1, 1)
        sd[f"{prefix}.conv.modulation.weight"] = randn(cin, style_dim)
        sd[f"{prefix}.conv.modulation.bias"] = torch.ones(cin) + randn(cin) * perturb
        sd[f"{prefix}.bias"] = randn(1, 3, 1, 1) * perturb
        if upsample:
            sd[f"{prefix}.upsample.kernel"] = make_kernel([1, 3, 3, 1]) * 4

    styled("conv1", ch[4], ch[4], False)
    rgb("to_rgb1", ch[4], False)
    cin = ch[4]
    for j, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        styled(f"convs.{2 * j}", cin, cout, True)
        styled(f"convs.{2 * j + 1}", cout, cout, False)
        rgb(f"to_rgbs.{j}", cout, True)
        cin = cout
    for l in range(num_layers):
        res = (l + 5) // 2
        sd[f"noises.noise_{l}"] = randn(1, 1, 2 ** res, 2 ** res)
    if noconst:  # LatentInput instead of ConstantInput (`models/stylegan2.py:281-288`); drawn last: other keys unchanged
        n = ch[4] * 16
        sd["input.input"] = randn(1)
        sd["input.linear.weight"] = randn(n, style_dim)
        sd["input.linear.bias"] = randn(n) * perturb
        sd["input.activate.bias"] = randn(n) * perturb
    return sd


# ----------------------------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------------------------

def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """Dense formulation of `op/upfirdn2d.py:159-200` on [N,C,H,W]."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    z = x.new_zeros(n, c, h * up, w * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(pad[0], 0), max(pad[1], 0), max(pad[0], 0), max(pad[1], 0)])
    z = z[:, :, max(-pad[0], 0): z.shape[2] - max(-pad[1], 0), max(-pad[0], 0): z.shape[3] - max(-pad[1], 0)]
    wk = torch.flip(kernel.to(x.dtype), [0, 1]).view(1, 1, kh, kw)
    o = F.conv2d(z.reshape(n * c, 1, z.shape[2], z.shape[3]), wk)
    o = o[:, :, ::down, ::down]
    return o.reshape(n, c, o.shape[2], o.shape[3])


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """`op/fused_act.py:86-97`; bias broadcasts along dim 1."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.view(*shape), negative_slope) * scale


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """`models/stylegan2.py:123-146`."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias * lr_mul)


def mapping(z, sd, n_mlp=8, lr_mlp=0.01):
    """PixelNorm + n_mlp EqualLinear(fused_lrelu) — `models/stylegan2.py:15-20,386-391`; 2-D input path."""
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(n_mlp):
        x = equal_linear(x, sd[f"style.{i + 1}.weight"], sd[f"style.{i + 1}.bias"], lr_mul=lr_mlp, activation=True)
    return x


def mapping_3d_cuda(z, sd, n_mlp=8, lr_mlp=0.01):
    """What `Generator(zs, map_latents=True)` returns per z on the reference's CUDA path (`models/stylegan2.py:506-509`
    feeds `s[None, None, :]`, a [1,1,512] tensor): PixelNorm reduces over dim 1 — a singleton — so every element is
    normalised by itself (`:20`), and the CUDA fused_bias_act indexes `b[(i / step_b) % size_b]` with
    `step_b = prod(x.shape[2:]) = 512` (`op/fused_bias_act_kernel.cu:29,67-69`), i.e. bias[0] for every feature.
    (The CPU fallback broadcasts differently and returns another shape, SURVEY.md §8(c) caveat 3; this restates the CUDA
    kernel and is checked on the GPU box against the compiled reference op, tests/test_gpu_plugins.py.)"""
    x = z * torch.rsqrt(z ** 2 + 1e-8)
    for i in range(n_mlp):
        w, b = sd[f"style.{i + 1}.weight"], sd[f"style.{i + 1}.bias"]
        scale = (1 / math.sqrt(w.shape[1])) * lr_mlp
        x = F.leaky_relu(F.linear(x, w * scale) + b[0] * lr_mlp, 0.2) * 2 ** 0.5
    return x


def modulated_conv2d(x, style_w, sd, prefix, demodulate=True, upsample=False):
    """`models/stylegan2.py:217-254`: per-sample modulated (+demodulated) weights, grouped conv;
    up path = stride-2 transposed conv to (2H+1)x(2W+1) followed by the 4x4 blur with pad (1,1)."""
    weight = sd[f"{prefix}.conv.weight"].to(x.dtype)
    _, cout, cin, k, _ = weight.shape
    b, _, h, w = x.shape
    s = equal_linear(style_w, sd[f"{prefix}.conv.modulation.weight"].to(x.dtype),
                     sd[f"{prefix}.conv.modulation.bias"].to(x.dtype)).view(b, 1, cin, 1, 1)
    wgt = (1 / math.sqrt(cin * k * k)) * weight * s
    if demodulate:
        d = torch.rsqrt(wgt.pow(2).sum([2, 3, 4]) + 1e-8)
        wgt = wgt * d.view(b, cout, 1, 1, 1)
    if upsample:
        wt = wgt.transpose(1, 2).reshape(b * cin, cout, k, k)
        out = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt, padding=0, stride=2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
        out = upfirdn2d(out, sd[f"{prefix}.conv.blur.kernel"], pad=(1, 1))
    else:
        out = F.conv2d(x.reshape(1, b * cin, h, w), wgt.view(b * cout, cin, k, k), padding=k // 2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
    return out


def styled_conv(x, style_w, noise, sd, prefix, upsample):
    """`models/stylegan2.py:338-343`: conv -> noise injection (`:262-266`) -> FusedLeakyReLU."""
    out = modulated_conv2d(x, style_w, sd, prefix, demodulate=True, upsample=upsample)
    out = out + sd[f"{prefix}.noise.weight"].to(x.dtype) * noise.to(x.dtype)
    return fused_leaky_relu(out, sd[f"{prefix}.activate.bias"].to(x.dtype))


def to_rgb(x, style_w, skip, sd, prefix):
    """`models/stylegan2.py:356-365`: 1x1 modulated conv (no demod) + bias + up-2 FIR of the skip."""
    out = modulated_conv2d(x, style_w, sd, prefix, demodulate=False) + sd[f"{prefix}.bias"].to(x.dtype)
    if skip is not None:
        out = out + upfirdn2d(skip, sd[f"{prefix}.upsample.kernel"], up=2, pad=(2, 1))
    return out


def apply_bends(x, layer, bends):
    """`models/stylegan2.py:302-307`."""
    for bend in bends or []:
        if bend["layer"] == layer:
            x = bend["transform"](x)
    return x


def generator_forward(sd, size, latent, noise, truncation, truncation_latent, channel_multiplier=2,
                      bends=None, dtype=torch.float32):
    """`Generator.forward(..., input_is_latent=True)` — `models/stylegan2.py:526-576`.

    latent [B, n_latent, 512] (or [B,512]); noise: list of num_layers tensors [B or 1,1,h,w] (None -> buffer);
    truncation: float or [B]; truncation_latent [1,512].  Returns (image, activation_maps)."""
    log_size, num_layers, n_latent = layout(size)
    latent = latent.to(dtype)
    if latent.dim() == 2:
        latent = latent[:, None, :].repeat(1, n_latent, 1)
    b = latent.shape[0]
    noise = list(noise) if noise is not None else [None] * num_layers
    for l in range(num_layers):
        if noise[l] is None:
            noise[l] = sd[f"noises.noise_{l}"]
    if not torch.is_tensor(truncation):
        truncation = torch.full((1,), float(truncation))
    tl = truncation_latent.to(dtype)
    latent = tl[None, ...] + truncation.to(dtype)[:, None, None] * (latent - tl[None, ...])

    acts = []
    if "input.linear.weight" in sd:
        # LatentInput.forward (`models/stylegan2.py:290-294`, --noconst): EqualLinear(fused_lrelu) of the (truncated) first
        # latent row, then a second FusedLeakyReLU with its own bias, reshaped to [B, C, 4, 4]
        out = equal_linear(latent[:, 0], sd["input.linear.weight"].to(dtype), sd["input.linear.bias"].to(dtype),
                           activation=True)
        out = fused_leaky_relu(out, sd["input.activate.bias"].to(dtype)).reshape(b, -1, 4, 4)
    else:
        out = sd["input.input"].to(dtype).repeat(b, 1, 1, 1)
    out = apply_bends(out, 0, bends)
    out = styled_conv(out, latent[:, 0], noise[0], sd, "conv1", False)
    out = apply_bends(out, 1, bends)
    acts.append(out)
    image = to_rgb(out, latent[:, 1], None, sd, "to_rgb1")
    i = 1
    layer_id = 1
    for j in range(log_size - 2):
        layer_id += 1
        out = styled_conv(out, latent[:, i], noise[1 + 2 * j], sd, f"convs.{2 * j}", True)
        out = apply_bends(out, layer_id, bends)
        acts.append(out)
        layer_id += 1
        out = styled_conv(out, latent[:, i + 1], noise[2 + 2 * j], sd, f"convs.{2 * j + 1}", False)
        out = apply_bends(out, layer_id, bends)
        acts.append(out)
        image = to_rgb(out, latent[:, i + 2], image, sd, f"to_rgbs.{j}")
        i += 2
    return image, acts


def frames_to_u8(image):
    """`render.py:40-43`: clamp, (x+1)*127.5, NHWC, truncate toward zero (numpy astype(uint8))."""
    x = (image.clamp(-1, 1) + 1) * 127.5
    return x.permute(0, 2, 3, 1).contiguous().numpy().astype(np.uint8)
