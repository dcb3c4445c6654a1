"""CPU experiment: error of reduced-precision conv operands through the whole synthesis network.

Emulates the tensor-core algebra conv(x*s, W*c) * d with the activation operand (x*s) and/or the weight operand (W*c)
rounded to a narrower format, fp32 accumulation, everything else fp32 (oracle).  Metric = the parity metric of the
tests: max|a-b| / max|ref| per activation tensor / image.
    python tools/exp_precision.py [size] [cm] [batch]
"""
import math
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import stylegan2_oracle as O


def rnd(x, fmt):
    if fmt == "f32":
        return x
    if fmt == "f16":
        return x.half().float()
    if fmt == "bf16":
        return x.bfloat16().float()
    if fmt == "f16x2":   # hi + lo fp16 pair
        hi = x.half().float()
        return hi + (x - hi).half().float()
    if fmt == "bf16x2":
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    if fmt == "tf32t":   # truncation to 10 explicit mantissa bits
        return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    raise ValueError(fmt)


def make_modconv(afmt, wfmt, min_res=0):
    """Layers whose OUTPUT is at least min_res wide use (afmt, wfmt); the others the bf16x2 pair scheme."""
    afmt0, wfmt0 = afmt, wfmt

    def modulated_conv2d(x, style_w, sd, prefix, demodulate=True, upsample=False):
        out_res = x.shape[-1] * (2 if upsample else 1)
        afmt, wfmt = (afmt0, wfmt0) if out_res >= min_res else ("bf16x2", "bf16x2")
        weight = sd[f"{prefix}.conv.weight"]
        _, cout, cin, k, _ = weight.shape
        b, _, h, w = x.shape
        s = O.equal_linear(style_w, sd[f"{prefix}.conv.modulation.weight"], sd[f"{prefix}.conv.modulation.bias"])
        c = 1 / math.sqrt(cin * k * k)
        wc = weight[0] * c
        if k == 1 or afmt is None:   # ToRGB stays fp32 (SIMT kernel in the product)
            xs, wq = x * s.view(b, cin, 1, 1), wc
        else:
            xs, wq = rnd(x * s.view(b, cin, 1, 1), afmt), rnd(wc, wfmt)
        if demodulate:
            wsq = (wc ** 2).sum([2, 3])
            d = torch.rsqrt((s ** 2) @ wsq.t() + 1e-8)
        if upsample:
            out = F.conv_transpose2d(xs, wq.transpose(0, 1), stride=2)
            if demodulate:
                out = out * d.view(b, cout, 1, 1)
            out = O.upfirdn2d(out, sd[f"{prefix}.conv.blur.kernel"], pad=(1, 1))
        else:
            out = F.conv2d(xs, wq, padding=k // 2)
            if demodulate:
                out = out * d.view(b, cout, 1, 1)
        return out
    return modulated_conv2d


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    torch.manual_seed(0)
    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=0)
    _, nl, nlat = O.layout(size)
    z = torch.randn(batch, 512)
    w = O.mapping(z, sd)
    latent = w[:, None].repeat(1, nlat, 1)
    noise = [torch.randn(batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2)) for l in range(nl)]
    tl = torch.zeros(1, 512)
    orig = O.modulated_conv2d
    with torch.no_grad():
        ref_img, ref_acts = O.generator_forward(sd, size, latent, noise, 0.8, tl, channel_multiplier=cm)
        variants = [("f32", "f32", 0), ("bf16x2", "bf16x2", 0), ("f16", "f16x2", 0), ("f16", "f16", 0), ("f16x2", "f16", 0),
                    ("tf32t", "tf32t", 0), ("bf16", "bf16", 0)]
        if len(sys.argv) > 4:   # mixed: f16 activations only from this output resolution on
            variants = [("f16", "f16x2", int(r)) for r in sys.argv[4].split(",")]
        for afmt, wfmt, min_res in variants:
            O.modulated_conv2d = make_modconv(afmt, wfmt, min_res)
            img, acts = O.generator_forward(sd, size, latent, noise, 0.8, tl, channel_multiplier=cm)
            O.modulated_conv2d = orig
            errs = [float((a - r).abs().max() / r.abs().max()) for a, r in zip(acts, ref_acts)]
            ie = float((img - ref_img).abs().max() / ref_img.abs().max())
            print(f"act={afmt:7s} w={wfmt:7s} res>={min_res:4d}  image {ie:.2e}  acts max {max(errs):.2e}  last {errs[-1]:.2e}  "
                  f"per-layer {' '.join(f'{e:.1e}' for e in errs)}", flush=True)


if __name__ == "__main__":
    main()
