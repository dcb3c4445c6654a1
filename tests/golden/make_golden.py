"""Generate golden vectors by running the UNMODIFIED reference from /root/reference on CPU.

Run once in the build container (the reference cannot travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/{ops_golden,generator_g32,generator_g128,audio_glue}.npz (`--plugins`: only plugins.npz).  Inputs are derived from
numpy PCG64 seeds (oracle.stylegan2_oracle.synth_state_dict), so the fixtures only store the reference's OUTPUTS
(images + strided samples of the activation maps) and small inputs.

The reference's JIT build of its CUDA ops is stubbed out (`torch.utils.cpp_extension.load`): on CPU the
reference dispatches to its own Python fallbacks (`op/upfirdn2d.py:146-149`, `op/fused_act.py:87-94`), which
is exactly the oracle the survey names.  librosa/madmom/kornia/matplotlib are absent and
are stubbed so that `import audioreactive` succeeds; only the pure torch/scipy functions are exercised.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    import torch.utils.cpp_extension as cpp

    cpp.load = lambda *a, **k: types.SimpleNamespace()  # CPU path never touches the extension
    for name in ["librosa", "librosa.display", "madmom", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "kornia", "kornia.augmentation", "kornia.geometry", "kornia.geometry.transform", "ffmpeg"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import op  # noqa
    from models import stylegan2 as ref_sg2
    import audioreactive as ar
    return op, ref_sg2, ar


def strided(t, n=6):
    """Deterministic sub-sample of an activation map: every k-th channel, all pixels up to 32x32 then strided."""
    c = t.shape[1]
    cs = max(c // n, 1)
    s = max(t.shape[2] // 32, 1)
    return t[:, ::cs, ::s, ::s].contiguous().numpy()


def gen_case(ref_sg2, size, cm, batch, seed, psi_lo, store_noise=True, noise_none=(2,), image_stride=1):
    """image_stride > 1 (1024^2 fixture): the image is stored as every image_stride-th pixel plus a full-resolution
    128x128 centre crop (`image_crop`), so that the fixture stays small."""
    from oracle import stylegan2_oracle as O

    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed)
    g = ref_sg2.Generator(size, 512, 8, channel_multiplier=cm, constant_input=True, output_size=size)
    missing, unexpected = g.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".kernel" in m or "noises" in m for m in missing), missing
    g.eval()
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    log_size, num_layers, n_latent = O.layout(size)
    z = torch.from_numpy(rng.standard_normal((batch, 512)).astype(np.float32))
    with torch.no_grad():
        w = g.get_latent(z)  # 2-D path of the mapping network (CPU == CUDA)
        latent = w[:, None, :].repeat(1, n_latent, 1)
        latent = latent + 0.05 * torch.from_numpy(rng.standard_normal(latent.shape).astype(np.float32))  # W+
        noise = [torch.from_numpy(rng.standard_normal((batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
                 for l in range(num_layers)]
        for l in noise_none:  # exercise the buffer path (randomize_noise=False)
            noise[l] = None
        psi = torch.from_numpy(rng.uniform(psi_lo, 1.0, batch).astype(np.float32))
        tl = w.mean(0, keepdim=True) * 0.5
        g.truncation_latent = tl
        image, acts = g(latent, noise=list(noise), truncation=psi, input_is_latent=True, randomize_noise=False,
                        return_activation_maps=True)
    out = {"size": size, "cm": cm, "seed": seed, "z": z.numpy(), "w": w.numpy(), "latent": latent.numpy(),
           "psi": psi.numpy(), "truncation_latent": tl.numpy(), "image": image.numpy(),
           "noise_none": np.array(list(noise_none))}
    if image_stride > 1:
        c0 = size // 2 - 64
        out["image"] = image[:, :, ::image_stride, ::image_stride].contiguous().numpy()
        out["image_crop"] = image[:, :, c0:c0 + 128, c0:c0 + 128].contiguous().numpy()
        out["image_stride"] = np.array(image_stride)
        out["image_absmax"] = np.array(image.abs().max().item(), np.float32)
    out["batch"] = batch
    for l, n in enumerate(noise):
        if n is not None and store_noise:   # otherwise the test regenerates it: same PCG64 stream, same draw order
            out[f"noise_{l}"] = n.numpy()
    for l, a in enumerate(acts):
        out[f"act_{l}"] = strided(a)
        out[f"act_{l}_absmax"] = np.array(a.abs().max().item(), np.float32)
    return out


def gen_case_1024(ref_sg2):
    """BASELINE configs[1] architecture (1024^2 config-f, cm=2), ONE frame through the unmodified reference on CPU; like the
    default hooks, noise maps wider than 256 are None (the generator's registered buffers are used, examples/default.py:29-30)."""
    from oracle import stylegan2_oracle as O

    _, num_layers, _ = O.layout(1024)
    none = tuple(l for l in range(num_layers) if 2 ** ((l + 5) // 2) > 256)
    return gen_case(ref_sg2, 1024, 2, 1, seed=11, psi_lo=0.7, store_noise=False, noise_none=none, image_stride=4)


class Affine(torch.nn.Module):
    """x * a + b — a stand-in for a user bend (any nn.Module is allowed, README.md:120-146 of the reference)."""

    def __init__(self, a, b):
        super().__init__()
        self.a, self.b = a, b

    def forward(self, x):
        return x * self.a + self.b


class FlipW(torch.nn.Module):
    def forward(self, x):
        return torch.flip(x, dims=(3,))


def bend_list():
    """The bends of the `bends` fixture: widen the constant input 4x4 -> 4x8 (layer 0, as tauceti.py / kelp.py do),
    rescale after the first up-conv (layer 2), mirror after the 16^2 conv (layer 5)."""
    return [{"layer": 0, "transform": torch.nn.ReplicationPad2d((2, 2, 0, 0))},
            {"layer": 2, "transform": Affine(0.5, 0.1)},
            {"layer": 5, "transform": FlipW()}]


def gen_bend_case(ref_sg2, size=32, cm=2, batch=2, seed=9):
    """Reference Generator.forward with transform_dict_list (models/stylegan2.py:297-307,548-569): pins WHERE each layer id
    is applied and the non-square (H x 2H) data flow.  Only outputs are stored; inputs are redrawn from the seed."""
    from oracle import stylegan2_oracle as O

    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed)
    g = ref_sg2.Generator(size, 512, 8, channel_multiplier=cm, constant_input=True, output_size=size)
    g.load_state_dict(sd, strict=False)
    g.eval()
    latent, noise, tl = bend_case_inputs(size, batch, seed)
    g.truncation_latent = tl
    image, acts = g(latent, noise=list(noise), truncation=torch.ones(batch), input_is_latent=True, randomize_noise=False,
                    return_activation_maps=True, transform_dict_list=bend_list())
    out = {"size": size, "cm": cm, "seed": seed, "batch": batch, "image": image.numpy()}
    for l, a in enumerate(acts):
        out[f"act_{l}"] = strided(a)
        out[f"act_{l}_absmax"] = np.array(a.abs().max().item(), np.float32)
        out[f"act_{l}_shape"] = np.array(a.shape)
    return out


def gen_noconst_case(ref_sg2, size=32, cm=2, batch=2, seed=13):
    """Reference forward with LatentInput (`--noconst`, models/stylegan2.py:281-294); outputs only, inputs from the seed."""
    from oracle import stylegan2_oracle as O

    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed, noconst=True)
    g = ref_sg2.Generator(size, 512, 8, channel_multiplier=cm, constant_input=False, output_size=size)
    missing, unexpected = g.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    g.eval()
    latent, noise, tl, psi = noconst_case_inputs(size, batch, seed)
    g.truncation_latent = tl
    image, acts = g(latent, noise=list(noise), truncation=psi, input_is_latent=True, randomize_noise=False,
                    return_activation_maps=True)
    out = {"size": size, "cm": cm, "seed": seed, "batch": batch, "image": image.numpy()}
    for l, a in enumerate(acts):
        out[f"act_{l}"] = strided(a)
        out[f"act_{l}_absmax"] = np.array(a.abs().max().item(), np.float32)
    return out


def noconst_case_inputs(size, batch, seed):
    from oracle import stylegan2_oracle as O

    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(seed + 3000))
    latent = torch.from_numpy(rng.standard_normal((batch, n_latent, 512)).astype(np.float32)) * 0.7
    noise = [torch.from_numpy(rng.standard_normal((batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    tl = torch.from_numpy(rng.standard_normal((1, 512)).astype(np.float32)) * 0.1
    psi = torch.from_numpy(rng.uniform(0.5, 1.0, batch).astype(np.float32))
    return latent, noise, tl, psi


def bend_case_inputs(size, batch, seed):
    from oracle import stylegan2_oracle as O

    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    latent = torch.from_numpy(rng.standard_normal((batch, n_latent, 512)).astype(np.float32)) * 0.7
    noise = [torch.from_numpy(rng.standard_normal((batch, 1, 2 ** ((l + 5) // 2), 2 * 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    return latent, noise, torch.zeros(1, 512)


def ops_cases(op):
    rng = np.random.Generator(np.random.PCG64(7))
    out = {}
    k4 = (np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64.0).astype(np.float32)
    sym6 = np.array([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
                     0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
                     -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148],
                    dtype=np.float64)
    k12 = np.outer(sym6, sym6).astype(np.float32)
    k3 = rng.standard_normal((3, 3)).astype(np.float32)
    k2 = rng.standard_normal((2, 2)).astype(np.float32)
    k43 = rng.standard_normal((4, 3)).astype(np.float32)
    cases = [
        ("blur9", (2, 3, 9, 9), k4 * 4, 1, 1, (1, 1)),
        ("blur9x17", (2, 3, 9, 17), k4 * 4, 1, 1, (1, 1)),
        ("blur65", (1, 5, 65, 65), k4 * 4, 1, 1, (1, 1)),
        ("blur129x33", (1, 2, 129, 33), k4 * 4, 1, 1, (1, 1)),
        ("up8", (2, 3, 8, 8), k4 * 4, 2, 1, (2, 1)),
        ("up4x8", (2, 3, 4, 8), k4 * 4, 2, 1, (2, 1)),
        ("up37", (1, 3, 37, 21), k4 * 4, 2, 1, (2, 1)),
        ("down16", (2, 4, 16, 16), k4, 1, 2, (1, 1)),
        ("down17pad2", (1, 2, 17, 23), k4, 1, 2, (2, 2)),
        ("k3same", (1, 3, 12, 10), k3, 1, 1, (1, 1)),
        ("k2up2", (1, 3, 7, 9), k2, 2, 1, (1, 0)),
        ("k43asym", (1, 2, 11, 13), k43, 1, 1, (2, 1)),
        ("negpad", (1, 2, 16, 16), k4, 1, 1, (-1, -2)),
        ("negpad_up", (1, 2, 10, 12), k4 * 4, 2, 1, (-1, 3)),
        ("sym6_up2", (1, 2, 20, 20), k12 * 4, 2, 1, (6, 5)),
        ("sym6_down2", (1, 2, 40, 40), k12, 1, 2, (5, 5)),
        ("up3down2", (1, 2, 9, 9), k4, 3, 2, (2, 2)),
        ("tiny1", (1, 1, 1, 1), k4 * 4, 2, 1, (2, 1)),
    ]
    names = []
    for name, shape, k, up, down, pad in cases:
        x = rng.standard_normal(shape).astype(np.float32)
        y = op.upfirdn2d(torch.from_numpy(x), torch.from_numpy(k), up=up, down=down, pad=pad).numpy()
        out[f"ufd_{name}_x"] = x
        out[f"ufd_{name}_k"] = k
        out[f"ufd_{name}_cfg"] = np.array([up, down, pad[0], pad[1]])
        out[f"ufd_{name}_y"] = y
        names.append(name)
    out["ufd_names"] = np.array(names)
    # fused_leaky_relu CPU fallback: 2-D and 4-D (`op/fused_act.py:87-94`)
    for name, shape in (("fl2d", (5, 512)), ("fl4d", (2, 6, 5, 7)), ("fl4d_big", (1, 32, 16, 16))):
        x = rng.standard_normal(shape).astype(np.float32)
        b = rng.standard_normal(shape[1]).astype(np.float32)
        y = op.fused_leaky_relu(torch.from_numpy(x), torch.from_numpy(b)).numpy()
        out[f"{name}_x"], out[f"{name}_b"], out[f"{name}_y"] = x, b, y
    return out


def audio_glue_cases(ar):
    """Pure torch/scipy functions of audioreactive/ that run without librosa."""
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    ar.set_SMF(1)
    x1 = torch.from_numpy(rng.standard_normal(200).astype(np.float32))
    x3 = torch.from_numpy(rng.standard_normal((120, 6, 16)).astype(np.float32))
    x4 = torch.from_numpy(rng.standard_normal((64, 1, 4, 8)).astype(np.float32))
    out["gf_x1"], out["gf_x3"], out["gf_x4"] = x1.numpy(), x3.numpy(), x4.numpy()
    out["gf_y1_s5_c0"] = ar.gaussian_filter(x1.clone(), 5, causal=0).numpy()
    out["gf_y1_s3"] = ar.gaussian_filter(x1.clone(), 3).numpy()
    out["gf_y3_s4"] = ar.gaussian_filter(x3.clone(), 4).numpy()
    out["gf_y3_s2_c02"] = ar.gaussian_filter(x3.clone(), 2, causal=0.2).numpy()
    out["gf_y4_s5"] = ar.gaussian_filter(x4.clone(), 5).numpy()
    out["gf_y4_s128"] = ar.gaussian_filter(x4.clone(), 128).numpy()  # radius > n_frames branch (:350-355)
    env = torch.from_numpy(np.abs(rng.standard_normal(300)).astype(np.float32))
    env = ar.gaussian_filter(env, 2)
    out["pc_x"] = env.numpy()
    out["pc_y97"] = ar.percentile_clip(env.clone(), 97).numpy()
    out["pc_y50"] = ar.percentile_clip(env.clone(), 50).numpy()
    chroma = torch.from_numpy(rng.uniform(0, 1, (50, 12)).astype(np.float32))
    chroma = chroma / chroma.sum(1, keepdim=True)
    sel = torch.from_numpy(rng.standard_normal((12, 6, 64)).astype(np.float32))
    out["cw_chroma"], out["cw_sel"] = chroma.numpy(), sel.numpy()
    out["cw_y"] = ar.chroma_weight_latents(chroma, sel).numpy()
    n = torch.from_numpy(rng.uniform(-1, 3, 77).astype(np.float32))
    out["norm_x"] = n.numpy()
    out["norm_y"] = ar.normalize(n.clone()).numpy()
    # compress / expand / percentile (signal.py:257-316)
    e = torch.from_numpy(rng.uniform(0, 1, 200).astype(np.float32))
    out["dyn_x"] = e.numpy()
    out["dyn_compress"] = ar.compress(e.clone(), 0.6, 0.25).numpy()
    out["dyn_compress_inv"] = ar.compress(e.clone(), 0.3, 0.5, invert=True).numpy()
    out["dyn_expand"] = ar.expand(e.clone(), 0.8, 10).numpy()
    out["dyn_percentiles"] = np.array([ar.percentile(e, p) for p in (0, 10, 50, 97, 100)], np.float32)
    return out


def plugin_cases(ar, ref_sg2=None, ref_op=None):
    """audioreactive/latent.py functions the example hook files call: perlin_noise
    (its `.cuda()` calls are made a no-op for this CPU run; the arithmetic is untouched), spline_loops, slerp."""
    out = {}
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, shape, res, tile in (("a", (12, 8, 16), (3, 2, 4), (True, False, False)),
                                       ("b", (16, 32, 32), (1, 1, 1), (True, False, False)),
                                       ("c", (24, 16, 8), (8, 4, 4), (True, True, False)),
                                       ("d", (10, 6, 9), (5, 3, 3), (False, False, True))):
            np.random.seed(100 + ord(name))
            y = ar.perlin_noise(shape=shape, res=res, tileable=tile)
            out[f"perlin_{name}_cfg"] = np.array(list(shape) + list(res) + [int(t) for t in tile] + [100 + ord(name)])
            out[f"perlin_{name}_y"] = y.numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    rng = np.random.Generator(np.random.PCG64(21))
    sel = rng.standard_normal((4, 6, 32)).astype(np.float32)
    out["spline_sel"] = sel
    out["spline_y_100_2"] = ar.spline_loops(sel, 100, 2).numpy()
    out["spline_y_97_3_noloop"] = ar.spline_loops(sel, 97, 3, loop=False).numpy()
    out["spline_y_64_half"] = ar.spline_loops(sel, 64, 0.5).numpy()      # kelp.py passes fractional n_loops
    a, b = rng.standard_normal(32), rng.standard_normal(32)
    out["slerp_a"], out["slerp_b"] = a, b
    out["slerp_y"] = np.stack([ar.slerp(v, a, b) for v in (0.0, 0.25, 0.5, 1.0)])
    out["slerp_same"] = ar.slerp(0.3, a, a)
    # generate_audiovisual.get_noise_range (generate_audiovisual.py:22-34): which noise scales get_noise is asked for
    import generate_audiovisual as ref_ga

    for o, g in ((1024, 1024), (1920, 1024), (1080, 1024), (512, 512), (1024, 512), (1920, 512), (512, 256)):
        for sg1 in (False, True):
            lo, hi, f = ref_ga.get_noise_range(o, g, sg1)
            out[f"noise_range_{o}_{g}_{int(sg1)}"] = np.array([lo, hi] + [f(s) for s in range(lo, hi)])
    # plugin surface: generate() parameters + defaults (generate_audiovisual.py:59-91) and the CLI flags (:235-260)
    import ast
    import inspect
    import json

    sig = inspect.signature(ref_ga.generate)
    out["generate_signature"] = np.array(json.dumps(
        [[n, None if p.default is inspect.Parameter.empty else repr(p.default)] for n, p in sig.parameters.items()]))
    flags = []
    for node in ast.walk(ast.parse(open(os.path.join(REF, "generate_audiovisual.py")).read())):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            kw = {k.arg: ast.literal_eval(k.value) if not isinstance(k.value, ast.Name) else k.value.id for k in node.keywords}
            flags.append([node.args[0].value, kw.get("type"), kw.get("default"), kw.get("action")])
    out["cli_flags"] = np.array(json.dumps(flags))
    # model surface: Generator constructor / forward signatures and the state_dict layout real checkpoints carry
    if ref_sg2 is not None:
        def sig(fn):  # defaults as repr, function addresses stripped (deterministic fixture)
            import re

            return [[n, None if p.default is inspect.Parameter.empty else re.sub(r" at 0x[0-9a-f]+", "", repr(p.default))]
                    for n, p in inspect.signature(fn).parameters.items()]

        out["generator_init_signature"] = np.array(json.dumps(sig(ref_sg2.Generator.__init__)))
        out["generator_forward_signature"] = np.array(json.dumps(sig(ref_sg2.Generator.forward)))
        for tag, kw in (("1024_cm2_const", dict(size=1024, channel_multiplier=2, constant_input=True)),
                        ("512_cm1_noconst", dict(size=512, channel_multiplier=1, constant_input=False)),
                        ("256_cm2_1920", dict(size=256, channel_multiplier=2, constant_input=True, output_size=1920))):
            g = ref_sg2.Generator(kw.pop("size"), 512, 8, **kw)
            out[f"state_dict_{tag}"] = np.array(json.dumps({k: list(v.shape) for k, v in g.state_dict().items()}))
            out[f"layout_{tag}"] = np.array([g.n_latent, g.num_layers, g.log_size])
            del g
    # helper-package / op / render call signatures (names, order, defaults)
    if ref_sg2 is not None:
        import render as ref_render

        api = {}
        for name in ("onsets", "rms", "raw_chroma", "chroma", "normalize", "percentile", "percentile_clip", "compress",
                     "expand", "gaussian_filter", "load_audio", "set_SMF", "chroma_weight_latents", "slerp", "slerp_loops",
                     "spline_loops", "wrapping_slice", "generate_latents", "save_latents", "load_latents", "perlin_noise"):
            api["ar." + name] = sig(getattr(ar, name))
        for name in ("NetworkBend", "AddNoise", "Translate", "Zoom", "Rotate"):
            api["ar." + name] = sig(getattr(ar, name).__init__)
        api["op.upfirdn2d"] = sig(ref_op.upfirdn2d)
        api["op.fused_leaky_relu"] = sig(ref_op.fused_leaky_relu)
        api["op.FusedLeakyReLU"] = sig(ref_op.FusedLeakyReLU.__init__)
        api["render.render"] = sig(ref_render.render)
        out["api_signatures"] = np.array(json.dumps(api))
    t = torch.arange(10)
    out["wrap_8_5"] = ar.wrapping_slice(t, 8, 5).numpy()
    out["wrap_2_4"] = ar.wrapping_slice(t, 2, 4).numpy()
    return out


def main():
    op, ref_sg2, ar = import_reference()
    if "--glue" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "audio_glue.npz"), **audio_glue_cases(ar))
        return
    if "--noconst" in sys.argv:
        torch.set_grad_enabled(False)
        np.savez_compressed(os.path.join(HERE, "generator_noconst.npz"), **gen_noconst_case(ref_sg2))
        return
    if "--bends" in sys.argv:
        torch.set_grad_enabled(False)
        np.savez_compressed(os.path.join(HERE, "generator_bends.npz"), **gen_bend_case(ref_sg2))
        return
    if "--g256" in sys.argv:
        torch.set_grad_enabled(False)
        np.savez_compressed(os.path.join(HERE, "generator_g256.npz"),
                            **gen_case(ref_sg2, 256, 2, 2, seed=7, psi_lo=0.5, store_noise=False))
        return
    if "--g1024" in sys.argv:
        torch.set_grad_enabled(False)
        np.savez_compressed(os.path.join(HERE, "generator_g1024.npz"), **gen_case_1024(ref_sg2))
        return
    if "--plugins" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "plugins.npz"), **plugin_cases(ar, ref_sg2, op))
        return
    torch.set_grad_enabled(False)
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **ops_cases(op))
    np.savez_compressed(os.path.join(HERE, "generator_g32.npz"), **gen_case(ref_sg2, 32, 2, 2, seed=3, psi_lo=0.5))
    np.savez_compressed(os.path.join(HERE, "generator_g128.npz"), **gen_case(ref_sg2, 128, 1, 1, seed=5, psi_lo=0.7))
    np.savez_compressed(os.path.join(HERE, "generator_g256.npz"),   # BASELINE configs[0] architecture (256^2, cm=2)
                        **gen_case(ref_sg2, 256, 2, 2, seed=7, psi_lo=0.5, store_noise=False))
    np.savez_compressed(os.path.join(HERE, "generator_g1024.npz"), **gen_case_1024(ref_sg2))   # BASELINE configs[1]
    np.savez_compressed(os.path.join(HERE, "audio_glue.npz"), **audio_glue_cases(ar))
    np.savez_compressed(os.path.join(HERE, "generator_bends.npz"), **gen_bend_case(ref_sg2))
    np.savez_compressed(os.path.join(HERE, "generator_noconst.npz"), **gen_noconst_case(ref_sg2))
    np.savez_compressed(os.path.join(HERE, "plugins.npz"), **plugin_cases(ar, ref_sg2, op))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
