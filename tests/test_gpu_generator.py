"""Network-level parity: our Generator (both conv implementations) vs the golden vectors produced by the unmodified
reference, and vs the CPU oracle at BASELINE.json's config[0] shape (256x256, channel_multiplier=2).

Metric (fixed once): max|a-b| / max|ref| per tensor, for the image and every activation map.
Tolerance: 1e-3 (north_star); the fp32 SIMT path is held to 2e-5, the 3-product tensor-core path to 3e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as O
from tests.util import GOLDEN, golden_inputs, make_generator, rel_err, strided

pytestmark = pytest.mark.gpu

TOL = {"simt": 2e-5, "tc": 3e-4, "mixed": 1e-3}   # "mixed" = precision="mixed" (fp16 activations at >= 512^2): north_star bar


def _run(g, gold, noise, **kw):
    g.truncation_latent = torch.from_numpy(gold["truncation_latent"]).cuda()
    with torch.no_grad():
        return g(torch.from_numpy(gold["latent"]).cuda(), noise=[n.cuda() if n is not None else None for n in noise],
                 truncation=torch.from_numpy(gold["psi"]).cuda(), input_is_latent=True, randomize_noise=False, **kw)


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("fname", ["generator_g32.npz", "generator_g128.npz"])
def test_generator_matches_reference_golden(fname, impl):
    gold, noise = golden_inputs(fname)
    g, sd = make_generator(int(gold["size"]), int(gold["cm"]), int(gold["seed"]), impl)
    # mapping network (2-D path)
    with torch.no_grad():
        w = g.get_latent(torch.from_numpy(gold["z"]).cuda())
    assert rel_err(w.cpu().numpy(), gold["w"]) < 1e-5
    image, acts = _run(g, gold, noise, return_activation_maps=True)
    errs = [rel_err(strided(a), gold[f"act_{l}"]) * float(np.abs(gold[f"act_{l}"]).max() / gold[f"act_{l}_absmax"])
            for l, a in enumerate(acts)]
    e_img = rel_err(image.cpu().numpy(), gold["image"])
    print(f"{fname} {impl}: image {e_img:.2e} acts {['%.1e' % e for e in errs]}")
    assert max(errs) < TOL[impl], errs
    assert e_img < TOL[impl], e_img
    # the fused path without activation maps must give the same image (different buffers / epilogue outputs)
    image2, none = _run(g, gold, noise)
    assert none is None
    assert rel_err(image2.cpu().numpy(), image.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_generator_256_config0_vs_oracle(impl):
    """BASELINE.json configs[0]: 256x256 random-init generator (cm=2); 3 of the 64 latents are checked against the
    CPU oracle (≈1 s/frame), with a per-sample truncation and buffer noise for one layer."""
    size, cm, seed, b = 256, 2, 0, 3
    g, sd = make_generator(size, cm, seed, impl)
    log_size, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(1))
    z = torch.from_numpy(rng.standard_normal((b, 512)).astype(np.float32))
    noise = [torch.from_numpy(rng.standard_normal((b, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    noise[4] = None
    psi = torch.from_numpy(rng.uniform(0.5, 1.0, b).astype(np.float32))
    with torch.no_grad():
        w = O.mapping(z, sd)
        latent = w[:, None, :].repeat(1, n_latent, 1)
        tl = w.mean(0, keepdim=True)
        ref_img, ref_acts = O.generator_forward(sd, size, latent, noise, psi, tl, channel_multiplier=cm)
        g.truncation_latent = tl.cuda()
        img, acts = g(latent.cuda(), noise=[n.cuda() if n is not None else None for n in noise], truncation=psi.cuda(),
                      input_is_latent=True, randomize_noise=False, return_activation_maps=True)
    errs = [rel_err(a.cpu().numpy(), r.numpy()) for a, r in zip(acts, ref_acts)]
    e_img = rel_err(img.cpu().numpy(), ref_img.numpy())
    print(f"256 {impl}: image {e_img:.2e} acts {['%.1e' % e for e in errs]}")
    assert max(errs) < TOL[impl] and e_img < TOL[impl]
    # bytes: device-side uint8 NHWC pack vs render.py:40-43 semantics on our own fp32 image
    from maua_stylegan2_b200.stylegan2 import frames_to_u8

    u8 = frames_to_u8(img).cpu().numpy()
    assert np.array_equal(u8, O.frames_to_u8(img.cpu()))
    diff = np.abs(u8.astype(np.int32) - O.frames_to_u8(ref_img).astype(np.int32))
    assert diff.max() <= 1  # truncation can flip a byte where the fp32 images differ by 1e-4


def test_generator_api_surface():
    g, sd = make_generator(32, 2, 3, "tc")
    z = torch.randn(4, 512, device="cuda")
    with torch.no_grad():
        assert g(z, map_latents=True).shape == (4, g.n_latent, 512)   # values: tests/test_gpu_plugins.py
        lat = g.get_latent(z)[:, None, :].repeat(1, g.n_latent, 1)
        g.truncation_latent = g.mean_latent(256)
        img, lat_out = g([z], return_latents=True, truncation=0.7, randomize_noise=False)
        assert img.shape == (4, 3, 32, 32) and lat_out.shape == (4, g.n_latent, 512)
        # float truncation == tensor truncation
        img2, _ = g(lat, input_is_latent=True, truncation=torch.full((4,), 0.7, device="cuda"), randomize_noise=False)
        assert rel_err(img2.cpu().numpy(), img.cpu().numpy()) < 1e-6
        # randomize_noise=True draws fresh noise: images differ, shape stays
        img3, _ = g(lat, input_is_latent=True, truncation=0.7)
        assert img3.shape == img.shape and not torch.equal(img3, img)
        u8, _ = g(lat, input_is_latent=True, truncation=0.7, randomize_noise=False, return_u8=True)
        assert u8.dtype == torch.uint8 and u8.shape == (4, 32, 32, 3)


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_network_bend_changes_shape_like_reference(impl):
    """A bend that doubles the width at layer 1, before the first ToRGB (kelp-style padding): every later layer
    and the RGB skip become non-square.
    Checked against the oracle running the same torch transform."""
    size, cm, seed, b = 32, 2, 3, 2
    g, sd = make_generator(size, cm, seed, impl)

    class Widen(torch.nn.Module):
        def forward(self, x):
            return torch.cat([x, torch.flip(x, [3])], 3)

    bends = [{"layer": 1, "transform": Widen()}, {"layer": 0, "transform": torch.nn.Identity()}]
    log_size, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(9))
    latent = torch.from_numpy(rng.standard_normal((b, n_latent, 512)).astype(np.float32)) * 0.5
    noise = []
    for l in range(num_layers):
        r = 2 ** ((l + 5) // 2)
        wmul = 2 if l >= 1 else 1
        noise.append(torch.from_numpy(rng.standard_normal((b, 1, r, r * wmul)).astype(np.float32)))
    tl = torch.zeros(1, 512)
    with torch.no_grad():
        ref_img, ref_acts = O.generator_forward(sd, size, latent, noise, 1.0, tl, channel_multiplier=cm, bends=bends)
        g.truncation_latent = tl.cuda()
        img, acts = g(latent.cuda(), noise=[n.cuda() for n in noise], truncation=1.0, input_is_latent=True,
                      transform_dict_list=bends, return_activation_maps=True)
    assert img.shape == ref_img.shape == (b, 3, 32, 64)
    assert rel_err(img.cpu().numpy(), ref_img.numpy()) < TOL[impl]
    for a, r in zip(acts, ref_acts):
        assert a.shape == r.shape and rel_err(a.cpu().numpy(), r.numpy()) < TOL[impl]


def test_generator_1024_configf_vs_oracle():
    """BASELINE.json configs[1] architecture (1024x1024 config-f, channel_multiplier=2) at full size, one frame:
    every layer of the bench workload (halo conv R/BN/concat policies, fused ToRGB, TMA blur) against the CPU oracle."""
    size, cm, seed, b = 1024, 2, 0, 1
    g, sd = make_generator(size, cm, seed, "tc")
    log_size, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(2))
    latent = torch.from_numpy(rng.standard_normal((b, n_latent, 512)).astype(np.float32)) * 0.5
    noise = [torch.from_numpy(rng.standard_normal((b, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             if 2 ** ((l + 5) // 2) <= 256 else None for l in range(num_layers)]   # default hooks: buffers above 256
    tl = torch.zeros(1, 512)
    with torch.no_grad():
        ref_img, ref_acts = O.generator_forward(sd, size, latent, noise, 0.9, tl, channel_multiplier=cm)
        g.truncation_latent = tl.cuda()
        img, acts = g(latent.cuda(), noise=[n.cuda() if n is not None else None for n in noise], truncation=0.9,
                      input_is_latent=True, randomize_noise=False, return_activation_maps=True)
        img2, _ = g(latent.cuda(), noise=[n.cuda() if n is not None else None for n in noise], truncation=0.9,
                    input_is_latent=True, randomize_noise=False)       # fused-ToRGB / no-fp32-map path
    errs = [rel_err(a.cpu().numpy(), r.numpy()) for a, r in zip(acts, ref_acts)]
    e1, e2 = rel_err(img.cpu().numpy(), ref_img.numpy()), rel_err(img2.cpu().numpy(), ref_img.numpy())
    print(f"1024 tc: image {e1:.2e} / fused {e2:.2e} acts {['%.1e' % e for e in errs]}")
    assert max(errs) < TOL["tc"] and e1 < TOL["tc"] and e2 < TOL["tc"]


@pytest.mark.parametrize("batch,precision", [(8, "bf16x3"), (16, "bf16x3"), (8, "mixed"), (16, "mixed")])
def test_generator_1024_bench_batches_vs_oracle(batch, precision):
    """The MEASURED configurations: BASELINE configs[1] (batch 8, what bench.py times) and configs[2] (batch 16).  The tile
    policy of the conv kernels is batch dependent (wide32 needs 16*B >= 120 items, the n_ctas < 120 fallbacks flip with
    B), so the full batch runs through the product and two strided samples of it are checked against the CPU oracle —
    every activation map and the image — plus the fused (no activation maps) path and the uint8 frames bench.py emits."""
    size, cm, seed = 1024, 2, 0
    g, sd = make_generator(size, cm, seed, "tc", precision=precision)
    tol = TOL["mixed" if precision == "mixed" else "tc"]
    log_size, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(100 + batch))
    latent = torch.from_numpy(rng.standard_normal((batch, n_latent, 512)).astype(np.float32)) * 0.5
    noise = [torch.from_numpy(rng.standard_normal((batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             if 2 ** ((l + 5) // 2) <= 256 else None for l in range(num_layers)]   # default hooks: buffers above 256
    psi = torch.from_numpy(rng.uniform(0.6, 1.0, batch).astype(np.float32))
    tl = torch.from_numpy(rng.standard_normal((1, 512)).astype(np.float32)) * 0.1
    pick = [1, batch - 3]
    with torch.no_grad():
        g.truncation_latent = tl.cuda()
        dn = [n.cuda() if n is not None else None for n in noise]
        img, acts = g(latent.cuda(), noise=dn, truncation=psi.cuda(), input_is_latent=True, randomize_noise=False,
                      return_activation_maps=True)
        acts = [a[pick].cpu() for a in acts]
        img = img[pick].cpu()
        fused, _ = g(latent.cuda(), noise=dn, truncation=psi.cuda(), input_is_latent=True, randomize_noise=False)
        fused = fused[pick].cpu()
        u8, _ = g(latent.cuda(), noise=dn, truncation=psi.cuda(), input_is_latent=True, randomize_noise=False,
                  return_u8=True)
        u8 = u8[pick].cpu().numpy()
        torch.cuda.empty_cache()
        ref_img, ref_acts = O.generator_forward(sd, size, latent[pick], [n[pick] if n is not None else None for n in noise],
                                                psi[pick], tl, channel_multiplier=cm)
    errs = [rel_err(a.numpy(), r.numpy()) for a, r in zip(acts, ref_acts)]
    e1, e2 = rel_err(img.numpy(), ref_img.numpy()), rel_err(fused.numpy(), ref_img.numpy())
    print(f"1024 tc {precision} batch {batch}: image {e1:.2e} / fused {e2:.2e} acts {['%.1e' % e for e in errs]}")
    assert max(errs) < tol and e1 < tol and e2 < tol
    diff = np.abs(u8.astype(np.int32) - O.frames_to_u8(ref_img).astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < (1e-2 if precision == "bf16x3" else 1e-1)   # truncation to uint8 flips a byte where fp32 differs by 1e-4


@pytest.mark.parametrize("impl", ["tc", "mixed"])
def test_generator_1024_matches_reference_golden(impl):
    """BASELINE configs[1] architecture against the UNMODIFIED reference's own CPU output (tests/golden/generator_g1024.npz,
    written by make_golden.py --g1024): strided activation maps, strided image and a full-resolution centre crop."""
    from tests.test_oracle_golden import regenerate_noise

    gold = np.load(os.path.join(GOLDEN, "generator_g1024.npz"))
    size, cm, seed = int(gold["size"]), int(gold["cm"]), int(gold["seed"])
    g, sd = make_generator(size, cm, seed, "tc", precision="mixed" if impl == "mixed" else "bf16x3")
    noise = regenerate_noise(gold, size, seed)
    g.truncation_latent = torch.from_numpy(gold["truncation_latent"]).cuda()
    with torch.no_grad():
        img, acts = g(torch.from_numpy(gold["latent"]).cuda(), noise=[n.cuda() if n is not None else None for n in noise],
                      truncation=torch.from_numpy(gold["psi"]).cuda(), input_is_latent=True, randomize_noise=False,
                      return_activation_maps=True)
    st, c0 = int(gold["image_stride"]), size // 2 - 64
    amax = float(gold["image_absmax"])
    img = img.cpu().numpy()
    e_img = max(np.abs(img[:, :, ::st, ::st] - gold["image"]).max(),
                np.abs(img[:, :, c0:c0 + 128, c0:c0 + 128] - gold["image_crop"]).max()) / amax
    errs = [float(np.abs(strided(a) - gold[f"act_{l}"]).max() / float(gold[f"act_{l}_absmax"])) for l, a in enumerate(acts)]
    print(f"1024 golden {impl}: image {e_img:.2e} acts {['%.1e' % e for e in errs]}")
    assert e_img < TOL[impl] and max(errs) < TOL[impl]


def test_generator_512_confige_bend_and_truncation_sweep():
    """BASELINE.json configs[3] shape: 512x512 config-e (channel_multiplier=1), a translation bend on layer 6 driven by a
    per-frame modulation and a per-frame truncation sweep; kornia is absent, so the bend is an integer torch.roll
    (same contract: a `transform(modulation_batch) -> nn.Module` factory), checked against the oracle."""
    size, cm, seed, b = 512, 1, 4, 2
    g, sd = make_generator(size, cm, seed, "tc")

    class Roll(torch.nn.Module):
        def __init__(self, shifts):
            super().__init__()
            self.shifts = shifts

        def forward(self, x):
            return torch.stack([torch.roll(xi, int(s), dims=2) for xi, s in zip(x, self.shifts)])

    modulation = torch.tensor([3.0, 11.0])
    bends = [{"layer": 6, "transform": Roll(modulation)}]
    log_size, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(5))
    latent = torch.from_numpy(rng.standard_normal((b, n_latent, 512)).astype(np.float32)) * 0.5
    noise = [torch.from_numpy(rng.standard_normal((b, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    psi = torch.tensor([0.55, 0.95])
    tl = torch.from_numpy(rng.standard_normal((1, 512)).astype(np.float32)) * 0.1
    with torch.no_grad():
        ref_img, ref_acts = O.generator_forward(sd, size, latent, noise, psi, tl, channel_multiplier=cm, bends=bends)
        g.truncation_latent = tl.cuda()
        img, acts = g(latent.cuda(), noise=[n.cuda() for n in noise], truncation=psi.cuda(), input_is_latent=True,
                      transform_dict_list=bends, return_activation_maps=True)
    assert img.shape == (b, 3, 512, 512)
    errs = [rel_err(a.cpu().numpy(), r.numpy()) for a, r in zip(acts, ref_acts)]
    assert max(errs) < TOL["tc"] and rel_err(img.cpu().numpy(), ref_img.numpy()) < TOL["tc"]


@pytest.mark.parametrize("size,cm", [(256, 1), (512, 1)])
def test_return_u8_equals_u8_of_the_fp32_image(size, cm):
    """`return_u8=True` ends in maua_rgb_finish_u8 (last ToRGB + bias + skip + clamp/scale/truncate in one pass, no fp32
    image): the bytes must equal the two-step path (fp32 image -> maua_rgb_to_u8_nhwc) and render.py:40-43's arithmetic."""
    from maua_stylegan2_b200.stylegan2 import frames_to_u8

    g, _ = make_generator(size, cm, 12, "tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    gen = torch.Generator().manual_seed(3)
    latent = (torch.randn(3, g.n_latent, 512, generator=gen) * 0.6).cuda()
    with torch.no_grad():
        img, _ = g(latent, truncation=0.9, input_is_latent=True, randomize_noise=False)
        u8, _ = g(latent, truncation=0.9, input_is_latent=True, randomize_noise=False, return_u8=True)
    assert u8.dtype == torch.uint8 and tuple(u8.shape) == (3, size, size, 3)
    assert torch.equal(u8, frames_to_u8(img))
    assert np.array_equal(u8.cpu().numpy(), O.frames_to_u8(img.cpu()))
