"""The oracle is pinned against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ops_oracle as OO
from oracle import stylegan2_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_upfirdn2d_oracle_matches_reference_native(golden_dir):
    g = _load(golden_dir, "ops_golden.npz")
    for name in g["ufd_names"]:
        x, k, cfg, y = g[f"ufd_{name}_x"], g[f"ufd_{name}_k"], g[f"ufd_{name}_cfg"], g[f"ufd_{name}_y"]
        up, down, p0, p1 = [int(v) for v in cfg]
        for fma in (True, False):
            mine = OO.upfirdn2d_nchw(x, k, up, down, (p0, p1), fma=fma)
            assert mine.shape == y.shape, name
            # summation order differs from F.conv2d: a few ulp of the largest term
            np.testing.assert_allclose(mine, y, rtol=0, atol=2e-6 * max(1.0, np.abs(y).max()), err_msg=name)
        n, c, h, w = x.shape
        dense = OO.upfirdn2d_dense(x.reshape(n * c, h, w), k, up, up, down, down, p0, p1, p0, p1).reshape(y.shape)
        np.testing.assert_allclose(dense, y, rtol=0, atol=2e-6 * max(1.0, np.abs(y).max()), err_msg=name)
        # torch restatement used inside the generator oracle
        t = O.upfirdn2d(torch.from_numpy(x), torch.from_numpy(k), up, down, (p0, p1)).numpy()
        np.testing.assert_allclose(t, y, rtol=0, atol=2e-6 * max(1.0, np.abs(y).max()), err_msg=name)


def test_fused_leaky_relu_oracle_matches_reference_fallback(golden_dir):
    g = _load(golden_dir, "ops_golden.npz")
    for name in ("fl2d", "fl4d", "fl4d_big"):
        x, b, y = g[f"{name}_x"], g[f"{name}_b"], g[f"{name}_y"]
        np.testing.assert_allclose(OO.fused_leaky_relu(x, b), y, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(O.fused_leaky_relu(torch.from_numpy(x), torch.from_numpy(b)).numpy(), y,
                                   rtol=1e-6, atol=1e-7)


def test_fused_bias_act_grad_modes():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    b = rng.standard_normal(3).astype(np.float32)
    ref = rng.standard_normal(x.shape).astype(np.float32)
    y1 = OO.fused_bias_act(x, b, ref, 3, 1, 0.2, 1.5)
    xb = x + b[None, :, None, None]
    np.testing.assert_allclose(y1, np.where(ref > 0, xb, xb * np.float32(0.2)) * np.float32(1.5), rtol=1e-6)
    assert not OO.fused_bias_act(x, b, ref, 3, 2, 0.2, 1.5).any()
    np.testing.assert_allclose(OO.fused_bias_act(x, None, None, 1, 0, 0.2, 2.0), x * 2)


def regenerate_noise(g, size, seed):
    """Fixtures written with store_noise=False (generator_g256.npz) hold only the reference's outputs: the noise maps are
    redrawn from the same PCG64 stream in make_golden.gen_case's order (z, W+ jitter, noise per layer, psi)."""
    batch = int(g["batch"])
    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    rng.standard_normal((batch, 512))
    rng.standard_normal((batch, n_latent, 512))
    noise = [torch.from_numpy(rng.standard_normal((batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    for l in g["noise_none"]:
        noise[int(l)] = None
    return noise


@pytest.mark.parametrize("fname", ["generator_g32.npz", "generator_g128.npz", "generator_g256.npz",
                                   "generator_g1024.npz"])
def test_generator_oracle_matches_reference(golden_dir, fname):
    g = _load(golden_dir, fname)
    size, cm, seed = int(g["size"]), int(g["cm"]), int(g["seed"])
    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed)
    _, num_layers, _ = O.layout(size)
    if any(k.startswith("noise_") and k != "noise_none" for k in g.files):
        noise = [torch.from_numpy(g[f"noise_{l}"]) if f"noise_{l}" in g.files else None for l in range(num_layers)]
    else:
        noise = regenerate_noise(g, size, seed)
    with torch.no_grad():
        w = O.mapping(torch.from_numpy(g["z"]), sd)
        np.testing.assert_allclose(w.numpy(), g["w"], rtol=1e-4, atol=1e-5)
        image, acts = O.generator_forward(sd, size, torch.from_numpy(g["latent"]), noise, torch.from_numpy(g["psi"]),
                                          torch.from_numpy(g["truncation_latent"]), channel_multiplier=cm)
    from tests.golden.make_golden import strided

    if "image_stride" in g.files:   # 1024^2 fixture: strided image + full-resolution centre crop
        st, c0, scale = int(g["image_stride"]), size // 2 - 64, float(g["image_absmax"])
        im = image.numpy()
        assert np.abs(im[:, :, ::st, ::st] - g["image"]).max() <= 2e-5 * scale
        assert np.abs(im[:, :, c0:c0 + 128, c0:c0 + 128] - g["image_crop"]).max() <= 2e-5 * scale
    else:
        scale = np.abs(g["image"]).max()
        assert np.abs(image.numpy() - g["image"]).max() <= 2e-5 * scale

    for l, a in enumerate(acts):
        ref = g[f"act_{l}"]
        assert np.abs(strided(a) - ref).max() <= 2e-5 * float(g[f"act_{l}_absmax"]), l


def test_generator_oracle_with_bends_matches_reference(golden_dir):
    """transform_dict_list: layer-id placement (0 = constant input, 2 = first up-conv, 5 = 16^2 conv) and the non-square
    H x 2H data flow, against the reference's own forward with the same bends (generator_bends.npz)."""
    from tests.golden.make_golden import bend_case_inputs, bend_list, strided

    g = _load(golden_dir, "generator_bends.npz")
    size, cm, seed, batch = int(g["size"]), int(g["cm"]), int(g["seed"]), int(g["batch"])
    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed)
    latent, noise, tl = bend_case_inputs(size, batch, seed)
    with torch.no_grad():
        image, acts = O.generator_forward(sd, size, latent, noise, torch.ones(batch), tl, channel_multiplier=cm,
                                          bends=bend_list())
    assert tuple(image.shape) == (batch, 3, size, 2 * size) == g["image"].shape
    assert np.abs(image.numpy() - g["image"]).max() <= 2e-5 * np.abs(g["image"]).max()
    for l, a in enumerate(acts):
        assert list(a.shape) == list(g[f"act_{l}_shape"]), l
        assert np.abs(strided(a) - g[f"act_{l}"]).max() <= 2e-5 * float(g[f"act_{l}_absmax"]), l


def test_generator_oracle_noconst_matches_reference(golden_dir):
    """`--noconst` (LatentInput, models/stylegan2.py:281-294): first latent row -> EqualLinear(fused_lrelu) ->
    FusedLeakyReLU -> [B,512,4,4], with per-sample truncation, against the reference's forward."""
    from tests.golden.make_golden import noconst_case_inputs, strided

    g = _load(golden_dir, "generator_noconst.npz")
    size, cm, seed, batch = int(g["size"]), int(g["cm"]), int(g["seed"]), int(g["batch"])
    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed, noconst=True)
    assert sd["input.linear.weight"].shape == (8192, 512) and sd["input.input"].shape == (1,)
    latent, noise, tl, psi = noconst_case_inputs(size, batch, seed)
    with torch.no_grad():
        image, acts = O.generator_forward(sd, size, latent, noise, psi, tl, channel_multiplier=cm)
    assert np.abs(image.numpy() - g["image"]).max() <= 2e-5 * np.abs(g["image"]).max()
    for l, a in enumerate(acts):
        assert np.abs(strided(a) - g[f"act_{l}"]).max() <= 2e-5 * float(g[f"act_{l}_absmax"]), l
