"""GPU parity of the §8(f) plugin-side kernels: perlin_noise, spline/slerp loops and the fused network-bend warp
(csrc/bend.cu) against the reference's golden vectors (tests/golden/plugins.npz) and oracle/plugin_oracle.py."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import plugin_oracle as P

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plugins.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_perlin_kernel_matches_reference_golden(name):
    from maua_stylegan2_b200 import audioreactive as ar

    cfg = G[f"perlin_{name}_cfg"]
    shape, res, tile, seed = tuple(int(v) for v in cfg[0:3]), tuple(int(v) for v in cfg[3:6]), tuple(bool(v) for v in cfg[6:9]), int(cfg[9])
    np.random.seed(seed)
    y = ar.perlin_noise(shape=shape, res=res, tileable=tile)
    assert y.is_cuda and y.dtype == torch.float64 and tuple(y.shape) == shape
    ref = G[f"perlin_{name}_y"]
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13      # same operation order in fp64; libm sin/cos on the host
    np.random.seed(seed)
    y32 = ar.perlin_noise(shape=shape, res=res, tileable=tile, dtype=torch.float32)
    assert np.array_equal(y32.cpu().numpy(), y.cpu().numpy().astype(np.float32))


def test_perlin_kernel_full_size_vs_oracle():
    """kelp.py-sized call: one 2-bar loop of 256x256 noise, res (8, 4, 4)."""
    from maua_stylegan2_b200 import audioreactive as ar

    shape, res = (112, 256, 256), (8, 4, 4)
    np.random.seed(7)
    y = ar.perlin_noise(shape=shape, res=res)
    np.random.seed(7)
    ref = P.perlin_noise(shape, res, P.perlin_gradients(res))
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-13
    with pytest.raises(ValueError):
        ar.perlin_noise(shape=(10, 8, 8), res=(3, 2, 2))


@pytest.mark.parametrize("key,args", [("spline_y_100_2", (100, 2, True)), ("spline_y_97_3_noloop", (97, 3, False)),
                                      ("spline_y_64_half", (64, 0.5, True))])
def test_spline_loops_matches_reference_golden(key, args):
    from maua_stylegan2_b200 import audioreactive as ar

    n_frames, n_loops, loop = args
    y = ar.spline_loops(torch.from_numpy(G["spline_sel"]), n_frames, n_loops, loop=loop)
    assert y.is_cuda and tuple(y.shape) == G[key].shape
    ref = G[key]
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()   # fp32 mixing of a float64 reference


def test_slerp_loops_vs_oracle():
    from maua_stylegan2_b200 import audioreactive as ar

    rng = np.random.Generator(np.random.PCG64(5))
    sel = rng.standard_normal((3, 18, 512)).astype(np.float32)
    ar.set_SMF(1)
    y = ar.slerp_loops(sel, 96, 2, smoothing=2)
    ref = P.slerp_loops(sel, 96, 2, smoothing=2)
    assert tuple(y.shape) == ref.shape == (96, 18, 512)
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.array_equal(ar.slerp(0.25, G["slerp_a"], G["slerp_b"]), G["slerp_y"][1])


def _feat(rng, b, c, h, w):
    return torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))


def test_translate_bend_integer_shift_is_exact():
    from maua_stylegan2_b200 import audioreactive as ar

    rng = np.random.Generator(np.random.PCG64(1))
    h, w = 16, 32
    x = _feat(rng, 3, 20, h, w)
    noise = 0.2 * _feat(rng, 1, 1, h, 5 * w)
    t = torch.tensor([[0.0, 0.0], [float(w), 0.0], [7.0, 0.0]])
    y = ar.Translate(t.cuda(), h, w, noise.cuda())(x.cuda()).cpu()
    p = torch.nn.functional.pad(x, (w // 2, w // 2, 0, 0), mode="reflect")
    p = torch.nn.functional.pad(p, (w, w, 0, 0), mode="reflect")
    p = torch.nn.functional.pad(p, (w, 0, 0, 0), mode="reflect") + noise
    for b, s in enumerate((0, w, 7)):
        assert torch.equal(y[b], p[b, :, :, 2 * w - s:3 * w - s])     # integer shifts: weights are exactly 0 / 1


@pytest.mark.parametrize("h,w", [(16, 32), (32, 32), (8, 16)])
def test_translate_bend_vs_oracle(h, w):
    from maua_stylegan2_b200 import audioreactive as ar

    rng = np.random.Generator(np.random.PCG64(2))
    x = _feat(rng, 4, 37, h, w)
    noise = 0.2 * _feat(rng, 1, 1, h, 5 * w)
    t = torch.from_numpy(np.stack([rng.uniform(0, w, 4), rng.uniform(-1.5, 1.5, 4)], 1).astype(np.float32))
    y = ar.Translate(t.cuda(), h, w, noise.cuda())(x.cuda()).cpu()
    ref = P.translate(x, t, h, w, noise)
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) <= 2e-4 * float(ref.abs().max())   # fp32 coordinates (kornia: parity unpinned)


@pytest.mark.parametrize("h,w", [(16, 16), (8, 16)])
def test_zoom_and_rotate_bends_vs_oracle(h, w):
    from maua_stylegan2_b200 import audioreactive as ar

    rng = np.random.Generator(np.random.PCG64(3))
    x = _feat(rng, 4, 24, h, w)
    from maua_stylegan2_b200._lib import MauaError

    s = torch.tensor([1.0, 1.3, 0.8, 2.0])
    if h == w:
        y = ar.Zoom(s.cuda(), h, w)(x.cuda()).cpu()
        ref = P.zoom(x, s, h, w)
        assert float((y - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
        assert torch.equal(y[0], x[0])                                       # scale 1 is the identity, exactly
    else:  # ReflectionPad2d(max(h, w) - 1) exceeds the short axis: the reference raises too (bend.py:82-84)
        with pytest.raises(MauaError):
            ar.Zoom(s.cuda(), h, w)(x.cuda())
    a = torch.tensor([0.0, 30.0, -45.0, 90.0])
    y = ar.Rotate(a.cuda(), h, w)(x.cuda()).cpu()
    ref = P.rotate(x, a, h, w)
    assert float((y - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    assert torch.equal(y[0], x[0])
    if h == w:
        assert float((y[3] - torch.rot90(x[3], 1, (1, 2))).abs().max()) <= 1e-4


def test_fused_warp_replicate_and_vertical_stages():
    """Generic chain: replicate pads on both axes (two stages in y), per-channel noise, identity affine, off-centre crop
    sizes — checked against torch's own F.pad."""
    from maua_stylegan2_b200.audioreactive.bend import FusedWarp

    rng = np.random.Generator(np.random.PCG64(4))
    x = _feat(rng, 2, 5, 6, 7)
    noise = _feat(rng, 2, 5, 6 + 1 + 2 + 3 + 0, 7 + 4 + 1)
    ident = lambda hp, wp: torch.tensor([[1.0, 0, 0, 0, 1.0, 0]]).repeat(2, 1).reshape(2, 2, 3)
    mod = FusedWarp([(4, 1)], [(1, 2), (3, 0)], noise.cuda(), ident, (8, 6), pad_mode="replicate")
    y = mod(x.cuda()).cpu()
    p = torch.nn.functional.pad(x, (4, 1, 1, 2), mode="replicate")
    p = torch.nn.functional.pad(p, (0, 0, 3, 0), mode="replicate") + noise
    hp, wp = p.shape[2:]
    y0, x0 = int(hp / 2 - 8 / 2), int(wp / 2 - 6 / 2)
    assert torch.equal(y, p[:, :, y0:y0 + 8, x0:x0 + 6])


def test_bend_in_generator_matches_oracle_translate():
    """configs[3]-style bend: Translate on layer 6 of a 64x64 generator through the public forward, vs the CPU oracle
    generator with the oracle translate at the same layer."""
    from maua_stylegan2_b200 import audioreactive as ar
    from oracle import stylegan2_oracle as O
    from tests.util import make_generator, rel_err

    size, cm = 64, 1
    g, sd = make_generator(size, cm, seed=11, impl="tc")
    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(12))
    B = 2
    latent = torch.from_numpy(rng.standard_normal((B, n_latent, 512)).astype(np.float32)) * 0.6
    noise = [torch.from_numpy(rng.standard_normal((B, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    h = w = 32
    bnoise = 0.2 * torch.from_numpy(rng.standard_normal((1, 1, h, 5 * w)).astype(np.float32))
    t = torch.tensor([[5.25, 0.0], [20.5, 0.0]])
    tl = torch.zeros(1, 512)
    g.truncation_latent = tl.cuda()
    bend = {"layer": 6, "transform": ar.Translate(t.cuda(), h, w, bnoise.cuda())}
    with torch.no_grad():
        img, _ = g(latent.cuda(), noise=[n.cuda() for n in noise], truncation=1.0, input_is_latent=True,
                   randomize_noise=False, transform_dict_list=[bend])
        ref, _ = O.generator_forward(sd, size, latent, noise, 1.0, tl, channel_multiplier=cm,
                                     bends=[{"layer": 6, "transform": lambda a: P.translate(a, t, h, w, bnoise)}])
    assert rel_err(img.cpu().numpy(), ref.numpy()) < 1e-3


def test_map_latents_matches_reference_cuda_3d_path():
    """generate_latents / Generator(map_latents=True): the reference's CUDA path on [1,1,512] inputs (singleton PixelNorm +
    bias[0] broadcast).  Checked against the oracle restatement AND against the chain rebuilt from the reference's own
    compiled fused_bias_act CUDA op (oracle/_ref) + torch fp32 matmul."""
    import math

    from oracle import stylegan2_oracle as O
    from oracle.build_ref import load_ref
    from tests.util import make_generator, rel_err

    g, sd = make_generator(32, 2, seed=3, impl="tc")
    rng = np.random.Generator(np.random.PCG64(8))
    with torch.no_grad():
        for i in range(1, 9):  # the synthetic state dict has zero mapping biases: make bias[0] != bias[n] matter
            b = torch.from_numpy(rng.standard_normal(512).astype(np.float32)) * 20
            sd[f"style.{i}.bias"] = b
            getattr(g.style, str(i)).bias.copy_(b.cuda())
    z = torch.from_numpy(rng.standard_normal((5, 512)).astype(np.float32))
    with torch.no_grad():
        lat = g(z.cuda(), map_latents=True)
    assert tuple(lat.shape) == (5, g.n_latent, 512)
    assert torch.equal(lat[:, 0], lat[:, -1])
    ref = O.mapping_3d_cuda(z, sd)
    assert rel_err(lat[:, 0].cpu().numpy(), ref.numpy()) < 1e-5
    fused = load_ref("fused_ref")
    if fused is not None:
        outs = []
        empty = torch.empty(0, device="cuda")
        for s in z.cuda():
            x = s[None, None, :]
            x = x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
            for i in range(1, 9):
                w, b = sd[f"style.{i}.weight"].cuda(), sd[f"style.{i}.bias"].cuda()
                out = torch.nn.functional.linear(x, w * (1 / math.sqrt(512)) * 0.01)
                x = fused.fused_bias_act(out, b * 0.01, empty, 3, 0, 0.2, 2 ** 0.5)
            outs.append(x)
        r = torch.cat(outs, 0)[:, 0]
        assert rel_err(lat[:, 0].cpu().numpy(), r.cpu().numpy()) < 1e-5
    os.environ["MAUA_MAP_LATENTS"] = "2d"
    try:
        with torch.no_grad():
            lat2 = g(z.cuda(), map_latents=True)
    finally:
        del os.environ["MAUA_MAP_LATENTS"]
    assert rel_err(lat2[:, 0].cpu().numpy(), O.mapping(z, sd).numpy()) < 1e-5
