"""The plugin-helper oracle (oracle/plugin_oracle.py) against golden vectors written by the UNMODIFIED reference
(tests/golden/plugins.npz <- tests/golden/make_golden.py --plugins), and the host-side latent helpers of the product."""
import os

import numpy as np
import pytest
import torch

from oracle import plugin_oracle as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plugins.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_perlin_oracle_matches_reference(name):
    cfg = G[f"perlin_{name}_cfg"]
    shape, res, tile, seed = tuple(cfg[0:3]), tuple(cfg[3:6]), tuple(bool(v) for v in cfg[6:9]), int(cfg[9])
    np.random.seed(seed)
    y = P.perlin_noise(shape, res, P.perlin_gradients(res, tile))
    ref = G[f"perlin_{name}_y"]
    assert y.shape == ref.shape and y.dtype == np.float64
    assert np.abs(y - ref).max() <= 1e-14
    if tile[0]:  # loops seamlessly in time: frame 0 continues the last frame
        assert np.abs(y[0] - y[-1]).max() < 2.5 * res[0] / shape[0] * 4


@pytest.mark.parametrize("key,args", [("spline_y_100_2", (100, 2, True)), ("spline_y_97_3_noloop", (97, 3, False)),
                                      ("spline_y_64_half", (64, 0.5, True))])
def test_spline_loops_oracle_matches_reference(key, args):
    n_frames, n_loops, loop = args
    y = P.spline_loops(G["spline_sel"], n_frames, n_loops, loop=loop)
    assert y.shape == G[key].shape
    assert np.abs(y - G[key]).max() <= 1e-12


def test_slerp_oracle_matches_reference():
    a, b = G["slerp_a"], G["slerp_b"]
    y = np.stack([P.slerp(v, a, b) for v in (0.0, 0.25, 0.5, 1.0)])
    assert np.array_equal(y, G["slerp_y"])
    assert np.array_equal(P.slerp(0.3, a, a), G["slerp_same"])


def test_slerp_loops_shape_and_period():
    rng = np.random.Generator(np.random.PCG64(5))
    sel = rng.standard_normal((3, 18, 16))
    y = P.slerp_loops(sel, 96, 2, smoothing=1)
    assert y.shape == (96, 18, 16)
    assert np.array_equal(y[:48], y[48:])            # two identical loops
    assert np.array_equal(y[:, 0], y[:, 17])          # layer 0 broadcast to all layers (latent.py:79)


def test_bend_oracle_integer_translation_is_a_roll():
    """Translate by an integer number of pixels == reading the 5x reflect-padded strip at an integer offset."""
    rng = np.random.Generator(np.random.PCG64(9))
    h, w = 4, 8
    x = torch.from_numpy(rng.standard_normal((2, 3, h, w)).astype(np.float32))
    noise = torch.zeros(1, 1, h, 5 * w)
    y0 = P.translate(x, torch.tensor([[0.0, 0.0], [float(w), 0.0]]), h, w, noise)
    # the centre crop of the 5w-wide strip is columns [2w, 3w); a translation of +t pixels reads [2w - t, 3w - t)
    p = torch.nn.functional.pad(x, (w // 2, w // 2, 0, 0), mode="reflect")
    p = torch.nn.functional.pad(p, (w, w, 0, 0), mode="reflect")
    p = torch.nn.functional.pad(p, (w, 0, 0, 0), mode="reflect")
    assert torch.allclose(y0[0], p[0, :, :, 2 * w:3 * w], atol=1e-6)
    assert torch.allclose(y0[1], p[1, :, :, w:2 * w], atol=1e-6)


def test_bend_oracle_identity_zoom_and_rotation():
    rng = np.random.Generator(np.random.PCG64(10))
    x = torch.from_numpy(rng.standard_normal((2, 2, 6, 6)).astype(np.float32))
    assert torch.allclose(P.zoom(x, torch.ones(2), 6, 6), x, atol=1e-5)
    assert torch.allclose(P.rotate(x, torch.zeros(2), 6, 6), x, atol=1e-5)
    r90 = P.rotate(x, torch.full((2,), 90.0), 6, 6)
    assert torch.allclose(r90, torch.rot90(x, 1, (2, 3)), atol=1e-4)  # positive angle = anti-clockwise
