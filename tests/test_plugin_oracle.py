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


def test_spline_weight_matrix_reproduces_reference_spline_loops():
    """Host logic of the product's spline_loops: the FITPACK-equivalent basis matrix times the selection == the reference's
    per-coordinate splrep/splev (golden), for both the looped (5 knots) and open (4 knots) case."""
    from maua_stylegan2_b200.audioreactive.latent import spline_weights

    sel = G["spline_sel"].astype(np.float64)
    loop = np.concatenate([sel, sel[[0]]])
    y = np.einsum("tn,nld->tld", spline_weights(5, 50), loop)
    assert np.abs(y - G["spline_y_100_2"][:50]).max() <= 1e-12
    assert np.array_equal(G["spline_y_100_2"][:50], G["spline_y_100_2"][50:])
    y = np.einsum("tn,nld->tld", spline_weights(4, 32), sel)
    assert np.abs(y - G["spline_y_97_3_noloop"][:32]).max() <= 1e-12
    with pytest.raises(ValueError):
        spline_weights(3, 10)


def test_log_filterbank_host_table_matches_oracle():
    """madmom-flavoured onsets: the product's filterbank table (host side) against the oracle restatement, and the band
    ranges the ComplexFlux mask kernel reads."""
    from maua_stylegan2_b200.audioreactive import filters
    from oracle import audio_oracle as A

    for sr, fmin, fmax in ((22050, 20, 8000), (44100, 20, 150), (44100, 500, 8000), (48000, 30, 17000)):
        fb, lo, hi = filters.log_filterbank(sr, 1024, 24, fmin, fmax)
        ref = A.mm_log_filterbank(sr, 1024, 24, fmin, fmax)
        assert fb.shape == ref.T.shape and np.array_equal(fb.T, ref)
        np.testing.assert_allclose(fb.sum(1), 1.0, atol=1e-5)            # norm_filters=True
        for b in range(fb.shape[0]):
            nz = np.nonzero(fb[b])[0]
            assert lo[b] == max(nz[0] - 1, 0) and hi[b] == min(nz[-1] + 2, 1024)


def test_madmom_onset_oracle_reacts_to_clicks():
    """Sanity of the restated detection functions: a click train produces peaks at the click frames."""
    from oracle import audio_oracle as A

    sr = 22050
    y = np.zeros(sr * 2, np.float32)
    clicks = [0.5, 1.0, 1.5]
    rng = np.random.Generator(np.random.PCG64(0))
    for c in clicks:
        i = int(c * sr)
        y[i:i + 200] += rng.standard_normal(200).astype(np.float32) * np.hanning(200).astype(np.float32)
    o = A.onset_strength_mm(y, sr, 20, 8000)
    assert o.shape == (int(np.ceil(len(y) / 441)),)
    peaks = sorted(np.argsort(o)[-3:])
    want = [int(round(c * sr / 441)) for c in clicks]
    assert all(abs(p - w) <= 3 for p, w in zip(peaks, want)), (peaks, want)


def test_get_noise_range_matches_reference():
    """generate_audiovisual.get_noise_range (generate_audiovisual.py:22-34) decides which (height, width) get_noise is
    called with; golden values come from the reference function itself."""
    from maua_stylegan2_b200.generate_audiovisual import get_noise_range

    keys = [k for k in G.files if k.startswith("noise_range_")]
    assert len(keys) == 14
    for k in keys:
        o, g, sg1 = (int(v) for v in k.split("_")[2:])
        lo, hi, f = get_noise_range(o, g, bool(sg1))
        assert [lo, hi] + [f(s) for s in range(lo, hi)] == list(G[k]), k


def test_plugin_surface_matches_reference():
    """generate() parameters (names, order, defaults) and the CLI flags are the reference's, 1:1 (goldens extracted from
    the reference's own signature / argparse calls).  Ours may only ADD keyword arguments at the end."""
    import inspect
    import json

    from maua_stylegan2_b200 import generate_audiovisual as GA

    ref_sig = json.loads(str(G["generate_signature"]))
    ours = [[n, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for n, p in inspect.signature(GA.generate).parameters.items()]
    for (rn, rd), (on, od) in zip(ref_sig, ours):
        assert rn == on, (rn, on)
        if rn != "audioreactive_file":          # ours defaults to the packaged device-path default hook file
            assert rd == od, (rn, rd, od)
    assert [n for n, _ in ours[len(ref_sig):]] == ["audio", "sink", "generator", "latent_selection"]

    ref_flags = json.loads(str(G["cli_flags"]))
    kinds = {"str": str, "int": int, "float": float}
    got = {f"--{flag}": (kind, default) for flag, kind, default in GA.CLI_FLAGS}
    assert list(got) == [f[0] for f in ref_flags]
    for flag, kind, default, action in ref_flags:
        okind, odefault = got[flag]
        if action == "store_true":
            assert okind is None and odefault is False
        else:
            assert okind is kinds[kind]
            if flag != "--audioreactive_file":
                assert odefault == default, flag
    args = GA.build_parser().parse_args(["--ckpt", "g.pt", "--audio_file", "a.wav", "--shuffle_latents", "--fps", "24"])
    assert args.ckpt == "g.pt" and args.shuffle_latents and args.fps == 24 and args.G_res == 1024


def test_hook_discovery_and_override(tmp_path):
    """Hooks are found by name in --audioreactive_file, missing ones reported and left None, OVERRIDE returned
    (generate_audiovisual.py:262-292)."""
    from maua_stylegan2_b200 import generate_audiovisual as GA

    f = tmp_path / "hooks.py"
    f.write_text("OVERRIDE = dict(fps=24, out_size=1920)\n"
                 "def initialize(args):\n    args.flag = 1\n    return args\n"
                 "def get_truncation(args):\n    return 0.7\n")
    funcs, override = GA.load_hooks(str(f))
    assert set(funcs) == set(GA.HOOKS)
    assert funcs["initialize"] is not None and funcs["get_truncation"](None) == 0.7
    assert funcs["get_latents"] is None and funcs["get_noise"] is None and funcs["get_bends"] is None
    assert override == {"fps": 24, "out_size": 1920}
    default_funcs, default_override = GA.load_hooks(GA.DEFAULT_HOOK_FILE)
    assert all(default_funcs[k] is not None for k in ("initialize", "get_latents", "get_noise")) and default_override == {}


def test_generator_surface_matches_reference():
    """Generator constructor / forward parameters and the full state_dict layout (every key and shape, including the
    noise buffers and FIR kernels) equal the reference's for config-f 1024, config-e 512 --noconst and a 1920-wide
    output (goldens from the reference classes themselves)."""
    import inspect
    import json

    from maua_stylegan2_b200.stylegan2 import Generator

    def sig(fn):
        return [[n, None if p.default is inspect.Parameter.empty else repr(p.default)]
                for n, p in inspect.signature(fn).parameters.items()]

    for key, fn, extra in (("generator_init_signature", Generator.__init__, ["impl", "precision"]),
                           ("generator_forward_signature", Generator.forward, ["return_u8"])):
        ref = json.loads(str(G[key]))
        ours = sig(fn)
        assert ours[:len(ref)] == ref, key
        assert [n for n, _ in ours[len(ref):]] == extra
    for tag, kw in (("1024_cm2_const", dict(size=1024, channel_multiplier=2, constant_input=True)),
                    ("512_cm1_noconst", dict(size=512, channel_multiplier=1, constant_input=False)),
                    ("256_cm2_1920", dict(size=256, channel_multiplier=2, constant_input=True, output_size=1920))):
        ref = json.loads(str(G[f"state_dict_{tag}"]))
        g = Generator(kw.pop("size"), 512, 8, **kw)
        ours = {k: list(v.shape) for k, v in g.state_dict().items()}
        assert ours == ref, (tag, set(ours) ^ set(ref))
        assert [g.n_latent, g.num_layers, g.log_size] == list(G[f"layout_{tag}"])


def test_helper_op_and_render_signatures_match_reference():
    """Every public helper of `audioreactive`, the `op` functions and `render.render` take the reference's parameters
    (names, order, defaults); ours may only append keyword arguments (`dtype`, `sink`)."""
    import inspect
    import json
    import re

    from maua_stylegan2_b200 import audioreactive as ar
    from maua_stylegan2_b200 import op, render

    def sig(fn):
        return [[n, None if p.default is inspect.Parameter.empty else re.sub(r" at 0x[0-9a-f]+", "", repr(p.default))]
                for n, p in inspect.signature(fn).parameters.items()]

    api = json.loads(str(G["api_signatures"]))
    assert len(api) == 30
    allowed_extra = {"ar.perlin_noise": ["dtype"], "render.render": ["sink"]}
    for key, ref in api.items():
        ref = [[n, None if d is None else re.sub(r" at 0x[0-9a-f]+", "", d)] for n, d in ref]
        mod, name = key.split(".")
        obj = getattr({"ar": ar, "op": op, "render": render}[mod], name)
        ours = sig(obj.__init__ if inspect.isclass(obj) else obj)
        assert ours[:len(ref)] == ref, key
        assert [n for n, _ in ours[len(ref):]] == allowed_extra.get(key, []), key


def test_dynamics_helpers_match_reference_on_host_tensors():
    """normalize / compress / expand / percentile are plain tensor expressions in the product too (no kernel), so they
    are checked HERE, on CPU tensors, against values the reference functions returned (audio_glue.npz)."""
    from maua_stylegan2_b200.audioreactive import signal as S

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_glue.npz"))
    x = torch.from_numpy(g["dyn_x"])
    np.testing.assert_allclose(S.compress(x.clone(), 0.6, 0.25).numpy(), g["dyn_compress"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(S.compress(x.clone(), 0.3, 0.5, invert=True).numpy(), g["dyn_compress_inv"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(S.expand(x.clone(), 0.8, 10).numpy(), g["dyn_expand"], rtol=1e-6, atol=1e-7)
    got = np.array([S.percentile(x, p) for p in (0, 10, 50, 97, 100)], np.float32)
    assert np.array_equal(got, g["dyn_percentiles"])
    np.testing.assert_allclose(S.normalize(torch.from_numpy(g["norm_x"]).clone()).numpy(), g["norm_y"], rtol=1e-6, atol=1e-7)
