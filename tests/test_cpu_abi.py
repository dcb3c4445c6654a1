"""CPU-side checks: the C-ABI library builds, loads without a GPU and exports every symbol include/maua_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from maua_stylegan2_b200 import build

    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "maua_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(maua_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/maua_b200.h but not exported"


def test_binding_matches_header(lib_path):
    from maua_stylegan2_b200 import _lib as L

    bound = set(L.SIGNATURES) | set(L._SPECIAL)
    assert bound == set(declared_symbols())
    assert L.lib().maua_abi_version() == 2
    assert L.launch_count() == 0  # nothing ran on this GPU-less box


def test_struct_layouts_match_c():
    from maua_stylegan2_b200 import _lib as L

    assert ctypes.sizeof(L.StyleJob) == 64
    assert ctypes.sizeof(L.ConvEpilogue) == 128


def test_cpu_tensor_branch_matches_reference_goldens():
    """The reference dispatches on `input.device.type` and CPU inputs keep working (op/upfirdn2d.py:146-149,
    op/fused_act.py:87-94): the product's CPU-tensor branch is held to the fixtures written by the unmodified reference."""
    import numpy as np
    import torch

    from maua_stylegan2_b200 import op

    g = np.load(os.path.join(ROOT, "tests", "golden", "ops_golden.npz"))
    for name in g["ufd_names"]:
        x, k, cfg, y = g[f"ufd_{name}_x"], g[f"ufd_{name}_k"], g[f"ufd_{name}_cfg"], g[f"ufd_{name}_y"]
        up, down, p0, p1 = [int(v) for v in cfg]
        out = op.upfirdn2d(torch.from_numpy(x), torch.from_numpy(k), up=up, down=down, pad=(p0, p1)).numpy()
        assert out.shape == y.shape, name
        np.testing.assert_allclose(out, y, rtol=0, atol=2e-6 * max(1.0, np.abs(y).max()), err_msg=name)
    for name in ("fl2d", "fl4d", "fl4d_big"):
        out = op.fused_leaky_relu(torch.from_numpy(g[f"{name}_x"]), torch.from_numpy(g[f"{name}_b"])).numpy()
        np.testing.assert_allclose(out, g[f"{name}_y"], rtol=1e-6, atol=1e-7)
    act = op.FusedLeakyReLU(6)
    with torch.no_grad():
        act.bias.copy_(torch.from_numpy(g["fl4d_b"]))
        np.testing.assert_allclose(act(torch.from_numpy(g["fl4d_x"])).numpy(), g["fl4d_y"], rtol=1e-6, atol=1e-7)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """The CUDA branch has no fallback: without libmaua_b200.so the binding raises instead of computing elsewhere."""
    from maua_stylegan2_b200 import _lib as L

    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libmaua_b200.so"))
    with pytest.raises(L.MauaError):
        L.lib()
    with pytest.raises(L.MauaError):
        L.call("maua_upfirdn2d_f32")


def test_state_dict_keys_match_reference_layout():
    """The drop-in must load the reference's g_ema key layout."""
    from oracle import stylegan2_oracle as O
    from maua_stylegan2_b200.stylegan2 import Generator

    g = Generator(32, 512, 8, channel_multiplier=2, constant_input=True, output_size=32)
    keys = set(g.state_dict().keys())
    sd = O.synth_state_dict(32)
    assert set(sd.keys()) <= keys
    extra = keys - set(sd.keys())
    assert all(k.endswith(".kernel") for k in extra), extra
    assert g.n_latent == 8 and g.num_layers == 7


def test_cpu_tensor_branch_randomised_vs_index_spec():
    """The CPU-tensor branch of op.upfirdn2d_raw on random native-ABI configurations (asymmetric rates, minor > 1, negative
    pads) against the oracle's index-level restatement of the reference kernel (fp32 summation order differs: 1e-5)."""
    import numpy as np
    import torch

    from maua_stylegan2_b200 import op
    from oracle import ops_oracle as OO

    rng = np.random.default_rng(2026)
    checked = 0
    for case in range(30):
        major, minor = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        in_h, in_w = int(rng.integers(1, 16)), int(rng.integers(1, 16))
        kh, kw = int(rng.integers(1, 7)), int(rng.integers(1, 6))
        up_x, up_y, down_x, down_y = (int(v) for v in rng.integers(1, 4, 4))
        px0, px1, py0, py1 = (int(v) for v in rng.integers(-2, 5, 4))
        x = rng.standard_normal((major, in_h, in_w, minor)).astype(np.float32)
        k = rng.standard_normal((kh, kw)).astype(np.float32)
        out_h = (in_h * up_y + py0 + py1 - kh + down_y) // down_y
        out_w = (in_w * up_x + px0 + px1 - kw + down_x) // down_x
        if out_h <= 0 or out_w <= 0 or in_h * up_y + py0 + py1 < kh or in_w * up_x + px0 + px1 < kw:
            continue
        got = op.upfirdn2d_raw(torch.from_numpy(x), torch.from_numpy(k), up_x, up_y, down_x, down_y, px0, px1, py0, py1).numpy()
        assert got.shape == (major, out_h, out_w, minor), case
        for m in range(minor):
            spec = OO.upfirdn2d_index(x[..., m], k, up_x, up_y, down_x, down_y, px0, px1, py0, py1, fma=False)
            np.testing.assert_allclose(got[..., m], spec, rtol=0, atol=1e-5 * max(1.0, np.abs(spec).max()), err_msg=str(case))
        checked += 1
    assert checked >= 15
