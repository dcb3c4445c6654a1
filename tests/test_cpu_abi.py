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
    assert L.lib().maua_abi_version() == 1
    assert L.launch_count() == 0  # nothing ran on this GPU-less box


def test_struct_layouts_match_c():
    from maua_stylegan2_b200 import _lib as L

    assert ctypes.sizeof(L.StyleJob) == 56
    assert ctypes.sizeof(L.ConvEpilogue) == 128


def test_cpu_tensors_fail_loudly():
    import torch

    from maua_stylegan2_b200 import _lib as L
    from maua_stylegan2_b200 import op

    with pytest.raises(L.MauaError):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    with pytest.raises(L.MauaError):
        op.fused_leaky_relu(torch.zeros(2, 3), torch.zeros(3))


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
