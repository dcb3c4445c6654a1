"""Shared helpers for the GPU parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle import stylegan2_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, ref):
    """max|a-ref| / max|ref| — the per-tensor metric of SURVEY.md §7 'Hard parts' #2."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def make_generator(size, cm, seed, impl, precision="bf16x3", device="cuda"):
    from maua_stylegan2_b200.stylegan2 import Generator

    sd = O.synth_state_dict(size, channel_multiplier=cm, seed=seed)
    g = Generator(size, 512, 8, channel_multiplier=cm, constant_input=True, output_size=size, impl=impl,
                  precision=precision)
    missing, unexpected = g.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".kernel" in m for m in missing), missing
    return g.to(device).eval(), sd


def golden_inputs(name):
    g = np.load(os.path.join(GOLDEN, name))
    size = int(g["size"])
    _, num_layers, _ = O.layout(size)
    noise = [torch.from_numpy(g[f"noise_{l}"]) if f"noise_{l}" in g.files else None for l in range(num_layers)]
    return g, noise


def strided(t, n=6):
    c = t.shape[1]
    cs = max(c // n, 1)
    s = max(t.shape[2] // 32, 1)
    return t[:, ::cs, ::s, ::s].contiguous().cpu().numpy()
