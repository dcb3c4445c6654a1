"""The reference's own audio glue (pure torch/scipy; runs without librosa) pins the oracle restatement:
gaussian_filter / percentile_clip / chroma_weight_latents — fixtures from tests/golden/make_golden.py."""
import os

import numpy as np

from oracle import audio_oracle as A
from tests.util import GOLDEN


def test_gaussian_filter_matches_reference():
    g = np.load(os.path.join(GOLDEN, "audio_glue.npz"))
    cases = [("gf_x1", "gf_y1_s5_c0", 5, 0), ("gf_x1", "gf_y1_s3", 3, None), ("gf_x3", "gf_y3_s4", 4, None),
             ("gf_x3", "gf_y3_s2_c02", 2, 0.2), ("gf_x4", "gf_y4_s5", 5, None), ("gf_x4", "gf_y4_s128", 128, None)]
    for xk, yk, sigma, causal in cases:
        out = A.gaussian_filter(g[xk], sigma, causal=causal)
        assert out.shape == g[yk].shape
        np.testing.assert_allclose(out, g[yk], rtol=0, atol=3e-6, err_msg=yk)


def test_percentile_clip_and_chroma_weight_match_reference():
    g = np.load(os.path.join(GOLDEN, "audio_glue.npz"))
    np.testing.assert_allclose(A.percentile_clip(g["pc_x"], 97), g["pc_y97"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(A.percentile_clip(g["pc_x"], 50), g["pc_y50"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(A.chroma_weight_latents(g["cw_chroma"], g["cw_sel"]), g["cw_y"], rtol=1e-5, atol=1e-6)


def test_feature_chain_sanity():
    """Unpinned internals: shape / range / invariance properties on 3 s of synthetic audio with a click track."""
    sr = 22050
    rng = np.random.default_rng(0)
    y = 0.01 * rng.standard_normal(3 * sr)
    clicks = np.arange(0.25, 3, 0.5)
    for c in clicks:
        i = int(c * sr)
        y[i:i + 200] += np.hanning(200) * rng.standard_normal(200)
    n_frames = 90
    o = A.onsets(y, sr, n_frames, fmin=500, smooth=1, clip=99, power=2)
    assert o.shape == (n_frames,) and o.min() >= 0 and abs(o.max() - 1) < 1e-6
    peaks = [int(np.argmax(o[max(0, int(c * 30) - 4):int(c * 30) + 5])) + max(0, int(c * 30) - 4) for c in clicks]
    assert all(o[p] > 0.2 for p in peaks), "every click must show up in the onset envelope"
    ch = A.chroma(np.sin(2 * np.pi * 440 * np.arange(3 * sr) / sr) + 0.001 * rng.standard_normal(3 * sr), sr, n_frames)
    assert ch.shape == (n_frames, 12)
    np.testing.assert_allclose(ch.sum(1), 1.0, atol=1e-5)
    S = A.stft(y)
    np.testing.assert_allclose(A.istft(S, len(y)), y, atol=1e-9)          # STFT/ISTFT round trip
    assert A.mel_filterbank(sr, fmin=20, fmax=8000).shape == (128, 1025)
    assert int(np.argmax(A.chroma_filterbank(sr)[:, round(440 / sr * 2048)])) == 9  # A -> pitch class 9 (base C)
