"""Audio-chain measurement: the default hooks of generate() — initialize (2 onset envelopes),
get_latents (chroma-weighted latents + onset blends), get_noise (13 maps <= 256 px: gaussian-filtered noise, sigma 5 / 128,
onset cross-fades) — on synthetic audio, device time (CUDA events around each hook, after one warm-up run) next to the numpy
restatement (oracle/audio_oracle.py) on the host.  Used by bench.py (`audio_chain`) and runnable alone:

    python tools/bench_audio.py [--seconds 30] [--no-oracle]
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FPS, SR = 30, 44100


def _events():
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def measure_device(seconds, reps=2, out_size=1024, g_res=1024):
    """ms per hook on the current CUDA device; audio = N(0,1)*0.1 at 44.1 kHz (BASELINE configs[1] / configs[4])."""
    import numpy as np
    import torch

    from maua_stylegan2_b200 import audioreactive as ar
    from maua_stylegan2_b200.generate_audiovisual import DEFAULT_HOOK_FILE, get_noise_range, load_hooks

    hooks, _ = load_hooks(DEFAULT_HOOK_FILE)
    rng = np.random.Generator(np.random.PCG64(0))
    audio = (rng.standard_normal(int(seconds * SR)) * 0.1).astype(np.float32)
    n_frames = int(round(seconds * FPS))
    selection = torch.from_numpy(rng.standard_normal((12, 18, 512)).astype(np.float32)).cuda()
    ar.set_SMF(1)
    first, last, to_log2 = get_noise_range(out_size, g_res, False)
    out = {}
    for rep in range(reps + 1):   # rep 0 = warm-up (cuFFT plans, filterbanks, allocator)
        args = types.SimpleNamespace(audio=audio, sr=SR, n_frames=n_frames, duration=seconds, fps=FPS)
        t = {}
        torch.cuda.synchronize()
        e0, e1 = _events()
        w0 = time.perf_counter()
        e0.record()
        args = hooks["initialize"](args)
        e1.record()
        torch.cuda.synchronize()
        t["initialize (2 onset envelopes)"] = (e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
        e0, e1 = _events()
        w0 = time.perf_counter()
        e0.record()
        latents = hooks["get_latents"](selection=selection, args=args)
        e1.record()
        torch.cuda.synchronize()
        t["get_latents (chroma + blends)"] = (e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
        e0, e1 = _events()
        w0 = time.perf_counter()
        e0.record()
        n_maps, n_bytes = 0, 0
        for scale in range(first, last):
            edge = 2 ** to_log2(scale)
            m = hooks["get_noise"](height=edge, width=edge, scale=scale - first, num_scales=last - first, args=args)
            if m is not None:
                n_maps += 1
                n_bytes += m.numel() * 4
            del m
        e1.record()
        torch.cuda.synchronize()
        t[f"get_noise ({n_maps} maps <= 256 px, {n_bytes / 1e9:.2f} GB)"] = (e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
        if rep > 0:
            for k, (dev_ms, wall_ms) in t.items():
                a = out.setdefault(k, [0.0, 0.0])
                a[0] += dev_ms / reps
                a[1] += wall_ms / reps
        assert tuple(latents.shape) == (n_frames, 18, 512)
        del latents, args
        torch.cuda.empty_cache()
    return {k: {"device_ms": round(v[0], 3), "wall_ms": round(v[1], 3)} for k, v in out.items()}, n_frames


def measure_oracle(seconds, noise_max_width=16):
    """Host ms of the numpy restatement: same features for the whole clip; the noise maps only up to `noise_max_width` px
    (the sigma-128 filter is a 1025-tap loop in numpy: the larger maps would take minutes) — a bounded sample, stated."""
    import numpy as np

    from oracle import audio_oracle as A

    rng = np.random.Generator(np.random.PCG64(0))
    audio = (rng.standard_normal(int(seconds * SR)) * 0.1).astype(np.float32)
    n_frames = int(round(seconds * FPS))
    sel = rng.standard_normal((12, 18, 512)).astype(np.float32)
    out = {}
    t0 = time.perf_counter()
    lo = A.onsets(audio, SR, n_frames, fmax=150, smooth=5, clip=97, power=2, type="mm")
    hi = A.onsets(audio, SR, n_frames, fmin=500, smooth=5, clip=99, power=2, type="mm")
    out["initialize (2 onset envelopes)"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ch = A.chroma(audio, SR, n_frames)
    A.default_get_latents(sel, ch, lo, hi)
    out["get_latents (chroma + blends)"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    n_elems = 0
    w = 4
    while w <= noise_max_width:
        for _ in range(1 if w == 4 else 2):
            x = rng.standard_normal((n_frames, 1, w, w)).astype(np.float32)
            A.gaussian_filter(x, 5)
            A.gaussian_filter(x, 128)
            n_elems += x.size
        w *= 2
    out[f"get_noise SAMPLE (maps <= {noise_max_width} px only: {n_elems * 4 / 1e6:.1f} MB of noise)"] = (time.perf_counter() - t0) * 1e3
    return {k: round(v, 1) for k, v in out.items()}


def measure(seconds=30, oracle=True):
    dev, n_frames = measure_device(seconds)
    res = {"audio": f"{seconds} s synthetic N(0,1)*0.1 @ {SR} Hz -> {n_frames} frames @ {FPS} fps, default hooks "
                    "(audioreactive/examples/default.py)", "device": dev,
           "device_total_ms": round(sum(v["device_ms"] for v in dev.values()), 2)}
    if oracle:
        res["numpy_oracle_host_ms"] = measure_oracle(seconds)
        res["oracle_note"] = "numpy/scipy restatement (librosa/madmom absent: parity unpinned), single process"
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=30)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    print(json.dumps(measure(a.seconds, not a.no_oracle)))
