"""The example hook files (audioreactive/examples/{temper,tauceti,kelp}.py, device path) through the public frame loop:
hook contracts (README.md:91-146 of the reference: shapes of latents / noise / bends), bends applied per batch with their
modulation, non-square 2:1 output."""
import argparse

import numpy as np
import pytest
import torch

from tests.test_gpu_audio import SR, _audio

pytestmark = pytest.mark.gpu


def _args(seconds, fps, seed):
    from maua_stylegan2_b200 import audioreactive as ar

    ar.set_SMF(fps / 30)
    y = _audio(seconds, seed=seed)
    return argparse.Namespace(audio=y, sr=SR, duration=seconds, fps=fps, n_frames=int(round(seconds * fps)),
                              offset=0)


def _render(g, latents, noise, bends, batch):
    from maua_stylegan2_b200.render import FramePipeline

    frames = []
    pipe = FramePipeline(g, latents, noise, batch, truncation=1.0, bends=bends)
    with torch.no_grad():
        pipe.warmup()
        pipe.run(lambda f: frames.append(f.copy()))
    return np.concatenate(frames)


def _noise_list(hooks, args, size, wide):
    out = []
    for l in range(2 * (int(np.log2(size)) - 2) + 1):
        r = 2 ** ((l + 5) // 2)
        out.append(hooks.get_noise(height=r, width=(2 if wide else 1) * r, scale=l, num_scales=17, args=args))
    return out


def test_temper_hooks_render():
    from maua_stylegan2_b200.audioreactive.examples import temper as hooks
    from tests.util import make_generator

    g, _ = make_generator(64, 1, seed=4, impl="tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    args = hooks.initialize(_args(2.0, 15, seed=5))
    sel = torch.randn(12, g.n_latent, 512, generator=torch.Generator().manual_seed(1))
    lat = hooks.get_latents(sel, args)
    assert lat.is_cuda and tuple(lat.shape) == (30, g.n_latent, 512) and bool(torch.isfinite(lat).all())
    noise = _noise_list(hooks, args, 64, wide=False)
    for l, n in enumerate(noise):
        r = 2 ** ((l + 5) // 2)
        assert tuple(n.shape) == (30, 1, r, r) and abs(float(n.std()) - 0.5) < 1e-3      # noise /= std * 2
    m = hooks.circular_mask(16, 16, radius=8, soft=2)
    assert tuple(m.shape) == (16, 16)
    frames = _render(g, lat, noise, [], 8)
    assert frames.shape == (30, 64, 64, 3) and len({f.tobytes() for f in frames}) == 30


def test_tauceti_hooks_bends_scroll_and_widen():
    from maua_stylegan2_b200.audioreactive.examples import tauceti as hooks
    from tests.util import make_generator

    # the drop window is defined on the reference's 5591-frame render: 186.4 s at 30 fps reproduces it 1:1
    full = argparse.Namespace(duration=5591 / 30, fps=30, n_frames=5591)
    tr = hooks.scroll_modulation(full, 32)
    start, end = int(5591 * 45 / full.duration), int(5591 * 135 / full.duration)
    assert tuple(tr.shape) == (5591, 2) and float(tr[:, 1].abs().max()) == 0
    assert float(tr[:start - 30, 0].abs().max()) == 0                      # still before the drop
    x = tr[start + 30:end, 0].numpy()
    ramp = np.linspace(0, 32, 180)
    assert np.allclose(x, ramp[(np.arange(start + 30, end) - start) % 180], atol=1e-5)   # 6 s sawtooth 0 -> w
    assert np.allclose(tr[end:, 0].numpy(), ramp[((end - start) % 180) + 1])               # held after the drop
    assert 0 < float(tr[start, 0]) < ramp[30]                                            # rounded corner

    g, _ = make_generator(64, 1, seed=6, impl="tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    args = hooks.initialize(_args(2.0, 15, seed=7))
    sel = torch.randn(16, g.n_latent, 512, generator=torch.Generator().manual_seed(2))
    lat = hooks.get_latents(sel, args)
    assert tuple(lat.shape) == (30, g.n_latent, 512)
    noise = _noise_list(hooks, args, 64, wide=True)
    assert tuple(noise[0].shape) == (30, 1, 4, 8) and tuple(noise[-1].shape) == (30, 1, 64, 128)
    bends = hooks.get_bends(args)
    assert [b["layer"] for b in bends] == [0, 4] and tuple(bends[1]["modulation"].shape) == (30, 2)
    bends[1]["modulation"][:, 0] = torch.linspace(0, 32, 30)               # make the short clip scroll
    frames = _render(g, lat, noise, bends, 8)
    assert frames.shape == (30, 64, 128, 3) and len({f.tobytes() for f in frames}) == 30


def test_kelp_hooks_perlin_loops():
    from maua_stylegan2_b200.audioreactive.examples import kelp as hooks
    from tests.util import make_generator

    g, _ = make_generator(64, 1, seed=8, impl="tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    args = hooks.initialize(_args(8.0, 15, seed=9))
    args.sections = ([0.0, 3.0, 8.0], [1, 5])
    sel = torch.randn(12, g.n_latent, 512, generator=torch.Generator().manual_seed(3))
    lat = hooks.get_latents(sel, args)
    assert tuple(lat.shape) == (120, g.n_latent, 512) and bool(torch.isfinite(lat).all())
    np.random.seed(0)
    noise = _noise_list(hooks, args, 64, wide=True)
    assert all(tuple(n.shape)[0] == 120 and n.dtype == torch.float32 for n in noise)
    assert all(bool(torch.isfinite(n).all()) for n in noise)   # (perlin * 2 - 1 is not confined to [-1, 1], as in the reference)
    frames = _render(g, lat[:16].contiguous(), [n[:16].contiguous() for n in noise], hooks.get_bends(args), 8)
    assert frames.shape == (16, 64, 128, 3)
    # without caller-supplied sections the hook segments the track itself, like the reference (ar.laplacian_segmentation)
    args.sections = None
    stamps, labels = hooks.sections(args)
    assert stamps[0] == 0.0 and stamps[-1] == args.duration and len(labels) == len(stamps) - 1
    lat2 = hooks.get_latents(sel, args)
    assert tuple(lat2.shape) == (120, g.n_latent, 512) and bool(torch.isfinite(lat2).all())
