"""Default audio-reactive hooks on the device path (behaviour of the reference's audioreactive/examples/default.py:6-45).

Latents follow the chromagram; the high / low onset envelopes pull them towards two fixed latents of the selection.
Noise is a slowly drifting field (gaussian over time, sigma 128) that the onsets cross-fade into a fast one (sigma 5):
low onsets drive the coarse scales (< 128 px), high onsets the fine ones (> 32 px).  All tensors stay on the GPU; the
blends are the `maua_envelope_blend_f32` / `maua_gaussian_filter_f32` kernels."""
import torch as th

from maua_stylegan2_b200 import audioreactive as ar

ONSET_BANDS = {"lo_onsets": dict(fmax=150, clip=97), "hi_onsets": dict(fmin=500, clip=99)}
HI_TARGET, LO_TARGET = -4, -7          # rows of the latent selection the onsets pull towards
SLOW_SIGMA, FAST_SIGMA = 128, 5
MAX_NOISE_WIDTH = 256                  # larger maps fall back to the generator's own noise buffers


def initialize(args):
    for name, band in ONSET_BANDS.items():
        setattr(args, name, ar.onsets(args.audio, args.sr, args.n_frames, smooth=5, power=2, **band))
    return args


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    weights = ar.chroma(args.audio, args.sr, args.n_frames)
    latents = ar.gaussian_filter(ar.chroma_weight_latents(weights, selection), 4)
    for envelope, row in ((args.hi_onsets, HI_TARGET), (args.lo_onsets, LO_TARGET)):
        latents = ar.envelope_blend(latents, envelope, selection[row])
    return ar.gaussian_filter(latents, 2, causal=0.2)


def _smoothed_noise(n_frames, height, width, sigma):
    return ar.gaussian_filter(th.randn((n_frames, 1, height, width), device="cuda"), sigma)


def get_noise(height, width, scale, num_scales, args):
    if width > MAX_NOISE_WIDTH:
        return None
    fast = _smoothed_noise(args.n_frames, height, width, FAST_SIGMA)
    noise = _smoothed_noise(args.n_frames, height, width, SLOW_SIGMA)
    fades = []
    if width < 128:
        fades.append(args.lo_onsets)
    if width > 32:
        fades.append(args.hi_onsets)
    for envelope in fades:
        e = envelope.cuda()[:, None, None, None]
        noise = e * fast + (1 - e) * noise
    noise /= noise.std() * 2.5
    return noise  # stays on the device (the reference returns .cpu() and re-uploads every batch)
