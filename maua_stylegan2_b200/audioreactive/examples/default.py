"""The reference's default audio-reactive hooks (audioreactive/examples/default.py:6-45) on the device path:
chroma-weighted latents blended towards two fixed latents by low/high onsets; onset-blended gaussian-filtered noise."""
import torch as th

from maua_stylegan2_b200 import audioreactive as ar


def initialize(args):
    args.lo_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmax=150, smooth=5, clip=97, power=2)
    args.hi_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmin=500, smooth=5, clip=99, power=2)
    return args


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    chroma = ar.chroma(args.audio, args.sr, args.n_frames)
    chroma_latents = ar.chroma_weight_latents(chroma, selection)
    latents = ar.gaussian_filter(chroma_latents, 4)
    latents = ar.envelope_blend(latents, args.hi_onsets, selection[-4])
    latents = ar.envelope_blend(latents, args.lo_onsets, selection[-7])
    latents = ar.gaussian_filter(latents, 2, causal=0.2)
    return latents


def get_noise(height, width, scale, num_scales, args):
    if width > 256:
        return None
    lo_onsets = args.lo_onsets[:, None, None, None].cuda()
    hi_onsets = args.hi_onsets[:, None, None, None].cuda()
    noise_noisy = ar.gaussian_filter(th.randn((args.n_frames, 1, height, width), device="cuda"), 5)
    noise = ar.gaussian_filter(th.randn((args.n_frames, 1, height, width), device="cuda"), 128)
    if width < 128:
        noise = lo_onsets * noise_noisy + (1 - lo_onsets) * noise
    if width > 32:
        noise = hi_onsets * noise_noisy + (1 - hi_onsets) * noise
    noise /= noise.std() * 2.5
    return noise  # stays on the device (the reference returns .cpu() and re-uploads every batch)
