"""Spatially masked noise (device-path counterpart of the reference's audioreactive/examples/temper.py).

Latents: chromagram-weighted selection pulled towards two fixed latents by the high / low onset envelopes.
Noise: a soft disc mask decides where the fast (sigma 5) and slow (sigma 128) noise fields react to the drums —
inside the disc the low onsets drive the coarse layers, outside it the high onsets drive the fine layers
(temper.py:45-85).  Everything is built and left on the GPU."""
import numpy as np
import torch as th

from maua_stylegan2_b200 import audioreactive as ar

OVERRIDE = dict(audio_file="audioreactive/examples/Wavefunk - Temper.mp3", out_size=1024)


def initialize(args):
    args.lo_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmax=150, smooth=5, clip=97, power=2)
    args.hi_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmin=500, smooth=5, clip=99, power=2)
    return args


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    latents = ar.chroma_weight_latents(ar.chroma(args.audio, args.sr, args.n_frames), selection)
    latents = ar.gaussian_filter(latents, 4)
    for envelope, target in ((args.hi_onsets, selection[-4]), (args.lo_onsets, selection[-7])):
        latents = ar.envelope_blend(latents, envelope, target)
    return ar.gaussian_filter(latents, 2, causal=0.2)


def circular_mask(h, w, center=None, radius=None, soft=0):
    """Disc of ones (temper.py:45-58), optionally blurred with scipy's gaussian (host, h*w floats, once per scale)."""
    import scipy.ndimage as ndi

    cx, cy = center if center is not None else (int(w / 2), int(h / 2))
    if radius is None:
        radius = min(cx, cy, w - cx, h - cy)
    yy, xx = np.ogrid[:h, :w]
    mask = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) <= radius
    if soft > 0:
        mask = ndi.gaussian_filter(mask, sigma=int(round(soft)))
    return th.from_numpy(mask)


def get_noise(height, width, scale, num_scales, args):
    if width > 256:
        return None
    lo = args.lo_onsets.cuda()[:, None, None, None]
    hi = args.hi_onsets.cuda()[:, None, None, None]
    mask = circular_mask(height, width, radius=int(width / 2), soft=2)[None, None].float().cuda()
    shape = (args.n_frames, 1, height, width)
    fast = ar.gaussian_filter(th.randn(shape, device="cuda"), 5)
    noise = ar.gaussian_filter(th.randn(shape, device="cuda"), 128)
    if width < 128:
        noise = 2 * mask * lo * fast + (1 - mask) * (1 - lo) * noise
    if width > 32:
        noise = 0.75 * (1 - mask) * hi * fast + mask * (1 - 0.75 * hi) * noise
    noise /= noise.std() * 2
    return noise
