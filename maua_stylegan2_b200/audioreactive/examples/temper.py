"""Spatially masked noise on the device path (behaviour of the reference's audioreactive/examples/temper.py).

Latents: chromagram-weighted selection, pulled towards two fixed latents by the high / low onset envelopes.
Noise: a disc in the middle of every noise map splits it into two zones.  On the coarse scales (< 128 px) the disc
flickers with the low onsets while the surround keeps the slowly drifting field; on the fine scales (> 32 px) the
surround flickers with the high onsets (temper.py:61-85).  Everything is built and left on the GPU."""
import numpy as np
import torch as th

from maua_stylegan2_b200 import audioreactive as ar

OVERRIDE = dict(audio_file="audioreactive/examples/Wavefunk - Temper.mp3", out_size=1024)


def initialize(args):
    common = dict(smooth=5, power=2)
    args.lo_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmax=150, clip=97, **common)
    args.hi_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmin=500, clip=99, **common)
    return args


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    latents = ar.chroma_weight_latents(ar.chroma(args.audio, args.sr, args.n_frames), selection)
    latents = ar.gaussian_filter(latents, 4)
    for envelope, target in ((args.hi_onsets, selection[-4]), (args.lo_onsets, selection[-7])):
        latents = ar.envelope_blend(latents, envelope, target)
    return ar.gaussian_filter(latents, 2, causal=0.2)


def circular_mask(h, w, center=None, radius=None, soft=0):
    """Boolean disc mask [h, w] (temper.py:45-58).  `soft` runs scipy's gaussian over the BOOLEAN mask like the
    reference does, so the result stays boolean (a re-thresholded blur), not a soft edge."""
    import scipy.ndimage as ndi

    cx, cy = (int(w / 2), int(h / 2)) if center is None else center
    if radius is None:
        radius = min(cx, cy, w - cx, h - cy)
    dist = np.hypot(np.arange(w)[None, :] - cx, np.arange(h)[:, None] - cy)
    disc = dist <= radius
    if soft > 0:
        disc = ndi.gaussian_filter(disc, sigma=int(round(soft)))
    return th.from_numpy(disc)


def get_noise(height, width, scale, num_scales, args):
    if width > 256:
        return None
    n = args.n_frames
    lo = args.lo_onsets.cuda().reshape(n, 1, 1, 1)
    hi = args.hi_onsets.cuda().reshape(n, 1, 1, 1)
    inside = circular_mask(height, width, radius=int(width / 2), soft=2).float().cuda().reshape(1, 1, height, width)
    outside = 1 - inside
    fast = ar.gaussian_filter(th.randn((n, 1, height, width), device="cuda"), 5)
    noise = ar.gaussian_filter(th.randn((n, 1, height, width), device="cuda"), 128)
    if width < 128:   # coarse scales: the disc follows the low onsets, the surround keeps the slow field
        noise = 2 * inside * lo * fast + outside * (1 - lo) * noise
    if width > 32:    # fine scales: the surround follows the high onsets
        noise = 0.75 * outside * hi * fast + inside * (1 - 0.75 * hi) * noise
    noise /= noise.std() * 2
    return noise
