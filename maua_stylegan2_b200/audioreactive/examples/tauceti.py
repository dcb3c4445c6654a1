"""Network-bending example (device-path counterpart of the reference's audioreactive/examples/tauceti.py).

A 2:1 constant (ReplicationPad on layer 0) renders 2048x1024 frames that the frame loop fits to 1920x1080; during the
"drop" (45 s - 135 s of the original track, scaled to the rendered duration) the upper latent layers step through four
colour latents and a Translate bend on layer 4 scrolls the 16x32 feature map endlessly.  The Translate here is the fused
warp kernel (audioreactive/bend.py) instead of three ReflectionPads + kornia."""
import os

import numpy as np
import torch as th

from maua_stylegan2_b200 import audioreactive as ar

OVERRIDE = dict(audio_file="audioreactive/examples/Wavefunk - Tau Ceti Alpha.mp3", out_size=1920, dataparallel=False,
                fps=30)
REF_FRAMES = 5591  # frame count of the reference render at 30 fps (tauceti.py:18,39)
COLOR_LATENTS = "workspace/cyphept-multicolor-latents.npy"
COLOR_LAYER = 9


def _drop(args):
    return int(REF_FRAMES * (45 / args.duration)), int(REF_FRAMES * (135 / args.duration))


def initialize(args):
    args.low_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmax=150, smooth=5, clip=97, power=2)
    args.high_onsets = ar.onsets(args.audio, args.sr, args.n_frames, fmin=500, smooth=5, clip=99, power=2)
    return args


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    latents = ar.chroma_weight_latents(ar.chroma(args.audio, args.sr, args.n_frames), selection[:12])
    latents = ar.gaussian_filter(latents, 5)
    latents = ar.envelope_blend(latents, args.high_onsets, selection[-4])
    latents = ar.envelope_blend(latents, args.low_onsets, selection[-7])
    latents = ar.gaussian_filter(latents, 5, causal=0)

    start, end = _drop(args)
    start, end = min(start, args.n_frames), min(end, args.n_frames)
    colors = (ar.load_latents(COLOR_LATENTS) if os.path.exists(COLOR_LATENTS) else selection.cpu()).to("cuda", th.float32)
    length = end - start
    section = int(length / 4)
    parts = [latents[:start, COLOR_LAYER:]]
    for i in range(4 if section > 0 else 0):
        parts.append(colors[[i % len(colors)], COLOR_LAYER:].expand(section, -1, -1))
    if length - 4 * section > 0:  # the last colour holds until the drop ends
        parts.append(colors[[3 % len(colors)], COLOR_LAYER:].expand(length - 4 * section, -1, -1))
    parts.append(latents[end:, COLOR_LAYER:])
    latents[:, COLOR_LAYER:] = ar.gaussian_filter(th.cat(parts, 0).contiguous(), 5)
    return latents


def get_noise(height, width, scale, num_scales, args):
    if width > 256:
        return None
    lo = 1.25 * args.low_onsets.cuda()[:, None, None, None]
    hi = 1.25 * args.high_onsets.cuda()[:, None, None, None]
    shape = (args.n_frames, 1, height, width)
    fast = ar.gaussian_filter(th.randn(shape, device="cuda"), 5)
    noise = ar.gaussian_filter(th.randn(shape, device="cuda"), 128)
    if width > 8:
        noise = lo * fast + (1 - lo) * noise
        noise = hi * fast + (1 - hi) * noise
    noise /= noise.std() * 2.5
    return noise


def scroll_modulation(args, width):
    """[n_frames, 2] translation in feature pixels: 0 before the drop, a 6 s sawtooth 0 -> width during it, then held
    (tauceti.py:108-140); the intro -> drop corner is rounded with a sigma-5 gaussian."""
    start, end = _drop(args)
    start, end = min(start, args.n_frames), min(end, args.n_frames)
    period = int(6 * args.fps)
    ramp = np.linspace(0, width, period)
    t = np.arange(max(end - start, 0))
    x = np.zeros(args.n_frames)
    x[start:end] = ramp[t % period]
    if end < args.n_frames:
        x[end:] = ramp[min(((end - start) % period) + 1, period - 1)]
    translation = th.tensor(np.stack([x, np.zeros_like(x)], 1)).float()
    fps = int(args.fps)
    if start - 5 * fps >= 0 and start + 5 * fps <= args.n_frames:
        window = ar.gaussian_filter(translation[start - 5 * fps:start + 5 * fps, 0].contiguous(), 5)
        translation[start - fps:start + fps, 0] = window[4 * fps:-4 * fps].cpu()
    return translation


def get_bends(args):
    widen = th.nn.Sequential(th.nn.ReplicationPad2d((2, 2, 0, 0)),
                             ar.AddNoise(0.025 * th.randn(size=(1, 1, 4, 8), device="cuda")))
    layer = 4
    h = 2 ** layer
    w = 2 * h
    noise = 0.2 * th.randn((1, 1, h, 5 * w), device="cuda")
    return [{"layer": 0, "transform": widen},
            {"layer": layer, "transform": lambda batch: ar.Translate(batch, h, w, noise),
             "modulation": scroll_modulation(args, w)}]
