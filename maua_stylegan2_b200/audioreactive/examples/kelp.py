"""Loop-based interpolation with looping Perlin noise (device-path counterpart of the reference's
audioreactive/examples/kelp.py).

Sections of the track get their own spline loop through four latents (two sets: calm / drop, blended by the RMS
envelope); noise is tileable-in-time Perlin noise that loops every two bars, a high-frequency field faded in by the RMS.
Like the reference, the sections come from `ar.laplacian_segmentation` (audioreactive/segmentation.py: device STFT /
constant-Q / MFCC front-end, host spectral clustering) unless `args.sections = (timestamps, labels)` is supplied; the
reference takes the drum envelopes from private multitrack stems, here they come from the mix through band-limited
onsets."""
import os

import numpy as np
import torch as th

from maua_stylegan2_b200 import audioreactive as ar

OVERRIDE = dict(audio_file="audioreactive/examples/Wavefunk - Dwelling in the Kelp.mp3", out_size=1920)
BPM = 130
DROP_LATENTS = "workspace/cyphept_kelp_drop_latents.npy"
COLOR_LAYER = 9


def initialize(args):
    rms = ar.rms(args.audio, args.sr, args.n_frames, smooth=10, clip=60, power=1)
    rms = ar.expand(rms, threshold=0.8, ratio=10)
    rms = ar.gaussian_filter(rms, 4)
    args.rms = ar.normalize(rms)
    args.kick_onsets = ar.onsets(args.audio, args.sr, args.n_frames, margin=1, fmax=150, smooth=4)
    args.snare_onsets = ar.onsets(args.audio, args.sr, args.n_frames, margin=1, fmin=500, smooth=4)
    return args


def sections(args):
    """(timestamps incl. the end of the clip, labels): `args.sections` if the caller supplies them, else the reference's
    `ar.laplacian_segmentation(args.audio, args.sr, k=7)` (kelp.py:44-47), else a fixed 16-bar grid (clips too short to
    segment)."""
    if getattr(args, "sections", None) is not None:
        return args.sections
    stamps, labels = ar.laplacian_segmentation(args.audio, args.sr, k=7)
    stamps = [t for t in stamps if t < args.duration]
    if len(stamps) >= 2:
        return stamps + [args.duration], labels[:len(stamps)]
    bar = 4 * 60 / BPM
    stamps = list(np.arange(0, args.duration, 16 * bar)) + [args.duration]
    return stamps, [i % 7 for i in range(len(stamps) - 1)]


def get_latents(selection, args):
    selection = selection.to("cuda", th.float32)
    drop_selection = (ar.load_latents(DROP_LATENTS).to("cuda", th.float32) if os.path.exists(DROP_LATENTS)
                      else selection.flip(0))
    rms = args.rms.cuda()[:, None, None]
    stamps, labels = sections(args)
    pieces = []
    for start, stop, label in zip(stamps, stamps[1:], labels):
        f0 = int(round(start / args.duration * args.n_frames))
        f1 = int(round(stop / args.duration * args.n_frames))
        if f1 <= f0:
            continue
        bars = (stop - start) * (BPM / 60) / 4

        def loop(latents, n_loops):
            seq = ar.spline_loops(ar.wrapping_slice(latents, label % len(latents), 4), n_frames=f1 - f0, n_loops=n_loops)
            seq[:, COLOR_LAYER:] = latents[[label % len(latents)], COLOR_LAYER:]
            return seq

        calm, drop = loop(selection, bars / 4), loop(drop_selection, bars / 2)
        pieces.append((1 - rms[f0:f1]) * calm + rms[f0:f1] * drop)
    have = sum(len(p) for p in pieces)
    if have < args.n_frames:
        pieces.append(pieces[-1][[-1]].expand(args.n_frames - have, -1, -1))
    latents = ar.gaussian_filter(th.cat(pieces)[:args.n_frames].float().contiguous(), 3)
    kick = 0.666 * args.kick_onsets.cuda()[:, None, None]
    snare = 0.666 * args.snare_onsets.cuda()[:, None, None]
    latents = kick * selection[[2]] + (1 - kick) * latents
    latents = snare * selection[[1]] + (1 - snare) * latents
    return ar.gaussian_filter(latents.contiguous(), 1, causal=0.2)


def get_noise(height, width, scale, num_scales, args):
    if width > 512:
        return None
    num_bars = max(int(round(args.duration * (BPM / 60) / 4)), 2)
    loop_frames = int(args.n_frames / num_bars * 2)  # one loop = two bars

    def looped(res):
        frames = max(loop_frames - loop_frames % res[0], res[0])   # perlin_noise needs shape % res == 0
        field = ar.perlin_noise(shape=(frames, height, width), res=res, dtype=th.float32)[:, None]
        field = field.repeat(-(-args.n_frames // frames), 1, 1, 1)[:args.n_frames]
        return field

    smooth, busy = looped((1, 1, 1)), looped((8, 4, 4))
    rms = args.rms.cuda()[:, None, None, None]
    return rms * busy + (1 - rms) * smooth


def get_bends(args):
    widen = th.nn.Sequential(th.nn.ReplicationPad2d((2, 2, 0, 0)),
                             ar.AddNoise(0.025 * th.randn(size=(1, 1, 4, 8), device="cuda")))
    return [{"layer": 0, "transform": widen}]
