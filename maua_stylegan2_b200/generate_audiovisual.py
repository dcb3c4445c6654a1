"""Drop-in for the reference's `generate_audiovisual.py` (plugin surface kept 1:1).

    python -m maua_stylegan2_b200.generate_audiovisual --ckpt G.pt --audio_file track.wav [--audioreactive_file hooks.py]

`generate()` takes the same arguments as the reference (generate_audiovisual.py:59-91) and calls the same six hooks
— initialize(args) / get_latents(selection, args) / get_noise(height, width, scale, num_scales, args) /
get_bends(args) / get_rewrites(args) / get_truncation(args) — discovered by name in `--audioreactive_file`, with the
`OVERRIDE` dict applied (:266-292).  `args` carries every flag plus `audio, sr, n_frames, duration`.

What differs: audio features, latents and noise are produced and kept ON DEVICE (no `.cpu()` / pin / re-upload per batch),
the generator is the B200 `Generator`, the frame loop is `render.FramePipeline`, and with `torchrun` frames are sharded
over the GPUs of the box (each rank delivers its own uint8 frames through a shared pinned host ring) instead of
`th.nn.DataParallel`.
Extra keyword arguments: `audio=(array, sr)` bypasses `load_audio` (synthetic audio), `sink=` replaces ffmpeg.
"""
import argparse
import gc
import importlib.util
import os
import random
import time
import traceback
import uuid

import numpy as np
import torch as th

from . import audioreactive as ar
from . import render
from .stylegan2 import Generator

HOOKS = ["initialize", "get_latents", "get_noise", "get_bends", "get_rewrites", "get_truncation"]
DEFAULT_HOOK_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "audioreactive", "examples", "default.py")


def get_noise_range(out_size, generator_resolution, is_stylegan1):
    """generate_audiovisual.py:22-34"""
    log_max_res = int(np.log2(out_size))
    log_min_res = 2 + (log_max_res - int(np.log2(generator_resolution)))
    if is_stylegan1:
        return log_min_res, log_max_res + 1, (lambda x: x)
    return 2 * log_min_res + 1, 2 * (log_max_res + 1), (lambda x: int(x / 2))


def load_generator(ckpt, is_stylegan1, G_res, out_size, noconst, latent_dim, n_mlp, channel_multiplier, dataparallel,
                   base_res_factor):
    """generate_audiovisual.py:37-56 (StyleGAN1 and DataParallel are not part of the B200 path)."""
    if is_stylegan1:
        raise NotImplementedError("StyleGAN1 (models/stylegan1.py) is out of scope for the B200 path (SURVEY.md §2 #16)")
    return Generator(G_res, latent_dim, n_mlp, channel_multiplier=channel_multiplier, constant_input=not noconst,
                     checkpoint=ckpt, output_size=out_size, base_res_factor=base_res_factor).cuda().eval()


def load_hooks(audioreactive_file):
    """(funcs, OVERRIDE) from a hook file path or dotted module path (generate_audiovisual.py:262-292)."""
    path = audioreactive_file
    if not os.path.exists(path) and os.path.exists(path.replace(".", "/") + ".py"):
        path = path.replace(".", "/") + ".py"
    spec = importlib.util.spec_from_file_location("maua_hooks_" + uuid.uuid4().hex[:6], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    funcs = {}
    for name in HOOKS:
        funcs[name] = getattr(mod, name, None)
        if funcs[name] is None:
            print(f"No '{name}' function found in --audioreactive_file, using default...")
    return funcs, dict(getattr(mod, "OVERRIDE", {}))


def _namespace(values, skip):
    """`args` as the reference builds it: every keyword of generate() as an attribute (generate_audiovisual.py:93-98)."""
    ns = argparse.Namespace()
    for key, value in values.items():
        if key not in skip:
            setattr(ns, key, value)
    return ns


def _audio_and_frames(args, audio, audio_file, offset, duration, fps):
    """Decode (or take) the audio, fix duration / n_frames and publish them on `args` (:103-108)."""
    if audio is None:
        samples, sr, duration = ar.load_audio(audio_file, offset, duration)
    else:
        samples, sr = audio
        if duration == -1:
            duration = len(samples) / sr
    args.audio, args.sr, args.duration = samples, sr, duration
    args.n_frames = int(round(duration * fps))
    return duration


def _noise_maps(get_noise, args, out_size, G_res, stylegan1):
    """One get_noise call per noise input of the generator, coarse to fine (:149-157); 1920 / 1080 outputs double the
    width / height of every map."""
    first, last, to_log2 = get_noise_range(out_size, G_res, stylegan1)
    tall, wide = (2 if out_size == 1080 else 1), (2 if out_size == 1920 else 1)
    maps = []
    for scale in range(first, last):
        edge = 2 ** to_log2(scale)
        m = get_noise(height=tall * edge, width=wide * edge, scale=scale - first, num_scales=last - first, args=args)
        if m is not None:
            print(list(m.shape), f"amplitude={m.std()}")
        maps.append(m)
    return maps


def generate(ckpt, audio_file, initialize=None, get_latents=None, get_noise=None, get_bends=None, get_rewrites=None,
             get_truncation=None, output_dir="./output", audioreactive_file=DEFAULT_HOOK_FILE, offset=0, duration=-1,
             latent_file=None, shuffle_latents=False, G_res=1024, out_size=1024, fps=30, latent_count=12, batch=8,
             dataparallel=False, truncation=1.0, stylegan1=False, noconst=False, latent_dim=512, n_mlp=8,
             channel_multiplier=2, randomize_noise=False, ffmpeg_preset="slow", base_res_factor=1, output_file=None,
             args=None, audio=None, sink=None, generator=None, latent_selection=None):
    started = time.time()
    # Multi-GPU (the reference's `--dataparallel`, generate_audiovisual.py:54-55,249, is single-process nn.DataParallel):
    # here one process per GPU under torchrun; every rank runs generate(), frames are sharded by render.render and only
    # rank 0 writes the video.  `dataparallel=True` without torchrun has a single rank and renders on one GPU.
    from . import parallel

    rank, world, local_rank = parallel.init_from_env()
    if world > 1:
        th.cuda.set_device(local_rank)
    elif dataparallel and th.cuda.device_count() > 1:
        print("--dataparallel: launch with `torchrun --nproc-per-node N -m maua_stylegan2_b200.generate_audiovisual ...` "
              "to shard the frames over N GPUs; this process renders on one GPU.")
    if args is None:
        args = _namespace(dict(locals()), skip=("audio", "sink", "generator", "latent_selection", "started"))
    ar.set_SMF(fps / 30)  # smoothing independent of the frame rate (:101)
    th.set_grad_enabled(False)
    duration = _audio_and_frames(args, audio, audio_file, offset, duration, fps)

    # hooks that were not supplied fall back to the shipped default hook file
    fallback = None
    if get_latents is None or get_noise is None:
        fallback, _ = load_hooks(DEFAULT_HOOK_FILE)
        if initialize is None and get_latents is None:
            initialize = fallback["initialize"]
        get_latents = get_latents or fallback["get_latents"]
        get_noise = get_noise or fallback["get_noise"]
    if initialize is not None:
        args = initialize(args)

    if latent_selection is None:
        latent_selection = (ar.load_latents(latent_file) if latent_file is not None else
                            ar.generate_latents(latent_count, ckpt, G_res, noconst, latent_dim, n_mlp, channel_multiplier))
    if shuffle_latents:
        latent_selection = latent_selection[random.sample(range(len(latent_selection)), len(latent_selection))]
    latents = get_latents(selection=latent_selection, args=args)
    print(f"{list(latents.shape)} amplitude={latents.std()}\n")

    noise = _noise_maps(get_noise, args, out_size, G_res, stylegan1)
    gc.collect()
    bends = [] if get_bends is None else get_bends(args=args)
    rewrites = {} if get_rewrites is None else get_rewrites(args=args)
    truncation = float(truncation) if get_truncation is None else get_truncation(args=args)

    if generator is None:
        generator = load_generator(ckpt, stylegan1, G_res, out_size, noconst, latent_dim, n_mlp, channel_multiplier,
                                   dataparallel, base_res_factor)
    print(f"\npreprocessing took {time.time() - started:.2f}s\n")
    print(f"rendering {args.n_frames} frames...")
    if output_file is None and sink is None and rank == 0:
        os.makedirs(output_dir, exist_ok=True)
        stem = lambda path: str(path).split("/")[-1].split(".")[0].lower()
        output_file = f"{output_dir}/{stem(audio_file)}_{stem(ckpt)}_{uuid.uuid4().hex[:8]}.mp4"
    pipe = render.render(generator=generator, latents=latents, noise=noise, audio_file=audio_file, offset=offset,
                         duration=duration, batch_size=batch, truncation=truncation, bends=bends, rewrites=rewrites,
                         out_size=out_size, output_file=output_file, randomize_noise=randomize_noise,
                         ffmpeg_preset=ffmpeg_preset, sink=sink)
    print(f"\ntotal time taken: {(time.time() - started) / 60:.2f} minutes")
    return pipe


# CLI flags 1:1 with the reference (generate_audiovisual.py:235-260): (flag, type or None for a switch, default)
CLI_FLAGS = [
    ("ckpt", str, None), ("audio_file", str, None), ("audioreactive_file", str, DEFAULT_HOOK_FILE),
    ("output_dir", str, "./output"), ("offset", float, 0), ("duration", float, -1), ("latent_file", str, None),
    ("shuffle_latents", None, False), ("G_res", int, 1024), ("out_size", int, 1024), ("fps", int, 30),
    ("latent_count", int, 12), ("batch", int, 8), ("dataparallel", None, False), ("truncation", float, 1.0),
    ("stylegan1", None, False), ("noconst", None, False), ("latent_dim", int, 512), ("n_mlp", int, 8),
    ("channel_multiplier", int, 2), ("randomize_noise", None, False), ("base_res_factor", float, 1),
    ("ffmpeg_preset", str, "slow"), ("output_file", str, None),
]


def build_parser():
    parser = argparse.ArgumentParser(description="audio-reactive StyleGAN2 video on the B200 path")
    for flag, kind, default in CLI_FLAGS:
        if kind is None:
            parser.add_argument(f"--{flag}", action="store_true")
        else:
            parser.add_argument(f"--{flag}", type=kind, default=default)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    os.makedirs(args.output_dir, exist_ok=True)
    try:
        hooks, override = load_hooks(args.audioreactive_file)
    except Exception:
        print("Error while loading --audioreactive_file...")
        traceback.print_exc()
        raise SystemExit(1)
    for key, value in override.items():  # the hook file's OVERRIDE dict wins over the command line (:284-292)
        setattr(args, key, value)
    kwargs = vars(args).copy()
    generate(ckpt=kwargs.pop("ckpt", None), audio_file=kwargs.pop("audio_file", None), **hooks, **kwargs, args=args)


if __name__ == "__main__":
    main()
