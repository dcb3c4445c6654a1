"""Drop-in for the reference's `generate_audiovisual.py` (plugin surface kept 1:1).

    python -m maua_stylegan2_b200.generate_audiovisual --ckpt G.pt --audio_file track.wav [--audioreactive_file hooks.py]

`generate()` takes the same arguments as the reference (generate_audiovisual.py:59-91) and calls the same six hooks
— initialize(args) / get_latents(selection, args) / get_noise(height, width, scale, num_scales, args) /
get_bends(args) / get_rewrites(args) / get_truncation(args) — discovered by name in `--audioreactive_file`, with the
`OVERRIDE` dict applied (:266-292).  `args` carries every flag plus `audio, sr, n_frames, duration`.

What differs: audio features, latents and noise are produced and kept ON DEVICE (no `.cpu()` / pin / re-upload per batch),
the generator is the B200 `Generator`, the frame loop is `render.FramePipeline`, and with `torchrun` frames are sharded
over the GPUs of the box (one all-gather of uint8 frames per step) instead of `th.nn.DataParallel`.
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


def generate(ckpt, audio_file, initialize=None, get_latents=None, get_noise=None, get_bends=None, get_rewrites=None,
             get_truncation=None, output_dir="./output", audioreactive_file=DEFAULT_HOOK_FILE, offset=0, duration=-1,
             latent_file=None, shuffle_latents=False, G_res=1024, out_size=1024, fps=30, latent_count=12, batch=8,
             dataparallel=False, truncation=1.0, stylegan1=False, noconst=False, latent_dim=512, n_mlp=8,
             channel_multiplier=2, randomize_noise=False, ffmpeg_preset="slow", base_res_factor=1, output_file=None,
             args=None, audio=None, sink=None, generator=None, latent_selection=None):
    if args is None:
        kwargs = dict(locals())
        args = argparse.Namespace()
        for k, v in kwargs.items():
            if k not in ("audio", "sink", "generator", "latent_selection", "kwargs"):
                setattr(args, k, v)

    ar.set_SMF(fps / 30)  # smoothing independent of frame rate (:101)
    time_taken = time.time()
    th.set_grad_enabled(False)

    if audio is not None:
        audio_arr, sr = audio
        duration = len(audio_arr) / sr if duration == -1 else duration
    else:
        audio_arr, sr, duration = ar.load_audio(audio_file, offset, duration)
    args.audio, args.sr = audio_arr, sr
    n_frames = int(round(duration * fps))
    args.duration, args.n_frames = duration, n_frames

    default_funcs = None
    if get_latents is None or get_noise is None:
        default_funcs, _ = load_hooks(DEFAULT_HOOK_FILE)
        if initialize is None and get_latents is None:
            initialize = default_funcs["initialize"]
    if initialize is not None:
        args = initialize(args)

    # ---- latents ---------------------------------------------------------------------------------------------------
    if get_latents is None:
        get_latents = default_funcs["get_latents"]
    if latent_selection is None:
        if latent_file is not None:
            latent_selection = ar.load_latents(latent_file)
        else:
            latent_selection = ar.generate_latents(latent_count, ckpt, G_res, noconst, latent_dim, n_mlp, channel_multiplier)
    if shuffle_latents:
        idx = random.sample(range(len(latent_selection)), len(latent_selection))
        latent_selection = latent_selection[idx]
    latents = get_latents(selection=latent_selection, args=args)
    print(f"{list(latents.shape)} amplitude={latents.std()}\n")

    # ---- noise -----------------------------------------------------------------------------------------------------
    if get_noise is None:
        get_noise = default_funcs["get_noise"]
    noise = []
    range_min, range_max, exponent = get_noise_range(out_size, G_res, stylegan1)
    for scale in range(range_min, range_max):
        h = (2 if out_size == 1080 else 1) * 2 ** exponent(scale)
        w = (2 if out_size == 1920 else 1) * 2 ** exponent(scale)
        noise.append(get_noise(height=h, width=w, scale=scale - range_min, num_scales=range_max - range_min, args=args))
        if noise[-1] is not None:
            print(list(noise[-1].shape), f"amplitude={noise[-1].std()}")
    gc.collect()

    bends = get_bends(args=args) if get_bends is not None else []
    rewrites = get_rewrites(args=args) if get_rewrites is not None else {}
    truncation = get_truncation(args=args) if get_truncation is not None else float(truncation)

    if generator is None:
        generator = load_generator(ckpt, stylegan1, G_res, out_size, noconst, latent_dim, n_mlp, channel_multiplier,
                                   dataparallel, base_res_factor)
    print(f"\npreprocessing took {time.time() - time_taken:.2f}s\n")
    print(f"rendering {n_frames} frames...")
    if output_file is None and sink is None:
        os.makedirs(output_dir, exist_ok=True)
        checkpoint_title = str(ckpt).split("/")[-1].split(".")[0].lower()
        track_title = str(audio_file).split("/")[-1].split(".")[0].lower()
        output_file = f"{output_dir}/{track_title}_{checkpoint_title}_{uuid.uuid4().hex[:8]}.mp4"
    pipe = render.render(generator=generator, latents=latents, noise=noise, audio_file=audio_file, offset=offset,
                         duration=duration, batch_size=batch, truncation=truncation, bends=bends, rewrites=rewrites,
                         out_size=out_size, output_file=output_file, randomize_noise=randomize_noise,
                         ffmpeg_preset=ffmpeg_preset, sink=sink)
    print(f"\ntotal time taken: {(time.time() - time_taken) / 60:.2f} minutes")
    return pipe


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--ckpt", type=str)
    parser.add_argument("--audio_file", type=str)
    parser.add_argument("--audioreactive_file", type=str, default=DEFAULT_HOOK_FILE)
    parser.add_argument("--output_dir", type=str, default="./output")
    parser.add_argument("--offset", type=float, default=0)
    parser.add_argument("--duration", type=float, default=-1)
    parser.add_argument("--latent_file", type=str, default=None)
    parser.add_argument("--shuffle_latents", action="store_true")
    parser.add_argument("--G_res", type=int, default=1024)
    parser.add_argument("--out_size", type=int, default=1024)
    parser.add_argument("--fps", type=int, default=30)
    parser.add_argument("--latent_count", type=int, default=12)
    parser.add_argument("--batch", type=int, default=8)
    parser.add_argument("--dataparallel", action="store_true")
    parser.add_argument("--truncation", type=float, default=1.0)
    parser.add_argument("--stylegan1", action="store_true")
    parser.add_argument("--noconst", action="store_true")
    parser.add_argument("--latent_dim", type=int, default=512)
    parser.add_argument("--n_mlp", type=int, default=8)
    parser.add_argument("--channel_multiplier", type=int, default=2)
    parser.add_argument("--randomize_noise", action="store_true")
    parser.add_argument("--base_res_factor", type=float, default=1)
    parser.add_argument("--ffmpeg_preset", type=str, default="slow")
    parser.add_argument("--output_file", type=str, default=None)
    args = parser.parse_args()
    os.makedirs(args.output_dir, exist_ok=True)
    try:
        funcs, override = load_hooks(args.audioreactive_file)
    except Exception:
        print("Error while loading --audioreactive_file...")
        traceback.print_exc()
        raise SystemExit(1)
    arg_dict = vars(args).copy()
    for k, v in override.items():
        arg_dict[k] = v
        setattr(args, k, v)
    ckpt = arg_dict.pop("ckpt", None)
    audio_file = arg_dict.pop("audio_file", None)
    generate(ckpt=ckpt, audio_file=audio_file, **funcs, **arg_dict, args=args)


if __name__ == "__main__":
    main()
