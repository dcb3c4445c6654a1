"""Worker of tests/test_gpu_multigpu.py: one rank of a torchrun job that renders a frame sequence through
render.FramePipeline with rank-strided batches + the NCCL all-gather, rank 0 saving the gathered uint8 frames."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_file, size, cm, n_frames, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    mode = sys.argv[6]
    from maua_stylegan2_b200.parallel import AllGatherFrames, init_from_env
    from maua_stylegan2_b200.render import FramePipeline
    from tests.test_gpu_multigpu import make_inputs
    from tests.util import make_generator

    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    g, _ = make_generator(size, cm, 4, "tc", device=f"cuda:{local_rank}")
    g.truncation_latent = torch.zeros(1, 512, device=f"cuda:{local_rank}")
    latents, noise, psi = make_inputs(g, n_frames)
    if rank != 0:   # only rank 0's inputs count: render() must broadcast them (hooks draw rank-local random noise)
        latents, psi = torch.zeros_like(latents), torch.ones_like(psi)
        noise = [None if n is None else torch.zeros_like(n) for n in noise]
    frames = []
    if mode == "render":   # the public API: render.render picks up torch.distributed (all-gather + sharded host ring)
        from maua_stylegan2_b200 import render

        pipe = render.render(g, latents, noise, 0, n_frames / 30.0, batch, size, None, truncation=psi,
                             sink=lambda f: frames.append(f.copy()))
    else:                  # the north-star data path alone: one NCCL all-gather per step, rank 0 copies everything
        from maua_stylegan2_b200 import parallel

        parallel.broadcast_inputs([latents, psi] + [n for n in noise if n is not None], device=f"cuda:{local_rank}")
        pipe = FramePipeline(g, latents, list(noise), batch, truncation=psi, rank=rank, world=world)
        pipe.warmup()
        with torch.no_grad():
            pipe.run((lambda f: frames.append(f.copy())) if rank == 0 else None, AllGatherFrames(world))
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out_file, np.concatenate(frames))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
