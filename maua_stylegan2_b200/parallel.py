"""Multi-GPU frame sharding: one process per GPU (torchrun), rank-strided batches, ONE NCCL all-gather of the finished
uint8 NHWC frames per step over NVLink.  The reference's only multi-GPU render mode is
single-process `th.nn.DataParallel` (generate_audiovisual.py:54-55), which re-broadcasts the 121 MB of weights every
step and gathers fp32 images to GPU 0; here weights are replicated once and 3 B/pixel cross the switch.

The frame path has no other exchange step, so there is no other collective (frames are independent given their
latents / noise / truncation rows)."""
import datetime
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, world, local_rank) from torchrun's environment; initialises the default process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        # a mismatched collective must abort the job, not hang the box
        dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
    return rank, world, local_rank


class AllGatherFrames:
    """gather(frames_u8 [B,H,W,3], slot) -> (work, out_u8 [world*B,H,W,3]); ping-pong output buffers so that the
    collective of step i can overlap the synthesis of step i+1 and the D2H of step i-1."""

    def __init__(self, world, group=None):
        self.world = world
        self.group = group
        self.out = [None, None]

    def __call__(self, frames, slot):
        shape = (self.world * frames.shape[0],) + tuple(frames.shape[1:])
        if self.out[slot] is None or self.out[slot].shape != shape:
            self.out[slot] = torch.empty(shape, dtype=frames.dtype, device=frames.device)
        frames = frames.contiguous()
        if frames.is_cuda:
            work = dist.all_gather_into_tensor(self.out[slot], frames, group=self.group, async_op=True)
        else:  # gloo (CPU tests): list form
            chunks = list(self.out[slot].chunk(self.world, 0))
            work = dist.all_gather(chunks, frames, group=self.group, async_op=True)
        return work, self.out[slot]


def shard_plan(n_frames, batch, world):
    """[(step, rank, first_frame, n_valid)] — which frames each rank renders at each step (host logic, testable on CPU)."""
    starts = list(range(0, n_frames, batch))
    nb = len(starts)
    steps = (nb + world - 1) // world
    plan = []
    for step in range(steps):
        for rank in range(world):
            idx = step * world + rank
            n = starts[min(idx, nb - 1)]
            valid = min(batch, n_frames - n) if idx < nb else 0
            plan.append((step, rank, n, valid))
    return plan
