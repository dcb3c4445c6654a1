"""World-size-2 gloo test of the multi-GPU host logic: rank-strided batches + one all-gather per
step must deliver every frame exactly once, in order, with the padded tail trimmed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maua_stylegan2_b200.parallel import AllGatherFrames, shard_plan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame(n):
    """A recognisable fake uint8 NHWC frame for global frame index n."""
    return torch.full((4, 6, 3), n % 251, dtype=torch.uint8)


def _worker(rank, world, port, n_frames, batch, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gather = AllGatherFrames(world)
    plan = [p for p in shard_plan(n_frames, batch, world) if p[1] == rank]
    got = []
    for step, _, first, valid in plan:
        idx = [min(first + j, n_frames - 1) for j in range(batch)]           # short tails padded by repetition
        frames = torch.stack([_frame(i) for i in idx])
        work, out = gather(frames, step & 1)
        work.wait()
        n_valid = min(n_frames - step * world * batch, world * batch)
        got.append(out[:n_valid].clone())
    frames = torch.cat(got)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), frames.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plan_covers_every_frame_once():
    for n_frames, batch, world in [(900, 8, 1), (900, 8, 2), (900, 8, 8), (61, 8, 4), (7, 8, 2), (64, 16, 4)]:
        seen = []
        for step, rank, first, valid in shard_plan(n_frames, batch, world):
            seen += list(range(first, first + valid))
        assert seen == list(range(n_frames)), (n_frames, batch, world)


def test_all_gather_frames_world2_gloo(tmp_path):
    n_frames, batch, world = 37, 4, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, batch, str(tmp_path)), nprocs=world, join=True)
    want = torch.stack([_frame(i) for i in range(n_frames)]).numpy()
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"rank {r}: frames out of order or duplicated"


def _ring_worker(rank, world, port, n_steps, batch, out_dir):
    from maua_stylegan2_b200.parallel import HostFrameRing

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ring = HostFrameRing(f"maua_test_ring_{port}", rank, world, batch, (4, 6, 3), timeout_s=60)
    got = []
    for step in range(n_steps):
        dst = ring.slot_for_write(step)                       # blocks until rank 0 has consumed step - 2
        for j in range(batch):
            dst[j] = _frame((step * world + rank) * batch + j)
        if step > 0:                                          # the frame loop publishes step i-1 after queueing step i
            ring.publish(step - 1)
            if rank == 0:
                got.append(ring.frames_of(step - 1).clone())
                ring.release(step - 1)
    ring.publish(n_steps - 1)
    if rank == 0:
        got.append(ring.frames_of(n_steps - 1).clone())
        ring.release(n_steps - 1)
        np.save(os.path.join(out_dir, "ring.npy"), torch.cat(got).numpy())
    ring.close()
    dist.barrier()
    dist.destroy_process_group()


def test_host_frame_ring_world2_gloo(tmp_path):
    """Sharded D2H path: every rank writes its shard of each step into the shared two-slot ring, rank 0 reads world*B
    consecutive frames per step; the counters keep writers from overwriting a slot rank 0 has not consumed."""
    n_steps, batch, world = 9, 3, 2
    port = _free_port()
    mp.spawn(_ring_worker, args=(world, port, n_steps, batch, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "ring.npy")
    want = torch.stack([_frame(i) for i in range(n_steps * world * batch)]).numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    assert not os.path.exists(f"/dev/shm/maua_test_ring_{port}")
