"""Multi-GPU frame sharding: one process per GPU (torchrun), rank-strided batches.  Frames are
independent given their latents / noise / truncation rows, so the data path has NO collective: every rank copies its own
uint8 NHWC frames device->host into a shared pinned ring (HostFrameRing) that rank 0's sink reads in frame order.  The
inputs are broadcast once before the loop (broadcast_inputs).  AllGatherFrames — ONE NCCL all-gather of the finished
frames per step over NVLink, 3 B/pixel — is the fallback when /dev/shm cannot hold the ring, and the option for
on-device consumers.  The reference's only multi-GPU render mode is single-process `th.nn.DataParallel`
(generate_audiovisual.py:54-55), which re-broadcasts the 121 MB of weights every step and gathers fp32 images to GPU 0."""
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
        # (Measured at 8 GPUs: capping NCCL at 4 CTAs / channels made things WORSE — 4.79 instead of 4.51 ms per step — the
        # all-gather then runs longer and overlaps more of the persistent conv kernels, whose one-CTA-per-SM grids lose a
        # whole wave when an SM is busy; NCCL's defaults finish the 176 MB exchange quickly and are left alone.)
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


def current():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_ranks_agree(flag, device=None):
    """Logical AND of a per-rank boolean over the default process group (every rank gets the same answer)."""
    if current()[1] == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32,
                     device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def broadcast_inputs(tensors, src=0, device=None):
    """Make per-frame inputs identical on every rank (hooks draw random noise: each rank would otherwise filter its own
    random stream and consecutive batches, rendered by different ranks, would not be temporally coherent).
    `tensors`: list of tensors, overwritten in place with rank `src`'s values.  CPU tensors travel through `device`
    (NCCL moves device memory only) in chunks of at most 256 MB."""
    if current()[1] == 1:
        return tensors
    nccl = dist.get_backend() == "nccl"
    for t in tensors:
        if not torch.is_tensor(t):
            continue
        if not t.is_contiguous():
            raise ValueError("broadcast_inputs needs contiguous tensors")
        if t.is_cuda or not nccl:
            dist.broadcast(t, src=src)
            continue
        flat = t.view(-1)
        step = max(1, (256 << 20) // max(t.element_size(), 1))
        for i in range(0, flat.numel(), step):
            buf = flat[i:i + step].to(device)
            dist.broadcast(buf, src=src)
            flat[i:i + step].copy_(buf)
    return tensors


class HostFrameRing:
    """SLOTS-deep ring of uint8 frames in POSIX shared memory, pinned (cudaHostRegister) by every rank of the box.

    Rank r copies ITS shard of step i device->host straight into slot (i & 1), position r — every GPU uses its own PCIe
    link — and rank 0 (the only process that owns the video sink) reads world*B consecutive frames from host memory.
    With a single gather-then-D2H on rank 0 all frames of the box cross ONE link: 201 MB per step at 8 x batch 8 of
    1024^2 = 37 GB/s at the round-1 frame rate, ~70 % of PCIe Gen5 x16 and the first thing to saturate when the kernels
    get faster.  Layout: [header 4096 B: int64 ready[world], consumed] [SLOTS][world][B,H,W,3].
    Flow control (monotonic counters): writer r waits for consumed >= i - SLOTS + 1 before overwriting slot i % SLOTS;
    ready[r] = i+1 is written BY THE GPU — an 8-byte device->host copy queued on the D2H stream right behind the frames
    (`publish_on_stream`), so no host thread has to wait for the copy before telling rank 0; rank 0 waits for every
    ready[r] >= i+1.  Four slots decouple the ranks: with two, every rank ran in lock-step with rank 0's consumer and each
    hand-over paid a host polling latency (8 GPUs: 0.6 ms of a 4.5 ms step)."""

    HEADER = 4096
    SLOTS = 4
    DIR = os.environ.get("MAUA_RING_DIR", "/dev/shm")

    @classmethod
    def fits(cls, world, batch, frame_shape):
        """True when the shared-memory file system has room for the ring (containers often cap /dev/shm at 64 MB; writing
        past the cap raises SIGBUS, so callers fall back to gather-then-D2H on rank 0 instead of trying)."""
        need = cls.HEADER + cls.SLOTS * world * batch * int(torch.tensor(frame_shape).prod()) + (16 << 20)
        try:
            st = os.statvfs(cls.DIR)
            return st.f_bavail * st.f_frsize >= need
        except OSError:
            return False

    def __init__(self, name, rank, world, batch, frame_shape, timeout_s=180.0):
        import numpy as np

        self.rank, self.world, self.batch, self.timeout_s = rank, world, batch, timeout_s
        self.frame_shape = tuple(frame_shape)
        self.shard_bytes = batch * int(np.prod(self.frame_shape))
        self.nbytes = self.HEADER + self.SLOTS * world * self.shard_bytes
        self.path = os.path.join(self.DIR, name)
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        if world > 1:
            dist.barrier()
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        self.counters = self.mm[:self.HEADER].view(np.int64)      # [0..world-1] ready, [world] consumed
        self.frames = torch.from_numpy(self.mm[self.HEADER:]).view(self.SLOTS, world, batch, *self.frame_shape)
        self.counters_t = torch.from_numpy(self.counters)
        self._seq = None   # device-side step counter behind publish_on_stream
        self.registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.mm.ctypes.data, self.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the frame ring failed: {rc}")
            self.registered = True
        if world > 1:
            dist.barrier()

    def _wait(self, index, value, what):
        import time

        t0 = time.monotonic()
        while int(self.counters[index]) < value:
            dt = time.monotonic() - t0
            if dt > self.timeout_s:
                raise RuntimeError(f"HostFrameRing: rank {self.rank} timed out waiting for {what} >= {value}")
            if dt > 0.0005:          # busy-poll the first half millisecond (hand-overs are usually imminent), then yield
                time.sleep(0.00005)

    def slot_for_write(self, step):
        """Pinned host tensor [B,H,W,3] this rank fills for `step` (blocks until rank 0 has consumed step - SLOTS)."""
        if step >= self.SLOTS:
            self._wait(self.world, step - self.SLOTS + 1, "consumed")
        return self.frames[step % self.SLOTS, self.rank]

    def publish(self, step):
        """Host-side publish (CPU tensors / tests): call after the shard of `step` is in the ring."""
        self.counters[self.rank] = step + 1

    def publish_on_stream(self, step, device):
        """GPU-side publish: queue `ready[rank] = step + 1` on the CURRENT stream, behind the frame copy of `step`."""
        if self._seq is None:
            self._seq = torch.zeros(1, dtype=torch.int64, device=device)
        self._seq.fill_(step + 1)
        self.counters_t[self.rank:self.rank + 1].copy_(self._seq, non_blocking=True)

    def frames_of(self, step):
        """Rank 0: all world*B frames of `step` in frame order (blocks until every rank has published it)."""
        for r in range(self.world):
            self._wait(r, step + 1, f"ready[{r}]")
        return self.frames[step % self.SLOTS].reshape(self.world * self.batch, *self.frame_shape)

    def release(self, step):
        self.counters[self.world] = step + 1

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.mm.ctypes.data)
            self.registered = False
        if self.world > 1 and dist.is_initialized():
            dist.barrier()
        frames, counters, counters_t, mm = self.frames, self.counters, self.counters_t, self.mm
        self.frames = self.counters = self.counters_t = self.mm = None
        del frames, counters, counters_t, mm
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def ring_name():
    """The same shared-memory name on every rank of one torchrun job."""
    return "maua_ring_%s_%s" % (os.environ.get("TORCHELASTIC_RUN_ID", "local"), os.environ.get("MASTER_PORT", "0"))
