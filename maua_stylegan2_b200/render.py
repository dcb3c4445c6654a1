"""Frame loop — drop-in for the reference's `render.render` (render.py:14-192) on the B200 pipeline.

Same call signature (generator, latents, noise, offset, duration, batch_size, out_size, output_file, audio_file,
truncation, bends, rewrites, randomize_noise, ffmpeg_preset) plus two optional extensions: `sink` (any callable
receiving a uint8 [n,H,W,3] numpy batch; default = an ffmpeg rawvideo pipe like the reference, render.py:58-91)
and `world`/`rank` for frame sharding over GPUs.

What changed underneath (SURVEY.md §2.1 "D2H + uint8 convert", §3.2):
  * frames are clamped / scaled / packed to uint8 NHWC ON DEVICE (maua_rgb_to_u8_nhwc) — 3 B/pixel cross PCIe
    instead of 12, and no per-image `.cpu().numpy().astype()` on the host;
  * latents / noise / truncation / bend modulation are pinned once, each batch is copied H2D on a copy stream while
    the previous batch is still rendering (double buffering), D2H lands in pinned ping-pong buffers;
  * no 5-second `queue.get` timeouts: the writer is a plain bounded queue + thread that cannot silently truncate
    the video when a batch is slow (render.py:37,97 hazard).
"""
import os
import queue
import subprocess
import threading

import numpy as np
import torch

from . import _lib as L
from .stylegan2 import frames_to_u8


def _pin(t):
    if t is None:
        return None
    t = t.float().contiguous()
    return t.pin_memory() if not t.is_cuda else t


class FramePipeline:
    """Batches -> generator -> uint8 NHWC frames in pinned host memory, double-buffered over two CUDA streams."""

    def __init__(self, generator, latents, noise, batch_size, truncation=1.0, bends=None, rewrites=None,
                 randomize_noise=False, device=None, rank=0, world=1, use_graph=True, fit_size=None):
        self.g = generator
        self.device = device or next(generator.parameters()).device
        self.batch = batch_size
        self.rank, self.world = rank, world
        self.latents = _pin(latents)
        self.noise = [_pin(n) for n in noise]
        self.truncation = truncation if isinstance(truncation, float) else _pin(truncation)
        self.bends = bends or []
        for bend in self.bends:
            if "modulation" in bend:
                bend["modulation"] = _pin(bend["modulation"])
        self.rewrites = rewrites or {}
        self.randomize_noise = randomize_noise
        self.fit_size = fit_size  # 1920 / 1080: 2048-wide (-tall) frames are cropped + resized on the device
        self.n_frames = len(self.latents)
        self.starts = list(range(0, self.n_frames, batch_size))
        self.copy_stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._host = [None, None]
        # CUDA graphs: the ~50 launches of a batch are captured once per ping-pong slot and replayed; the host then only
        # issues the input copies + one graph launch per step (Python + ctypes per launch would otherwise bound e2e).
        # Per-batch Python objects (bend modulation, rewrites, fresh random noise) cannot be captured -> eager.
        self.use_graph = (use_graph and not self.bends and not self.rewrites and not randomize_noise)
        self._graphs = [None, None]
        self._slot_busy = [None, None]  # per ping-pong slot: (d2h-done event, collective work) still reading its frames
        self.kernels_per_step = 0       # library kernels launched per batch (graph: counted while capturing)

    def _graph_slot(self, slot, n):
        """Static inputs + captured forward of one ping-pong slot (created on first use: one eager run, then capture)."""
        gs = self._graphs[slot]
        if gs is not None:
            return gs
        sl = slice(n, n + self.batch)
        gs = {"latent": torch.empty((self.batch,) + tuple(self.latents.shape[1:]), device=self.device),
              "noise": [None if ns is None else (ns.to(self.device) if ns.shape[0] == 1 else
                                                 torch.empty((self.batch,) + tuple(ns.shape[1:]), device=self.device))
                        for ns in self.noise],
              "trunc": self.truncation if isinstance(self.truncation, float) else
              torch.empty(self.batch, device=self.device),
              "ready": None, "consumed": None}
        self._graphs[slot] = gs
        self._fill_static(gs, sl)
        run = lambda: self.g(styles=gs["latent"], noise=gs["noise"], truncation=gs["trunc"], transform_dict_list=[],
                             randomize_noise=False, input_is_latent=True, return_u8=True)[0]
        run()
        torch.cuda.current_stream(self.device).synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = L.launch_count()
        with torch.cuda.graph(graph):
            gs["out"] = run()
        self.kernels_per_step = L.launch_count() - l0
        gs["graph"] = graph
        return gs

    def _graph_prefetch(self, n, slot):
        """H2D of batch n into the slot's static inputs on the COPY stream: it overlaps the other slot's replay and only
        waits for this slot's previous replay (which read the same buffers)."""
        gs = self._graph_slot(slot, n)
        with torch.cuda.stream(self.copy_stream):
            if gs["consumed"] is not None:
                self.copy_stream.wait_event(gs["consumed"])
            self._fill_static(gs, slice(n, n + self.batch))
            gs["ready"] = torch.cuda.Event()
            gs["ready"].record(self.copy_stream)

    def _graph_step(self, n, slot):
        """Replay the slot's captured forward on batch n (prefetched, or copied now); returns the static uint8 frames."""
        cur = torch.cuda.current_stream(self.device)
        busy = self._slot_busy[slot]
        if busy is not None:  # the static output of this slot may still be read by the D2H copy / all-gather of step i-2
            if busy[0] is not None:
                cur.wait_event(busy[0])
            if busy[1] is not None:
                busy[1].wait()
            self._slot_busy[slot] = None
        gs = self._graph_slot(slot, n)
        if gs["ready"] is None:
            self._graph_prefetch(n, slot)
        cur.wait_event(gs["ready"])
        gs["ready"] = None
        gs["graph"].replay()
        gs["consumed"] = torch.cuda.Event()
        gs["consumed"].record(cur)
        return gs["out"]

    def warmup(self):
        """Capture the CUDA graphs of both ping-pong slots (or run one eager batch) before the frame loop starts."""
        if self.n_frames == 0:
            return
        with torch.no_grad():
            if self.use_graph and self.n_frames >= self.batch:
                for slot in (0, 1):
                    self._graph_step(0, slot)
            else:
                self._render(self._stage(0), 0)
        torch.cuda.current_stream(self.device).synchronize()
        self.h2d_bytes = 0

    def _fill_static(self, gs, sl):
        gs["latent"].copy_(self.latents[sl], non_blocking=True)
        self.h2d_bytes += 0 if self.latents.is_cuda else gs["latent"].numel() * 4
        for dst, ns in zip(gs["noise"], self.noise):
            if ns is not None and ns.shape[0] != 1:
                dst.copy_(ns[sl], non_blocking=True)
                self.h2d_bytes += 0 if ns.is_cuda else dst.numel() * 4
        if not isinstance(self.truncation, float):
            gs["trunc"].copy_(self.truncation[sl], non_blocking=True)
            self.h2d_bytes += gs["trunc"].numel() * 4

    def prepare_host_buffers(self, frame_shape, n_per_step=None):
        """Pin the two ping-pong frame buffers up front (pinning 25-200 MB takes milliseconds; keep it out of the loop)."""
        n = n_per_step or self.batch * self.world
        for slot in (0, 1):
            self._host[slot] = torch.empty((n,) + tuple(frame_shape), dtype=torch.uint8).pin_memory()

    # -- one batch -------------------------------------------------------------------------------------------------
    def _stage(self, n):
        """H2D of batch n on the copy stream (pinned -> device, non_blocking); returns a dict of device tensors."""
        sl = slice(n, n + self.batch)
        with torch.cuda.stream(self.copy_stream):
            item = {"latent": self.latents[sl].to(self.device, non_blocking=True)}
            self.h2d_bytes += 0 if self.latents.is_cuda else item["latent"].numel() * 4
            item["noise"] = []
            for ns in self.noise:
                if ns is None:
                    item["noise"].append(None)
                elif ns.shape[0] == 1:
                    item["noise"].append(ns.to(self.device, non_blocking=True))
                else:
                    t = ns[sl].to(self.device, non_blocking=True)
                    self.h2d_bytes += 0 if ns.is_cuda else t.numel() * 4
                    item["noise"].append(t)
            if isinstance(self.truncation, float):
                item["truncation"] = self.truncation
            else:
                item["truncation"] = self.truncation[sl].to(self.device, non_blocking=True)
                self.h2d_bytes += item["truncation"].numel() * 4
            item["bends"] = []
            staged = [item["latent"]] + [t for t in item["noise"] if t is not None]
            if torch.is_tensor(item["truncation"]):
                staged.append(item["truncation"])
            for bend in self.bends:
                if "modulation" in bend:
                    mod = bend["modulation"][sl].to(self.device, non_blocking=True)
                    self.h2d_bytes += mod.numel() * 4
                    staged.append(mod)
                    transform = bend["transform"](mod)
                    # tensors the transform factory derived from the modulation on the copy stream (module parameters /
                    # buffers / attributes) are read by the compute stream as well
                    if isinstance(transform, torch.nn.Module):
                        staged += [t for t in list(transform.parameters()) + list(transform.buffers()) if t.is_cuda]
                        for m in transform.modules():
                            staged += [v for v in vars(m).values() if torch.is_tensor(v) and v.is_cuda]
                    item["bends"].append({"layer": bend["layer"], "transform": transform})
                else:
                    item["bends"].append({"layer": bend["layer"], "transform": bend["transform"]})
            item["staged"] = staged
            item["ready"] = torch.cuda.Event()
            item["ready"].record(self.copy_stream)
        return item

    def _apply_rewrites(self, n):
        # render.py:160-167 (with the Tensor.copy() bug fixed: originals are cloned once)
        if not self.rewrites:
            return
        if not hasattr(self, "_orig"):
            params = dict(self.g.named_parameters())
            self._orig = {k: params[k].detach().clone() for k in self.rewrites}
        for name, (rewrite, modulation) in self.rewrites.items():
            transform = rewrite(modulation[n:n + self.batch])
            new = transform(self._orig[name]).to(self.device)
            mod = self.g
            parts = name.split(".")
            for a in parts[:-1]:
                mod = getattr(mod, a)
            setattr(mod, parts[-1], torch.nn.Parameter(new, requires_grad=False))

    def _render(self, item, n):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(item["ready"])
        # The staged tensors were allocated on the COPY stream: tell the caching allocator that the compute stream reads
        # them, otherwise dropping `item` hands their blocks straight back to the copy stream's pool and the H2D of
        # batch i+2 overwrites noise maps that batch i's forward (still running) has not read yet.
        for t in item["staged"]:
            t.record_stream(cur)
        self._apply_rewrites(n)
        l0 = L.launch_count()
        frames, _ = self.g(styles=item["latent"], noise=item["noise"], truncation=item["truncation"],
                           transform_dict_list=item["bends"], randomize_noise=self.randomize_noise,
                           input_is_latent=True, return_u8=True)
        self.kernels_per_step = L.launch_count() - l0
        return frames  # uint8 [b,H,W,3] on device

    def run(self, consume, gather=None, ring=None):
        """Render every frame; `consume(np.uint8 [n,H,W,3])` is called in frame order on ranks where it is not None.

        world == 1: step i renders batch i.   world > 1: step i renders batches i*world + rank (rank-strided, so one
        step yields world*B CONSECUTIVE frames).  Two ways to bring them to the sink:
          gather(frames_u8) -> (work, out_u8): ONE NCCL all-gather per step (parallel.AllGatherFrames) — every rank ends up
              with all frames on the device; without `ring`, rank 0 copies all of them to the host.
          ring (parallel.HostFrameRing): every rank copies ITS shard device->host into a shared pinned ring over its own
              PCIe link and rank 0's sink reads the consecutive frames from host memory (no single-link ceiling).
        Both may be given (the collective then serves on-device consumers only); with neither, at world > 1 every rank
        keeps its own frames (rank 0's `consume` then sees only rank 0's batches).  Short tails are padded by repeating the
        last frame and trimmed before `consume`.  H2D of step i+1 (copy stream) and D2H of step i-1 (d2h stream) overlap
        the synthesis of step i (compute stream); the compute stream never waits for a collective."""
        world, rank = self.world, self.rank   # (world > 1 without gather/ring: the frames stay sharded in each rank's HBM)
        nb = len(self.starts)
        steps = (nb + world - 1) // world
        cur = torch.cuda.current_stream(self.device)
        d2h = torch.cuda.Stream(self.device)

        def batch_start(step):
            return self.starts[min(step * world + rank, nb - 1)]

        def finish(p):
            if ring is not None:
                # (the GPU published ready[rank] itself, behind the frame copy: nothing to wait for on the writer side)
                if consume is not None:
                    consume(ring.frames_of(p["step"]).numpy()[:p["valid"]])
                    ring.release(p["step"])
                return
            p["done"].synchronize()
            if consume is not None:
                consume(p["host"].numpy()[:p["valid"]])

        pending = None
        graph_ok = self.use_graph
        nxt = self._stage(batch_start(0)) if (steps and not graph_ok) else None
        for i in range(steps):
            n = batch_start(i)
            full = n + self.batch <= self.n_frames
            if graph_ok and full:
                frames = self._graph_step(n, i & 1)
                if i + 1 < steps and batch_start(i + 1) + self.batch <= self.n_frames:
                    self._graph_prefetch(batch_start(i + 1), (i + 1) & 1)   # overlaps this step's replay
            else:
                item = nxt if nxt is not None else self._stage(n)
                nxt = self._stage(batch_start(i + 1)) if (i + 1 < steps and not graph_ok) else None
                frames = self._render(item, n)
            if self.fit_size in (1920, 1080):
                frames = fit_frames(frames, self.fit_size)
            if frames.shape[0] < self.batch:  # short tail: pad by repeating the last frame (equal-size collective)
                frames = torch.cat([frames, frames[-1:].expand(self.batch - frames.shape[0], -1, -1, -1)], 0)
            valid = min(self.n_frames - i * world * self.batch, world * self.batch)
            if gather is None and ring is None:
                valid = min(self.n_frames - n, self.batch)      # unexchanged: this rank's own batch only
            work = None
            if gather is not None:
                work, out = gather(frames, i & 1)   # async collective on NCCL's stream; compute does not wait for it
            else:
                out = frames
            if ring is not None:
                # sharded D2H: this rank's own frames -> its position in the shared pinned ring (all ranks take part)
                dst = ring.slot_for_write(i)
                if pending is not None:
                    finish(pending)                 # step i-1: publish (and, on rank 0, consume) before queueing more
                    pending = None
                ready = torch.cuda.Event()
                ready.record(cur)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(ready)
                    dst.copy_(frames, non_blocking=True)
                    frames.record_stream(d2h)
                    ring.publish_on_stream(i, self.device)     # ready[rank] = i + 1, written by the GPU after the frames
                    done = torch.cuda.Event()
                    done.record(d2h)
                self.d2h_bytes += frames.numel()
                pending = {"done": done, "valid": valid, "step": i}
                if graph_ok and full:
                    self._slot_busy[i & 1] = (done, work)
            elif consume is not None:
                slot = i & 1
                if self._host[slot] is None or self._host[slot].shape != out.shape:
                    self._host[slot] = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
                ready = torch.cuda.Event()
                ready.record(cur)
                with torch.cuda.stream(d2h):
                    if work is not None:
                        work.wait()                 # the d2h stream (not the compute stream) waits for the gather
                    d2h.wait_event(ready)
                    self._host[slot].copy_(out, non_blocking=True)
                    out.record_stream(d2h)
                    done = torch.cuda.Event()
                    done.record(d2h)
                self.d2h_bytes += out.numel()
                if pending is not None:
                    finish(pending)                 # frames of step i-1 (other ping-pong slot)
                pending = {"done": done, "host": self._host[slot], "valid": valid, "step": i}
                if graph_ok and full:
                    self._slot_busy[i & 1] = (done if gather is None else None, work)
            elif graph_ok and full and work is not None:
                self._slot_busy[i & 1] = (None, work)
        if pending is not None:
            if ring is not None:
                pending["done"].synchronize()   # this rank's last shard is in the ring before run() returns
            finish(pending)
        for slot in (0, 1):   # the last collectives are part of the job: the compute stream joins them before run() returns
            busy = self._slot_busy[slot]
            if busy is not None and busy[1] is not None:
                busy[1].wait()
                self._slot_busy[slot] = (busy[0], None)


class FFmpegSink:
    """rawvideo rgb24 on stdin -> libx264 yuv420p (+ 320k audio mux), the reference's wire format (render.py:58-91)."""

    def __init__(self, output_file, width, height, fps, audio_file=None, offset=0, duration=None, preset="slow",
                 vcodec=None):
        # libx264 like the reference; MAUA_VCODEC=h264_nvenc (or vcodec=) hands the encode to NVENC when the installed
        # ffmpeg was built with it — at >1000 frames/s of synthesis libx264 `slow` is the end-to-end limit
        vcodec = vcodec or os.environ.get("MAUA_VCODEC", "libx264")
        cmd = ["ffmpeg", "-hide_banner", "-y", "-v", "warning", "-f", "rawvideo", "-pix_fmt", "rgb24", "-framerate",
               str(fps), "-s", f"{width}x{height}", "-i", "pipe:"]
        if audio_file is not None:
            cmd += ["-ss", str(offset)] + (["-t", str(duration)] if duration else []) + ["-i", audio_file]
        cmd += ["-framerate", str(fps), "-vcodec", vcodec, "-pix_fmt", "yuv420p", "-preset", preset]
        if audio_file is not None:
            cmd += ["-b:a", "320K", "-ac", "2"]
        cmd += [output_file]
        self.proc = subprocess.Popen(cmd, stdin=subprocess.PIPE)
        self.q = queue.Queue(maxsize=8)
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        while True:
            frames = self.q.get()
            if frames is None:
                break
            self.proc.stdin.write(frames.tobytes())
        self.proc.stdin.close()
        self.proc.wait()

    def __call__(self, frames):
        self.q.put(np.ascontiguousarray(frames).copy())

    def close(self):
        self.q.put(None)
        self.t.join()


def pillow_bilinear_coeffs(in_size, out_size):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` (src/libImaging/Resample.c) for the BILINEAR filter
    (support 1.0, triangle): per output index the first source index, the tap count and 22-bit fixed-point taps.
    Host-side table construction only (a few KB); checked against PIL itself in tests/test_fit_frames.py."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - xmin
        v = np.abs((np.arange(n) + xmin - center + 0.5) * inv)
        w = np.where(v < 1.0, 1.0 - v, 0.0)
        total = w.sum()
        kk[xx, :n] = w / total if total != 0.0 else w
        bounds[xx] = (xmin, n)
    fixed = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << 22)), np.trunc(0.5 + kk * (1 << 22))).astype(np.int32)
    return bounds, fixed


_FIT_TABLES = {}


def fit_frames(frames, out_size):
    """render.py:98-105 on the device: uint8 [n,H,W,3] frames of a 2048-wide (-tall) generator -> crop 112 px off both
    ends of the long axis -> PIL-exact bilinear resize to 1920x1080 (1080x1920).  Other shapes pass through."""
    from . import _lib as L

    n, h, w, _ = frames.shape
    if w == 2048:
        crop, out_hw = (0, 112, h, w - 224), (1080, 1920)
    elif h == 2048:
        crop, out_hw = (112, 0, h - 224, w), (1920, 1080)
    else:
        return frames
    key = (crop, out_hw, frames.device)
    if key not in _FIT_TABLES:
        bx, kx = pillow_bilinear_coeffs(crop[3], out_hw[1])
        by, ky = pillow_bilinear_coeffs(crop[2], out_hw[0])
        _FIT_TABLES[key] = tuple(torch.from_numpy(np.ascontiguousarray(t)).to(frames.device) for t in (bx, kx, by, ky))
    bx, kx, by, ky = _FIT_TABLES[key]
    frames = frames.contiguous()
    out = torch.empty((n, out_hw[0], out_hw[1], 3), dtype=torch.uint8, device=frames.device)
    L.call("maua_fit_frames_u8", frames.data_ptr(), out.data_ptr(), n, h, w, crop[0], crop[1], crop[2], crop[3],
           out_hw[0], out_hw[1], bx.data_ptr(), kx.data_ptr(), kx.shape[1], by.data_ptr(), ky.data_ptr(), ky.shape[1],
           L.stream_ptr(frames.device))
    return out


def render(generator, latents, noise, offset, duration, batch_size, out_size, output_file, audio_file=None,
           truncation=1.0, bends=[], rewrites={}, randomize_noise=False, ffmpeg_preset="slow", sink=None):
    """Drop-in for the reference's `render.render` (render.py:14-192).  Under torchrun (torch.distributed initialised,
    see parallel.init_from_env — what `generate(dataparallel=True)` does) the frames are sharded over the ranks of the
    box: inputs are made identical on every rank (broadcast from rank 0), rank r renders batches i*world + r, every rank
    copies its own uint8 frames device->host into a shared pinned ring (no data-path collective; when /dev/shm has no room,
    or MAUA_FRAME_EXCHANGE=gather: one NCCL all-gather per step and rank 0 copies), and only rank 0 owns the sink / ffmpeg
    process.  Returns the FramePipeline (its h2d/d2h byte counters are per rank)."""
    from . import parallel

    sizes = {512: (512, 512), 1024: (1024, 1024), 1920: (1920, 1080), 1080: (1080, 1920)}
    if out_size not in sizes:
        raise Exception("The only output sizes currently supported are: 512, 1024, 1080, or 1920")
    w, h = sizes[out_size]
    rank, world = parallel.current()
    own_sink = sink is None and rank == 0
    if own_sink:
        sink = FFmpegSink(output_file, w, h, len(latents) / duration, audio_file, offset, duration, ffmpeg_preset)
    if hasattr(generator, "module"):  # th.nn.DataParallel wrapper of the reference CLI (generate_audiovisual.py:54-55)
        generator = generator.module
    noise = list(noise)
    gather = ring = None
    if world > 1:
        device = next(generator.parameters()).device
        shared = [latents] + noise + [truncation] + [b.get("modulation") for b in bends]
        shared += [m for _, m in (rewrites or {}).values()]
        parallel.broadcast_inputs([t for t in shared if torch.is_tensor(t)], device=device)
        want_ring = os.environ.get("MAUA_FRAME_EXCHANGE", "ring") != "gather"
        if want_ring and parallel.all_ranks_agree(parallel.HostFrameRing.fits(world, batch_size, (h, w, 3)), device):
            # frames are independent and the only consumer is rank 0's sink, which reads host memory: no collective at all
            ring = parallel.HostFrameRing(parallel.ring_name(), rank, world, batch_size, (h, w, 3))
        else:
            # no room in /dev/shm (or MAUA_FRAME_EXCHANGE=gather): one NCCL all-gather per step and rank 0 copies the
            # gathered frames itself — one PCIe link instead of `world`
            gather = parallel.AllGatherFrames(world)
    pipe = FramePipeline(generator, latents, noise, batch_size, truncation, bends, rewrites, randomize_noise,
                         fit_size=out_size, rank=rank, world=world)
    pipe.warmup()

    def consume(frames):
        assert frames.shape[2] == w and frames.shape[1] == h, (
            f"generator's output image size does not match specified output size: \n"
            f"got: {frames.shape[2]}x{frames.shape[1]}\t\tshould be {w}x{h}")
        sink(frames)

    try:
        with torch.no_grad():
            pipe.run(consume if rank == 0 else None, gather, ring)
    finally:
        if ring is not None:
            ring.close()
    if own_sink:
        sink.close()
    return pipe
