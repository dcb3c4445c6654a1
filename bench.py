#!/usr/bin/env python
"""bench.py — frames/s of the StyleGAN2 synthesis hot path at BASELINE.json configs[1]
(1024x1024 config-f generator, 30 s of audio @30 fps = 900 frames, default-hook shaped latents/noise, batch 8).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

A step = one pass of the hot path over one batch of `--batch` frames (per GPU; weak scaling): style prologue, 17
modulated convs (tcgen05 path; precision per `dtype`), blur/noise/bias/lrelu epilogues, 9 ToRGB, uint8 NHWC pack.  For
N > 1 the frames are sharded rank-strided with no data-path collective (frames are independent); e2e brings
them to rank 0's sink through a shared pinned host ring, every rank over its own PCIe link.
  value    : frames/s with latents/noise resident in HBM, CUDA events around exactly K steps, max over ranks.
  e2e      : frames/s through the public frame loop (maua_stylegan2_b200.render.FramePipeline): pinned host
             latents/noise -> H2D every step, synthesis, uint8 frames -> D2H into pinned memory, inside the timed region.
  roofline : dominant kernel = maua_modconv_tc (tensor-bound); achieved = algorithmic conv FLOPs / summed kernel time
             measured live with CUDA events around every launch of K extra instrumented steps.
  cpu_baseline / --impl reference : the oracle port of the reference's CPU path (oracle/stylegan2_oracle.py,
             torch CPU fp32, all host threads) on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE, CM, FPS, AUDIO_S = 1024, 2, 30, 30
DEFAULT_PRECISION = "mixed"
DTYPES = {"bf16x3": "bf16x3 (split-bf16 tensor-core products hi*hi + hi*lo + lo*hi, fp32 accumulate; fp32-grade: network "
                    "error ~5e-5 of the tensor max)",
          "mixed": "bf16x3 (split-bf16, 3 products) below 512^2; fp16 activations x fp16 (hi, lo) weight pair at >= 512^2 (one "
                   "tensor-core pass); fp32 accumulate everywhere; measured network error 4.9e-4 of the tensor max on the "
                   "activation maps (parity bar 1e-3; --precision bf16x3: 5e-5)",
          "bf16": "bf16 (single product, fast preview)"}
CONV_GFLOP_PER_FRAME = 148.52  # BASELINE.md §3 (2*MAC, algorithmic)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            m = json.load(open(f))
            p.update({k: float(m[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
            p["source"] = "measured"
        except Exception:
            pass
    return p


def ncu_traffic(kernel):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch set of `kernel`, from the newest committed
    `ncu --set full` capture of this same step (profiles/rNN_traffic.json, written by tools/collect_profiles.py)."""
    d = os.path.join(ROOT, "profiles")
    try:
        for f in sorted((x for x in os.listdir(d) if x.endswith("_traffic.json")), reverse=True):
            t = json.load(open(os.path.join(d, f)))
            if kernel in t:
                return t[kernel]["dram_bytes"]
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------------
# workload (synthetic, shape-faithful to audioreactive/examples/default.py)
# ----------------------------------------------------------------------------------------------------------------

def make_workload(n_frames, seed=0):
    """Host tensors: latents [n_frames,18,512] (smooth mixture of 12 'selected' W+ rows, like chroma-weighted latents),
    noise: per-frame maps for widths <= 256 (examples/default.py:29-30 returns None above -> buffer noise)."""
    import numpy as np
    import torch

    rng = np.random.Generator(np.random.PCG64(seed))
    sel = torch.from_numpy(rng.standard_normal((12, 18, 512)).astype(np.float32))
    wgt = torch.from_numpy(rng.uniform(0, 1, (n_frames, 12)).astype(np.float32))
    k = torch.hann_window(31, periodic=False)
    k = (k / k.sum())[None, None]
    wgt = torch.nn.functional.conv1d(torch.nn.functional.pad(wgt.t()[:, None], (15, 15), mode="circular"), k)[:, 0].t()
    wgt = wgt / wgt.sum(1, keepdim=True)
    latents = torch.einsum("tn,nld->tld", wgt, sel).contiguous()
    noise = []
    num_layers = (10 - 2) * 2 + 1
    g = torch.Generator().manual_seed(seed)
    for l in range(num_layers):
        r = 2 ** ((l + 5) // 2)
        noise.append(torch.randn(n_frames, 1, r, r, generator=g) if r <= 256 else None)
    return latents, noise


def make_generator(device, impl="tc", precision="bf16x3", seed=0):
    import torch

    from maua_stylegan2_b200.stylegan2 import Generator

    torch.manual_seed(seed)
    g = Generator(SIZE, 512, 8, channel_multiplier=CM, constant_input=True, output_size=SIZE, impl=impl,
                  precision=precision)
    with torch.no_grad():  # zero-initialised parameters get N(0, 0.1^2) so noise/bias paths do real work
        for name, prm in g.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias") or (name.startswith("to_rgb") and name.endswith(".bias")):
                prm.normal_(0, 0.1)
    g = g.to(device).eval()
    g.truncation_latent = torch.zeros(1, 512, device=device)
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "50", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ----------------------------------------------------------------------------------------------------------------

def cpu_port_run(steps, warmup, budget_s=150.0, frames_per_step=1):
    import numpy as np
    import torch

    from oracle import stylegan2_oracle as O

    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if "MAUA_CPU_THREADS" in os.environ:
        torch.set_num_threads(int(os.environ["MAUA_CPU_THREADS"]))
    else:
        # "all the host threads it can use": oneDNN grouped convs get SLOWER when oversubscribed (128 threads were 13x
        # slower than 16 on the pool's hosts), so pick the fastest count on a quick 256x256 probe of the same code
        sd_p = O.synth_state_dict(256, channel_multiplier=CM, seed=0)
        _, nl_p, nlat_p = O.layout(256)
        lat_p = torch.zeros(1, nlat_p, 512)
        nz_p = [torch.zeros(1, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2)) for l in range(nl_p)]
        best = (1e30, avail)
        for t in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
            torch.set_num_threads(t)
            with torch.no_grad():
                O.generator_forward(sd_p, 256, lat_p, nz_p, 1.0, torch.zeros(1, 512), channel_multiplier=CM)
                t0 = time.perf_counter()
                O.generator_forward(sd_p, 256, lat_p, nz_p, 1.0, torch.zeros(1, 512), channel_multiplier=CM)
                dt = time.perf_counter() - t0
            if dt < best[0]:
                best = (dt, t)
            if dt > 4 * best[0]:
                break
        torch.set_num_threads(best[1])
        del sd_p
    cores = torch.get_num_threads()
    sd = O.synth_state_dict(SIZE, channel_multiplier=CM, seed=0)
    _, num_layers, n_latent = O.layout(SIZE)
    rng = np.random.Generator(np.random.PCG64(1))
    latent = torch.from_numpy(rng.standard_normal((frames_per_step, n_latent, 512)).astype(np.float32)) * 0.5
    noise = [torch.from_numpy(rng.standard_normal((frames_per_step, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             if 2 ** ((l + 5) // 2) <= 256 else None for l in range(num_layers)]
    tl = torch.zeros(1, 512)

    def one():
        with torch.no_grad():
            img, _ = O.generator_forward(sd, SIZE, latent, noise, 1.0, tl, channel_multiplier=CM)
            O.frames_to_u8(img)

    t_w = time.perf_counter()
    for _ in range(max(1, warmup)):
        one()
        if time.perf_counter() - t_w > budget_s / 3:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"fps": done * frames_per_step / dt, "cores": cores, "steps_done": done, "seconds": dt,
            "sample": f"{done} step(s) x {frames_per_step} frame of the 1024x1024 config-f batch (oracle port, torch CPU fp32, "
                      f"{cores} threads)"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------------------------
def _emit(line):
    """The ONE JSON line on the real stdout (see main(): fd 1 is pointed at stderr while the bench runs)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries print to stdout behind Python's back (NCCL: "NCCL version 2.28.9+cuda12.9" on the first collective);
    # the contract is ONE JSON line, so everything else written to fd 1 is sent to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--conv", default="tc", choices=["tc", "simt"])
    ap.add_argument("--precision", default=DEFAULT_PRECISION, choices=["bf16x3", "mixed", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-audio-chain", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "configs[1]: 1024x1024 config-f generator (channel_multiplier=2), 30 s audio @30 fps = 900 "
                          "frames, default-hook shaped latents [900,18,512] + per-frame noise for widths <= 256",
              "batch_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": f"frames sharded dp{world}",
              "l2": "per-step working set (activations ~2.4 GB at batch 8) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_port_run(args.steps, args.warmup)
        config = dict(config, reference_sample="each step = ONE frame of the batch-8 workload (a bounded sample: the CPU path "
                                               "needs ~0.6-2 s per frame); kind 'port' = oracle restatement of the reference's CPU "
                                               "path (oracle/stylegan2_oracle.py), thread count probed for the fastest setting")
        line = {"impl": "reference", "metric": "1024x1024 frames/sec", "value": r["fps"], "unit": "frames/s",
                "n_gpus": args.gpus, "steps": r["steps_done"], "warmup": args.warmup,
                "ms_per_step": 1000.0 * r["seconds"] / max(r["steps_done"], 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"], "cpu": cpu_model()},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    import torch
    import torch.distributed as dist

    from maua_stylegan2_b200 import _lib as L
    from maua_stylegan2_b200.parallel import AllGatherFrames, HostFrameRing, all_ranks_agree, init_from_env, ring_name
    from maua_stylegan2_b200.render import FramePipeline

    rank, world, local_rank = init_from_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pk = peaks()
    B = args.batch
    n_frames = FPS * AUDIO_S
    latents_h, noise_h = make_workload(n_frames)
    g = make_generator(device, impl=args.conv, precision=args.precision)
    latents_d = latents_h.to(device)
    noise_d = [n.to(device) if n is not None else None for n in noise_h]
    nb = n_frames // B

    def step_local(i):
        # rank-strided batches; inputs are already resident in HBM
        n = ((i * world + rank) % nb) * B
        frames, _ = g(latents_d[n:n + B], noise=[x[n:n + B] if x is not None else None for x in noise_d],
                      truncation=1.0, input_is_latent=True, randomize_noise=False, return_u8=True)
        return frames

    def barrier():
        if world > 1:
            dist.barrier()

    def pipeline_run(lat, nz, n_steps, offset, to_host):
        """K steps of the public frame loop (render.FramePipeline: CUDA-graph replay of the captured forward, rank-strided
        batches; `value` keeps each rank's frames in its own HBM, `e2e` moves them to rank 0's sink through the shared
        pinned host ring — no data-path collective).  Graph capture and
        first-touch setup happen in warmup(), outside the timed region.  Returns (device ms, wall s, pipe)."""
        lo = offset * B * world
        idx = torch.tensor([j % n_frames for j in range(lo, lo + n_steps * B * world)]).to(lat.device)
        pipe = FramePipeline(g, lat[idx], [x[idx] if x is not None else None for x in nz], B, truncation=1.0, rank=rank,
                             world=world)
        ring = gather = None
        if to_host and world > 1 and all_ranks_agree(HostFrameRing.fits(world, B, (SIZE, SIZE, 3)), device):
            # every rank copies ITS frames device->host into a shared pinned ring over its own PCIe link; rank 0's sink
            # reads world*B consecutive frames per step from host memory (render.render does the same): frames are
            # independent, so there is no data-path collective
            ring = HostFrameRing(f"{ring_name()}_{offset}_{n_steps}", rank, world, B, (SIZE, SIZE, 3))
        elif to_host and world > 1:   # /dev/shm has no room: one NCCL all-gather per step, rank 0 copies the gathered frames
            gather = AllGatherFrames(world)
        if to_host and ring is None and rank == 0:
            pipe.prepare_host_buffers((SIZE, SIZE, 3))
        pipe.warmup()
        sink_bytes = [0]

        def consume(frames):
            sink_bytes[0] += int(frames[:, ::64, ::64].sum()) * 0 + frames.nbytes  # touch the host frames

        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        pipe.run(consume if (to_host and rank == 0) else None, gather, ring)
        e1.record()      # (run() makes the compute stream wait for the last collectives before returning)
        torch.cuda.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        if ring is not None:
            ring.close()
        return e0.elapsed_time(e1), wall, pipe

    with torch.no_grad():
        sampler = ClockSampler(local_rank) if rank == 0 else None
        # ---- value: inputs resident in HBM, exactly K steps, CUDA events, max over ranks --------------------------
        pipeline_run(latents_d, noise_d, args.warmup, 0, False)
        l0 = L.launch_count()
        ms_dev, _, vpipe = pipeline_run(latents_d, noise_d, args.steps, args.warmup, False)
        launches = vpipe.kernels_per_step * args.steps
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([ms_dev], device=device)
        lt = torch.tensor([float(launches)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        ms_total = float(ms.item())
        value = args.steps * B * world / (ms_total / 1000.0)

        # ---- e2e through the public frame loop (host buffers, H2D + D2H inside the timed region) -----------------
        pipeline_run(latents_h, noise_h, args.warmup, 0, True)
        _, dt, pipe = pipeline_run(latents_h, noise_h, args.steps, args.warmup, True)
        dtt = torch.tensor([dt], device=device)
        traffic = torch.tensor([float(pipe.h2d_bytes), float(pipe.d2h_bytes)], device=device)
        if world > 1:
            dist.all_reduce(dtt, op=dist.ReduceOp.MAX)
            dist.all_reduce(traffic, op=dist.ReduceOp.SUM)
        e2e = {"value": args.steps * B * world / float(dtt.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(traffic[0].item()) // args.steps,
               "d2h_bytes_per_step": int(traffic[1].item()) // max(args.steps, 1),
               "api": "maua_stylegan2_b200.render.FramePipeline (pinned host latents/noise -> uint8 frames in pinned host memory)"
                      + ("; bytes summed over ranks: every rank uploads its own batches and copies its own frames into a "
                         "shared pinned host ring (parallel.HostFrameRing) read by rank 0's sink" if world > 1 else "")}

        # ---- roofline pass: CUDA events around every launch of the hot kernels (rank 0, a few extra steps) --------
        roof = roof_ufd = shares = None
        if rank == 0:
            names = {"maua_modconv_tc", "maua_modconv_simt_f32", "maua_blur_act_nhwc", "maua_torgb_f32", "maua_rgb_finish_u8",
                     "maua_modulate_split_nhwc", "maua_style_prologue_f32", "maua_rgb_to_u8_nhwc", "maua_upfirdn2d_f32",
                     "maua_noise_bias_act_f32", "maua_rgb_finish_f32", "maua_rgb_weights_f32"}
            step_local(999)    # un-instrumented warm-up of the per-operator path (plan build, first-use attribute calls)
            torch.cuda.synchronize()
            L.PROFILE = {"names": names, "events": []}
            nprof = 5
            for i in range(nprof):
                step_local(1000 + i)   # no collective here: only rank 0 runs the instrumented pass
            torch.cuda.synchronize()
            ev = L.PROFILE["events"]
            L.PROFILE = None
            # the k-th launch of every instrumented step is the same kernel on the same shapes: take the MEDIAN over the
            # steps (a host-side hiccup between the start event and the launch — e.g. a Python GC pause — lands inside
            # one launch's event pair and would otherwise inflate that layer's mean)
            per_step = len(ev) // nprof
            times = [statistics.median(ev[k + j * per_step][2].elapsed_time(ev[k + j * per_step][3]) for j in range(nprof))
                     for k in range(per_step)]
            tot = {}
            conv_ms, conv_fl, conv_issued = 0.0, 0.0, 0.0
            per_layer = {}
            for (name, tag, _, _), t in zip(ev[:per_step], times):
                tot[name] = tot.get(name, 0.0) + t
                t *= nprof   # (the sums below are divided by nprof)
                if name in ("maua_modconv_tc", "maua_modconv_simt_f32") and tag is not None:
                    conv_ms += t
                    conv_fl += tag["flops"] * nprof
                    conv_issued += tag["flops"] * tag.get("nprod", 1) * nprof
                    key = f"L{tag['layer']}:{tag['cin']}->{tag['cout']}@{tag['h']}{'up' if tag['up'] else ''}"
                    a = per_layer.setdefault(key, [0.0, 0.0])
                    a[0] += t
                    a[1] += tag["flops"] * nprof
            shares = {k: round(v, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}
            if conv_ms > 0:
                ach = conv_fl / (conv_ms * 1e-3) / 1e12
                nprod = 3 if (args.precision in ("bf16x3", "mixed") and args.conv == "tc") else 1
                roof = {"kernel": "maua_modconv_tc" if args.conv == "tc" else "maua_modconv_simt_f32", "bound": "tensor",
                        "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_tflops_sustained"], "traffic": ncu_traffic("maua_modconv_tc") if args.conv == "tc" and B == 8 else None,
                        "traffic_unit": "bytes of DRAM read+write per launch set (17 conv launches of one batch-8 step)",
                        "traffic_source": "committed `ncu --set full` capture of this same step (profiles/*_traffic.json, "
                                          "tools/collect_profiles.py) — not measured live by this run",
                        "peak_source": pk["source"] + " sustained bf16",
                        "algorithmic_gflop_per_launch_set": conv_fl / nprof / 1e9,
                        "products_per_mac": nprod if args.precision != "mixed" else "3 below 512^2, 2 (fp16 a*(w_hi+w_lo)) at >= 512^2",
                        "issued_frac": conv_issued / (conv_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                        "ms_per_step": conv_ms / nprof,
                        # the instrumented pass is eager and runs after the timed loops (its clock under the power cap can
                        # differ by several % between box visits); the kernel's SHARE of the step is stable, so the same
                        # figure is also given scaled to the graph-replayed step of the timed region
                        "share_of_step": conv_ms / nprof / sum(tot.values()),
                        "achieved_in_timed_region": conv_fl / nprof / (conv_ms / nprof / sum(tot.values()) * ms_total / args.steps * 1e-3) / 1e12,
                        "per_layer_tflops": {k: round(v[1] / (v[0] * 1e-3) / 1e12, 2) for k, v in per_layer.items()},
                        "per_layer_ms": {k: round(v[0] / nprof, 4) for k, v in per_layer.items()}}
            # standalone upfirdn2d (the public op) on the largest Blur of the frame: [B,32,2049,2049] -> [B,32,2048,2048]
            from maua_stylegan2_b200 import op

            bb = min(B, 4)
            x = torch.randn(bb, 32, 2049, 2049, device=device)
            kk = torch.tensor([1.0, 3.0, 3.0, 1.0], device=device)
            k4 = kk[None] * kk[:, None] / 16
            for _ in range(3):
                y = op.upfirdn2d(x, k4, pad=(1, 1))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            s.record()
            for _ in range(reps):
                y = op.upfirdn2d(x, k4, pad=(1, 1))
            e.record()
            torch.cuda.synchronize()
            by = 4.0 * (x.numel() + y.numel())
            gbs = by / (s.elapsed_time(e) / reps * 1e-3) / 1e9
            roof_ufd = {"kernel": "maua_upfirdn2d_f32 (blur_tile_pipe_kernel)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"],
                        "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": ncu_traffic("maua_upfirdn2d_f32") if bb == 4 else None,
                        "algorithmic_bytes": by,
                        "shape": f"[{bb},32,2049,2049]->[{bb},32,2048,2048] fp32 (in+out {by / 1e9:.2f} GB > L2)"}
            del x, y

    # ---- GPU comparator (SURVEY §8(d) / BASELINE.md §4.5): the reference's own CUDA path — its compiled op/ extension
    # (oracle/_ref) + cuDNN grouped convs on per-sample modulated weights — on this same GPU, after (never inside) the repo
    # arm's timed regions.  Reported next to cpu_baseline; not the target, the honest comparator.
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        try:
            from oracle import gpu_reference as GR

            torch.cuda.empty_cache()
            r = GR.time_forward(SIZE, CM, B, steps=3, warmup=2)
            gpu_ref = {"unit": "frames/s", "batch": B,
                       "tf32": round(r["tf32"]["frames_per_s"], 2), "fp32": round(r["fp32"]["frames_per_s"], 2),
                       "ms_per_step_tf32": round(r["tf32"]["ms_per_step"], 2),
                       "ms_per_step_fp32": round(r["fp32"]["ms_per_step"], 2),
                       "what": "reference CUDA path on this GPU: oracle/_ref upfirdn2d + fused_bias_act extensions (the "
                               "reference's own op/*.cu) + cuDNN grouped F.conv2d / F.conv_transpose2d (groups = batch) on "
                               "materialised per-sample weights, models/stylegan2.py:217-254; cudnn.benchmark on; "
                               "tf32 = torch default cudnn.allow_tf32, fp32 = allow_tf32 off"}
        except Exception as e:  # oracle/_ref absent: say so instead of inventing a number
            gpu_ref = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # ---- audio chain: the default hooks of generate() on configs[1]'s 30 s of synthetic audio —
    # device ms per hook next to the numpy restatement on the host (bounded sample for the noise maps, stated in the keys)
    audio_chain = None
    if rank == 0 and world == 1 and not args.no_audio_chain:
        try:
            from tools import bench_audio

            torch.cuda.empty_cache()
            audio_chain = bench_audio.measure(AUDIO_S, oracle=not args.no_cpu_baseline)
        except Exception as e:
            audio_chain = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(steps=2, warmup=1, budget_s=40.0)
        cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "cpu": cpu_model()}

    if rank == 0:
        line = {"metric": "1024x1024 frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPES[args.precision] if args.conv == "tc" else "f32",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(lt.item()), "roofline": roof, "roofline_upfirdn2d": roof_ufd,
                "kernel_ms_per_step": shares, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "audio_chain": audio_chain,
                "conv_gflop_per_frame": CONV_GFLOP_PER_FRAME}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
