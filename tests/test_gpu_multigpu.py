"""Hardware multi-GPU test: frames rendered by 2 ranks (rank-strided batches, one NCCL
all-gather of uint8 frames per step) must be BYTE-IDENTICAL to the single-GPU run, in order, padded tail trimmed.
Skipped on boxes with fewer than 2 GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_inputs(g, n_frames, seed=11):
    gen = torch.Generator().manual_seed(seed)
    latents = torch.randn(n_frames, g.n_latent, 512, generator=gen) * 0.5
    noise = []
    for l in range(g.num_layers):
        r = 2 ** ((l + 5) // 2)
        noise.append(torch.randn(n_frames, 1, r, r, generator=gen) if r <= 64 else None)
    psi = torch.linspace(0.6, 1.0, n_frames)
    return latents, noise, psi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,mode", [(2, "render"), (2, "gather")])
def test_nccl_sharded_frames_equal_single_gpu(tmp_path, world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    from maua_stylegan2_b200.render import FramePipeline
    from tests.util import make_generator

    size, cm, n_frames, batch = 512, 1, 45, 4     # 11 full batches + a tail of 1: uneven over 2 ranks
    out_file = str(tmp_path / "frames.npy")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_frames_worker.py"), out_file,
           str(size), str(cm), str(n_frames), str(batch), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    multi = np.load(out_file)

    g, _ = make_generator(size, cm, 4, "tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    latents, noise, psi = make_inputs(g, n_frames)
    frames = []
    pipe = FramePipeline(g, latents, list(noise), batch, truncation=psi)
    pipe.warmup()
    with torch.no_grad():
        pipe.run(lambda f: frames.append(f.copy()))
    single = np.concatenate(frames)
    assert single.shape == (n_frames, size, size, 3) == multi.shape
    assert np.array_equal(single, multi), "NCCL-sharded frames differ from the single-GPU run"
