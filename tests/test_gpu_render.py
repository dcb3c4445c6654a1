"""Frame loop: CUDA-graph replay must produce byte-identical frames to the eager path, in order, including the padded
tail; frames through FramePipeline equal a direct Generator call."""
import numpy as np
import pytest
import torch

from tests.util import make_generator

pytestmark = pytest.mark.gpu


def _inputs(g, n_frames, seed=0):
    gen = torch.Generator().manual_seed(seed)
    latents = torch.randn(n_frames, g.n_latent, 512, generator=gen) * 0.5
    noise = []
    for l in range(g.num_layers):
        r = 2 ** ((l + 5) // 2)
        noise.append(torch.randn(n_frames, 1, r, r, generator=gen) if r <= 16 else None)
    return latents, noise


def test_graph_replay_equals_eager_and_direct_forward():
    from maua_stylegan2_b200.render import FramePipeline

    g, _ = make_generator(32, 2, 3, "tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    n_frames, batch = 29, 4          # 7 full batches + a tail of 1
    latents, noise = _inputs(g, n_frames)
    psi = torch.linspace(0.5, 1.0, n_frames)
    out = {}
    for mode in (True, False):
        frames = []
        pipe = FramePipeline(g, latents, list(noise), batch, truncation=psi, use_graph=mode)
        pipe.warmup()
        with torch.no_grad():
            pipe.run(lambda f: frames.append(f.copy()))
        out[mode] = np.concatenate(frames)
        assert out[mode].shape == (n_frames, 32, 32, 3)
        assert pipe.use_graph == mode
    assert np.array_equal(out[True], out[False]), "graph replay differs from eager"
    with torch.no_grad():
        direct, _ = g(latents[8:12].cuda(), noise=[n[8:12].cuda() if n is not None else None for n in noise],
                      truncation=psi[8:12].cuda(), input_is_latent=True, randomize_noise=False, return_u8=True)
    assert np.array_equal(out[True][8:12], direct.cpu().numpy())


def test_eager_loop_gpu_bound_with_host_noise_equals_direct_forward():
    """Eager frame loop (use_graph=False: what bends / rewrites / randomize_noise select) on a generator large enough to be
    GPU-bound (512^2: the forward takes longer than staging the next two batches), with PER-FRAME host noise at every
    resolution.  The H2D staging of batch i+2 must not reuse device memory that batch i's forward still reads (the
    staged tensors are allocated on the copy stream; round-1 advisor finding) — every batch must equal a direct forward."""
    from maua_stylegan2_b200.render import FramePipeline

    size, batch, n_frames = 512, 4, 24
    g, _ = make_generator(size, 1, 4, "tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    gen = torch.Generator().manual_seed(5)
    latents = torch.randn(n_frames, g.n_latent, 512, generator=gen) * 0.5
    noise = [torch.randn(n_frames, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2), generator=gen) for l in range(g.num_layers)]
    frames = []
    pipe = FramePipeline(g, latents, list(noise), batch, truncation=1.0, use_graph=False)
    pipe.warmup()
    with torch.no_grad():
        pipe.run(lambda f: frames.append(f.copy()))
    got = np.concatenate(frames)
    assert got.shape == (n_frames, size, size, 3)
    with torch.no_grad():
        for n in range(0, n_frames, batch):
            direct, _ = g(latents[n:n + batch].cuda(), noise=[x[n:n + batch].cuda() for x in noise], truncation=1.0,
                          input_is_latent=True, randomize_noise=False, return_u8=True)
            assert np.array_equal(got[n:n + batch], direct.cpu().numpy()), f"batch at frame {n} was corrupted"
