"""The whole-forward C ABI (maua_synth_* / csrc/synth.cu) against the per-operator orchestration (synthesis.py): both
sequence the same kernels, so images, uint8 frames and truncated latents must be BIT-IDENTICAL; plus the handle's own
contract (rebinding on a batch-size change, weight rewrites, CUDA-graph capture)."""
import numpy as np
import pytest
import torch

from tests.util import make_generator

pytestmark = pytest.mark.gpu


def _inputs(g, batch, seed, per_frame_upto=64):
    gen = torch.Generator().manual_seed(seed)
    latent = (torch.randn(batch, g.n_latent, 512, generator=gen) * 0.5).cuda()
    noise = []
    for l in range(g.num_layers):
        r = 2 ** ((l + 5) // 2)
        noise.append(torch.randn(batch, 1, r, r, generator=gen).cuda() if r <= per_frame_upto else None)
    psi = torch.linspace(0.6, 1.0, batch).cuda()
    return latent, noise, psi


def _forward(g, use_handle, latent, noise, psi, **kw):
    from maua_stylegan2_b200 import synthesis

    synthesis._USE_HANDLE = use_handle
    try:
        with torch.no_grad():
            return g(latent, noise=list(noise), truncation=psi, input_is_latent=True, randomize_noise=False, **kw)
    finally:
        synthesis._USE_HANDLE = True


@pytest.mark.parametrize("size,cm,precision", [(128, 1, "bf16x3"), (256, 2, "bf16x3"), (1024, 2, "bf16x3"), (1024, 2, "mixed"),
                                               (512, 1, "mixed")])
def test_handle_forward_is_bit_identical_to_operator_path(size, cm, precision):
    from maua_stylegan2_b200 import _lib as L

    g, _ = make_generator(size, cm, 6, "tc", precision=precision)
    g.truncation_latent = torch.randn(1, 512, device="cuda") * 0.1
    for batch in (2, 3):   # the second batch size forces a re-bind of the workspace
        latent, noise, psi = _inputs(g, batch, 40 + batch)
        l0 = L.launch_count()
        img_h, lat_h = _forward(g, True, latent, noise, psi, return_latents=True)
        n_handle = L.launch_count() - l0
        assert getattr(g, "_synth_handle", None) is not None and g._synth_handle.batch == batch
        img_o, lat_o = _forward(g, False, latent, noise, psi, return_latents=True)
        n_ops = L.launch_count() - l0 - n_handle
        assert torch.equal(img_h, img_o), f"image differs: {(img_h - img_o).abs().max().item()}"
        assert torch.equal(lat_h, lat_o)
        assert n_handle == n_ops, (n_handle, n_ops)
        u8_h, _ = _forward(g, True, latent, noise, psi, return_u8=True)
        u8_o, _ = _forward(g, False, latent, noise, psi, return_u8=True)
        assert u8_h.dtype == torch.uint8 and torch.equal(u8_h, u8_o)
        # float truncation and buffer noise everywhere (noise=None entries)
        img_h2, _ = _forward(g, True, latent, [None] * g.num_layers, 0.7)
        img_o2, _ = _forward(g, False, latent, [None] * g.num_layers, 0.7)
        assert torch.equal(img_h2, img_o2)


def test_handle_follows_weight_changes_and_graph_capture():
    g, _ = make_generator(128, 1, 8, "tc")
    g.truncation_latent = torch.zeros(1, 512, device="cuda")
    latent, noise, psi = _inputs(g, 2, 9)
    a, _ = _forward(g, True, latent, noise, psi)
    with torch.no_grad():
        g.convs[1].conv.weight.mul_(1.5)          # in-place edit bumps _version -> new handle, re-packed weights
    b, _ = _forward(g, True, latent, noise, psi)
    b_ops, _ = _forward(g, False, latent, noise, psi)
    assert not torch.equal(a, b) and torch.equal(b, b_ops)
    # capture a forward (the bind happened in the eager call above) and replay it on new inputs
    static_lat = latent.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        out, _ = g(static_lat, noise=list(noise), truncation=psi, input_is_latent=True, randomize_noise=False, return_u8=True)
    lat2, _, _ = _inputs(g, 2, 10)
    static_lat.copy_(lat2)
    graph.replay()
    torch.cuda.synchronize()
    want, _ = _forward(g, False, lat2, noise, psi, return_u8=True)
    assert torch.equal(out, want)
    # the frame loop's short tail: an eager forward at ANOTHER batch size re-binds the handle; the captured graph's
    # workspace must stay alive and valid (the handle keeps the two most recent batch sizes)
    ws2 = g._synth_handle.workspace.data_ptr()
    lat3, noise3, psi3 = _inputs(g, 3, 11)
    tail_h, _ = _forward(g, True, lat3, noise3, psi3, return_u8=True)
    tail_o, _ = _forward(g, False, lat3, noise3, psi3, return_u8=True)
    assert torch.equal(tail_h, tail_o) and g._synth_handle.batch == 3
    lat4, _, _ = _inputs(g, 2, 12)
    static_lat.copy_(lat4)
    graph.replay()
    torch.cuda.synchronize()
    want4, _ = _forward(g, False, lat4, noise, psi, return_u8=True)
    assert torch.equal(out, want4)
    _forward(g, True, lat4, noise, psi)            # eager again at batch 2: the cached workspace (same addresses) is re-bound
    assert g._synth_handle.batch == 2 and g._synth_handle.workspace.data_ptr() == ws2
