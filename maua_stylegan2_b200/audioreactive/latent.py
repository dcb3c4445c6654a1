"""Device-side drop-in for the parts of `audioreactive/latent.py` on the hot path:
`chroma_weight_latents`, `generate_latents`, `load_latents`, `save_latents`, `wrapping_slice`."""
import numpy as np
import torch as th

from .. import _lib as L


def chroma_weight_latents(chroma, latents):
    """latent.py:15-26: (chroma[..., None, None] * latents[None, ...]).sum(1) as one kernel."""
    dev = th.device("cuda", th.cuda.current_device())
    c = chroma.to(device=dev, dtype=th.float32).contiguous()
    lat = latents.to(device=dev, dtype=th.float32).contiguous()
    out = th.empty((c.shape[0],) + tuple(lat.shape[1:]), device=dev, dtype=th.float32)
    L.call("maua_chroma_weight_latents_f32", c.data_ptr(), lat.data_ptr(), out.data_ptr(), c.shape[0], lat.shape[0],
           lat[0].numel(), L.stream_ptr(dev))
    return out


def envelope_blend(x, envelope, target):
    """x = envelope[:,None,None] * target + (1 - envelope[:,None,None]) * x   (examples/default.py:20-21), in place."""
    e = envelope.to(device=x.device, dtype=th.float32).contiguous()
    tg = target.to(device=x.device, dtype=th.float32).contiguous()
    L.call("maua_envelope_blend_f32", x.data_ptr(), e.data_ptr(), tg.data_ptr(), x.shape[0], x[0].numel(),
           L.stream_ptr(x.device))
    return x


def wrapping_slice(tensor, start, length, return_indices=False):
    """latent.py:113-133"""
    if start + length <= tensor.shape[0]:
        indices = th.arange(start, start + length)
    else:
        indices = th.cat((th.arange(start, tensor.shape[0]), th.arange(0, (start + length) % tensor.shape[0])))
    if tensor.shape[0] == 1:
        indices = th.zeros(1, dtype=th.int64)
    if return_indices:
        return indices
    return tensor[indices]


def generate_latents(n_latents, ckpt, G_res, noconst=False, latent_dim=512, n_mlp=8, channel_multiplier=2):
    """latent.py:136-159: random z -> mapping network -> [n, n_latent, 512]."""
    from ..stylegan2 import Generator

    generator = Generator(G_res, latent_dim, n_mlp, channel_multiplier=channel_multiplier, constant_input=not noconst,
                          checkpoint=ckpt).cuda()
    zs = th.randn((n_latents, latent_dim), device="cuda")
    latent_selection = generator(zs, map_latents=True)
    del generator, zs
    return latent_selection


def save_latents(latents, filename):
    np.save(filename, latents.cpu() if th.is_tensor(latents) else latents)


def load_latents(filename):
    return th.from_numpy(np.load(filename))
