"""Device-side drop-in for `audioreactive/latent.py`: `chroma_weight_latents`,
`slerp`, `slerp_loops`, `spline_loops`, `wrapping_slice`, `generate_latents`, `save_latents`, `load_latents`,
`perlin_noise` — same names and arguments; sequences are built and kept on the device."""
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


def slerp(val, low, high):
    """latent.py:29-45 (host scalars / numpy vectors, as in the reference)."""
    omega = np.arccos(np.clip(np.dot(low / np.linalg.norm(low), high / np.linalg.norm(high)), -1, 1))
    so = np.sin(omega)
    if so == 0:
        return (1.0 - val) * low + val * high  # L'Hopital's rule/LERP
    return np.sin((1.0 - val) * omega) / so * low + np.sin(val * omega) / so * high


def _host(x):
    return x.detach().cpu().numpy() if th.is_tensor(x) else np.asarray(x)


def slerp_loops(latent_selection, n_frames, n_loops, smoothing=1, loop=True):
    """latent.py:48-82: geodesic interpolation between the selection's layer-0 latents, gaussian-smoothed, tiled to
    n_frames and broadcast to 18 layers.  (The reference raises a dtype error in its own gaussian_filter call — float64
    slerp output against float32 taps; here the sequence is cast to fp32 and filtered on the device.)"""
    from .signal import gaussian_filter

    sel = _host(latent_selection).astype(np.float64)
    if loop:
        sel = np.concatenate([sel, sel[[0]]])
    per = int(n_frames // max(1, n_loops) // len(sel))
    vals = np.linspace(0.0, 1.0, per)
    base = np.stack([slerp(v, sel[n % len(sel)][0], sel[(n + 1) % len(sel)][0]) for n in range(len(sel)) for v in vals])
    base = gaussian_filter(th.from_numpy(base.astype(np.float32)), smoothing)
    base = th.cat([base] * int(n_frames / len(base)), axis=0)
    base = th.cat([base[:, None, :]] * 18, axis=1)
    if n_frames - len(base) != 0:
        base = th.cat([base, base[0:n_frames - len(base)]])
    return base


def spline_weights(n_points, n_out):
    """[n_out, n_points] float64 matrix W with  W @ y == splev(linspace(0,1,n_out), splrep(linspace(0,1,n_points), y))
    for every column y: FITPACK's interpolating cubic spline (s=0) is linear in y, and its knot vector
    (x0 x4, x[2:-2], x_end x4) is the not-a-knot one of `make_interp_spline`."""
    from scipy import interpolate

    xs = np.linspace(0, 1, n_points)
    k = min(3, n_points - 1)
    if n_points <= 3:
        raise ValueError("spline_loops needs at least 4 latents (FITPACK: m > k)")
    spl = interpolate.make_interp_spline(xs, np.eye(n_points), k=k)
    return spl(np.linspace(0, 1, n_out))


def spline_loops(latent_selection, n_frames, n_loops, loop=True):
    """latent.py:85-110.  The reference fits layers*512 scalar splines in a Python loop; the fit is linear in the data, so
    the whole sequence is ONE [loop_len, n_sel] x [n_sel, layers*512] mixing launch (the chroma-weighting kernel)."""
    dev = th.device("cuda", th.cuda.current_device())
    sel = latent_selection if th.is_tensor(latent_selection) else th.from_numpy(np.asarray(latent_selection))
    sel = sel.to(device=dev, dtype=th.float32)
    if loop:
        sel = th.cat([sel, sel[[0]]])
    loop_len = int(n_frames // max(1, n_loops))
    w = th.from_numpy(spline_weights(sel.shape[0], loop_len).astype(np.float32))
    base = chroma_weight_latents(w, sel)
    base = th.cat([base] * int(n_frames / len(base)), axis=0)
    if n_frames - len(base) > 0:
        base = th.cat([base, base[0:n_frames - len(base)]])
    return base[:n_frames]


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


def _perlinterpolant(t):
    return t * t * t * (t * (t * 6 - 15) + 10)


def perlin_noise(shape, res, tileable=(True, False, False), interpolant=_perlinterpolant, dtype=th.float64):
    """latent.py:188-246: 3-D Perlin noise [shape] with `res` periods per axis, values stretched to [-1, 1].
    Lattice gradients come from np.random exactly like the reference (theta, then phi: same seed -> same noise); every
    voxel is then evaluated in one kernel (maua_perlin_noise) instead of ~40 full-size float64 temporaries.
    `dtype`: float64 like the reference (default) or float32 (half the HBM bytes; arithmetic stays fp64)."""
    if interpolant is not _perlinterpolant:
        raise L.MauaError("perlin_noise: only the default interpolant t^3(t(6t-15)+10) is compiled into the kernel")
    if any(s % r for s, r in zip(shape, res)):
        raise ValueError("shape must be a multiple of res")
    dev = th.device("cuda", th.cuda.current_device())
    theta = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    phi = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    gradients = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        gradients[-1, :, :] = gradients[0, :, :]
    if tileable[1]:
        gradients[:, -1, :] = gradients[:, 0, :]
    if tileable[2]:
        gradients[:, :, -1] = gradients[:, :, 0]
    g = th.from_numpy(np.ascontiguousarray(gradients)).to(dev)
    if dtype not in (th.float64, th.float32):
        raise ValueError("dtype must be float64 or float32")
    out = th.empty(tuple(int(s) for s in shape), device=dev, dtype=dtype)
    L.call("maua_perlin_noise", g.data_ptr(), out.data_ptr(), int(shape[0]), int(shape[1]), int(shape[2]), int(res[0]),
           int(res[1]), int(res[2]), 1 if dtype == th.float64 else 0, L.stream_ptr(dev))
    return out


def save_latents(latents, filename):
    np.save(filename, latents.cpu() if th.is_tensor(latents) else latents)


def load_latents(filename):
    return th.from_numpy(np.load(filename))
