"""TEST INFRASTRUCTURE — CPU restatement of the reference's plugin-side helpers that §8(f) brings onto the device:
`perlin_noise`, `spline_loops`, `slerp`, `slerp_loops` (audioreactive/latent.py) and the Translate / Zoom / Rotate network
bends (audioreactive/bend.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Pinned: perlin_noise / spline_loops / slerp against tests/golden/plugins.npz (written by the UNMODIFIED reference,
tests/golden/make_golden.py --plugins).
PARITY UNPINNED: the bends.  kornia (kT.Translate / kT.Scale / kT.Rotate, kA.CenterCrop — un-vendored, unpinned in
requirements.txt:1-11, absent here) holds their arithmetic; this file restates kornia's published semantics
(warp_affine: dst = M·src, bilinear, zeros padding, align_corners=True; get_rotation_matrix2d = OpenCV convention,
centre ((W-1)/2, (H-1)/2); center crop start = int(src/2 - dst/2)) through torch's own F.pad / F.grid_sample.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy import interpolate


# ---- audioreactive/latent.py:184-246 ----------------------------------------------------------------------------------
def perlin_gradients(res, tileable=(True, False, False)):
    """latent.py:209-218: lattice gradients from np.random (theta first, then phi), tile wrap applied."""
    theta = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    phi = 2 * np.pi * np.random.rand(res[0] + 1, res[1] + 1, res[2] + 1)
    g = np.stack((np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)), axis=3)
    if tileable[0]:
        g[-1, :, :] = g[0, :, :]
    if tileable[1]:
        g[:, -1, :] = g[:, 0, :]
    if tileable[2]:
        g[:, :, -1] = g[:, :, 0]
    return g


def perlin_noise(shape, res, gradients):
    """latent.py:203-246 with the lattice gradients given (float64 throughout, same operation order)."""
    delta = (res[0] / shape[0], res[1] / shape[1], res[2] / shape[2])
    d = (shape[0] // res[0], shape[1] // res[1], shape[2] // res[2])
    grid = np.mgrid[0:res[0]:delta[0], 0:res[1]:delta[1], 0:res[2]:delta[2]].transpose(1, 2, 3, 0) % 1
    g = gradients.repeat(d[0], 0).repeat(d[1], 1).repeat(d[2], 2)
    sl = [(slice(None, -d[0]), slice(d[0], None)), (slice(None, -d[1]), slice(d[1], None)),
          (slice(None, -d[2]), slice(d[2], None))]
    n = {}
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                gg = g[sl[0][a], sl[1][b], sl[2][c]]
                n[a, b, c] = ((grid[..., 0] - a) * gg[..., 0] + (grid[..., 1] - b) * gg[..., 1]) + (grid[..., 2] - c) * gg[..., 2]
    t = grid * grid * grid * (grid * (grid * 6 - 15) + 10)
    n00 = n[0, 0, 0] * (1 - t[..., 0]) + t[..., 0] * n[1, 0, 0]
    n10 = n[0, 1, 0] * (1 - t[..., 0]) + t[..., 0] * n[1, 1, 0]
    n01 = n[0, 0, 1] * (1 - t[..., 0]) + t[..., 0] * n[1, 0, 1]
    n11 = n[0, 1, 1] * (1 - t[..., 0]) + t[..., 0] * n[1, 1, 1]
    n0 = (1 - t[..., 1]) * n00 + t[..., 1] * n10
    n1 = (1 - t[..., 1]) * n01 + t[..., 1] * n11
    return ((1 - t[..., 2]) * n0 + t[..., 2] * n1) * 2 - 1


# ---- audioreactive/latent.py:29-110 -----------------------------------------------------------------------------------
def slerp(val, low, high):
    """latent.py:29-45"""
    omega = np.arccos(np.clip(np.dot(low / np.linalg.norm(low), high / np.linalg.norm(high)), -1, 1))
    so = np.sin(omega)
    if so == 0:
        return (1.0 - val) * low + val * high
    return np.sin((1.0 - val) * omega) / so * low + np.sin(val * omega) / so * high


def gaussian_filter_np(x, sigma, smf=1.0):
    """signal.py:319-368 (symmetric case) in float64 numpy: circular correlation along axis 0."""
    T = x.shape[0]
    radius = min(int(sigma * 4 * smf), 3 * T)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    g = np.exp(-0.5 / sigma ** 2 * k ** 2)
    g /= g.sum()
    y = np.zeros_like(x, dtype=np.float64)
    if radius <= T:
        for j, kk in enumerate(range(-radius, radius + 1)):
            y += g[j] * np.roll(x, -kk, axis=0)
    else:
        xp = np.concatenate([np.zeros((radius - T,) + x.shape[1:]), x, x, x, np.zeros((radius - T,) + x.shape[1:])])
        for t in range(T):
            y[t] = np.tensordot(g, xp[t:t + 2 * radius + 1], axes=(0, 0))
    return y


def slerp_loops(latent_selection, n_frames, n_loops, smoothing=1, loop=True, n_layers=18, smf=1.0):
    """latent.py:48-82.  The reference crashes in its gaussian_filter (float64 latents vs float32 taps);
    this is the evident intent: the same sequence with the filter applied in one dtype."""
    sel = np.asarray(latent_selection, dtype=np.float64)
    if loop:
        sel = np.concatenate([sel, sel[[0]]])
    base = []
    for n in range(len(sel)):
        for val in np.linspace(0.0, 1.0, int(n_frames // max(1, n_loops) // len(sel))):
            base.append(slerp(val, sel[n % len(sel)][0], sel[(n + 1) % len(sel)][0]))
    base = gaussian_filter_np(np.stack(base), smoothing, smf)
    base = np.concatenate([base] * int(n_frames / len(base)), axis=0)
    base = np.concatenate([base[:, None, :]] * n_layers, axis=1)
    if n_frames - len(base) != 0:
        base = np.concatenate([base, base[0:n_frames - len(base)]])
    return base


def spline_loops(latent_selection, n_frames, n_loops, loop=True):
    """latent.py:85-110: one interpolating cubic B-spline (FITPACK splrep, s=0) per latent coordinate."""
    sel = np.asarray(latent_selection)
    if loop:
        sel = np.concatenate([sel, sel[[0]]])
    x = np.linspace(0, 1, int(n_frames // max(1, n_loops)))
    base = np.zeros((len(x), *sel.shape[1:]))
    xs = np.linspace(0, 1, sel.shape[0])
    for lay in range(sel.shape[1]):
        for lat in range(sel.shape[2]):
            base[:, lay, lat] = interpolate.splev(x, interpolate.splrep(xs, sel[:, lay, lat]))
    base = np.concatenate([base] * int(n_frames / len(base)), axis=0)
    if n_frames - len(base) > 0:
        base = np.concatenate([base, base[0:n_frames - len(base)]])
    return base[:n_frames]


# ---- audioreactive/bend.py:51-102 -------------------------------------------------------------------------------------
def _warp_crop(p, m_fwd, out_hw):
    """kornia warp_affine(p, M, dsize=p.shape[-2:], bilinear, zeros, align_corners=True) then CenterCrop(out_hw)."""
    B, C, Hp, Wp = p.shape
    m = torch.cat([m_fwd.double(), torch.tensor([[[0.0, 0.0, 1.0]]], dtype=torch.float64).expand(B, 1, 3)], 1)
    minv = torch.linalg.inv(m)[:, :2]                                      # dst pixel -> src pixel
    ys, xs = torch.meshgrid(torch.arange(Hp, dtype=torch.float64), torch.arange(Wp, dtype=torch.float64), indexing="ij")
    dst = torch.stack([xs, ys, torch.ones_like(xs)], -1).reshape(1, -1, 3)  # [1, Hp*Wp, 3]
    src = dst @ minv.transpose(1, 2)                                        # [B, Hp*Wp, 2]
    gx = src[..., 0] / max(Wp - 1, 1) * 2 - 1
    gy = src[..., 1] / max(Hp - 1, 1) * 2 - 1
    grid = torch.stack([gx, gy], -1).reshape(B, Hp, Wp, 2)
    out = F.grid_sample(p.double(), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    h, w = out_hw
    y0, x0 = int(Hp / 2 - h / 2), int(Wp / 2 - w / 2)
    return out[:, :, y0:y0 + h, x0:x0 + w].float()


def translate(x, translation, h, w, noise):
    """bend.py:61-71: three reflection pads (to 5x width), AddNoise, kT.Translate(b), CenterCrop((h, w))."""
    p = F.pad(x, (int(w / 2), int(w / 2), 0, 0), mode="reflect")
    p = F.pad(p, (w, w, 0, 0), mode="reflect")
    p = F.pad(p, (w, 0, 0, 0), mode="reflect")
    p = p + noise
    B = x.shape[0]
    m = torch.zeros(B, 2, 3, dtype=torch.float64)
    m[:, 0, 0] = m[:, 1, 1] = 1
    m[:, :, 2] = translation.double()
    return _warp_crop(p, m, (h, w))


def zoom(x, scale, h, w):
    """bend.py:82-85: ReflectionPad2d(max(h, w) - 1), kT.Scale(b) about the centre, CenterCrop((h, w))."""
    pad = int(max(h, w)) - 1
    p = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    B, _, Hp, Wp = p.shape
    s = scale.double().reshape(B, -1)
    sx, sy = s[:, 0], s[:, -1]
    cx, cy = (Wp - 1) / 2, (Hp - 1) / 2
    m = torch.zeros(B, 2, 3, dtype=torch.float64)
    m[:, 0, 0], m[:, 1, 1] = sx, sy
    m[:, 0, 2], m[:, 1, 2] = (1 - sx) * cx, (1 - sy) * cy
    return _warp_crop(p, m, (h, w))


def rotate(x, angle_deg, h, w):
    """bend.py:97-102: ReflectionPad2d(int(max(h, w) * (1 - sqrt(2)/2))), kT.Rotate(b) (degrees, anti-clockwise,
    OpenCV matrix), CenterCrop((h, w))."""
    pad = int(max(h, w) * (1 - math.sqrt(2) / 2))
    p = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    B, _, Hp, Wp = p.shape
    a = angle_deg.double().reshape(B) * math.pi / 180
    al, be = torch.cos(a), torch.sin(a)
    cx, cy = (Wp - 1) / 2, (Hp - 1) / 2
    m = torch.zeros(B, 2, 3, dtype=torch.float64)
    m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = al, be, (1 - al) * cx - be * cy
    m[:, 1, 0], m[:, 1, 1], m[:, 1, 2] = -be, al, be * cx + (1 - al) * cy
    return _warp_crop(p, m, (h, w))
