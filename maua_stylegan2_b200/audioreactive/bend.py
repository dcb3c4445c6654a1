"""Network bends — drop-in for the reference's `audioreactive/bend.py` without kornia.

Same classes and constructor arguments (`NetworkBend`, `AddNoise`, `Print`, `Translate`, `Zoom`, `Rotate`).  The
reference builds nn.Sequential(ReflectionPad2d..., AddNoise, kornia transform, CenterCrop) (bend.py:61-102): up to six
full passes over a tensor padded to 5-9x its size.  Here the whole chain is ONE launch of `maua_bend_warp_f32`
(csrc/bend.cu), evaluated per output pixel; `self.sequential` is that fused module.  kornia is un-vendored and absent,
so its arithmetic is restated from its published semantics — parity unpinned (oracle/plugin_oracle.py header).
"""
import ctypes as C
import math

import torch as th

from .. import _lib as L


class NetworkBend(th.nn.Module):
    """bend.py:12-25: `sequential_fn(modulation)` builds the module applied to the intermediate features."""

    def __init__(self, sequential_fn, modulation):
        super().__init__()
        self.sequential = sequential_fn(modulation)

    def forward(self, x):
        return self.sequential(x)


class AddNoise(th.nn.Module):
    """bend.py:28-40"""

    def __init__(self, noise):
        super().__init__()
        self.noise = noise

    def forward(self, x):
        return x + self.noise.to(x.device)


class Print(th.nn.Module):
    """bend.py:43-48"""

    def forward(self, x):
        print(x.shape, [x.min().item(), x.mean().item(), x.max().item()], th.std(x).item())
        return x


class FusedWarp(th.nn.Module):
    """pad stages (reflect | replicate) -> + noise -> affine warp (bilinear, zeros) -> centre crop, in one kernel.

    pads_x / pads_y: lists of (before, after) pads applied in order; minv_fn(Hp, Wp) -> [B,2,3] INVERSE affine on the
    device (padded-frame pixel coordinates, x first)."""

    def __init__(self, pads_x, pads_y, noise, minv_fn, out_hw, pad_mode="reflect"):
        super().__init__()
        self.pads_x, self.pads_y = list(pads_x), list(pads_y)
        self.noise, self.minv_fn, self.out_hw = noise, minv_fn, out_hw
        self.pad_mode = {"reflect": 0, "replicate": 1}[pad_mode]

    def forward(self, x):
        if not x.is_cuda:
            raise L.MauaError("network bends run on the CUDA path only (no CPU fallback)")
        x = x.float().contiguous()
        B, Cn, H, W = x.shape
        Hp = H + sum(a + b for a, b in self.pads_y)
        Wp = W + sum(a + b for a, b in self.pads_x)
        oh, ow = self.out_hw
        y0, x0 = int(Hp / 2 - oh / 2), int(Wp / 2 - ow / 2)   # kornia CenterCrop
        minv = self.minv_fn(Hp, Wp).to(device=x.device, dtype=th.float32).reshape(-1, 6).contiguous()
        if minv.shape[0] != B:
            raise L.MauaError(f"bend modulation has {minv.shape[0]} rows for a batch of {B}")
        noise, nb, nc = None, 1, 1
        if self.noise is not None:
            noise = self.noise.to(device=x.device, dtype=th.float32)
            while noise.dim() < 4:
                noise = noise[None]
            if noise.shape[0] not in (1, B) or noise.shape[1] not in (1, Cn) or tuple(noise.shape[2:]) != (Hp, Wp):
                noise = noise.expand(B if noise.shape[0] != 1 else 1, Cn if noise.shape[1] != 1 else 1, Hp, Wp)
            noise = noise.contiguous()
            nb, nc = noise.shape[0], noise.shape[1]
        px = (C.c_int * max(2 * len(self.pads_x), 1))(*[v for p in self.pads_x for v in p])
        py = (C.c_int * max(2 * len(self.pads_y), 1))(*[v for p in self.pads_y for v in p])
        y = th.empty((B, Cn, oh, ow), device=x.device, dtype=th.float32)
        L.call("maua_bend_warp_f32", x.data_ptr(), y.data_ptr(), L.ptr(noise), minv.data_ptr(), B, Cn, H, W, px,
               len(self.pads_x), py, len(self.pads_y), self.pad_mode, nb, nc, oh, ow, y0, x0, L.stream_ptr(x.device))
        return y


def _rows(mod, cols):
    m = th.as_tensor(mod).float()
    return m.reshape(m.shape[0], -1)[:, :cols] if m.dim() > 0 else m.reshape(1, 1)


class Translate(NetworkBend):
    """bend.py:51-71: reflect out to 5x width, add noise, translate by `modulation` [B,2] pixels (x, y), centre crop.
    kT.Translate: out(x, y) = in(x - tx, y - ty)."""

    def __init__(self, modulation, h, w, noise):
        def sequential_fn(b):
            t = _rows(b, 2)

            def minv(Hp, Wp):
                m = th.zeros(t.shape[0], 2, 3, device=t.device)
                m[:, 0, 0] = m[:, 1, 1] = 1
                m[:, 0, 2], m[:, 1, 2] = -t[:, 0], -t[:, 1]
                return m

            return FusedWarp([(int(w / 2), int(w / 2)), (w, w), (w, 0)], [], noise, minv, (h, w))

        super().__init__(sequential_fn, modulation)


class Zoom(NetworkBend):
    """bend.py:73-85: ReflectionPad2d(max(h, w) - 1), kT.Scale(modulation) about the centre, centre crop."""

    def __init__(self, modulation, h, w):
        padding = int(max(h, w)) - 1

        def sequential_fn(b):
            s = _rows(b, 2)

            def minv(Hp, Wp):
                sx, sy = s[:, 0], s[:, -1]
                cx, cy = (Wp - 1) / 2, (Hp - 1) / 2
                m = th.zeros(s.shape[0], 2, 3, device=s.device)
                m[:, 0, 0], m[:, 1, 1] = 1 / sx, 1 / sy
                m[:, 0, 2], m[:, 1, 2] = cx * (1 - 1 / sx), cy * (1 - 1 / sy)
                return m

            return FusedWarp([(padding, padding)], [(padding, padding)], None, minv, (h, w))

        super().__init__(sequential_fn, modulation)


class Rotate(NetworkBend):
    """bend.py:88-102: ReflectionPad2d(int(max(h, w) * (1 - sqrt(2) / 2))), kT.Rotate(modulation degrees,
    anti-clockwise about the centre), centre crop."""

    def __init__(self, modulation, h, w):
        padding = int(max(h, w) * (1 - math.sqrt(2) / 2))

        def sequential_fn(b):
            a = _rows(b, 1)[:, 0] * (math.pi / 180)

            def minv(Hp, Wp):
                al, be = th.cos(a), th.sin(a)
                cx, cy = (Wp - 1) / 2, (Hp - 1) / 2
                m = th.zeros(a.shape[0], 2, 3, device=a.device)
                m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = al, -be, cx - al * cx + be * cy
                m[:, 1, 0], m[:, 1, 1], m[:, 1, 2] = be, al, cy - be * cx - al * cy
                return m

            return FusedWarp([(padding, padding)], [(padding, padding)], None, minv, (h, w))

        super().__init__(sequential_fn, modulation)
