"""Operator API of the reference's `op/` package (op/__init__.py:1-2) on top of libmaua_b200.so.

    upfirdn2d(input[N,C,H,W], kernel[kh,kw], up=1, down=1, pad=(p0,p1))      op/upfirdn2d.py:145-156
    fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)           op/fused_act.py:86-97
    FusedLeakyReLU(channel, negative_slope=0.2, scale=2**0.5)  (.bias)        op/fused_act.py:74-83
    fused_bias_act(input, bias, refer, act, grad, alpha, scale)               op/fused_bias_act.cpp:11-21
    upfirdn2d_raw(input[major,H,W,minor], kernel, up_x, ..., pad_y1)          op/upfirdn2d.cpp:12-23

Inference only (no autograd; the reference renders under `torch.set_grad_enabled(False)`, render.py:10).
CUDA tensors only: this package has no CPU path — a CPU tensor raises instead of silently computing elsewhere.
"""
import torch
from torch import nn

from . import _lib as L


def _require_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise L.MauaError(f"{what}: expected a CUDA tensor (this build has no CPU path), got "
                          f"{t.device if torch.is_tensor(t) else type(t)}")
    if t.dtype != torch.float32:
        raise L.MauaError(f"{what}: only float32 is supported, got {t.dtype}")


def upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Native-ABI form: input [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]."""
    _require_cuda(input, "upfirdn2d")
    _require_cuda(kernel, "upfirdn2d(kernel)")
    x = input.contiguous()
    k = kernel.contiguous()
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    y = torch.empty((major, max(out_h, 0), max(out_w, 0), minor), device=x.device, dtype=x.dtype)
    if y.numel():
        with torch.cuda.device(x.device):
            L.call("maua_upfirdn2d_f32", x.data_ptr(), y.data_ptr(), k.data_ptr(), major, in_h, in_w, minor, kh, kw,
                   up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, L.stream_ptr(x.device))
    return y


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    n, c, in_h, in_w = input.shape
    out = upfirdn2d_raw(input.reshape(-1, in_h, in_w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.view(n, c, out.shape[1], out.shape[2])


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    _require_cuda(input, "fused_bias_act")
    x = input.contiguous()
    b = bias.contiguous() if (bias is not None and bias.numel()) else None
    r = refer.contiguous() if (refer is not None and refer.numel()) else None
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.size(i)
    y = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            L.call("maua_fused_bias_act_f32", x.data_ptr(), L.ptr(b), L.ptr(r), y.data_ptr(), x.numel(), step_b,
                   b.numel() if b is not None else 0, int(act), int(grad), float(alpha), float(scale),
                   L.stream_ptr(x.device))
    return y


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
