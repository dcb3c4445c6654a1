"""Operator API of the reference's `op/` package (op/__init__.py:1-2) on top of libmaua_b200.so.

    upfirdn2d(input[N,C,H,W], kernel[kh,kw], up=1, down=1, pad=(p0,p1))      op/upfirdn2d.py:145-156
    fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)           op/fused_act.py:86-97
    FusedLeakyReLU(channel, negative_slope=0.2, scale=2**0.5)  (.bias)        op/fused_act.py:74-83
    fused_bias_act(input, bias, refer, act, grad, alpha, scale)               op/fused_bias_act.cpp:11-21
    upfirdn2d_raw(input[major,H,W,minor], kernel, up_x, ..., pad_y1)          op/upfirdn2d.cpp:12-23

Inference only (no autograd; the reference renders under `torch.set_grad_enabled(False)`, render.py:10).

Device dispatch follows the reference (`op/upfirdn2d.py:146-149`, `op/fused_act.py:87-94`): the operators look at
`input.device.type`.  CUDA tensors ALWAYS go through libmaua_b200.so (a missing library raises MauaError — there is no
fallback for them); CPU tensors are evaluated by the torch expressions below, because the reference's contract is that
CPU inputs keep working (its CPU path is what its golden vectors are made of).  The two branches never substitute for
each other.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L


def _require_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise L.MauaError(f"{what}: expected a CUDA tensor, got {t.device if torch.is_tensor(t) else type(t)}")
    if t.dtype != torch.float32:
        raise L.MauaError(f"{what}: only float32 is supported, got {t.dtype}")


def _upfirdn2d_cpu(x, k, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """CPU-tensor branch: [major, H, W, minor] -> zero-stuff by `up`, pad (negative pads crop), correlate with the
    flipped kernel, keep every `down`-th sample — the definition the reference's `upfirdn2d_native` implements
    (op/upfirdn2d.py:159-200), written as one conv2d over the (major*minor) planes."""
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    planes = x.permute(0, 3, 1, 2).reshape(major * minor, 1, in_h, in_w)
    z = planes.new_zeros(major * minor, 1, in_h * up_y, in_w * up_x)
    z[:, :, ::up_y, ::up_x] = planes
    z = F.pad(z, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    z = z[:, :, max(-pad_y0, 0): z.shape[2] - max(-pad_y1, 0), max(-pad_x0, 0): z.shape[3] - max(-pad_x1, 0)]
    out = F.conv2d(z, torch.flip(k, [0, 1]).to(z.dtype).view(1, 1, kh, kw))[:, :, ::down_y, ::down_x]
    return out.reshape(major, minor, out.shape[2], out.shape[3]).permute(0, 2, 3, 1)


def upfirdn2d_raw(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Native-ABI form: input [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]."""
    if torch.is_tensor(input) and input.device.type == "cpu":
        return _upfirdn2d_cpu(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1).contiguous()
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


def _fused_bias_act_cpu(x, bias, refer, act, grad, alpha, scale):
    """CPU-tensor branch of the fused op: bias broadcast along dim 1 (op/fused_act.py:89-93), act 1 = linear, 3 = leaky
    relu; grad 0 forward, 1 gated by the sign of `refer`, 2 zeros (op/fused_bias_act_kernel.cu:18-49)."""
    if bias is not None and bias.numel():
        x = x + bias.view(1, -1, *([1] * (x.ndim - 2)))
    if grad == 2:
        return torch.zeros_like(x)
    if act == 3:
        gate = x if grad == 0 else refer
        x = torch.where(gate > 0, x, x * alpha)
    return x * scale


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    if torch.is_tensor(input) and input.device.type == "cpu":
        return _fused_bias_act_cpu(input, bias, refer, int(act), int(grad), float(alpha), float(scale))
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
