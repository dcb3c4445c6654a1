"""GPU comparator: the reference's OWN CUDA path for the synthesis forward, timed on the same B200.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `gpu_reference` leg and tests): never on the product path.

What the reference runs on a GPU: `models/stylegan2.py:217-254` materialises the
per-sample modulated weights and calls cuDNN grouped convs (`F.conv2d` / `F.conv_transpose2d`, groups = batch), its
`op/upfirdn2d_kernel.cu` for Blur / Upsample (`op/upfirdn2d.py:88-156`) and `op/fused_bias_act_kernel.cu` for
FusedLeakyReLU (`op/fused_act.py:82-97`).  Here that is the oracle's functional restatement of the forward
(oracle/stylegan2_oracle.py, pinned to the reference's goldens) executed on CUDA tensors with its two operators swapped
for the reference's compiled extensions from oracle/_ref (built by oracle/build_ref.py from the sources where they lie).
torch's default `cudnn.allow_tf32 = True` makes the reference's convs TF32 on this GPU; both settings are timed.
"""
import contextlib

import torch

from . import stylegan2_oracle as O
from .build_ref import load_ref


class RefOps:
    def __init__(self):
        self.ufd = load_ref("upfirdn2d_ref")
        self.fused = load_ref("fused_ref")
        if self.ufd is None or self.fused is None:
            raise RuntimeError("oracle/_ref is not built (oracle/build_ref.py needs /root/reference)")

    def upfirdn2d(self, x, kernel, up=1, down=1, pad=(0, 0)):
        """op/upfirdn2d.py:145-156 -> UpFirDn2d.forward (:88-142): [N,C,H,W] viewed as [N*C,H,W,1]."""
        n, c, h, w = x.shape
        out = self.ufd.upfirdn2d(x.reshape(-1, h, w, 1), kernel.to(x.device), up, up, down, down, pad[0], pad[1], pad[0],
                                 pad[1])
        return out.view(n, c, out.shape[1], out.shape[2])

    def fused_leaky_relu(self, x, bias, negative_slope=0.2, scale=2 ** 0.5):
        """op/fused_act.py:96-97 -> FusedLeakyReLUFunction.forward (:56-61)."""
        return self.fused.fused_bias_act(x.contiguous(), bias.to(x.device), x.new_empty(0), 3, 0, negative_slope, scale)


@contextlib.contextmanager
def reference_cuda_ops(ops=None):
    """Inside the block the oracle's generator_forward uses the reference's compiled CUDA operators."""
    ops = ops or RefOps()
    saved = (O.upfirdn2d, O.fused_leaky_relu)
    O.upfirdn2d, O.fused_leaky_relu = ops.upfirdn2d, ops.fused_leaky_relu
    try:
        yield ops
    finally:
        O.upfirdn2d, O.fused_leaky_relu = saved


def forward(sd_cuda, size, latent, noise, truncation, truncation_latent, channel_multiplier=2, allow_tf32=True, ops=None):
    """One reference-CUDA-path forward (image, activation maps); all tensors on the GPU."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    try:
        with reference_cuda_ops(ops), torch.no_grad():
            return O.generator_forward(sd_cuda, size, latent, noise, truncation, truncation_latent,
                                       channel_multiplier=channel_multiplier)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def time_forward(size, channel_multiplier, batch, steps=3, warmup=2, seed=0, device="cuda"):
    """frames/s of the reference CUDA path at `batch`, TF32 on and off (CUDA events, cudnn.benchmark on like render.py:11)."""
    import numpy as np

    ops = RefOps()
    sd = {k: v.to(device) for k, v in O.synth_state_dict(size, channel_multiplier=channel_multiplier, seed=seed).items()}
    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(1))
    latent = torch.from_numpy(rng.standard_normal((batch, n_latent, 512)).astype(np.float32)).to(device) * 0.5
    noise = [torch.randn(batch, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2), device=device) if 2 ** ((l + 5) // 2) <= 256
             else None for l in range(num_layers)]
    tl = torch.zeros(1, 512, device=device)
    psi = torch.ones(batch, device=device)
    prev_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    out = {}
    try:
        for tf32 in (True, False):
            for _ in range(warmup):
                forward(sd, size, latent, noise, psi, tl, channel_multiplier, tf32, ops)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                img, _ = forward(sd, size, latent, noise, psi, tl, channel_multiplier, tf32, ops)
                u8 = ((img.clamp(-1, 1) + 1) * 127.5).permute(0, 2, 3, 1).to(torch.uint8)   # render.py:40-43 on device
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out["tf32" if tf32 else "fp32"] = {"frames_per_s": batch / (ms * 1e-3), "ms_per_step": ms}
            del img, u8
    finally:
        torch.backends.cudnn.benchmark = prev_bench
    torch.cuda.empty_cache()
    return out
