"""Python side of the whole-forward C ABI (maua_synth_* in include/maua_b200.h): one ctypes call per Generator.forward
instead of ~50.  `Generator.forward` takes this path whenever the call needs nothing but the image (no bends, no
activation maps, ConstantInput, every layer inside the tensor-core shape set); everything else stays on the per-operator
orchestration of synthesis.py.  Both paths launch the same kernels in the same order, so they are bit-identical
(tests/test_gpu_synth_handle.py)."""
import ctypes as C

import torch

from . import _lib as L
from .plan import F16_MIN_RES, tc_supported

PRECISIONS = {"bf16x3": 0, "mixed": 1, "bf16": 2}


class SynthHandle:
    def __init__(self, g):
        from .stylegan2 import ConstantInput

        if not isinstance(g.input, ConstantInput) or g.impl != "tc":
            raise L.MauaError("the whole-forward handle needs ConstantInput and impl='tc'")
        if not all(tc_supported(sp.cin, sp.cout) for sp in g._specs):
            raise L.MauaError("a layer is outside the tensor-core shape set")
        self.device = g.input.input.device
        self.key = g._plan_key()
        n = len(g._specs)
        self._layers = (L.SynthLayer * n)()
        self._keep = []   # tensors whose storage the descriptor points into

        def ptr(t):
            t = t.detach()
            if not t.is_contiguous() or t.dtype != torch.float32:
                t = t.float().contiguous()
            self._keep.append(t)
            return t.data_ptr()

        for i, sp in enumerate(g._specs):
            l, conv = self._layers[i], sp.mod.conv
            l.conv_weight, l.mod_weight, l.mod_bias = ptr(conv.weight), ptr(conv.modulation.weight), ptr(conv.modulation.bias)
            l.noise_weight, l.act_bias = ptr(sp.mod.noise.weight), ptr(sp.mod.activate.bias)
            l.noise_buffer = ptr(getattr(g.noises, f"noise_{sp.noise_index}"))
            l.blur_kernel = ptr(conv.blur.kernel) if sp.up else None
            l.cin, l.cout, l.up, l.latent_index, l.rgb_latent_index = sp.cin, sp.cout, int(sp.up), sp.latent_index, -1
            if sp.rgb is not None:
                rc = sp.rgb.conv
                l.rgb_weight, l.rgb_mod_weight, l.rgb_mod_bias = ptr(rc.weight), ptr(rc.modulation.weight), ptr(rc.modulation.bias)
                l.rgb_bias = ptr(sp.rgb.bias)
                l.rgb_up_kernel = ptr(sp.rgb.upsample.kernel) if hasattr(sp.rgb, "upsample") else None
                l.rgb_latent_index = sp.rgb_latent_index
        desc = L.SynthDesc()
        desc.layers = C.cast(self._layers, C.POINTER(L.SynthLayer))
        desc.const_input = ptr(g.input.input)
        desc.n_layers, desc.in_h, desc.in_w = n, g.input.input.shape[2], g.input.input.shape[3]
        desc.style_dim, desc.n_latent = g.style_dim, g.n_latent
        desc.precision, desc.f16_min_res, desc.min_rgb_size = PRECISIONS[g.precision], F16_MIN_RES, g.min_rgb_size
        self.n_layers = n
        self.h = C.c_void_p()
        L.check(L.lib().maua_synth_create(C.byref(desc), C.byref(self.h)), "maua_synth_create")
        with torch.cuda.device(self.device):
            self.plan = torch.empty(L.lib().maua_synth_plan_bytes(self.h), dtype=torch.uint8, device=self.device)
            L.call("maua_synth_prepare", self.h, self.plan.data_ptr(), self.plan.numel(), L.stream_ptr(self.device))
        self.batch = 0
        self.workspace = None
        self._workspaces = {}   # batch size -> workspace tensor (at most two, most recently bound last)

    def __del__(self):
        try:
            if self.h:
                L.lib().maua_synth_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def bind(self, batch):
        """Lay the workspace out for `batch`.  The two most recent batch sizes keep their workspaces (a frame loop alternates
        between full batches — possibly baked into captured CUDA graphs — and one short tail batch): re-binding a cached
        size re-uses the SAME device addresses, so graphs captured for it stay valid."""
        if batch == self.batch:
            return
        if torch.cuda.is_current_stream_capturing():
            raise L.MauaError("maua_synth_bind inside CUDA-graph capture: run one eager forward at this batch size first")
        with torch.cuda.device(self.device):
            ws = self._workspaces.pop(batch, None)
            if ws is None:
                while len(self._workspaces) >= 2:
                    self._workspaces.pop(next(iter(self._workspaces)))   # drop the least recently bound size
                nbytes = L.lib().maua_synth_workspace_bytes(self.h, batch)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._workspaces[batch] = ws                                  # most recent last
            self.workspace = ws
            L.call("maua_synth_bind", self.h, ws.data_ptr(), ws.numel(), batch, L.stream_ptr(self.device))
        self.batch = batch

    def forward(self, latent, noise, mean, psi_t, psi_s, out_hw, want_u8, want_rgb=True):
        """latent [B,rows,D] fp32 cuda; noise: list of (tensor or None, bstride); returns (image or None, u8 or None)."""
        batch = latent.shape[0]
        self.bind(batch)
        ptrs = (C.c_void_p * self.n_layers)()
        strides = (C.c_longlong * self.n_layers)()
        for i, (t, bs) in enumerate(noise):
            ptrs[i] = t.data_ptr() if t is not None else None
            strides[i] = bs
        h, w = out_hw
        image = torch.empty((batch, 3, h, w), device=self.device, dtype=torch.float32) if want_rgb else None
        u8 = torch.empty((batch, h, w, 3), device=self.device, dtype=torch.uint8) if want_u8 else None
        with torch.cuda.device(self.device):
            L.call("maua_synth_forward", self.h, latent.data_ptr(), latent.shape[1], ptrs, strides, L.ptr(mean), L.ptr(psi_t),
                   float(psi_s), batch, L.ptr(image), L.ptr(u8), L.stream_ptr(self.device))
        return image, u8

    def truncated_latents(self, batch, rows, dim):
        """Copy of the [batch, rows, dim] truncated latents the last forward left in the workspace."""
        off = L.lib().maua_synth_truncated_latents(self.h) - self.workspace.data_ptr()
        return self.workspace[off:off + batch * rows * dim * 4].view(torch.float32).view(batch, rows, dim).clone()
