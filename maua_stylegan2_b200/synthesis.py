"""Orchestration of one Generator forward over the C ABI (the body of models/stylegan2.py:537-576).

Per batch: 1 style-prologue launch, then per StyledConv
   impl "tc":   [modulate_split if the input is fp32 NCHW]  ->  modconv_tc (epilogue fused)
                up layers: modconv_tc(up=1) -> blur_act_nhwc
   impl "simt": modconv_simt -> [upfirdn2d blur] -> noise_bias_act
and a torgb launch per resolution.  Python bends (`transform_dict_list`) are applied on real fp32 NCHW tensors
between layers, exactly where the reference applies them (ManipulationLayer, models/stylegan2.py:297-307).
"""
import ctypes as C
import os

import torch

from . import _lib as L
from .plan import batch_buffers
from .stylegan2 import SQRT2, _modconv_simt, _noise_bias_act, _prep_noise, _torgb, frames_to_u8
from .op import upfirdn2d


_FUSE_RGB = os.environ.get("MAUA_FUSE_RGB", "1") == "1"


def _bend(x, layer_id, bends):
    for t in bends:
        if t["layer"] == layer_id:
            x = t["transform"].to(x.device)(x)
    return x


def _has_bend(layer_id, bends):
    return any(t["layer"] == layer_id for t in bends)


def _modulate_split(x, bstride, s, batch, fmt="bf16x3"):
    """fp32 NCHW (optionally batch-broadcast) * s[b,c] -> (hi, lo) bf16 NHWC, or (fp16 plane, None) for fmt "f16"."""
    c, h, w = x.shape[1], x.shape[2], x.shape[3]
    if fmt == "f16":
        hi = torch.empty((batch, h, w, c), device=x.device, dtype=torch.float16)
        L.call("maua_modulate_f16_nhwc", x.data_ptr(), bstride, L.ptr(s), hi.data_ptr(), batch, c, h, w,
               L.stream_ptr(x.device))
        return hi, None
    hi = torch.empty((batch, h, w, c), device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    L.call("maua_modulate_split_nhwc", x.data_ptr(), bstride, L.ptr(s), hi.data_ptr(), lo.data_ptr(), batch, c, h, w,
           L.stream_ptr(x.device))
    return hi, lo


_USE_HANDLE = os.environ.get("MAUA_SYNTH_C", "1") == "1"


def _synthesize_handle(g, latent, noise, truncation, want_u8):
    """The whole-forward C ABI (one maua_synth_forward call, csrc/synth.cu) when the call is eligible, else None."""
    from .stylegan2 import ConstantInput
    from .synth_handle import SynthHandle

    if not (_USE_HANDLE and g.impl == "tc" and isinstance(g.input, ConstantInput) and L.PROFILE is None):
        return None
    device, batch = latent.device, latent.shape[0]
    mean = g.truncation_latent
    if mean is None or mean.numel() != g.style_dim or latent.shape[1] < g.n_latent or latent.shape[2] != g.style_dim:
        return None
    h = getattr(g, "_synth_handle", None)
    if h is None or h.key != g._plan_key():
        try:
            h = SynthHandle(g)
        except L.MauaError:
            g._synth_handle_failed = g._plan_key()
            return None
        g._synth_handle = h
    psi_t, psi_s = None, 1.0
    if torch.is_tensor(truncation):
        psi_t = truncation.to(device=device, dtype=torch.float32).contiguous()
        if psi_t.numel() == 1:
            psi_t = psi_t.reshape(1).expand(batch).contiguous()
        elif psi_t.numel() != batch:
            raise L.MauaError(f"truncation has {psi_t.numel()} entries for a batch of {batch}")
    else:
        psi_s = float(truncation)
    hh, ww = g.input.input.shape[2], g.input.input.shape[3]
    prepared = []
    for sp in g._specs:
        if sp.up:
            hh, ww = 2 * hh, 2 * ww
        nz = noise[sp.noise_index]
        if nz is None:  # randomize_noise=True: fresh N(0,1) per layer (models/stylegan2.py:263-265)
            nz = torch.randn(batch, 1, hh, ww, device=device)
        prepared.append(_prep_noise(nz, device, batch, hh * ww))
    mean = mean.to(device=device, dtype=torch.float32).contiguous()
    image, u8 = h.forward(latent, prepared, mean, psi_t, psi_s, (hh, ww), want_u8, want_rgb=not want_u8)
    latent_t = _LazyLatents(h, batch, latent.shape[1], g.style_dim)
    return (u8 if want_u8 else image), latent_t, []


class _LazyLatents:
    """Truncated latents of a handle forward, copied out of the workspace only if the caller asks (return_latents)."""

    def __init__(self, h, batch, rows, dim):
        self.args = (h, batch, rows, dim)

    def get(self):
        h, batch, rows, dim = self.args
        return h.truncated_latents(batch, rows, dim)


def synthesize(g, latent, noise, truncation, bends, want_acts=False, want_u8=False):
    device = latent.device
    batch = latent.shape[0]
    if not bends and not want_acts:
        fast = _synthesize_handle(g, latent, noise, truncation, want_u8)
        if fast is not None:
            return fast
    plan = g._get_plan()
    with torch.cuda.device(device):
        bc = batch_buffers(g, plan, batch)
        stream = L.stream_ptr(device)

        # ---- style prologue: truncation + all affines + demod (models/stylegan2.py:541-543, :220-225) --------------
        mean = g.truncation_latent.to(device=device, dtype=torch.float32).contiguous()
        psi_t, psi_s = None, 1.0
        if torch.is_tensor(truncation):
            psi_t = truncation.to(device=device, dtype=torch.float32).contiguous()
            if psi_t.numel() == 1:
                psi_t = psi_t.reshape(1).expand(batch).contiguous()
            elif psi_t.numel() != batch:
                raise L.MauaError(f"truncation has {psi_t.numel()} entries for a batch of {batch}")
        else:
            psi_s = float(truncation)
        if latent.shape[1] < g.n_latent or latent.shape[2] != g.style_dim:
            raise L.MauaError(f"latents of shape {tuple(latent.shape)}: need at least {g.n_latent} rows of {g.style_dim}")
        if mean.numel() != g.style_dim:
            # a per-layer mean latent ([n_latent, D], broadcast by the reference's `truncation_latent[None, ...]`,
            # models/stylegan2.py:541-543): the lerp is evaluated here with the reference's own expression and the
            # prologue kernel runs without truncation
            psi_b = psi_t[:, None, None] if psi_t is not None else psi_s
            latent = (mean[None, ...] + psi_b * (latent - mean[None, ...])).contiguous()
            latent_t = latent
            L.call("maua_style_prologue_f32", bc["table"].data_ptr(), plan["n_jobs"], latent.data_ptr(), None, None, 1.0,
                   None, batch, latent.shape[1], g.style_dim, stream)
        else:
            # (rows beyond n_latent feed no layer: the kernel never writes them, so they are carried over untruncated)
            latent_t = torch.empty_like(latent) if latent.shape[1] == g.n_latent else latent.clone()
            L.call("maua_style_prologue_f32", bc["table"].data_ptr(), plan["n_jobs"], latent.data_ptr(), mean.data_ptr(),
                   L.ptr(psi_t), psi_s, latent_t.data_ptr(), batch, latent.shape[1], g.style_dim, stream)

        # ---- input (models/stylegan2.py:547-548) --------------------------------------------------------------------
        from .stylegan2 import ConstantInput

        if isinstance(g.input, ConstantInput):
            x = g.input.input
            x_bstride = 0
        else:
            x = g.input(latent_t)
            x_bstride = x[0].numel()
        if _has_bend(0, bends):
            x = _bend(x.expand(batch, -1, -1, -1).contiguous() if x.shape[0] == 1 else x, 0, bends)
            x_bstride = x[0].numel()
        split = None  # (hi, lo) bf16 NHWC pre-scaled by the consuming layer's style, when available

        acts = []
        image = None
        layers = plan["layers"]
        current_size = 2  # doubled by every up layer; conv1 runs at 4
        for li, lp in enumerate(layers):
            sp = lp.spec
            conv = sp.mod.conv
            s, d = bc["views"][lp.job]
            nxt = layers[li + 1] if li + 1 < len(layers) else None
            bend_here = _has_bend(sp.layer_id, bends)
            nz = noise[sp.noise_index]
            in_h, in_w = (x.shape[2], x.shape[3]) if split is None else (split[0].shape[1], split[0].shape[2])
            out_h, out_w = (2 * in_h, 2 * in_w) if sp.up else (in_h, in_w)
            if nz is None:  # randomize_noise=True: fresh N(0,1) per layer (models/stylegan2.py:263-265)
                nz = torch.randn(batch, 1, out_h, out_w, device=device)

            # algorithmic work of this layer: conv FLOPs = 2*H_in*W_in*Cin*Cout*9 per sample
            L.TAG = {"layer": li, "up": sp.up, "cin": sp.cin, "cout": sp.cout, "h": in_h, "w": in_w,
                     "flops": 2.0 * in_h * in_w * sp.cin * sp.cout * 9 * batch,
                     "nprod": {"bf16": 1, "f16": 2, "bf16x3": 3}[lp.fmt] if lp.tc_ok else 1,
                     "out_elems": float(batch) * sp.cout * out_h * out_w}
            if lp.tc_ok:
                nprod = {"bf16": 1, "f16": 2, "bf16x3": 3}[lp.fmt]
                if lp.fmt == "f16" and (in_h + (1 if sp.up else 0) < 64 or in_w + (1 if sp.up else 0) < 32):
                    raise L.MauaError(f"precision='mixed': layer {li} runs at {in_h}x{in_w}, below the 64x32 grid the fp16 "
                                      "halo kernel needs (a bend shrank it?) - use precision='bf16x3'")
                if split is None:
                    split = _modulate_split(x.contiguous(), x_bstride, s, batch, lp.fmt)
                want_split = nxt is not None and nxt.tc_ok and not bend_here
                # ToRGB fused into the conv epilogue when one CTA sees every output channel (Cout <= 128, halo kernel):
                # the epilogue emits 3 partial sums per pixel instead of the Cout-channel fp32 map that torgb would re-read
                fuse_rgb = (sp.rgb is not None and not sp.up and not want_acts and not bend_here and sp.cout <= 128
                            and in_h >= 64 and in_w >= 32 and g.min_rgb_size <= out_h
                            and _FUSE_RGB)
                want_f32 = want_acts or bend_here or (sp.rgb is not None and not fuse_rgb) or (nxt is not None and not want_split)
                y = torch.empty((batch, sp.cout, out_h, out_w), device=device, dtype=torch.float32) if want_f32 else None
                o_hi = o_lo = None
                if want_split and nxt.fmt == "f16":
                    o_hi = torch.empty((batch, out_h, out_w, sp.cout), device=device, dtype=torch.float16)
                elif want_split:
                    o_hi = torch.empty((batch, out_h, out_w, sp.cout), device=device, dtype=torch.bfloat16)
                    o_lo = torch.empty_like(o_hi)
                nzc, nz_bs = _prep_noise(nz, device, batch, out_h * out_w)
                ep = L.ConvEpilogue()
                ep.noise, ep.noise_weight, ep.noise_bstride = nzc.data_ptr(), sp.mod.noise.weight.data_ptr(), nz_bs
                ep.bias = sp.mod.activate.bias.data_ptr()
                ep.s_next = bc["views"][nxt.job][0].data_ptr() if want_split else None
                ep.out_hi, ep.out_lo = L.ptr(o_hi), L.ptr(o_lo)
                ep.out_fmt = 1 if (want_split and nxt.fmt == "f16") else 0
                ep.out_f32_nchw = L.ptr(y)
                ep.slope, ep.act_scale, ep.activate = 0.2, SQRT2, 1
                ws = plan["workspace"]
                ep.workspace, ep.workspace_bytes = L.ptr(ws), (ws.numel() if ws is not None else 0)
                rgb_partial = None
                if fuse_rgb:
                    rs_f, _ = bc["views"][lp.rgb_job]
                    wr = torch.empty((batch, 3, sp.cout), device=device, dtype=torch.float32)
                    L.call("maua_rgb_weights_f32", sp.rgb.conv.weight.data_ptr(), rs_f.data_ptr(), wr.data_ptr(), batch,
                           sp.cout, float(sp.rgb.conv.scale), stream)
                    rgb_partial = torch.empty((batch, 3, out_h, out_w), device=device, dtype=torch.float32)
                    ep.rgb_w, ep.rgb_out = wr.data_ptr(), rgb_partial.data_ptr()
                if not sp.up:
                    ep.d = d.data_ptr()
                    L.call("maua_modconv_tc", split[0].data_ptr(), L.ptr(split[1]), lp.w_hi.data_ptr(),
                           lp.w_lo.data_ptr(), C.byref(ep), batch, sp.cin, sp.cout, in_h, in_w, 0, nprod, stream)
                else:
                    u = torch.empty((batch, 2 * in_h + 1, 2 * in_w + 1, sp.cout), device=device, dtype=torch.float32)
                    # demodulation commutes with the per-channel FIR: the conv streams raw phases, blur_act multiplies by d
                    ep_raw = L.ConvEpilogue()
                    ep_raw.out_raw_nhwc, ep_raw.activate = u.data_ptr(), 0
                    ep.d = d.data_ptr()
                    ep_raw.workspace, ep_raw.workspace_bytes = ep.workspace, ep.workspace_bytes
                    L.call("maua_modconv_tc", split[0].data_ptr(), L.ptr(split[1]), lp.w_hi.data_ptr(),
                           lp.w_lo.data_ptr(), C.byref(ep_raw), batch, sp.cin, sp.cout, in_h, in_w, 1, nprod, stream)
                    L.call("maua_blur_act_nhwc", u.data_ptr(), conv.blur.kernel.data_ptr(), C.byref(ep), batch, sp.cout,
                           2 * in_h + 1, 2 * in_w + 1, stream)
                split = (o_hi, o_lo) if want_split else None
                x, x_bstride = y, (y[0].numel() if y is not None else 0)
            else:
                # fp32 SIMT path, reference op order: conv -> [blur] -> noise -> bias+lrelu
                if split is not None:
                    raise L.MauaError("internal: split activations reached a SIMT layer")
                xin = x.expand(batch, -1, -1, -1).contiguous() if x.shape[0] != batch else x
                y = _modconv_simt(xin, conv.weight, s, d, conv.scale, 3, sp.up)
                if sp.up:
                    y = upfirdn2d(y, conv.blur.kernel, pad=conv.blur.pad)
                y = _noise_bias_act(y, nz, sp.mod.noise.weight, sp.mod.activate.bias, 0.2, SQRT2)
                x, x_bstride = y, y[0].numel()

            if bend_here:
                x = _bend(x, sp.layer_id, bends).contiguous()
                x_bstride = x[0].numel()
            if want_acts:
                acts.append(x)
            current_size *= 2 if (sp.up or li == 0) else 1
            last_rgb = sp.rgb is not None and not any(l2.spec.rgb is not None for l2 in layers[li + 1:])
            if sp.rgb is not None and lp.tc_ok and rgb_partial is not None and want_u8 and last_rgb and out_w % 4 == 0:
                # last ToRGB straight to uint8 NHWC (the full-resolution fp32 image is never written)
                u8 = torch.empty((batch, out_h, out_w, 3), device=device, dtype=torch.uint8)
                L.call("maua_rgb_finish_u8", rgb_partial.data_ptr(), sp.rgb.bias.data_ptr(), L.ptr(image),
                       sp.rgb.upsample.kernel.data_ptr() if image is not None else None, u8.data_ptr(), batch, out_h,
                       out_w, stream)
                image, want_u8 = u8, False
            elif sp.rgb is not None and lp.tc_ok and rgb_partial is not None:
                new_image = torch.empty_like(rgb_partial)
                L.call("maua_rgb_finish_f32", rgb_partial.data_ptr(), sp.rgb.bias.data_ptr(), L.ptr(image),
                       sp.rgb.upsample.kernel.data_ptr() if image is not None else None, new_image.data_ptr(), batch,
                       out_h, out_w, stream)
                image = new_image
            elif sp.rgb is not None:
                rs, _ = bc["views"][lp.rgb_job]
                if g.min_rgb_size <= current_size:
                    image = _torgb(x, sp.rgb.conv.weight, rs, sp.rgb.bias, image,
                                   sp.rgb.upsample.kernel if image is not None else None, sp.rgb.conv.scale)

        if want_u8 and image is not None:
            image = frames_to_u8(image)
    return image, latent_t, acts
