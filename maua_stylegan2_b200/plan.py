"""Weight-derived execution plan of a Generator: everything that depends only on the parameters
(not on the batch) is computed once and cached until a parameter object changes (render.py's `rewrites`
replace nn.Parameters between batches, render.py:160-167 — the cache key tracks data_ptr/_version).

  * Wsq[co,ci] = c^2 * sum_k W^2                    (demodulation without per-sample weights)
  * packed tensor-core weights  [tap][Cout][Cin] bf16 hi/lo with c folded in  (impl == "tc")
  * the style-prologue job table (one MauaStyleJob per modulated layer) per batch size
"""
import ctypes as C
import os

import torch

from . import _lib as L


SPLITK_WORKSPACE_BYTES = 64 << 20


# precision="mixed": layers whose OUTPUT is at least this wide take their activations as ONE fp16 plane (the "f16"
# format, include/maua_b200.h) instead of the bf16 (hi, lo) pair.  512 selects the four layers of the 1024^2 generator that
# are bound by the A-operand fetch of small-N MMAs (Cout <= 64) and carry 47 % of its conv time.
F16_MIN_RES = int(os.environ.get("MAUA_F16_MIN_RES", "512"))


class LayerPlan:
    __slots__ = ("spec", "job", "rgb_job", "wsq", "w_hi", "w_lo", "s_off", "d_off", "rgb_s_off", "tc_ok", "fmt", "out_res")


def _pack_tc(weight, scale, fmt="bf16x3"):
    cout, cin, k = weight.shape[1], weight.shape[2], weight.shape[3]
    dt = torch.float16 if fmt == "f16" else torch.bfloat16
    hi = torch.empty((k * k, cout, cin), device=weight.device, dtype=dt)
    lo = torch.empty_like(hi)
    with torch.cuda.device(weight.device):
        L.call("maua_pack_weight_f16x2" if fmt == "f16" else "maua_pack_weight_bf16x2", weight.data_ptr(), hi.data_ptr(),
               lo.data_ptr(), cout, cin, k, float(scale), L.stream_ptr(weight.device))
    return hi, lo


def tc_supported(cin, cout):
    return cin % 32 == 0 and cout % 16 == 0 and cout >= 16


def build_plan(g, key):
    from .stylegan2 import _weight_sq

    device = g.input.input.device
    layers = []
    s_total = 0
    d_total = 0
    n_jobs = 0
    res = 2
    for li, sp in enumerate(g._specs):
        lp = LayerPlan()
        lp.spec = sp
        conv = sp.mod.conv
        lp.wsq = _weight_sq(conv.weight, conv.scale)
        lp.tc_ok = g.impl == "tc" and tc_supported(sp.cin, sp.cout)
        res *= 2 if (sp.up or li == 0) else 1
        lp.out_res = res  # nominal (bends may change it at run time; synthesize() checks the real shape)
        lp.fmt = "f16" if (lp.tc_ok and g.precision == "mixed" and res >= F16_MIN_RES) else \
            ("bf16" if g.precision == "bf16" else "bf16x3")
        lp.w_hi = lp.w_lo = None
        if lp.tc_ok:
            lp.w_hi, lp.w_lo = _pack_tc(conv.weight, conv.scale, lp.fmt)
        lp.job = n_jobs
        n_jobs += 1
        lp.s_off = s_total
        s_total += sp.cin
        lp.d_off = d_total
        d_total += sp.cout
        lp.rgb_job = -1
        lp.rgb_s_off = -1
        if sp.rgb is not None:
            lp.rgb_job = n_jobs
            n_jobs += 1
            lp.rgb_s_off = s_total
            s_total += sp.rgb.conv.in_channel
        layers.append(lp)
    # scratch of the deterministic split-K (tiny layers): zeroed once, every launch hands it back zeroed; launches that
    # share it must be stream-ordered (one Generator forward at a time per device -- the frame loop's single compute stream)
    workspace = torch.zeros(SPLITK_WORKSPACE_BYTES, device=device, dtype=torch.uint8) if g.impl == "tc" else None
    return {"key": key, "layers": layers, "s_total": s_total, "d_total": d_total, "n_jobs": n_jobs,
            "device": device, "batch": {}, "workspace": workspace}


def batch_buffers(g, plan, batch):
    """Persistent per-batch-size buffers: s [sum_cin * B], d [sum_cout * B] and the device job table."""
    bc = plan["batch"].get(batch)
    if bc is not None:
        return bc
    device = plan["device"]
    s_buf = torch.empty(plan["s_total"] * batch, device=device, dtype=torch.float32)
    d_buf = torch.empty(plan["d_total"] * batch, device=device, dtype=torch.float32)
    jobs = (L.StyleJob * plan["n_jobs"])()
    views = {}
    keep = []
    for lp in plan["layers"]:
        sp = lp.spec
        conv = sp.mod.conv
        s_view = s_buf[lp.s_off * batch:(lp.s_off + sp.cin) * batch].view(batch, sp.cin)
        d_view = d_buf[lp.d_off * batch:(lp.d_off + sp.cout) * batch].view(batch, sp.cout)
        j = jobs[lp.job]
        j.mod_w, j.mod_b = conv.modulation.weight.data_ptr(), conv.modulation.bias.data_ptr()
        j.wsq = lp.wsq.data_ptr() if conv.demodulate else None
        j.s_out, j.d_out = s_view.data_ptr(), d_view.data_ptr()
        j.cin, j.cout, j.latent_index = sp.cin, sp.cout, sp.latent_index
        if lp.fmt == "f16" and conv.demodulate:
            # fp16 consumers take range-normalised styles (|s| < 1, the power of two moved into d: MauaStyleJob.s_norm_out)
            s_norm = torch.empty_like(s_view)
            keep.append(s_norm)
            j.s_norm_out = s_norm.data_ptr()
            s_view = s_norm
        views[lp.job] = (s_view, d_view if conv.demodulate else None)
        if sp.rgb is not None:
            rc = sp.rgb.conv
            rs_view = s_buf[lp.rgb_s_off * batch:(lp.rgb_s_off + rc.in_channel) * batch].view(batch, rc.in_channel)
            r = jobs[lp.rgb_job]
            r.mod_w, r.mod_b, r.wsq = rc.modulation.weight.data_ptr(), rc.modulation.bias.data_ptr(), None
            r.s_out, r.d_out = rs_view.data_ptr(), None
            r.cin, r.cout, r.latent_index = rc.in_channel, 3, sp.rgb_latent_index
            views[lp.rgb_job] = (rs_view, None)
    raw = bytes(jobs)
    table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    bc = {"s_buf": s_buf, "d_buf": d_buf, "table": table, "views": views, "keep": keep}
    plan["batch"][batch] = bc
    return bc


def style_single(modulation, style, wsq=None, cout=0):
    """s (and d) for ONE modulated layer through the prologue kernel (used by stand-alone module forwards)."""
    style = style.contiguous().float()
    b, dim = style.shape
    cin = modulation.weight.shape[0]
    s = torch.empty((b, cin), device=style.device, dtype=torch.float32)
    d = torch.empty((b, cout), device=style.device, dtype=torch.float32) if wsq is not None else None
    job = (L.StyleJob * 1)()
    job[0].mod_w, job[0].mod_b = modulation.weight.data_ptr(), modulation.bias.data_ptr()
    job[0].wsq = wsq.data_ptr() if wsq is not None else None
    job[0].s_out, job[0].d_out = s.data_ptr(), (d.data_ptr() if d is not None else None)
    job[0].cin, job[0].cout, job[0].latent_index = cin, cout, 0
    table = torch.frombuffer(bytearray(bytes(job)), dtype=torch.uint8).to(style.device)
    with torch.cuda.device(style.device):
        L.call("maua_style_prologue_f32", table.data_ptr(), 1, style.data_ptr(), None, None, 1.0, None, b, 1, dim,
               L.stream_ptr(style.device))
    return s, d
