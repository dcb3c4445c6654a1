"""Unit parity of the tensor-core modulated conv (maua_modconv_tc + blur_act_nhwc + layout kernels) against an
fp64 CPU evaluation of the same formula.  Tolerance: 2e-4 of the tensor max for the
3-product split-bf16 mode (north_star bar is 1e-3 on activations), 3e-2 for the single-product fast mode."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _pack(w, scale, fmt="bf16x3"):
    from maua_stylegan2_b200.plan import _pack_tc

    return _pack_tc(w[None].contiguous(), scale, fmt)


def _run_tc(x, w, s, d, noise, nw, bias, s_next, up, nprod, scale, workspace=None, out_f16=False, d_in_conv=False):
    """nprod 3 / 1: bf16 (hi, lo) activations; nprod 2: the "f16" format (one fp16 activation plane, fp16 weight pair).
    out_f16: the epilogue writes the consumer's activations as one fp16 plane (returned as o_hi, o_lo = zeros)."""
    from maua_stylegan2_b200 import _lib as L
    from maua_stylegan2_b200.synthesis import _modulate_split

    b, cin, h, wd = x.shape
    cout = w.shape[0]
    fmt = "f16" if nprod == 2 else "bf16x3"
    w_hi, w_lo = _pack(w, scale, fmt)
    hi, lo = _modulate_split(x, x[0].numel(), s, b, fmt)
    if lo is None:
        lo = hi   # ignored by the kernel
    stream = L.stream_ptr(x.device)
    oh, ow = (2 * h, 2 * wd) if up else (h, wd)
    y = torch.full((b, cout, oh, ow), float("nan"), device="cuda")
    o_hi = torch.zeros((b, oh, ow, cout), device="cuda", dtype=torch.float16 if out_f16 else torch.bfloat16)
    o_lo = torch.zeros_like(o_hi)
    ep = L.ConvEpilogue()
    ep.out_fmt = 1 if out_f16 else 0
    ep.noise, ep.noise_weight = noise.data_ptr(), nw.data_ptr()
    ep.noise_bstride = oh * ow if noise.shape[0] == b else 0
    ep.bias, ep.s_next = bias.data_ptr(), s_next.data_ptr()
    ep.out_hi, ep.out_lo, ep.out_f32_nchw = o_hi.data_ptr(), o_lo.data_ptr(), y.data_ptr()
    ep.slope, ep.act_scale, ep.activate = 0.2, 2 ** 0.5, 1
    if workspace is not None:
        ep.workspace, ep.workspace_bytes = workspace.data_ptr(), workspace.numel()
    u = None
    if not up:
        ep.d = d.data_ptr()
        L.call("maua_modconv_tc", hi.data_ptr(), lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C.byref(ep), b, cin,
               cout, h, wd, 0, nprod, stream)
    else:
        u = torch.full((b, 2 * h + 1, 2 * wd + 1, cout), float("nan"), device="cuda")
        er = L.ConvEpilogue()
        er.out_raw_nhwc, er.activate = u.data_ptr(), 0
        if d_in_conv:      # stand-alone form: the conv epilogue applies the demodulation itself
            er.d = d.data_ptr()
        else:              # product form: raw phases out of the conv, d applied by blur_act (commutes with the FIR)
            ep.d = d.data_ptr()
        er.workspace, er.workspace_bytes = ep.workspace, ep.workspace_bytes
        L.call("maua_modconv_tc", hi.data_ptr(), lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C.byref(er), b, cin,
               cout, h, wd, 1, nprod, stream)
        k = torch.tensor([1.0, 3.0, 3.0, 1.0])
        k4 = (k[None] * k[:, None] / 16).cuda()
        L.call("maua_blur_act_nhwc", u.data_ptr(), k4.data_ptr(), C.byref(ep), b, cout, 2 * h + 1, 2 * wd + 1, stream)
    torch.cuda.synchronize()
    return y, o_hi, o_lo, u


def _reference(x, w, s, d, noise, nw, bias, up, scale):
    x, w, s, d, noise, nw, bias = (t.double().cpu() for t in (x, w, s, d, noise, nw, bias))
    xs = x * s[:, :, None, None]
    if not up:
        o = F.conv2d(xs, w * scale, padding=1)
        raw = None
    else:
        o = F.conv_transpose2d(xs, (w * scale).transpose(0, 1), stride=2)     # [B,Cout,2H+1,2W+1]
        raw = o.permute(0, 2, 3, 1)                                           # raw phases (no demodulation)
        k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
        k4 = (k[None] * k[:, None] / 16)
        c = o.shape[1]
        o = F.conv2d(F.pad(o, [1, 1, 1, 1]), torch.flip(k4, [0, 1])[None, None].repeat(c, 1, 1, 1), groups=c)
    o = o * d[:, :, None, None]
    o = o + nw * noise + bias[None, :, None, None]
    return F.leaky_relu(o, 0.2) * 2 ** 0.5, raw


CASES = [
    # b, cin, cout, h, w, up
    (2, 64, 64, 16, 16, False),
    (3, 32, 32, 20, 12, False),      # KC=32 path, non power-of-two image
    (8, 512, 512, 4, 4, False),      # batch folded into the M tile
    (2, 128, 256, 32, 32, False),    # BN=256
    (1, 64, 32, 40, 24, False),
    (2, 64, 64, 8, 8, True),
    (3, 128, 64, 16, 12, True),
    (8, 512, 512, 4, 4, True),
    (1, 64, 32, 32, 32, True),       # KC=64, BN=32, 4 phases
    (2, 32, 16, 9, 7, True),         # odd sizes
    (8, 512, 512, 16, 16, True),     # 17x17 grid -> 17x7 tiles (TMA box with fewer than 128 rows)
    (4, 256, 256, 32, 32, True),     # 33x33 grid -> 11x11 tiles
    (3, 64, 64, 5, 3, True),         # whole images folded over the batch, ragged last tile
]


@pytest.mark.parametrize("b,cin,cout,h,w,up", CASES)
def test_tc_conv_matches_fp64_reference(b, cin, cout, h, w, up):
    torch.manual_seed(b * 1000 + cin + cout + h)
    x = torch.randn(b, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda")
    s = 1 + 0.5 * torch.randn(b, cin, device="cuda")
    d = 0.5 + torch.rand(b, cout, device="cuda")
    oh, ow = (2 * h, 2 * w) if up else (h, w)
    noise = torch.randn(b, 1, oh, ow, device="cuda")
    nw = torch.tensor([0.3], device="cuda")
    bias = 0.1 * torch.randn(cout, device="cuda")
    s_next = 1 + 0.5 * torch.randn(b, cout, device="cuda")
    scale = 1 / (cin * 9) ** 0.5
    ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
    y, o_hi, o_lo, u = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
    _check(ref, raw, y, o_hi, o_lo, u, s_next, up)
    # same layer with the split-K scratch available: identical maths, S partial sums reduced in a fixed order
    ws = torch.zeros(64 << 20, device="cuda", dtype=torch.uint8)
    y2, o_hi2, o_lo2, u2 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale, workspace=ws)
    _check(ref, raw, y2, o_hi2, o_lo2, u2, s_next, up)
    assert not ws[:4096].any(), "split-K must hand the tile counters back zeroed"
    y3, _, _, u3 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale, workspace=ws)
    assert torch.equal(y2, y3), "split-K reduction must be bitwise reproducible"
    if up:
        assert torch.equal(u2, u3)


# ---------------------------------------------------------------------------------------------------------------
# The persistent halo kernel (v2, csrc/modconv_tc2.cu) — the kernel that carries ~85 % of the conv time of the bench.
# It only takes layers with a GEMM grid >= 64 x 32 (or the wide 32^2 case), which the small cases above never reach.
# ---------------------------------------------------------------------------------------------------------------
V2_CASES = [
    # b, cin, cout, h, w, up          what the default tile policy picks (asserted through maua_modconv_tc_last_config)
    (8, 512, 512, 32, 32, False),    # "wide32": R=1 BN=256 needs 16*B >= 120 items
    (2, 128, 128, 256, 256, False),  # R=2 BN=128 (3 MMAs per K-step)
    (1, 32, 32, 1024, 1024, False),  # R=4 BN=32 concat, weights resident in shared memory
    (2, 64, 64, 128, 96, False),     # BN=64 concat, non-square
    (1, 256, 256, 64, 64, False),    # BN=256, few items -> policy fallbacks (n_ctas < 120)
    (2, 64, 32, 512, 512, True),     # up, BN=32 concat, 4 phases in one item
    (2, 128, 64, 256, 256, True),    # up, R=2 BN=64, 2 phase groups
    (2, 256, 128, 128, 128, True),   # up, BN=128, 2 phase groups
    (1, 512, 256, 64, 64, True),     # up, BN=256, 4 phase groups (one phase per item)
    (3, 64, 48, 70, 40, False),      # Cout not a power of two (BN=16), ragged tiles
]


def _case_tensors(b, cin, cout, h, w, up, seed):
    torch.manual_seed(seed)
    x = torch.randn(b, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda")
    s = 1 + 0.5 * torch.randn(b, cin, device="cuda")
    d = 0.5 + torch.rand(b, cout, device="cuda")
    oh, ow = (2 * h, 2 * w) if up else (h, w)
    noise = torch.randn(b, 1, oh, ow, device="cuda")
    nw = torch.tensor([0.3], device="cuda")
    bias = 0.1 * torch.randn(cout, device="cuda")
    s_next = 1 + 0.5 * torch.randn(b, cout, device="cuda")
    return x, wt, s, d, noise, nw, bias, s_next, 1 / (cin * 9) ** 0.5


@pytest.mark.parametrize("b,cin,cout,h,w,up", V2_CASES)
def test_tc2_halo_kernel_matches_fp64_reference(b, cin, cout, h, w, up):
    from maua_stylegan2_b200 import _lib as L

    x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, b * 1000 + cin + cout + h)
    ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
    y, o_hi, o_lo, u = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
    cfg = L.last_conv_config()
    assert cfg.startswith("v2 "), f"{(b, cin, cout, h, w, up)} was routed to {cfg!r}, not to the halo kernel"
    print(cfg)
    _check(ref, raw, y, o_hi, o_lo, u, s_next, up)
    y2, o_hi2, o_lo2, u2 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
    assert torch.equal(y, y2) and torch.equal(o_hi, o_hi2) and torch.equal(o_lo, o_lo2), "v2 must be deterministic"


# Forced configurations "R,BN,cat,groups": every tap table of modconv_tc2.cu (c_taps[0..7] = same-res; up with 1 / 2 / 4
# phase groups), both product modes (three N=BN MMAs / concat) and every R, on shapes small enough for the fp64 reference.
FORCED = [
    ((2, 64, 32, 64, 40, False), ["1,32,0,1", "2,32,0,1", "4,32,0,1", "1,32,1,1", "2,16,1,1", "4,32,1,1", "4,16,0,1"]),
    ((2, 64, 64, 48, 24, False), ["1,64,0,1", "2,64,1,1", "1,32,1,1"]),
    ((1, 128, 256, 32, 32, False), ["1,256,0,1", "2,128,0,1", "1,128,0,1"]),
    ((2, 64, 32, 56, 24, True), ["1,32,0,1", "1,32,1,1", "2,32,0,1", "1,32,0,2", "2,32,1,2", "4,32,0,2", "1,32,0,4",
                                 "2,32,1,4", "4,32,0,4", "4,16,1,4"]),
    ((1, 128, 128, 32, 32, True), ["1,128,0,1", "1,128,0,2", "2,128,0,4", "1,64,1,2"]),
    ((1, 64, 256, 33, 17, True), ["1,256,0,4", "1,256,0,2", "2,128,0,2"]),
]


@pytest.mark.parametrize("case,forces", FORCED)
def test_tc2_forced_configurations(case, forces, monkeypatch):
    from maua_stylegan2_b200 import _lib as L

    b, cin, cout, h, w, up = case
    x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, 77 + cin + h)
    ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
    for f in forces:
        monkeypatch.setenv("MAUA_TC_FORCE", f)
        y, o_hi, o_lo, u = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
        cfg = L.last_conv_config()
        r, bn, cat, groups = (int(v) for v in f.split(","))
        assert f"v2 up={int(up)} R={r} BN={bn} cat={cat} groups={groups} " in cfg, (f, cfg)
        _check(ref, raw, y, o_hi, o_lo, u, s_next, up)
    monkeypatch.delenv("MAUA_TC_FORCE")
    # an infeasible forced configuration is an error, never a silent fallback to another kernel
    monkeypatch.setenv("MAUA_TC_FORCE", "4,256,1,1")
    with pytest.raises(L.MauaError):
        _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)


# CTA pairs (cta_group::2, MAUA_TC_PAIR): M = 256 MMAs over two CTAs of a cluster, the weight tile split between them.
# MAUA_TC_PAIR=2 pairs every split-bf16 non-concat configuration with BN >= 32; results must equal the unpaired kernel's
# bit for bit (same MMAs, same accumulation order per output element) — and match the fp64 reference.
PAIR_CASES = [
    ((2, 128, 256, 64, 64, False), [None, "1,256,0,1", "1,128,0,1", "2,128,0,1", "2,64,0,1", "4,32,0,1"]),
    ((1, 64, 128, 72, 40, False), [None, "2,128,0,1", "1,64,0,1"]),                  # ragged tiles, odd pixel-tile count
    ((3, 64, 64, 48, 24, False), ["1,64,0,1", "2,32,0,1"]),                          # 27 pixel tiles at R=1: dummy peer tile
    ((1, 128, 256, 33, 17, True), [None, "1,256,0,4", "1,256,0,2", "2,128,0,2", "1,128,0,1"]),
    ((2, 64, 128, 64, 32, True), [None, "1,128,0,2", "2,64,0,2", "4,32,0,4"]),
]


@pytest.mark.parametrize("case,forces", PAIR_CASES)
def test_tc2_cta_pairs_match_single_cta_kernel(case, forces, monkeypatch):
    from maua_stylegan2_b200 import _lib as L

    b, cin, cout, h, w, up = case
    x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, 501 + cout + h)
    ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
    for f in forces:
        if f is None:
            monkeypatch.delenv("MAUA_TC_FORCE", raising=False)
        else:
            monkeypatch.setenv("MAUA_TC_FORCE", f)
        monkeypatch.setenv("MAUA_TC_PAIR", "0")
        y0, h0, l0, u0 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
        cfg0 = L.last_conv_config()
        if not cfg0.startswith("v2 "):
            continue   # (the default policy may route a small case to v1: nothing to pair)
        assert cfg0.endswith("pair=0"), cfg0
        monkeypatch.setenv("MAUA_TC_PAIR", "2")
        y1, h1, l1, u1 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale)
        cfg1 = L.last_conv_config()
        assert cfg1.endswith("pair=1"), (f, cfg1)
        print(cfg1)
        _check(ref, raw, y1, h1, l1, u1, s_next, up)
        for a, c in ((y0, y1), (h0, h1), (l0, l1), (u0, u1)):
            if a is not None:
                assert torch.equal(a, c), f"pair kernel differs from the single-CTA kernel ({f})"


# The "f16" activation format (precision="mixed": the >= 512^2 layers): one fp16 activation plane, fp16 (hi, lo) weights.
# The only rounding is the 11-bit activation operand: per-layer error ~2e-4 of the tensor max (bar: 6e-4 here, 1e-3 on
# the network); the fp16 OUTPUT plane adds the consumer's own operand rounding (2^-11 per element).
F16_CASES = [
    ((2, 64, 32, 128, 64, False), None),         # mode 4 (concat N = 2*BN), R=4
    ((1, 64, 64, 96, 64, False), None),          # mode 4, BN=64
    ((1, 128, 128, 64, 64, False), None),        # mode 3 (two N=BN MMAs into one accumulator)
    ((1, 64, 256, 64, 32, False), None),         # mode 3, BN=256
    ((2, 64, 32, 72, 40, True), None),           # up, mode 4
    ((1, 128, 64, 64, 48, True), None),          # up, BN=64
    ((1, 128, 128, 64, 32, True), None),         # up, mode 3
    ((2, 64, 32, 64, 40, False), ["1,32,1,1", "2,32,0,1", "4,16,1,1"]),
    ((2, 64, 32, 56, 24, True), ["1,32,1,1", "2,32,1,2", "4,32,0,4", "2,32,0,1"]),
    # collapsed taps (one phase group, no concat, BN <= 64: 4 MMAs of N = 4BN / 2BN / 2BN / BN per K-step and plane)
    ((2, 64, 32, 56, 24, True), ["1,32,0,1", "4,32,0,1", "1,16,0,1", "2,16,0,1"]),
    ((1, 128, 64, 64, 48, True), ["1,64,0,1", "2,32,0,1", "1,32,0,1"]),
    ((3, 32, 64, 33, 17, True), ["1,64,0,1", "1,32,0,1"]),          # ragged tiles, one K chunk
]


@pytest.mark.parametrize("case,forces", F16_CASES)
def test_tc2_f16_activation_format(case, forces, monkeypatch):
    from maua_stylegan2_b200 import _lib as L

    b, cin, cout, h, w, up = case
    x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, 300 + cin + cout + h)
    ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
    monkeypatch.setenv("MAUA_TC_COLL", "1")   # the collapsed-tap form is opt-in (measured slower): cover it here
    for f in (forces or [None]):
        if f is not None:
            monkeypatch.setenv("MAUA_TC_FORCE", f)
        for out_f16 in (False, True):
            y, o_hi, o_lo, u = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 2, scale, out_f16=out_f16)
            cfg = L.last_conv_config()
            assert cfg.startswith("v2 ") and " prod=2 " in cfg, cfg
            if up and f is not None and f.endswith(",0,1") and int(f.split(",")[1]) <= 64:
                assert " coll=1 " in cfg, cfg     # eligible forced configurations take the collapsed-tap form
            if up:
                assert rel_err(u.cpu().numpy(), raw.numpy()) < 6e-4, "raw transposed-conv phases"
            assert not torch.isnan(y).any()
            e = rel_err(y.cpu().numpy(), ref.numpy())
            rec = (o_hi.float() + (0 if out_f16 else o_lo.float())).permute(0, 3, 1, 2).cpu().double() \
                / s_next.cpu().double()[:, :, None, None]
            e2 = rel_err(rec.numpy(), ref.numpy())
            print(f"{case} force={f} out_f16={out_f16}: {cfg.split(' items')[0]}  fp32 out {e:.2e}  split out {e2:.2e}")
            assert e < 6e-4, f"fp32 NCHW output rel err {e}"
            assert e2 < (1.2e-3 if out_f16 else 6e-4), f"NHWC output rel err {e2}"
    # the fp16 format needs the halo kernel: tiny layers must refuse it loudly instead of running something else
    monkeypatch.delenv("MAUA_TC_FORCE", raising=False)
    xs, wts, ss, ds, ns, nws, bs, sns, sc = _case_tensors(2, 64, 64, 16, 16, False, 1)
    with pytest.raises(L.MauaError):
        _run_tc(xs, wts, ss, ds, ns, nws, bs, sns, False, 2, sc)


def test_tc2_collapsed_taps_agree_with_per_tap_form(monkeypatch):
    """MAUA_TC_COLL=0 runs the same configuration tap by tap (9 MMAs of N = BN per K-step and plane): same products, a
    different fp32 accumulation order — the raw phases must agree to fp32 rounding, far below the fp16-operand error."""
    from maua_stylegan2_b200 import _lib as L

    b, cin, cout, h, w, up = 2, 96, 32, 56, 40, True
    x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, 4242)
    for f in ("2,32,0,1", "1,16,0,1"):
        monkeypatch.setenv("MAUA_TC_FORCE", f)
        monkeypatch.setenv("MAUA_TC_COLL", "0")
        y0, _, _, u0 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 2, scale)
        assert " coll=0 " in L.last_conv_config()
        monkeypatch.setenv("MAUA_TC_COLL", "1")
        y1, _, _, u1 = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 2, scale)
        assert " coll=1 " in L.last_conv_config()
        assert not torch.isnan(u1).any() and not torch.isnan(y1).any()
        assert rel_err(u1.cpu().numpy(), u0.cpu().numpy()) < 5e-6, f
        assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 5e-6, f


def test_f16_layout_kernels():
    from maua_stylegan2_b200.synthesis import _modulate_split

    torch.manual_seed(8)
    x = torch.randn(3, 40, 5, 7, device="cuda") * 50
    x[0, 0, 0, 0] = 1e6          # saturates to the largest finite fp16 instead of overflowing to inf
    s = torch.randn(3, 40, device="cuda")
    s[0, 0] = 1.0
    hi, lo = _modulate_split(x, x[0].numel(), s, 3, "f16")
    assert lo is None and hi.dtype == torch.float16
    want = (x * s[:, :, None, None]).permute(0, 2, 3, 1).clamp(-65504, 65504)
    assert torch.equal(hi, want.to(torch.float16))
    assert torch.isfinite(hi.float()).all() and hi[0, 0, 0, 0] == 65504
    w = torch.randn(1, 24, 40, 3, 3, device="cuda")
    w_hi, w_lo = _pack(w[0], 0.25, "f16")
    wantw = (w[0] * 0.25).permute(2, 3, 0, 1).reshape(9, 24, 40)
    assert w_hi.dtype == torch.float16 and torch.equal(w_hi, wantw.to(torch.float16))
    assert (w_hi.float() + w_lo.float() - wantw).abs().max() <= wantw.abs().max() * 2 ** -20


def test_up_conv_with_demodulation_in_the_conv_epilogue():
    """Stand-alone form of the transposed conv: ep.d given to the conv (u = d * convT(x)), none to blur_act."""
    for case in ((2, 64, 32, 72, 40, True), (8, 512, 512, 4, 4, True)):
        b, cin, cout, h, w, up = case
        x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, up, 900 + cin)
        ref, raw = _reference(x, wt, s, d, noise, nw, bias, up, scale)
        y, o_hi, o_lo, u = _run_tc(x, wt, s, d, noise, nw, bias, s_next, up, 3, scale, d_in_conv=True)
        assert rel_err(u.cpu().numpy(), (raw * d.double().cpu()[:, None, None, :]).numpy()) < 2e-4
        assert rel_err(y.cpu().numpy(), ref.numpy()) < 2e-4


def test_tc2_fused_torgb_partial_sums():
    """Fused ToRGB epilogue (Cout <= 128, same-res): rgb_out[b,k] = sum_c rgb_w[b,k,c] * act[b,c]."""
    from maua_stylegan2_b200 import _lib as L
    from maua_stylegan2_b200.synthesis import _modulate_split

    for (b, cin, cout, h, w) in ((2, 64, 32, 128, 64), (1, 128, 128, 64, 64), (2, 64, 64, 80, 40)):
        x, wt, s, d, noise, nw, bias, s_next, scale = _case_tensors(b, cin, cout, h, w, False, 5 + cout)
        ref, _ = _reference(x, wt, s, d, noise, nw, bias, False, scale)
        rgb_w = torch.randn(b, 3, cout, device="cuda")
        w_hi, w_lo = _pack(wt, scale)
        hi, lo = _modulate_split(x, x[0].numel(), s, b)
        rgb = torch.full((b, 3, h, w), float("nan"), device="cuda")
        o_hi = torch.zeros((b, h, w, cout), device="cuda", dtype=torch.bfloat16)
        o_lo = torch.zeros_like(o_hi)
        ep = L.ConvEpilogue()
        ep.d, ep.noise, ep.noise_weight, ep.noise_bstride = d.data_ptr(), noise.data_ptr(), nw.data_ptr(), h * w
        ep.bias, ep.s_next = bias.data_ptr(), s_next.data_ptr()
        ep.out_hi, ep.out_lo = o_hi.data_ptr(), o_lo.data_ptr()
        ep.rgb_w, ep.rgb_out = rgb_w.data_ptr(), rgb.data_ptr()
        ep.slope, ep.act_scale, ep.activate = 0.2, 2 ** 0.5, 1
        L.call("maua_modconv_tc", hi.data_ptr(), lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C.byref(ep), b, cin, cout,
               h, w, 0, 3, L.stream_ptr(x.device))
        torch.cuda.synchronize()
        assert L.last_conv_config().startswith("v2 ")
        want = torch.einsum("bkc,bchw->bkhw", rgb_w.double().cpu(), ref)
        assert rel_err(rgb.cpu().numpy(), want.numpy()) < 2e-4
        rec = (o_hi.float() + o_lo.float()).permute(0, 3, 1, 2).cpu().double() / s_next.cpu().double()[:, :, None, None]
        assert rel_err(rec.numpy(), ref.numpy()) < 3e-4


def _check(ref, raw, y, o_hi, o_lo, u, s_next, up):
    if up:
        assert rel_err(u.cpu().numpy(), raw.numpy()) < 2e-4, "raw transposed-conv phases"
    assert not torch.isnan(y).any()
    e = rel_err(y.cpu().numpy(), ref.numpy())
    assert e < 2e-4, f"fp32 NCHW output rel err {e}"
    rec = (o_hi.float() + o_lo.float()).permute(0, 3, 1, 2).cpu().double() / s_next.cpu().double()[:, :, None, None]
    e2 = rel_err(rec.numpy(), ref.numpy())
    assert e2 < 3e-4, f"split NHWC output rel err {e2}"


def test_tc_conv_single_product_mode_is_bf16_grade():
    torch.manual_seed(5)
    b, cin, cout, h, w = 2, 64, 64, 16, 16
    x = torch.randn(b, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, 3, 3, device="cuda")
    s = torch.ones(b, cin, device="cuda")
    d = torch.ones(b, cout, device="cuda")
    noise = torch.zeros(1, 1, h, w, device="cuda")
    nw = torch.zeros(1, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    scale = 1 / (cin * 9) ** 0.5
    ref, _ = _reference(x, wt, s, d, noise, nw, bias, False, scale)
    y, *_ = _run_tc(x, wt, s, d, noise, nw, bias, torch.ones(b, cout, device="cuda"), False, 1, scale)
    e = rel_err(y.cpu().numpy(), ref.numpy())
    assert 1e-5 < e < 3e-2, e


def test_layout_kernels():
    from maua_stylegan2_b200.synthesis import _modulate_split

    torch.manual_seed(6)
    x = torch.randn(3, 40, 5, 7, device="cuda")
    s = torch.randn(3, 40, device="cuda")
    hi, lo = _modulate_split(x, x[0].numel(), s, 3)
    want = (x * s[:, :, None, None]).permute(0, 2, 3, 1)
    assert torch.equal(hi, want.to(torch.bfloat16))
    assert (hi.float() + lo.float() - want).abs().max() <= want.abs().max() * 2 ** -16
    # broadcast (constant input) form
    hi1, lo1 = _modulate_split(x[:1], 0, s, 3)
    want1 = (x[:1] * s[:, :, None, None]).permute(0, 2, 3, 1)
    assert torch.equal(hi1, want1.to(torch.bfloat16))
    w = torch.randn(1, 24, 40, 3, 3, device="cuda")
    w_hi, w_lo = _pack(w[0], 0.25)
    wantw = (w[0] * 0.25).permute(2, 3, 0, 1).reshape(9, 24, 40)
    assert torch.equal(w_hi, wantw.to(torch.bfloat16))
    assert (w_hi.float() + w_lo.float() - wantw).abs().max() <= wantw.abs().max() * 2 ** -16
