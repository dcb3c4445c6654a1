"""Exhaustive hardware sweep of the halo kernel's configuration space (R stacked tiles, BN, concat mode, phase groups)
for every >= 64^2 layer of a generator config.  Prints, per layer, the policy's own pick and the measured ranking; the
cost model in csrc/modconv_tc2.cu is fitted to this table.   python tools/tune_tc2.py [--size 1024] [--batch 8]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maua_stylegan2_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--cm", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--only", default=None, help="'cin,cout,h,up': run just this layer with the policy's configuration "
                                             "(or --force 'R,BN,cat,groups'), no sweep — the ncu capture target")
ap.add_argument("--force", default=None)
ap.add_argument("--prod", type=int, default=3, help="3 = split bf16, 2 = fp16 activation format, 1 = single bf16 product")
ap.add_argument("--min-res", type=int, default=64, help="only layers whose output is at least this wide")
ap.add_argument("--out-f16", type=int, default=-1, help="epilogue output format (default: fp16 plane iff --prod 2)")
ap.add_argument("--up-d", action="store_true", help="transposed layers: demodulate in the conv epilogue (stand-alone form); "
                                                    "default = the product form (raw phases, d applied by blur_act)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
chan = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * a.cm, 128: 128 * a.cm, 256: 64 * a.cm, 512: 32 * a.cm,
        1024: 16 * a.cm}
layers = []  # (cin, cout, h, up, fuse_rgb, last)
res = 64
while res <= a.size:
    layers.append((chan[res // 2], chan[res], res // 2, True, False, False))
    layers.append((chan[res], chan[res], res, False, chan[res] <= 128, res == a.size))
    res *= 2
B = a.batch
stream = L.stream_ptr(dev)


def run(cin, cout, h, up, fuse, last, force):
    if force is None:
        os.environ.pop("MAUA_TC_FORCE", None)
    else:
        os.environ["MAUA_TC_FORCE"] = "%d,%d,%d,%d" % force
    dt = torch.float16 if a.prod == 2 else torch.bfloat16
    x_hi = torch.randn(B, h, h, cin, device=dev).to(dt)
    x_lo = (torch.randn(B, h, h, cin, device=dev) * 2 ** -9).to(dt)
    w_hi = torch.randn(9, cout, cin, device=dev).to(dt)
    w_lo = (torch.randn(9, cout, cin, device=dev) * 2 ** -9).to(dt)
    out_f16 = (a.prod == 2) if a.out_f16 < 0 else bool(a.out_f16)
    d = torch.rand(B, cout, device=dev) + 0.5
    ep = L.ConvEpilogue()
    if not up or a.up_d:
        ep.d = d.data_ptr()
    keep = [d]
    if up:
        u = torch.empty(B, 2 * h + 1, 2 * h + 1, cout, device=dev)
        ep.out_raw_nhwc, ep.activate = u.data_ptr(), 0
        keep.append(u)
    else:
        nz = torch.randn(B, 1, h, h, device=dev)
        nw = torch.tensor([0.1], device=dev)
        bias = torch.zeros(cout, device=dev)
        sn = torch.ones(B, cout, device=dev)
        o_hi = torch.empty(B, h, h, cout, device=dev, dtype=torch.float16 if out_f16 else torch.bfloat16)
        o_lo = torch.empty_like(o_hi)
        ep.out_fmt = 1 if out_f16 else 0
        ep.noise, ep.noise_weight, ep.noise_bstride = nz.data_ptr(), nw.data_ptr(), h * h
        ep.bias = bias.data_ptr()
        if not last:  # the last layer only feeds ToRGB
            ep.s_next, ep.out_hi, ep.out_lo = sn.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr()
        if not fuse:  # the separate ToRGB kernel reads the fp32 NCHW activation
            y = torch.empty(B, cout, h, h, device=dev)
            ep.out_f32_nchw = y.data_ptr()
            keep.append(y)
        ep.slope, ep.act_scale, ep.activate = 0.2, 2 ** 0.5, 1
        keep += [nz, nw, bias, sn, o_hi, o_lo]
        if fuse:
            wr = torch.randn(B, 3, cout, device=dev)
            ro = torch.empty(B, 3, h, h, device=dev)
            ep.rgb_w, ep.rgb_out = wr.data_ptr(), ro.data_ptr()
            keep += [wr, ro]

    def launch():
        L.call("maua_modconv_tc", x_hi.data_ptr(), x_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), C.byref(ep), B, cin,
               cout, h, h, 1 if up else 0, a.prod, stream)

    try:
        launch()
        launch()
    except Exception:
        return None
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.reps):
        launch()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / a.reps


if a.only:
    cin, cout, h, up = (int(v) for v in a.only.split(","))
    force = tuple(int(v) for v in a.force.split(",")) if a.force else None
    t = run(cin, cout, h, bool(up), (not up) and cout <= 128, (not up) and h == a.size, force)
    print(f"{cin}->{cout} @{h} {'up' if up else 'same'} force={force}: {t:.4f} ms "
          f"({2.0 * h * h * cin * cout * 9 * B / t / 1e9:.0f} TF/s)")
    sys.exit(0)

for cin, cout, h, up, fuse, last in layers:
    if (2 * h if up else h) < a.min_res:
        continue
    flops = 2.0 * h * h * cin * cout * 9 * B
    base = run(cin, cout, h, up, fuse, last, None)
    rows = []
    for groups in ((1, 2, 4) if up else (1,)):
        for r in (1, 2, 4):
            for bn in (16, 32, 64, 128, 256):
                if cout % bn or (fuse and bn != cout):
                    continue
                for cat in ((0, 1) if bn <= 64 else (0,)):
                    t = run(cin, cout, h, up, fuse, last, (r, bn, cat, groups))
                    if t is not None:
                        rows.append((t, r, bn, cat, groups))
    rows.sort()
    print(f"== {cin}->{cout} @{h} {'up' if up else 'same'}{' +rgb' if fuse else ''}: policy {base:.4f} ms "
          f"({flops / base / 1e9:.0f} TF/s)")
    for t, r, bn, cat, groups in rows[:6]:
        print(f"   {t:.4f} ms ({flops / t / 1e9:.0f} TF/s)  R={r} BN={bn} cat={cat} groups={groups}")
    sys.stdout.flush()
