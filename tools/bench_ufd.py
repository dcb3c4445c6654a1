"""Micro-benchmark of the public upfirdn2d op on the largest Blur shape of a 1024^2 frame (HBM roofline)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_stylegan2_b200 import op
import bench
pk = bench.peaks()
x = torch.randn(4, 32, 2049, 2049, device="cuda")
kk = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
k4 = kk[None] * kk[:, None] / 16
for _ in range(3):
    y = op.upfirdn2d(x, k4, pad=(1, 1))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    y = op.upfirdn2d(x, k4, pad=(1, 1))
e.record()
torch.cuda.synchronize()
gbs = 4.0 * (x.numel() + y.numel()) / (s.elapsed_time(e) / 10 * 1e-3) / 1e9
print(json.dumps({"upfirdn2d_gbs": gbs, "frac_of_measured_hbm": gbs / pk["hbm_gbs"]}))
