"""Minimal driver for ncu: W warm-up + K steps of the bench workload's hot path (same generator, batch and inputs as
bench.py `value`), nothing else.  Optional --ufd adds the standalone upfirdn2d op on the largest Blur shape."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--precision", default=bench.DEFAULT_PRECISION)
ap.add_argument("--ufd", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lat, noise = bench.make_workload(64)
g = bench.make_generator(dev, precision=a.precision)
lat = lat.to(dev)
noise = [n.to(dev) if n is not None else None for n in noise]
with torch.no_grad():
    for i in range(a.warmup + a.steps):
        n = (i * a.batch) % 56
        g(lat[n:n + a.batch], noise=[x[n:n + a.batch] if x is not None else None for x in noise], truncation=1.0,
          input_is_latent=True, randomize_noise=False, return_u8=True)
    if a.ufd:
        from maua_stylegan2_b200 import op

        x = torch.randn(4, 32, 2049, 2049, device=dev)
        kk = torch.tensor([1.0, 3.0, 3.0, 1.0], device=dev)
        for _ in range(3):
            op.upfirdn2d(x, kk[None] * kk[:, None] / 16, pad=(1, 1))
torch.cuda.synchronize()
print("done")
