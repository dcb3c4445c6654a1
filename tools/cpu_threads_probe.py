"""Which torch thread count is fastest for the oracle port on this host? (bench.py's cpu_baseline uses the result)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import stylegan2_oracle as O
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd = O.synth_state_dict(size, channel_multiplier=2, seed=0)
_, nl, nlat = O.layout(size)
lat = torch.randn(1, nlat, 512) * 0.5
noise = [torch.randn(1, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2)) for l in range(nl)]
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), "default threads", torch.get_num_threads())
for t in (8, 16, 32, 64, 128):
    torch.set_num_threads(t)
    with torch.no_grad():
        O.generator_forward(sd, size, lat, noise, 1.0, torch.zeros(1, 512))
        t0 = time.perf_counter()
        O.generator_forward(sd, size, lat, noise, 1.0, torch.zeros(1, 512))
    print(t, "threads:", round(time.perf_counter() - t0, 2), "s/frame at", size)
