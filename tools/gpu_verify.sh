#!/bin/bash
# Short GPU-box visit: the whole -m gpu suite, smoke, and the two bench lines (no ncu).  Outputs: gpurun_out/<tag>_*.
TAG=${1:-ver}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 700 python bench.py --steps 40 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 400 python bench.py --steps 40 --warmup 5 --precision bf16x3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > $O/${TAG}_bench_bf16x3.json 2> $O/${TAG}_bench_bf16x3.err
tail -4 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log
python - <<PY
import json
for f in ("bench", "bench_bf16x3"):
    d = json.load(open("$O/${TAG}_%s.json" % f))
    print(f, round(d["value"], 1), "fps", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 1), d["clocks"], d["roofline"]["frac"])
    print(d["kernel_ms_per_step"])
    print({k.split(":")[0]: v for k, v in d["roofline"]["per_layer_ms"].items()})
PY
