#!/bin/bash
# Quick GPU check after a kernel change: conv/generator parity tests, per-layer timings, bench line (no ncu).
TAG=${1:-quick}
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_generator.py tests/test_gpu_render.py -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest.log
python tools/tune_tc2.py --only 32,32,1024,0 --reps 4 > $O/${TAG}_layers.log 2>&1
python tools/tune_tc2.py --only 64,64,512,0 --reps 4 >> $O/${TAG}_layers.log 2>&1
python tools/tune_tc2.py --only 64,32,512,1 --reps 4 >> $O/${TAG}_layers.log 2>&1
python tools/tune_tc2.py --only 256,256,128,0 --reps 4 >> $O/${TAG}_layers.log 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_pytest.log $O/${TAG}_layers.log
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"conv ms",d["roofline"]["ms_per_step"],"frac",d["roofline"]["frac"])
print(d["kernel_ms_per_step"]); print(d["roofline"]["per_layer_ms"])
PY
