set -x
mkdir -p gpurun_out
T=${TAG:-r2c}
python -m pytest tests/test_gpu_synth_handle.py -x -q -rs > gpurun_out/${T}_pytest_handle.log 2>&1; echo pytest-handle rc=$?
tail -25 gpurun_out/${T}_pytest_handle.log
python -m pytest tests -m gpu -q -rs -x --deselect tests/test_gpu_tc_conv.py --deselect tests/test_gpu_synth_handle.py > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/${T}_pytest.log
for V in 0 3 4 5; do MAUA_UFD_VARIANT=$V python tools/bench_ufd.py; done 2>&1 | tee gpurun_out/${T}_ufd_variants.log
python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/${T}_ops_v0.log 2>&1; tail -2 gpurun_out/${T}_ops_v0.log
MAUA_UFD_VARIANT=3 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/${T}_ops_v3.log 2>&1; tail -2 gpurun_out/${T}_ops_v3.log
bash tools/prof_layer.sh ${T}_l16 "32,32,1024,0" - 2
bash tools/prof_layer.sh ${T}_l15 "64,32,512,1" - 2
python bench.py --steps 20 --warmup 3 --precision mixed --no-cpu-baseline --no-gpu-reference > gpurun_out/${T}_bench_mixed.json 2> gpurun_out/${T}_bench_mixed.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_mixed.json"))
r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4), "launches", d["gpu_launches"])
PY
find gpurun_out -name '*.ncu-rep' -size +40M -delete
