set -x
mkdir -p gpurun_out
T=${TAG:-r2b}
python -m pytest tests/test_gpu_tc_conv.py -x -q -rs -s > gpurun_out/${T}_pytest_conv.log 2>&1; echo pytest-conv rc=$?
tail -15 gpurun_out/${T}_pytest_conv.log
python -m pytest tests -m gpu -q -rs --durations=10 --deselect tests/test_gpu_tc_conv.py > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -40 gpurun_out/${T}_pytest.log
for P in bf16x3 mixed; do
python bench.py --steps 20 --warmup 3 --precision $P > gpurun_out/${T}_bench_$P.json 2> gpurun_out/${T}_bench_$P.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_$P.json"))
r=d["roofline"]
print("$P", "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
print(r["per_layer_ms"])
print(d["kernel_ms_per_step"])
print(d.get("gpu_reference"))
PY
done
python tools/tune_tc2.py --prod 2 --min-res 512 > gpurun_out/${T}_tune_f16.log 2>&1; tail -40 gpurun_out/${T}_tune_f16.log
