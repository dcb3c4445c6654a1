set -x
mkdir -p gpurun_out
T=${TAG:-r2g}
python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_generator.py tests/test_gpu_synth_handle.py -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for RPT in 8 4; do
MAUA_BLUR_RPT=$RPT python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench_rpt$RPT.json 2> gpurun_out/${T}_bench_rpt$RPT.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_rpt$RPT.json"))
r=d["roofline"]
print("rpt$RPT value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
print(r["per_layer_ms"])
print(d["kernel_ms_per_step"])
PY
done
