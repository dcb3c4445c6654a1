set -x
mkdir -p gpurun_out
T=${TAG:-r2d}
python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_synth_handle.py -x -q -rs > gpurun_out/${T}_pytest_conv.log 2>&1; echo pytest-conv rc=$?
tail -12 gpurun_out/${T}_pytest_conv.log
python -m pytest tests -m gpu -q -rs -x --deselect tests/test_gpu_tc_conv.py --deselect tests/test_gpu_synth_handle.py > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/${T}_pytest.log
for P in mixed bf16x3; do
python bench.py --steps 20 --warmup 3 --precision $P --no-cpu-baseline --no-gpu-reference > gpurun_out/${T}_bench_$P.json 2> gpurun_out/${T}_bench_$P.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_$P.json"))
r=d["roofline"]
print("$P value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4), "launches", d["gpu_launches"])
print(r["per_layer_ms"])
print(d["kernel_ms_per_step"])
PY
done
bash tools/prof_layer.sh ${T}_l16 "32,32,1024,0" - 2
bash tools/prof_layer.sh ${T}_l15 "64,32,512,1" - 2
bash tools/prof_layer.sh ${T}_l14 "64,64,512,0" - 2
python tools/tune_tc2.py --prod 2 --min-res 512 > gpurun_out/${T}_tune_f16.log 2>&1; tail -32 gpurun_out/${T}_tune_f16.log
find gpurun_out -name '*.ncu-rep' -size +40M -delete
