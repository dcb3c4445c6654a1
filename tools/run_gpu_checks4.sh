set -x
mkdir -p gpurun_out
T=${TAG:-r2e}
python -m pytest tests -m gpu -q -rs -x > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -12 gpurun_out/${T}_pytest.log
for P in mixed bf16x3; do
python bench.py --steps 20 --warmup 3 --precision $P --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench_$P.json 2> gpurun_out/${T}_bench_$P.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_$P.json"))
r=d["roofline"]
print("$P value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4), "launches", d["gpu_launches"])
print(r["per_layer_ms"])
print(d["kernel_ms_per_step"])
print(d["roofline_upfirdn2d"])
PY
done
python tools/bench_audio.py --seconds 30 > gpurun_out/${T}_audio30.json 2> gpurun_out/${T}_audio30.err; cat gpurun_out/${T}_audio30.json; tail -3 gpurun_out/${T}_audio30.err
bash tools/prof_layer.sh ${T}_l16 "32,32,1024,0" - 2
find gpurun_out -name '*.ncu-rep' -size +40M -delete
