set -x
mkdir -p gpurun_out
T=${TAG:-r2j}
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
print(r["per_layer_ms"])
print(d["kernel_ms_per_step"])
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:blur_act --launch-skip 15 -c 1 -o gpurun_out/${T}_blur python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/${T}_blur.log 2>&1
ncu -i gpurun_out/${T}_blur.ncu-rep --page source --csv > gpurun_out/${T}_blur_source.csv 2>/dev/null
ncu -i gpurun_out/${T}_blur.ncu-rep --page details > gpurun_out/${T}_blur_details.txt 2>/dev/null
bash tools/prof_layer.sh ${T}_l16 "32,32,1024,0" - 2
find gpurun_out -name '*.ncu-rep' -size +40M -delete
