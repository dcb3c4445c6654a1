set -x
mkdir -p gpurun_out
T=${TAG:-r2k}
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python tools/bench_audio.py --seconds 30 --no-oracle > gpurun_out/${T}_audio30.json 2> gpurun_out/${T}_audio30.err; cat gpurun_out/${T}_audio30.json
python tools/bench_audio.py --seconds 300 --no-oracle > gpurun_out/${T}_audio300.json 2> gpurun_out/${T}_audio300.err; cat gpurun_out/${T}_audio300.json; tail -2 gpurun_out/${T}_audio300.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
PY
