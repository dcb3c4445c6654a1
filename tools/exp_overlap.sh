# A/B of the side-stream ToRGB branch (MAUA_SYNTH_OVERLAP) + the parity tests that guard it
set -x
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 600 python -m pytest tests/test_gpu_synth_handle.py tests/test_gpu_generator.py tests/test_gpu_render.py -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/${T}_pytest.log
for ov in 0 1 0 1; do
  MAUA_SYNTH_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_ov${ov}.json 2> gpurun_out/${T}_ov${ov}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_ov${ov}.json"))
print("overlap=$ov value",round(d["value"],1),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1), d["kernel_ms_per_step"])
PY
done
