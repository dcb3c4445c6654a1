# collapsed-tap transposed conv (MAUA_TC_COLL): parity tests, then A/B of the step
set -x
mkdir -p gpurun_out
T=${TAG:-r2c}
timeout 200 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_generator.py tests/test_gpu_synth_handle.py -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${T}_pytest.log
for cl in 1 0; do
  MAUA_TC_COLL=$cl timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_coll${cl}.json 2> gpurun_out/${T}_coll${cl}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_coll${cl}.json"))
pl=d["roofline"]["per_layer_ms"]
print("coll=$cl value",round(d["value"],1),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1), {k.split(":")[0]:v for k,v in pl.items() if k.startswith(("L13","L15","L9:","L11"))}, d["kernel_ms_per_step"]["maua_modconv_tc"])
PY
done
