set -x
mkdir -p gpurun_out
T=${TAG:-r2i}
python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_generator.py tests/test_gpu_synth_handle.py tests/test_gpu_reference_goldens.py -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
r=d["roofline"]
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
print(d["kernel_ms_per_step"])
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_memcheck.log 2>&1; echo memcheck rc=$?
tail -6 gpurun_out/${T}_memcheck.log
