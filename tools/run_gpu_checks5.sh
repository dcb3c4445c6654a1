set -x
mkdir -p gpurun_out
T=${TAG:-r2f}
python -m pytest tests/test_gpu_generator.py -q -x -s -k "1024" 2>&1 | grep -E "1024 |passed|failed" | tee gpurun_out/${T}_errors_1024.log
python -m pytest tests/test_gpu_comparator.py -q -x -s 2>&1 | grep -E "reference CUDA|passed|failed" | tee -a gpurun_out/${T}_errors_1024.log
for LIBTAG in default epi4; do
  if [ $LIBTAG = epi4 ]; then export MAUA_LIB_PATH=$PWD/maua_stylegan2_b200/libmaua_b200_epi4.so; else unset MAUA_LIB_PATH; fi
  python -m pytest tests/test_gpu_tc_conv.py -q -x > gpurun_out/${T}_conv_$LIBTAG.log 2>&1; tail -2 gpurun_out/${T}_conv_$LIBTAG.log
  python bench.py --steps 20 --warmup 3 --precision mixed --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_bench_mixed_$LIBTAG.json 2> gpurun_out/${T}_bench_mixed_$LIBTAG.err; echo bench rc=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_mixed_$LIBTAG.json"))
r=d["roofline"]
print("$LIBTAG mixed value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3),"conv ms",round(r["ms_per_step"],3),"frac",round(r["frac"],4))
print(r["per_layer_ms"])
PY
done
unset MAUA_LIB_PATH
