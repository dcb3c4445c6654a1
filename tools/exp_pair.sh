# CTA-pair (cta_group::2) conv: parity tests, then A/B of the whole step and of the per-layer times
set -x
mkdir -p gpurun_out
T=${TAG:-r2p}
timeout 900 python -m pytest tests/test_gpu_tc_conv.py -x -q -m gpu -k "pairs" -s > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?
tail -25 gpurun_out/${T}_pytest.log
if ! grep -q " passed" gpurun_out/${T}_pytest.log || grep -q "failed" gpurun_out/${T}_pytest.log; then echo "PAIR TESTS FAILED"; exit 1; fi
for pr in 0 1 ${EXTRA_PAIR}; do
  MAUA_TC_PAIR=$pr timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_pair${pr}.json 2> gpurun_out/${T}_pair${pr}.err
  tail -3 gpurun_out/${T}_pair${pr}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_pair${pr}.json"))
print("pair=$pr value",round(d["value"],1),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1))
print({k.split(":")[0]:v for k,v in d["roofline"]["per_layer_ms"].items()})
PY
done
