# programmatic dependent launch A/B: smoke (hang detector), the -m gpu suite, then bench with MAUA_PDL=1 / 0
set -x
mkdir -p gpurun_out
T=${TAG:-r2q}
timeout 150 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/${T}_smoke.log
grep -q "smoke:" gpurun_out/${T}_smoke.log || { echo "SMOKE FAILED"; exit 1; }
( timeout 400 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
for pd in 1 0 1 0; do
  MAUA_PDL=$pd timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_pdl${pd}.json 2> gpurun_out/${T}_pdl${pd}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_pdl${pd}.json"))
print("pdl=$pd value",round(d["value"],1),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1), d["clocks"]["reasons"])
PY
done
