# per-layer time of the <= 32^2 layers (v1 kernel) under forced split-K factors: bash tools/exp_small_splitk.sh TAG
T=${1:-sk}
mkdir -p gpurun_out
for S in 0 1 2 4 8 16; do
  if [ $S = 0 ]; then unset MAUA_TC_SPLITK; else export MAUA_TC_SPLITK=$S; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > gpurun_out/${T}_sk$S.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_sk$S.json"))
pl=d["roofline"]["per_layer_ms"]
print("S=$S", " ".join(f"{k.split(':')[0]}={v:.4f}" for k,v in list(pl.items())[:8]), "step", round(d["ms_per_step"],3))
PY
done
