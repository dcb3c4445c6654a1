set -x
mkdir -p gpurun_out
T=${TAG:-r2m}
N=${NGPU:-2}
nvidia-smi -L
python -m pytest tests/test_gpu_multigpu.py -q -rs -x -s > gpurun_out/${T}_pytest_multi.log 2>&1; echo pytest-multi rc=$?
tail -15 gpurun_out/${T}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --precision ${PREC:-mixed} > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo bench rc=$?
tail -5 gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_${N}gpu.json"))
print("N=$N value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms",round(d["ms_per_step"],3), d["e2e"])
PY
