set -x
mkdir -p gpurun_out
T=${TAG:-r2t}
timeout 500 python tools/tune_tc2.py --prod 3 --min-res 64 > gpurun_out/${T}_tune_bf16x3.log 2>&1
cat gpurun_out/${T}_tune_bf16x3.log
timeout 300 python tools/tune_tc2.py --prod 2 --min-res 512 > gpurun_out/${T}_tune_f16.log 2>&1
cat gpurun_out/${T}_tune_f16.log
