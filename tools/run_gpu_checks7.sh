set -x
mkdir -p gpurun_out
T=${TAG:-r2h}
timeout 300 ncu --set full --import-source on --clock-control none -k regex:blur_act --launch-skip 8 -c 8 -o gpurun_out/${T}_blur python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/${T}_blur.log 2>&1
ncu -i gpurun_out/${T}_blur.ncu-rep --page raw --csv > gpurun_out/${T}_blur_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${T}_blur_raw.csv | tee gpurun_out/${T}_blur_summary.txt
ncu -i gpurun_out/${T}_blur.ncu-rep --page source --csv --kernel-name regex:blur_act > gpurun_out/${T}_blur_source_all.csv 2>/dev/null
ncu -i gpurun_out/${T}_blur.ncu-rep --page details > gpurun_out/${T}_blur_details.txt 2>/dev/null
find gpurun_out -name '*.ncu-rep' -size +40M -delete
