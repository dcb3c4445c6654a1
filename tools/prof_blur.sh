#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --import-source on --clock-control none -k regex:blur_act --launch-skip 15 -c 1 -o gpurun_out/blur python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/blur.log 2>&1
ncu -i gpurun_out/blur.ncu-rep --page source --csv > gpurun_out/blur_source.csv 2>/dev/null
ncu -i gpurun_out/blur.ncu-rep --page details > gpurun_out/blur_details.txt 2>/dev/null
tail -2 gpurun_out/blur.log
