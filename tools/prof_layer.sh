#!/bin/bash
# ncu --set full + source counters of ONE conv layer: bash tools/prof_layer.sh TAG "cin,cout,h,up" [force|-] [prod]
TAG=$1; LAYER=$2; FORCE=""; if [ -n "$3" ] && [ "$3" != "-" ]; then FORCE="--force $3"; fi; PROD=${4:-3}
cd ${GRAFT_REPO_ROOT:-.}
timeout 200 ncu --set full --import-source on --clock-control none -k regex:modconv_tc2 --launch-skip 2 -c 1 -o gpurun_out/$TAG python tools/tune_tc2.py --only $LAYER $FORCE --reps 1 --prod $PROD > gpurun_out/$TAG.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
tail -2 gpurun_out/$TAG.log
