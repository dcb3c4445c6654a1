#!/bin/bash
# ncu --set full + source counters of the LAST (512->1024, C=32) maua_blur_act_nhwc launch of one step: bash tools/prof_blur.sh TAG
TAG=${1:-blur}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
# profile_step runs the per-operator path: 8 blur_act launches per step, the last one is the 1024^2 layer
timeout 300 ncu --set full --import-source on --clock-control none -k regex:blur_act_nhwc --launch-skip 15 -c 1 -o gpurun_out/$TAG \
    python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/$TAG.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
python tools/ncu_source_top.py gpurun_out/${TAG}_source.csv 30 > gpurun_out/${TAG}_top.txt 2>&1
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Eligible Warps|Warp Cycles Per Issued|Registers Per|Mem Pipes Busy|Theoretical Occupancy|Achieved Occupancy|bank conflict" gpurun_out/${TAG}_details.txt | head -20
cat gpurun_out/${TAG}_top.txt
