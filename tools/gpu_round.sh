#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full capture.  Outputs land in gpurun_out/<tag>_*.
# usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01b'
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 40 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python tools/bench_ufd.py > $O/${TAG}_ufd.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python tools/profile_step.py --steps 1 --warmup 1 --ufd > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'modconv_tc|blur_act|blur_tile|torgb|rgb_finish' \
    --launch-skip 0 -c 60 -o $O/${TAG}_full python tools/profile_step.py --steps 1 --warmup 0 --ufd > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full_raw.csv > $O/${TAG}_full_summary.txt 2>&1
find $O -name '*.ncu-rep' -size +30M -delete
ls -la $O | head -40
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; cat $O/${TAG}_ufd.json
