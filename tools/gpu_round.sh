#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture (+ audio chain).  Outputs: gpurun_out/<tag>_*.
# usage (here): gpurun --timeout 1700 -- 'bash tools/gpu_round.sh r02a'   then   python tools/collect_profiles.py r02a r02
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log
( timeout 300 python __graft_entry__.py smoke ) > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 700 python bench.py --steps 40 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 400 python bench.py --steps 40 --warmup 5 --precision bf16x3 --no-cpu-baseline --no-gpu-reference --no-audio-chain > $O/${TAG}_bench_bf16x3.json 2> $O/${TAG}_bench_bf16x3.err
timeout 300 python tools/bench_ufd.py > $O/${TAG}_ufd.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python tools/profile_step.py --steps 1 --warmup 1 --ufd > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'modconv_tc|blur_act|blur_tile|torgb|rgb_finish|style_' \
    --launch-skip 0 -c 70 -o $O/${TAG}_full python tools/profile_step.py --steps 1 --warmup 0 --ufd > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full_raw.csv > $O/${TAG}_full_summary.txt 2>&1
# audio chain: every kernel of the default hooks on 30 s of audio (second pass of bench_audio = warm)
timeout 400 ncu --set full --clock-control none -k regex:'hpss|gaussian_filter|nn_filter|mm_filt|mm_onset|filterbank|frame_|overlap_add|cens_|resample|percentile|envelope|chroma_weight' \
    --launch-skip 100 -c 110 -o $O/${TAG}_audio python tools/bench_audio.py --seconds 30 --no-oracle > $O/${TAG}_audio.log 2>&1
ncu -i $O/${TAG}_audio.ncu-rep --page raw --csv > $O/${TAG}_audio_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_audio_raw.csv > $O/${TAG}_audio_summary.txt 2>&1
# per-layer source-level captures: "cin,cout,h,up:name:products" (3 = split bf16, CTA pairs for BN >= 128; 2 = fp16 format)
for L in "32,32,1024,0:l16:2" "64,32,512,1:l15:2" "64,64,512,0:l14:2" "128,64,256,1:l13:2" "256,256,128,0:l10:3" "512,512,64,0:l08:3"; do
  SHAPE=${L%%:*}; REST=${L#*:}; NAME=${REST%%:*}; PROD=${REST#*:}
  bash tools/prof_layer.sh ${TAG}_${NAME} "$SHAPE" - $PROD
  python tools/ncu_source_top.py $O/${TAG}_${NAME}_source.csv 14 > $O/${TAG}_${NAME}_top.txt 2>&1
done
find $O -name '*.ncu-rep' -size +30M -delete
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json | head -c 1500; cat $O/${TAG}_ufd.json
