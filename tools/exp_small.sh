for l in 512,512,32,0 512,512,32,1 512,512,16,0 512,512,16,1; do
  python tools/tune_tc2.py --only $l --reps 6
  for f in 1,256,0,1 1,128,0,1 2,128,0,1 2,256,0,1 1,64,0,1 1,256,0,4 1,128,0,2 1,256,0,2 1,128,0,4; do
    MAUA_TC2_MIN_TILES=1 python tools/tune_tc2.py --only $l --reps 6 --force $f 2>&1 | tail -1
  done
done
