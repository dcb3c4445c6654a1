for d in 0 1 4 8; do echo "== MAUA_TC_DBG=$d"; for l in 32,32,1024,0 64,64,512,0 64,32,512,1 256,256,128,0; do MAUA_TC_DBG=$d python tools/tune_tc2.py --only $l --reps 4; done; done
