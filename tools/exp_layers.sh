for l in 32,32,1024,0 64,64,512,0 128,128,256,0 64,32,512,1 256,256,128,0 128,64,256,1; do python tools/tune_tc2.py --only $l --reps 4; done
