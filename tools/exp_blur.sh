python -m pytest tests/test_gpu_generator.py tests/test_gpu_tc_conv.py -x -q 2>&1 | tail -2
for r in 8 4; do MAUA_BLUR_RPT=$r python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/blur_rpt$r.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/blur_rpt$r.json"))
print("RPT=$r value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"blur ms",d["kernel_ms_per_step"]["maua_blur_act_nhwc"],"conv",d["kernel_ms_per_step"]["maua_modconv_tc"])
PY
done
