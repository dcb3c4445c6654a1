"""Copy the judged summaries of one GPU-box visit (tools/gpu_round.sh TAG -> gpurun_out/TAG_*) into profiles/ROUND_*:
bench line, ncu launch list (+ per-kernel share summary), ncu --set full per-launch summary, and the DRAM traffic of the
dominant kernels per step (bench.py reads profiles/ROUND_traffic.json for roofline.traffic).
usage: python tools/collect_profiles.py r01b r01"""
import collections
import csv
import json
import os
import shutil
import sys

tag, rnd = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
shutil.copy(f"{G}/{tag}_bench.json", f"{P}/{rnd}_bench_1gpu.json")
for extra in ("bench_bf16x3.json", "audio_summary.txt", "smoke.log"):
    if os.path.exists(f"{G}/{tag}_{extra}"):
        shutil.copy(f"{G}/{tag}_{extra}", f"{P}/{rnd}_{extra.replace('bench_bf16x3', 'bench_1gpu_bf16x3')}")
shutil.copy(f"{G}/{tag}_launches.csv", f"{P}/{rnd}_ncu_launches.csv")
shutil.copy(f"{G}/{tag}_full_summary.txt", f"{P}/{rnd}_ncu_full_summary.txt")
if os.path.exists(f"{G}/{tag}_ufd.json"):
    shutil.copy(f"{G}/{tag}_ufd.json", f"{P}/{rnd}_upfirdn2d.json")
# per-layer source-level captures (tools/prof_layer.sh): top stall lines + the headline counters
with open(f"{P}/{rnd}_ncu_layers.txt", "w") as f:
    f.write("ncu --set full --import-source on --clock-control none, ONE launch of maua_modconv_tc per layer (batch 8, "
            "operand formats of precision 'mixed'; transposed layers in the product form: raw phases, no d), tools/prof_layer.sh; top stall lines of the SASS view + counters\n")
    for name, what in (("l08", "512->512 @64 (split bf16, CTA pairs)"), ("l10", "256->256 @128 (split bf16, CTA pairs)"),
                       ("l13", "128->64 @256 up (fp16 format)"), ("l14", "64->64 @512 +rgb (fp16 format)"),
                       ("l15", "64->32 @512 up (fp16 format)"), ("l16", "32->32 @1024 +rgb (fp16 format)")):
        top, det = f"{G}/{tag}_{name}_top.txt", f"{G}/{tag}_{name}_details.txt"
        if not os.path.exists(top):
            continue
        f.write(f"\n===== {what} =====\n")
        if os.path.exists(det):
            for line in open(det):
                if any(k in line for k in ("Duration", "Elapsed Cycles", "highest-utilized", "Issue Slots Busy", "Eligible Warps",
                                            "Warp Cycles Per Issued", "Registers Per Thread", "DRAM Throughput",
                                            "L2 Cache Throughput")):
                    f.write(line)
        f.write(open(top).read())

rows = list(csv.reader(open(f"{G}/{tag}_launches.csv")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[h]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    n = r[ki].split("(")[0].replace("void ", "")
    if n.startswith("at::"):
        n = n.split("<")[0] + " (torch: synthetic-input setup, not on the step)"
    a = d.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in d.values())
with open(f"{P}/{rnd}_ncu_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none  python tools/profile_step.py --steps 1 --warmup 1 --ufd\n"
            "2 steps (1 warm-up incl. weight packing + 1) of batch 8 of the 1024x1024 config-f hot path (default precision) + 3\n"
            "standalone upfirdn2d calls; cold-cache, serialised: compare SHARES\n\n")
    for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us {100 * t / tot:5.1f} %  x{c:<3d} {n}\n")
    f.write(f"{tot:10.1f} us  total of {sum(v[0] for v in d.values())} launches\n")

full = [l.split(" | ") for l in open(f"{G}/{tag}_full_summary.txt").read().strip().split("\n")[1:]]
traffic = {}
for key, pat in (("maua_modconv_tc", "modconv"), ("maua_blur_act_nhwc", "blur_act"), ("maua_upfirdn2d_f32", "blur_tile_pipe")):
    sel = [r for r in full if pat in r[0]]
    if not sel:
        continue
    per = 3 if pat.startswith("blur_tile") else 1   # blur_tile: three identical standalone calls in the capture
    traffic[key] = {"launches": len(sel) // per, "dram_bytes": sum(float(r[6]) + float(r[7]) for r in sel) * 1e6 / per,
                    "us": sum(float(r[2]) for r in sel) / per}
traffic["source"] = f"ncu --set full --clock-control none, one step of batch 8 (profiles/{rnd}_ncu_full_summary.txt)"
json.dump(traffic, open(f"{P}/{rnd}_traffic.json", "w"), indent=1)
print(open(f"{P}/{rnd}_ncu_launches_summary.txt").read())
print(traffic)
