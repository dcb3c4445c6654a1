"""Summarise an ncu --set full report (raw page CSV) into one line per launch: the metrics DESIGN.md / bench.py cite.
usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_summary.py X.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
C = {h: i for i, h in enumerate(hdr)}
cols = [
    ("kernel", "Kernel Name"), ("grid", "Grid Size"), ("us", "gpu__time_duration.sum"),
    ("tc_pipe%", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("hmma_ops%", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("smemKB", "launch__shared_mem_per_block_dynamic"),
    ("waves", "launch__waves_per_multiprocessor"),
]
units = rows[1]
print(" | ".join(n for n, _ in cols))
for r in data:
    out = []
    for n, c in cols:
        if c not in C:
            out.append("-")
            continue
        v, u = r[C[c]], units[C[c]]
        if n == "kernel":
            v = v.replace("void ", "").replace("maua::", "")[:34]
        else:
            try:
                f = float(v.replace(",", ""))
                if n == "us":
                    f = f / 1e3 if u == "ns" else (f * 1e3 if u == "ms" else f)
                if n.endswith("_MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * f
                if n == "smemKB":
                    f = {"byte": 1 / 1024, "Kbyte": 1.0}.get(u, 1 / 1024) * f
                v = f"{f:.1f}"
            except ValueError:
                pass
        out.append(v)
    print(" | ".join(out))
