"""Summarise an ncu --set full report (raw CSV page) into a per-launch markdown table.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv > profiles/....md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
def g(r, name, default=""):
    return r[hdr.index(name)] if name in hdr else default
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor smem-operand path %"),
        ("dram__bytes_read.sum", "dram rd (GB)"), ("dram__bytes_write.sum", "dram wr (GB)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem wavefronts (tensor)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (lsu)"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
print("| " + " | ".join(c[1] + (" [" + units[hdr.index(c[0])] + "]" if c[0] in hdr and units[hdr.index(c[0])] else "") for c in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    out = []
    for name, _ in cols:
        v = g(r, name)
        if name == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f}" if f < 1000 else f"{f:.3e}"
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
