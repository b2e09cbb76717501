"""One forward step of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py next to the live
CUDA-event times of the same launches (the `layers` key of a plain bench.py run of the same command).
usage: python tools/launch_vs_live.py launches.csv bench.json [step_index] > profiles/x.md"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, bi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
data = [r for r in rows[h + 1:] if len(r) > vi]
bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
layers = [l for l in bench["layers"] if l["launch"] != "head"]      # the 1x1 head is fused into the last conv
starts = [i for i, r in enumerate(data) if "prologue_kernel" in r[ki]]
s = starts[int(sys.argv[3]) if len(sys.argv) > 3 else 4]
step = data[s:s + len(layers)]
ncu_ms = [float(r[vi].replace(",", "")) / 1e6 for r in step]
live_ms = [l["ms"] for l in layers]
print("| # | kernel | grid | block | ncu ms | ncu share | live ms (bench.py) | live share |\n|---|---|---|---|---|---|---|---|")
for i, (r, l) in enumerate(zip(step, layers)):
    name = r[ki].split("(")[0].replace("void ", "")
    print(f"| {i} | {name} ({l['launch']}) | {r[gi]} | {r[bi]} | {ncu_ms[i]:.4f} | {100 * ncu_ms[i] / sum(ncu_ms):.1f} % | "
          f"{live_ms[i]:.4f} | {100 * live_ms[i] / sum(live_ms):.1f} % |")
print(f"| | **total** | | | {sum(ncu_ms):.3f} | | {sum(live_ms):.3f} | |")
