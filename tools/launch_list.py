"""Print the kernels of the LAST call bracket in an ncu `--metrics gpu__time_duration.sum --csv` launch list (tools; not product).
usage: launch_list.py file.csv marker_kernel_substring  — prints the launches from the last occurrence of the marker (minus `back`) on"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"], float(r["Metric Value"]) / 1000.0, r["Grid Size"], r["Block Size"]) for r in rows]
mark = sys.argv[2]
back = int(sys.argv[3]) if len(sys.argv) > 3 else 0
idx = [i for i, n in enumerate(names) if mark in n[0]]
s = idx[-1] - back
tot = 0.0
for n in names[s:]:
    short = n[0].replace("<unnamed>::", "").replace("void ", "")
    print(f"{n[1]:10.2f} us  {short[:70]:70s} grid {n[2]} block {n[3]}")
    tot += n[1]
print(f"total {tot:.1f} us over {len(names) - s} launches")
