"""Aggregate an ncu source page by CUDA source line: stall samples and executed instructions.  usage: ncu_lines.py REP kernel-regex [top]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
seen_fn = set()
first_fn = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]
        if first_fn is None: first_fn = fn
        cur_fn = fn
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur_fn != first_fn: continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] != "-":   # SASS rows carry an address; CUDA rows have "-" and already hold the aggregate
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, line, r[1][:110])
    s = int(d.get("Warp Stall Sampling (All Samples)", "0") or 0)
    ins = int(d.get("Instructions Executed", "0") or 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += s; a[1] += ins
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
print(f"kernel: {first_fn}\ntotal stall samples {tot_s}, warp instructions {tot_i}\n")
print("| samples % | instr % | file:line | source |\n|---|---|---|---|")
for (f, l, src), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"| {100 * s / tot_s:.1f} | {100 * ins / tot_i:.1f} | {f}:{l} | `{src.strip()}` |")
