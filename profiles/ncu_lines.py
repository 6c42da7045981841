"""Per-source-line instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py <csv> <kernel index 0..> [top N]   -- prints the hottest CUDA source lines of one kernel"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
want = int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# kernels are delimited by the first "File Path" row of each group; functions repeat per file
kern = -1; last_fn = None; cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if r[1] != last_fn: kern += 1; last_fn = r[1]
        continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if kern != want or hdr is None or r[0] == "": continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    g = lambda k: float(r[hdr[k]]) if r[hdr[k]] not in ("-", "") else 0.0
    agg[(cur_file, line)] = (g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples"), r[1].strip()[:110])
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print(f"kernel {want}: {last_fn if kern==want else ''} total warp instructions {tot:.3g}, samples {tots:.0f}")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:4d} inst {100*v[0]/tot:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}  samples {100*v[2]/max(tots,1):5.1f}%  {v[3]}")
