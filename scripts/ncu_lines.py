"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > dump.csv; python scripts/ncu_lines.py dump.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname = None; hdr = None; seen_kernels = 0
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue          # SASS rows
    key = (fname, int(r[0]))
    i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
    try: s = int(r[i_s]); ins = int(r[i_i])
    except ValueError: continue
    if key in agg: continue           # second launch repeats the same table
    agg[key] = (s, ins, r[1].strip()[:110])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}  total warp-instructions {ti}")
for (f, l), (s, ins, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*ins/ti:5.1f}% ins  {f}:{l:<4d} {src}")
