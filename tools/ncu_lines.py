"""Summarise an ncu --page source --print-source cuda,sass CSV by CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iline = 0; isrc = 1; iaddr = 2
iexec = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples")
per = collections.OrderedDict(); cur = None
for r in rows[hi + 1:]:
    if len(r) <= iexec: continue
    if r[iline].strip() == 'Line No': continue
    if r[iline].strip().isdigit():
        cur = (int(r[iline]), r[isrc].strip()); per.setdefault(cur, [0, 0])
    if r[iaddr].strip() and cur is not None:
        try:
            per[cur][0] += int(r[iexec]); per[cur][1] += int(r[ismp])
        except ValueError:
            pass
tot = sum(v[0] for v in per.values()); stot = sum(v[1] for v in per.values())
print("total inst", tot, "samples", stot)
for (ln, src), (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{ln:5d} {n/tot*100:5.1f}% inst  {s/max(stot,1)*100:5.1f}% smp  {src[:110]}")
