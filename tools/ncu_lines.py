"""Aggregate an ncu source page (cuda,sass) by source line (development aid):
   ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv, sys, os, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None
per_line = []
hdr = None
for r in rows:
	if len(r) == 2 and r[0] == "File Path":
		cur = os.path.basename(r[1]); continue
	if len(r) == 2: continue
	if r and r[0] == "Line No":
		hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
	if hdr is None or not r or r[0] == "": continue
	try:
		per_line.append((cur, int(r[0]), r[1].strip()[:90], int(r[iI]), int(r[iS])))
	except ValueError:
		pass
totI = sum(x[3] for x in per_line); totS = sum(x[4] for x in per_line)
print("total inst", totI, "samples", totS)
byfile = collections.Counter(); byfileS = collections.Counter()
for f, l, s, i, sm in per_line:
	byfile[f] += i; byfileS[f] += sm
for f, i in byfile.most_common():
	print(f"{f:24s} inst {100*i/totI:5.1f}%  samples {100*byfileS[f]/totS:5.1f}%")
print()
for f, l, s, i, sm in sorted(per_line, key=lambda x: -x[4])[:top]:
	print(f"{f:18s}:{l:4d} inst {100*i/totI:5.2f}% samp {100*sm/totS:5.2f}%  {s}")
