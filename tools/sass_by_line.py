"""Attribute ncu's per-SASS-instruction counts to source lines (development aid).
usage: python tools/sass_by_line.py <ncu source-page csv> <nvdisasm -g -c output> <kernel mangled-name substring>"""
import csv, re, sys, collections
src_csv, sass_file, kname = sys.argv[1:4]
# 1. nvdisasm: sequence of (offset, file:line, inlined chain)
lines = open(sass_file).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = "?"
off2line = {}
for l in lines[start + 1:]:
	if l.startswith("\t.section") or l.startswith("//-----"):
		break
	m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
	if m:
		cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
		continue
	m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
	if m:
		off2line[int(m.group(1), 16)] = (cur, m.group(2))
# 2. ncu: address, source, samples, inst executed
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, isamp, iex, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
pipe = collections.Counter()
tot = [0, 0, 0]
for r in rows[2:]:
	off = int(r[ia], 16) - base
	(loc, _) = off2line.get(off, (("?", 0, False), ""))
	ex, th, sa = int(r[iex]), int(r[ithr]), int(r[isamp])
	a = agg[(loc[0], loc[1])]
	a[0] += ex; a[1] += th; a[2] += sa; a[3] += 1
	tot[0] += ex; tot[1] += th; tot[2] += sa
	op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
	pipe[op.split(".")[0]] += ex
print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} samples {tot[2]}")
print("top source lines by warp instructions executed:")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
	print(f"  {f}:{ln:<5d} inst {a[0]/tot[0]*100:5.1f}%  samples {a[2]/max(tot[2],1)*100:5.1f}%  avg-threads {a[1]/max(a[0],1):5.1f}  sass {a[3]}")
print("by file:")
byf = collections.defaultdict(lambda: [0, 0])
for (f, ln), a in agg.items():
	byf[f][0] += a[0]; byf[f][1] += a[2]
for f, a in sorted(byf.items(), key=lambda kv: -kv[1][0]):
	print(f"  {f:20s} inst {a[0]/tot[0]*100:5.1f}%  samples {a[1]/max(tot[2],1)*100:5.1f}%")
print("top opcodes:")
for op, c in pipe.most_common(25):
	print(f"  {op:10s} {c/tot[0]*100:5.1f}%")

# --- stall breakdown by source line
if "stall_long_sb" in hdr:
	cols = {n: hdr.index(n) for n in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_selected")}
	by = collections.defaultdict(lambda: collections.Counter())
	for r in rows[2:]:
		off = int(r[ia], 16) - base
		(loc, _) = off2line.get(off, (("?", 0, False), ""))
		for n, i in cols.items():
			by[(loc[0], loc[1])][n] += int(r[i] or 0)
	for reason in ("stall_long_sb", "stall_wait"):
		print(f"top lines by {reason}:")
		for key, c in sorted(by.items(), key=lambda kv: -kv[1][reason])[:14]:
			print(f"  {key[0]}:{key[1]:<5d} {c[reason]/max(tot[2],1)*100:5.1f}% of all samples")
