"""Instruction budget of a kernel by source file and opcode class, from an ncu source page with SASS (development aid):
   ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_opclasses.py x.csv [attempts]"""
import csv, os, sys, collections

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H", "DMNMX", "F2F.F64", "F2F.F32.F64", "I2F.F64", "F2I.F64", "DMMA")
MEM = ("LDG", "STG", "LDS", "STS", "LDGSTS", "LDC", "LDCU", "LDL", "STL", "ATOM", "RED", "LDGDEPBAR", "DEPBAR", "UBLKCP", "SYNCS")
CTRL = ("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "EXIT", "CALL", "RET", "BREAK", "NOP", "YIELD", "BMOV", "JMP", "USETMAXREG")
WARP = ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX", "POPC", "BREV", "FLO", "R2UR", "S2R", "S2UR", "CS2R")
MOVE = ("MOV", "UMOV", "SEL", "PRMT", "FSEL", "USEL", "CSEL", "P2R", "R2P", "PLOP3", "UPLOP3")

def classify(op):
	base = op.split()[0]
	if base.startswith("@"):
		base = op.split()[1]
	for name, group in (("fp64", FP64), ("memory", MEM), ("control", CTRL), ("warp", WARP), ("move", MOVE)):
		if any(base == g or base.startswith(g + ".") for g in group):
			return name
	if base[0] in "FH" and not base.startswith("FLO"):
		return "fp32"
	return "integer"

rows = csv.reader(open(sys.argv[1]))
attempts = float(sys.argv[2]) if len(sys.argv) > 2 else None
cur, hdr = None, None
table = collections.defaultdict(collections.Counter)
ops = collections.Counter()
for r in rows:
	if len(r) == 2 and r[0] == "File Path":
		cur = os.path.basename(r[1]); continue
	if len(r) == 2:
		continue
	if r and r[0] == "Line No":
		hdr = r; iI = hdr.index("Instructions Executed"); continue
	if hdr is None or not r or r[0] != "" or not r[2].startswith("0x"):
		continue
	n = int(r[iI]) if r[iI].isdigit() else 0
	c = classify(r[3].strip())
	table[cur][c] += n
	ops[(c, r[3].strip().split()[0] if not r[3].strip().startswith("@") else r[3].strip().split()[1])] += n
total = sum(sum(v.values()) for v in table.values())
classes = ["fp64", "integer", "memory", "move", "warp", "control", "fp32"]
scale = 32.0 / attempts if attempts else 100.0 / total
unit = "warp instructions per 32 attempts" if attempts else "% of all instructions"
print(f"total {total:.4g} warp instructions; table in {unit}")
print(f"{'file':26s}" + "".join(f"{c:>9s}" for c in classes) + f"{'sum':>9s}")
for f, cnt in sorted(table.items(), key=lambda kv: -sum(kv[1].values())):
	print(f"{f:26s}" + "".join(f"{cnt[c] * scale:9.1f}" for c in classes) + f"{sum(cnt.values()) * scale:9.1f}")
print(f"{'all':26s}" + "".join(f"{sum(t[c] for t in table.values()) * scale:9.1f}" for c in classes) + f"{total * scale:9.1f}")
print("\ntop opcodes:")
for (c, o), n in ops.most_common(40):
	print(f"  {o:22s} {c:8s} {n * scale:8.1f}")
