"""Stall breakdown per warp role (producer / consumer) of the warp-specialised integrate kernel (development aid)."""
import csv, os, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None
role_stall = {'producer': collections.Counter(), 'consumer': collections.Counter()}
inst = collections.Counter()
for r in rows:
	if len(r) == 2 and r[0] == "File Path": cur = os.path.basename(r[1]); continue
	if len(r) == 2: continue
	if r and r[0] == "Line No": hdr = r; continue
	if hdr is None or not r or r[0] == "": continue
	try: line = int(r[0])
	except ValueError: continue
	role = 'producer' if cur in ('duo.cuh', 'mt19937.cuh', 'gauss.cuh') or cur.startswith('sm_30') else 'consumer'
	if cur == 'duo.cuh' and 'PairConsumer' in ''.join(r[1:2]): role = 'consumer'
	if cur == 'duo.cuh' and 70 <= line <= 105: role = 'consumer'
	for h, v in zip(hdr, r):
		if h.startswith('stall_') and 'Not Issued' not in h and v.isdigit():
			role_stall[role][h] += int(v)
	executed = r[hdr.index("Instructions Executed")]
	inst[role] += int(executed) if executed.isdigit() else 0
for role in role_stall:
	tot = sum(role_stall[role].values())
	print(role, 'inst', inst[role], 'samples', tot)
	print("   " + "  ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in role_stall[role].most_common(9)))
