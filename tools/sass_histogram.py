"""Development aid: opcode histogram of a built model library's kernels (cuobjdump -sass), for profiles/.
usage: python tools/sass_histogram.py <model name> [kernel-name substring ...]"""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from jitcsde_b200 import _build, models
name = sys.argv[1]
spec = models.fhn_1000() if name == "fhn_1000" else models.ALL[name]
lib = _build.build(spec)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
wanted = sys.argv[2:] or ["k_integrate", "k_wide_integrate"]
print(f"# cuobjdump -sass {os.path.relpath(lib)}  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false)")
for block in out.split("\tFunction : ")[1:]:
	fn = block.split("\n", 1)[0].strip()
	if not any(w in fn for w in wanted):
		continue
	arch = re.search(r"EF_CUDA_SM(\d+)", block)
	ops = collections.Counter()
	for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", block):
		ops[m.group(1)] += 1
	total = sum(ops.values())
	key = ("DFMA", "DMUL", "DADD", "DSETP", "DMMA", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "USETMAXREG", "BAR", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "LOP3", "IMAD", "SHF")
	print(f"\n{fn}\n  {total} instructions; " + ", ".join(f"{k} {ops[k]}" for k in key if ops[k]))
	print("  top: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
