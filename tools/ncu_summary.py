"""Print the handful of ncu metrics we track from a .ncu-rep (development aid): python tools/ncu_summary.py rep [out.txt] [note]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','dram__bytes_read.sum.per_second','dram__bytes_write.sum.per_second','gpu__dram_throughput.avg.pct','launch__registers_per_thread ','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum ','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct','smsp__average_warps_issue_stalled','sm__cycles_elapsed.max ','l1tex__data_pipe_lsu_wavefronts.avg.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ', 'smsp__inst_executed_op_shared', 'lts__throughput.avg.pct']
out = []
if len(sys.argv) > 3: out.append("# " + sys.argv[3])
for vals in rows[2:]:
	for h, u, v in zip(hdr, units, vals):
		if any((h + " ").startswith(w) or h.startswith(w.strip()) for w in want):
			try:
				if abs(float(v)) < 5e-3 and 'stalled' in h: continue
			except ValueError: pass
			out.append(f"{h} [{u}] = {v}")
	out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2: open(sys.argv[2], "w").write(text + "\n")
