"""Development aid: the metrics DESIGN.md quotes, from an ncu report: python tools/ncu_summary.py report.ncu-rep "header line" > profiles/x.txt"""
import csv, subprocess, sys
rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
	"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
	"launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
	"lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
	"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
	"smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tensor.sum",
	"sm__inst_executed_pipe_tensor_op_dmma.sum")
print("#", header)
for h, u, v in zip(hdr, units, vals):
	if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) >= 0.05):
		print(f"{h} [{u}] = {v}")
