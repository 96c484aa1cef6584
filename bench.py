#!/usr/bin/env python
"""
Benchmark of the hot path: accepted trajectory-steps / second (FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2-ensemble 20] [--t-end 100]

Workload (BASELINE.json configs[1], SURVEY.md §8 d config 2): noisy Lorenz, additive noise g=(0.1,0.1,0.1) (SRA),
B = 2^20 trajectories per GPU, y0_k = default_rng(42).random((B_total,3))[k], seed_k = k (global index k), one
`integrate(100.0)` from t=0 with the reference's default controller parameters.  One "step" = one such pass.

* `value`  — Σ accepted steps ÷ device time of the integrate kernel (CUDA events on the launching stream), inputs
             resident in HBM; max over ranks.
* `e2e`    — the same passes timed through the public API (`jitcsde.set_initial_value` → `integrate`): host→device
             copy of the initial states, re-seeding, the kernel, device→host copy of the final states, and (N>1) the
             NCCL all-reduce of the ensemble moments — wall clock bracketed by barriers, max over ranks.
* `roofline` — the kernel is bound by the FP64 pipe, not by HBM or tensor cores: achieved = algorithmic FLOP per
             accepted step (SURVEY.md §8 d: ≈235 for SRA-Lorenz) × accepted/s; peak = a dependent-chain DFMA kernel
             measured live on this GPU (MEASURED_PEAKS.json carries no FP64 figure).  The HBM view (≈235 B/attempt
             against MEASURED_PEAKS.json's copy bandwidth) is reported beside it.
* `cpu_baseline` / `--impl reference` — the reference's own C core (oracle/_ref, the reference's shipping flags)
             driven by its Python controller, one process per host core, on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ACCEPTED = 235.0   # SRA-Lorenz, one fresh draw per attempt, 1.04 attempts per accepted step
BYTES_PER_ATTEMPT = 235.0   # generator state + one noise item streamed through HBM
FP64_NOMINAL_TFLOPS = 37.2  # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz
# DRAM traffic of the integrate kernel per attempted step, from one `ncu --set full` capture of the steady state
# (profiles/r01_duo_final_ncu_summary.txt: dram__bytes_read.sum 23.66 GB + dram__bytes_write.sum 16.46 GB for
# 2.7335e8 attempts at B=2^18).  A full-size launch (17 s) cannot be replayed under ncu within the GPU budget, so
# `roofline.traffic` is this per-attempt figure times the attempts of one launch.
NCU_DRAM_BYTES_PER_ATTEMPT = (23.661648e9 + 16.464983e9) / 273349641


def measured_peaks():
	try:
		with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
			return json.load(fh), "measured"
	except Exception:
		return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
	"""nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
	Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
		"clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

	def __init__(self, index):
		self.index = index
		self.rows = []
		self.proc = None

	def start(self):
		try:
			self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			self.thread = threading.Thread(target=self._read, daemon=True)
			self.thread.start()
		except Exception:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			parts = [p.strip() for p in line.split(",")]
			if len(parts) >= 7:
				self.rows.append(parts)

	def stop(self):
		if not self.proc:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except Exception:
			self.proc.kill()
		sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
		mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
		pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
		names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
		reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
		return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
			"power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
	"""One process per core: the reference C core + its (transcribed) Python controller on `count` trajectories."""
	core, first, count, t_end, kind, flavour = args
	try:
		os.sched_setaffinity(0, {core})
	except Exception:
		pass
	sys.path.insert(0, ROOT)
	from jitcsde_b200 import models
	spec = models.lorenz_additive
	rng_rows = np.random.default_rng(42).random((first + count, 3))[first:]
	accepted = 0
	t0 = time.perf_counter()
	if kind == "reference":
		from oracle import build_ref, ref_driver
		mod = build_ref.load(spec, flavour)
		for j in range(count):
			R = ref_driver.RefIntegrator(mod, rng_rows[j], 0.0, first + j)
			R.integrate(t_end)
			accepted += R.accepted
	else:
		from oracle import port
		res = port.run_ensemble(spec, rng_rows, np.arange(first, first + count, dtype=np.uint32), 0.0, t_end, flavour=flavour)
		accepted = int(res["accepted"].sum())
	return accepted, time.perf_counter() - t0


def cpu_reference_pass(per_core, t_end, pool, cores, kind, flavour):
	"""All host cores, `per_core` trajectories each; returns (accepted, wall seconds)."""
	jobs = [(c, c * per_core, per_core, t_end, kind, flavour) for c in range(cores)]
	t0 = time.perf_counter()
	out = pool.map(_cpu_worker, jobs)
	wall = time.perf_counter() - t0
	return sum(a for a, _ in out), wall


def cpu_kind():
	from jitcsde_b200 import models
	from oracle import build_ref
	if build_ref.available(models.lorenz_additive, "default"):
		return "reference", "default"
	return "port", "default"


def host_cores():
	try:
		return len(os.sched_getaffinity(0))
	except Exception:
		return os.cpu_count() or 1


def run_reference_arm(args):
	import multiprocessing as mp
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	cores = host_cores()
	kind, flavour = cpu_kind()
	per_core = args.cpu_per_core
	ctx = mp.get_context("fork")
	with ctx.Pool(cores) as pool:
		for _ in range(args.warmup):
			cpu_reference_pass(per_core, args.t_end, pool, cores, kind, flavour)
		acc = 0
		t0 = time.perf_counter()
		for _ in range(args.steps):
			a, _ = cpu_reference_pass(per_core, args.t_end, pool, cores, kind, flavour)
			acc += a
		wall = time.perf_counter() - t0
	value = acc / wall
	sample = f"{cores * per_core} trajectories per step ({per_core} per core, global indices 0..{cores * per_core - 1}) integrated to t={args.t_end}"
	line = {
		"impl": "reference", "metric": "accepted trajectory-steps/sec (FP64)", "value": value, "unit": "steps/s",
		"n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
		"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": workload_config(args),
		"cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample,
			"flags": "reference default (-Ofast -march=native)" if flavour == "default" else "strict IEEE"},
		"e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	}
	print(json.dumps(line), flush=True)


def workload_config(args):
	return {"workload": f"noisy Lorenz ensemble (BASELINE.json configs[1]): additive noise g=0.1 (SRA), 2^{args.log2_ensemble} trajectories per GPU, "
		f"t=0..{args.t_end:g}, default controller, y0=default_rng(42).random, seed_k=k",
		"ensemble_per_gpu": 1 << args.log2_ensemble, "t_end": args.t_end, "n": 3,
		"cache": "working set (2.6 GB generator state per 2^20 trajectories) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
	import torch
	import torch.distributed as dist

	from jitcsde_b200 import _abi, jitcsde, models

	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback in the product path")
	torch.cuda.set_device(local)
	if world > 1:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local))

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	B = 1 << args.log2_ensemble
	spec = models.lorenz_additive
	# global index k = rank*B + j decides y0 and seed, so results do not depend on the number of GPUs
	from jitcsde_b200 import sharding
	lo, hi = sharding.shard_range(world * B, rank, world)  # weak scaling: B per GPU
	y0 = sharding.global_initial_states(lo, hi, 3)
	y0_pinned = torch.from_numpy(np.ascontiguousarray(y0)).pin_memory()
	seeds = sharding.global_seeds(lo, hi)

	SDE = jitcsde(spec=spec, ensemble=B, device=local, verbose=False)
	SDE.set_seed(seeds)
	SDE.set_integration_parameters()
	lib = _abi.load_library(spec)
	fp64_peak = _abi.measure_fp64_peak(lib, local)

	out_host = torch.empty((B, 3), dtype=torch.float64).pin_memory()

	def one_step():
		"""public API, host buffers in, host buffers out; returns (stats, moments)"""
		SDE.set_initial_value(y0_pinned.numpy(), 0.0)
		SDE._initiate()
		stats = SDE.SDE.integrate(args.t_end)
		SDE.SDE.get_state(out=out_host.numpy())
		mom = SDE.ensemble_moments() if world > 1 else None
		return stats, mom

	for _ in range(args.warmup):
		one_step()
	sampler = ClockSampler(local)
	if rank == 0:
		sampler.start()
	barrier()
	t0 = time.perf_counter()
	kernel_ms = 0.0
	accepted = rejected = launches = 0
	for _ in range(args.steps):
		stats, mom = one_step()
		kernel_ms += stats["kernel_ms"]
		accepted += stats["accepted"]
		rejected += stats["rejected"]
	barrier()
	wall = time.perf_counter() - t0
	clocks = sampler.stop() if rank == 0 else None
	# kernels launched per step: set_state (rows->cols, fill t, seed) + integrate + get_state (cols->rows) [+ moments]
	launches = args.steps * (5 + (1 if world > 1 else 0))
	failed = stats["n_min_step"] + stats["n_overflow"]

	# aggregate over ranks: totals summed, times max
	vec = torch.tensor([accepted, rejected, kernel_ms, wall, failed], dtype=torch.float64, device="cuda")
	if world > 1:
		tot = vec.clone()
		dist.all_reduce(tot, op=dist.ReduceOp.SUM)
		mx = vec.clone()
		dist.all_reduce(mx, op=dist.ReduceOp.MAX)
		accepted, rejected, failed = int(tot[0].item()), int(tot[1].item()), int(tot[4].item())
		kernel_ms, wall = mx[2].item(), mx[3].item()

	if rank == 0:
		peaks, peaks_src = measured_peaks()
		value = accepted / (kernel_ms * 1e-3)
		e2e = accepted / wall
		attempts = accepted + rejected
		achieved_tflops = value * FLOP_PER_ACCEPTED / 1e12 / world
		hbm_gbs = attempts / (kernel_ms * 1e-3) * BYTES_PER_ATTEMPT / 1e9 / world
		line = {
			"metric": "accepted trajectory-steps/sec (FP64)", "value": value, "unit": "steps/s", "n_gpus": world,
			"steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True,
			"scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
			"e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": int(B * 3 * 8 + B * 4), "d2h_bytes_per_step": int(B * 3 * 8),
				"ms_per_step": wall / args.steps * 1e3},
			"gpu_launches": launches, "clocks": clocks,
			"accepted_per_step": accepted // args.steps, "rejected_per_step": rejected // args.steps, "failed_trajectories": failed,
			"attempts_per_s": attempts / (kernel_ms * 1e-3), "accept_ratio": accepted / attempts if attempts else None,
			"roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp64_peak,
				"traffic": NCU_DRAM_BYTES_PER_ATTEMPT * attempts / args.steps / world,
				"traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per attempt (B=2^18 steady-state capture, "
					"profiles/r01_duo_final_ncu_summary.txt) x attempts of one launch; algorithmic bytes per launch = "
					f"{BYTES_PER_ATTEMPT:g} B x attempts",
				"per_gpu": True, "flop_per_accepted_step": FLOP_PER_ACCEPTED,
				"peak_source": "dependent-chain DFMA kernel measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry); nominal 37.2",
				"hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0),
					"bytes_per_attempt": BYTES_PER_ATTEMPT, "peak_source": peaks_src}},
		}
		if world == 1 and not args.no_cpu_baseline:
			line["cpu_baseline"] = cpu_baseline(args)
		print(json.dumps(line), flush=True)
	if world > 1:
		dist.destroy_process_group()


def cpu_baseline(args):
	import multiprocessing as mp
	cores = host_cores()
	kind, flavour = cpu_kind()
	per_core = args.cpu_per_core
	ctx = mp.get_context("fork")
	with ctx.Pool(cores) as pool:
		acc, wall = cpu_reference_pass(per_core, args.t_end, pool, cores, kind, flavour)
	return {"value": acc / wall, "unit": "steps/s", "cores": cores, "kind": kind,
		"sample": f"{cores * per_core} trajectories ({per_core} per core) of the same workload integrated to t={args.t_end:g}; "
			f"{'reference C core + its Python controller, reference default flags' if kind == 'reference' else 'C restatement, C controller'}",
		"seconds": wall}


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=2)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
	ap.add_argument("--log2-ensemble", type=int, default=20)
	ap.add_argument("--t-end", type=float, default=100.0)
	ap.add_argument("--cpu-per-core", type=int, default=192, help="trajectories per host core in one CPU-reference pass")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	args = ap.parse_args()
	if args.impl == "reference":
		run_reference_arm(args)
	else:
		run_gpu_arm(args)


if __name__ == "__main__":
	main()
