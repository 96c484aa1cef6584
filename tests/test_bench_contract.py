"""The driver parses ONE JSON line from bench.py; this checks the reference arm (CPU only: the reference's C core and
its Python controller on a tiny sample) carries every key the contract names, with the right types."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
	res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
		"--cpu-per-core", "1", "--t-end", "1"], capture_output=True, text=True, timeout=600)
	assert res.returncode == 0, res.stderr[-2000:]
	lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
	assert len(lines) == 1
	d = json.loads(lines[0])
	assert d["impl"] == "reference" and d["metric"] == "accepted trajectory-steps/sec (FP64)" and d["unit"] == "steps/s"
	assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
	assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
	assert "workload" in d["config"] and "model" not in d["config"]
	cb = d["cpu_baseline"]
	assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
	e2e = d["e2e"]
	assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_b200_arm_fails_loudly_without_a_gpu():
	"""No CPU fallback: on a machine without a CUDA device the product arm must exit non-zero, not print a number."""
	import torch
	if torch.cuda.is_available():
		import pytest
		pytest.skip("a GPU is present")
	res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--log2-ensemble", "5", "--t-end", "0.1",
		"--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
	assert res.returncode != 0
	assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
