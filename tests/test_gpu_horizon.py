"""GPU: parity at the CONFIGURED sizes and horizons of BASELINE.json (round-1 verdict: the headline checks stopped at T = 2
while the benchmark runs to T = 100).

* configs[1] (headline): 2^20 additive-Lorenz trajectories to T = 100 — a strided 128-trajectory sample bit-identical to
  the oracle (states, accepted and rejected counts), ensemble mean/variance within 3 standard errors of an independent
  oracle ensemble of 4096 trajectories (SURVEY.md §8 c ladder rungs 4-5);
* T = 100 goldens generated from the real reference core (tests/golden/reference_core_T100.json), SRA and SRI;
* configs[0] verbatim: examples/noisy_lorenz.py:89-96 — 10 000 samples on arange(0, 100, 0.01), one trajectory;
* configs[4] at its per-GPU size: 2^11 trajectories of the n = 1000 network to T = 1.
"""

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
from oracle import port
from util import GOLDEN_DIR, assert_bits_equal, unhex

pytestmark = pytest.mark.gpu


def golden_long():
	with open(os.path.join(GOLDEN_DIR, "reference_core_T100.json")) as fh:
		return json.load(fh)


def oracle_ensemble(spec, y0, seeds, T, workers=None):
	"""port.run_ensemble on all host cores (ctypes releases the GIL)"""
	workers = workers or min(32, len(os.sched_getaffinity(0)))
	chunks = np.array_split(np.arange(len(seeds)), workers)
	with ThreadPoolExecutor(workers) as pool:
		parts = list(pool.map(lambda idx: port.run_ensemble(spec, y0[idx], seeds[idx], 0.0, T), [c for c in chunks if len(c)]))
	return {key: np.concatenate([p[key] for p in parts]) for key in parts[0]}


def test_headline_config_at_full_size_and_horizon():
	spec, B, T = models.lorenz_additive, 1 << 20, 100.0
	y0 = sharding.global_initial_states(0, B, 3)
	seeds = sharding.global_seeds(0, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	stats = E.integrate(T)
	y, t, status = E.get_state(), E.t, E.status
	dt, acc, rej = E.controller_state()
	assert (status == 0).all() and (t == T).all() and np.isfinite(y).all()
	assert int(acc.sum()) == stats["accepted"] and int(rej.sum()) == stats["rejected"]
	# rung 4: the strided sample is the oracle's, bit for bit, after ~1.8e5 steps each
	idx = np.arange(0, B, B // 128)
	ref = oracle_ensemble(spec, y0[idx], seeds[idx], T)
	assert_bits_equal(y[idx], ref["y"], "sampled trajectories at T=100")
	assert (acc[idx].astype(np.int64) == ref["accepted"]).all()
	assert (rej[idx].astype(np.int64) == ref["rejected"]).all()
	# rung 5: moments against an independent oracle ensemble (other seeds, same initial-state rule)
	Bo = 4096
	ind = oracle_ensemble(spec, sharding.global_initial_states(0, Bo, 3), np.arange(Bo, dtype=np.uint32) + np.uint32(3_000_000), T)["y"]
	for i in range(3):
		se = np.sqrt(y[:, i].var() / B + ind[:, i].var() / Bo)
		assert abs(y[:, i].mean() - ind[:, i].mean()) <= 3 * se, f"mean {i}: {y[:, i].mean()} vs {ind[:, i].mean()} (SE {se})"
		vg, vo = y[:, i].var(ddof=1), ind[:, i].var(ddof=1)
		m4g = ((y[:, i] - y[:, i].mean()) ** 4).mean()
		m4o = ((ind[:, i] - ind[:, i].mean()) ** 4).mean()
		se_v = np.sqrt((m4g - vg ** 2) / B + (m4o - vo ** 2) / Bo)
		assert abs(vg - vo) <= 3 * se_v, f"variance {i}: {vg} vs {vo} (SE {se_v})"


@pytest.mark.parametrize("case", [0, 1])
def test_golden_trajectories_to_t100(case):
	g = golden_long()["trajectories"][case]
	spec = models.ALL[g["model"]]
	B, T = g["B"], unhex(g["T"])
	y0 = np.random.default_rng(42).random((B, spec.n))
	E = Ensemble(spec, B)
	E.seed(np.arange(B, dtype=np.uint32))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.integrate(T)
	y = E.get_state()
	dt, acc, rej = E.controller_state()
	for k, row in enumerate(g["rows"]):
		assert_bits_equal(y[k], unhex(row["y"]), f"{g['model']} trajectory {k}")
		assert int(acc[k]) == row["accepted"] and int(rej[k]) == row["rejected"]
		assert_bits_equal(dt[k], unhex(row["dt"]), "dt")


def test_config1_example_verbatim_full_grid():
	"""examples/noisy_lorenz.py:89-96: one trajectory, multiplicative noise (SRI), 10 000 integrate() calls — in one launch"""
	g = golden_long()["config1"]
	spec = models.ALL[g["model"]]
	y0 = np.random.default_rng(seed=42).random(3)
	times = np.arange(0.0, 100.0, 0.01)
	assert len(times) == g["n_times"]
	assert_bits_equal(y0, unhex(g["y0"]), "initial state")
	E = Ensemble(spec, 1)
	E.seed(np.array([g["seed"]], dtype=np.uint32))
	E.set_state(y0[None, :], 0.0)
	E.set_integration_parameters(ControllerParameters())
	ys = E.integrate_grid(times)[:, 0, :]
	dt, acc, rej = E.controller_state()
	# against the real reference core: every 250th sample, the last one, the step counts, an xor checksum of all samples
	assert_bits_equal(ys[g["kept"]], unhex(g["ys"]), "kept samples")
	assert int(acc[0]) == g["accepted"] and int(rej[0]) == g["rejected"]
	assert format(int(np.bitwise_xor.reduce(ys.view(np.uint64).ravel())), "016x") == g["checksum_xor"]
	# and against the oracle port: all 10 000 samples
	P = port.PortCore(spec, y0, 0.0, g["seed"])
	ref = np.array([P.integrate(float(T)) for T in times])
	assert_bits_equal(ys, ref, "all samples")


def test_config5_network_at_its_per_gpu_size():
	"""BASELINE.json configs[4]: n = 1000 FitzHugh–Nagumo network, 2^11 trajectories (one GPU's share of 2^14), T = 1"""
	spec, B, T = models.fhn_1000(), 1 << 11, 1.0
	y0 = sharding.global_initial_states(0, B, spec.n)
	seeds = sharding.global_seeds(0, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	stats = E.integrate(T)
	y, t, status = E.get_state(), E.t, E.status
	dt, acc, rej = E.controller_state()
	assert (status == 0).all() and (t == T).all() and np.isfinite(y).all()
	assert int(acc.sum()) == stats["accepted"] and int(rej.sum()) == stats["rejected"]
	idx = np.arange(0, B, B // 8)
	ref = oracle_ensemble(spec, y0[idx], seeds[idx], T, workers=8)
	assert_bits_equal(y[idx], ref["y"], "sampled trajectories")
	assert (acc[idx].astype(np.int64) == ref["accepted"]).all()
	assert (rej[idx].astype(np.int64) == ref["rejected"]).all()
