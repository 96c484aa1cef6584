"""GPU: behaviour of the boundary itself (include/jsde.h) — the round-1 review's hygiene items.

* jump tapes are uploaded explicitly: a second tape of the same shape, or an in-place edit followed by a new upload, is
  what the device uses (round 1 cached uploads by host-pointer identity);
* a tape that runs out is counted (jsde_stats.n_tape_exhausted) and the front end warns;
* `jitcsde_jump.integrate_grid` applies the jumps (round 1 inherited the jump-free grid call);
* generated jumps in sampling-grid mode; a tape call after a generated call uses the tape again;
* per-trajectory control parameters can be set over and over without growing the device footprint;
* ensemble moments of the n = 1000 network (the shared-memory footprint must not grow with n).
"""

import warnings

import numpy as np
import pytest
import sympy

from jitcsde_b200 import jitcsde_jump, models, xi, y
from jitcsde_b200._abi import ControllerParameters, Ensemble, JsdeError
from oracle import port
from util import assert_bits_equal, ensemble_inputs, jump_tape

pytestmark = pytest.mark.gpu


def make(spec, B, **ctrl):
	y0, seeds = ensemble_inputs(spec.n, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**ctrl))
	return E, y0, seeds


def oracle_jumps(spec, y0, seeds, T, taus, xis):
	return np.array([port.PortCore(spec, y0[k], 0.0, int(seeds[k])).integrate_jumps(T, taus[k], xis[k]) for k in range(len(seeds))])


def oracle_jumps_sampled(spec, y0, seeds, times, taus, xis):
	"""the sampling loop `for t in times: integrate(t)` on the oracle: last sample of every trajectory"""
	out = []
	for k in range(len(seeds)):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		for tt in times:
			last = P.integrate_jumps(float(tt), taus[k], xis[k])
		out.append(last)
	return np.array(out)


def test_a_second_tape_of_equal_shape_replaces_the_first():
	spec, B, L, T = models.lorenz_jumpy, 32, 12, 1.5
	E, y0, seeds = make(spec, B)
	taus1, xis1 = jump_tape(B, L)
	E.integrate_jumps(T, taus1, xis1)
	assert_bits_equal(E.get_state(), oracle_jumps(spec, y0, seeds, T, taus1, xis1), "first tape")
	# same shapes, other values, and — what round 1 got wrong — possibly the same host addresses
	for attempt in range(3):
		taus2, xis2 = taus1 * (0.5 + 0.1 * attempt), -xis1
		E.set_state(y0, 0.0)
		E.set_integration_parameters(ControllerParameters())  # dt survives set_state (it lives on the Python object in the reference)
		E.integrate_jumps(T, taus2, xis2)
		assert_bits_equal(E.get_state(), oracle_jumps(spec, y0, seeds, T, taus2, xis2), f"tape {attempt + 2}")
		del taus2, xis2
	# in-place edit + explicit upload
	taus1 *= 2.0
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.set_jump_tape(taus1, xis1)
	E.integrate_jumps(T)  # continues on the tape set before
	assert_bits_equal(E.get_state(), oracle_jumps(spec, y0, seeds, T, taus1, xis1), "edited tape")


def test_integrate_jumps_without_a_tape_is_refused():
	E, _, _ = make(models.lorenz_jumpy, 4)
	with pytest.raises(JsdeError, match="no jump tape"):
		E.integrate_jumps(1.0)


def test_tape_exhaustion_is_counted():
	spec, B, T = models.lorenz_jumpy, 64, 3.0
	E, y0, seeds = make(spec, B)
	taus = np.full((B, 2), 0.4)
	taus[:8] = 10.0  # these eight never reach their second jump before T
	xis = np.full((B, 2), 0.1)
	stats = E.integrate_jumps(T, taus, xis)
	assert stats["jumps"] == 2 * (B - 8)
	assert stats["n_tape_exhausted"] == B - 8
	assert_bits_equal(E.get_state(), oracle_jumps(spec, y0, seeds, T, taus, xis), "jumps stop at the tape's end like the oracle's tape closures")


F = [10 * (y(1) - y(0)), y(0) * (28 - y(2)) - y(1), y(0) * y(1) - sympy.Rational(8, 3) * y(2)]
Gm = [0.1 * y(i) for i in range(3)]
AMP = [0, 0, 0.0 + sympy.Abs(y(2)) * xi]


def test_front_end_warns_when_the_tape_runs_out():
	B = 8
	SDE = jitcsde_jump(np.full((B, 3), 0.2), AMP, F, Gm, xis=np.full((B, 3), 0.05), ensemble=B, verbose=False)
	SDE.set_seed(1)
	SDE.set_initial_value(np.random.default_rng(42).random((B, 3)), 0.0)
	with pytest.warns(RuntimeWarning, match="ran out"):
		SDE.integrate(2.0)
	assert SDE.tape_exhausted == B and SDE.last_stats["jumps"] == 3 * B
	with warnings.catch_warnings():
		warnings.simplefilter("error")
		SDE.integrate(2.5)  # already reported; nothing new runs out
	assert SDE.tape_exhausted == B


def test_front_end_grid_applies_jumps():
	B, L = 16, 24
	taus, xis = jump_tape(B, L)
	times = np.linspace(0.0, 2.0, 9)
	y0 = np.random.default_rng(42).random((B, 3))

	def fresh():
		SDE = jitcsde_jump(taus, AMP, F, Gm, xis=xis, ensemble=B, verbose=False)
		SDE.set_seed(0)
		SDE.set_initial_value(y0, 0.0)
		return SDE
	A, Bq = fresh(), fresh()
	grid = A.integrate_grid(times)
	loop = np.array([Bq.integrate(T) for T in times])
	assert A.last_stats["jumps"] > 0
	assert_bits_equal(grid, loop, "grid vs loop of integrate() calls")
	assert_bits_equal(grid[-1], oracle_jumps_sampled(A.spec, y0, np.arange(B), times, taus, xis), "last sample vs the oracle's sampling loop")


def test_generated_jumps_in_grid_mode_and_tape_after_generated():
	spec, B, T = models.lorenz_jumpy, 48, 1.5
	times = np.linspace(0.0, T, 7)
	E, y0, seeds = make(spec, B)
	grid = E.integrate_grid(times, jumps="generated", seed=77, first_index=5)
	E2, _, _ = make(spec, B)
	loop = []
	for tt in times:
		E2.integrate_jumps_generated(tt, 77, 5)
		loop.append(E2.get_state())
	assert_bits_equal(grid, np.array(loop), "generated jumps: grid vs loop")
	assert E.last_stats["jumps"] > 0
	# the same handle now with a tape: must not keep drawing generated variates
	taus, xis = jump_tape(B, 16)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	g2 = E.integrate_grid(times, taus, xis)
	assert_bits_equal(g2[-1], oracle_jumps_sampled(spec, y0, seeds, times, taus, xis), "tape after generated")


def test_per_trajectory_parameters_reuse_their_buffer():
	import torch
	spec, B, T = models.gbm_pars, 4096, 0.5
	E, y0, seeds = make(spec, B)
	rng = np.random.default_rng(3)
	free0 = None
	for it in range(40):
		pars = np.column_stack([rng.uniform(0.0, 0.1, B), rng.uniform(0.1, 0.3, B)])
		E.set_parameters(pars)
		if it == 2:
			free0 = torch.cuda.mem_get_info()[0]
	assert torch.cuda.mem_get_info()[0] >= free0 - (1 << 20), "device memory shrinks with every set_parameters call"
	E.integrate(T)
	yy = E.get_state()
	for k in (0, 17, B - 1):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.set_parameters(*pars[k])
		assert_bits_equal(yy[k], P.integrate(T), f"trajectory {k}")


def test_moments_of_the_n1000_network():
	import torch
	spec, B = models.fhn_1000(), 96
	y0 = np.random.default_rng(1).random((B, spec.n))
	E = Ensemble(spec, B)
	E.seed(np.arange(B, dtype=np.uint32))
	E.set_state(y0, 0.0)
	shift = np.full(spec.n, 0.5)
	acc = torch.zeros(1 + 2 * spec.n, dtype=torch.float64, device="cuda")
	E.moments_into(acc.data_ptr(), shift)
	torch.cuda.synchronize()
	got = acc.cpu().numpy()
	assert got[0] == B
	np.testing.assert_allclose(got[1:1 + spec.n], (y0 - shift).sum(axis=0), rtol=0, atol=1e-11)
	np.testing.assert_allclose(got[1 + spec.n:], ((y0 - shift) ** 2).sum(axis=0), rtol=1e-12)
