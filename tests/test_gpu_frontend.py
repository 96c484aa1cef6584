"""GPU: the public API (`jitcsde`, `jitcsde_jump`) end to end — symbolic input → generated CUDA → integrate — against the
oracle built from the same printed expressions; the reference's own API-level tests are the model
(tests/test_jitcsde.py:27-74, :221-240; tests/test_jumps.py:27-59)."""

import numpy as np
import pytest
import sympy

from jitcsde_b200 import UnsuccessfulIntegration, jitcsde, jitcsde_jump, xi, y
from oracle import port
from util import assert_bits_equal, jump_tape

pytestmark = pytest.mark.gpu

f = [10 * (y(1) - y(0)), y(0) * (28 - y(2)) - y(1), y(0) * y(1) - sympy.Rational(8, 3) * y(2)]
g = [0.1 * y(i) for i in range(3)]
y0 = np.array([0.77395605, 0.43887844, 0.85859792])


def test_single_trajectory_like_the_reference_example():
	"""examples/noisy_lorenz.py:89-96 with set_seed(42): shapes and values of a B=1 run."""
	SDE = jitcsde(f, g, verbose=False)
	SDE.set_seed(42)
	SDE.set_initial_value(y0, 0.0)
	data = np.array([SDE.integrate(T) for T in np.arange(0.0, 0.3, 0.01)])
	assert data.shape == (30, 3) and isinstance(SDE.t, float)
	P = port.PortCore(SDE.spec, y0, 0.0, 42)
	ref = np.array([P.integrate(T) for T in np.arange(0.0, 0.3, 0.01)])
	assert_bits_equal(data, ref, "time series")
	assert SDE.t == P.t and SDE.dt == P.dt


def test_integrate_grid_equals_the_loop_of_calls():
	SDE = jitcsde(f, g, verbose=False)
	SDE.set_seed(42)
	SDE.set_initial_value(y0, 0.0)
	times = np.arange(0.0, 0.3, 0.01)
	data = SDE.integrate_grid(times)
	assert data.shape == (30, 3)
	P = port.PortCore(SDE.spec, y0, 0.0, 42)
	assert_bits_equal(data, np.array([P.integrate(T) for T in times]), "time series")
	ens = jitcsde(f, g, ensemble=5, verbose=False)
	ens.set_seed(42)
	ens.set_initial_value(y0, 0.0)
	assert_bits_equal(ens.integrate_grid(times)[:, 0, :], data, "ensemble member 0")


def test_ensemble_seeding_rule_and_reset():
	B = 40
	SDE = jitcsde(f, g, ensemble=B, verbose=False)
	SDE.set_seed(1000)
	SDE.set_initial_value(y0, 0.0)   # one state broadcast to all trajectories
	out = SDE.integrate(0.2)
	assert out.shape == (B, 3)
	for k in (0, 7, 39):
		P = port.PortCore(SDE.spec, y0, 0.0, 1000 + k)
		assert_bits_equal(out[k], P.integrate(0.2), f"trajectory {k}")
	# setting the initial value again restarts noise and generator (reference _jitcsde.py:179-182,242-246) but keeps dt
	dt_before = SDE.dt.copy()
	SDE.set_initial_value(y0, 0.0)
	again = SDE.integrate(0.2)
	P = port.PortCore(SDE.spec, y0, 0.0, 1000)
	P.integrate(0.2)
	P2 = port.PortCore(SDE.spec, y0, 0.0, 1000, first_step=float(dt_before[0]))
	assert_bits_equal(again[0], P2.integrate(0.2), "second run starts from the kept dt")
	assert not np.array_equal(again, out) or dt_before[0] == 1.0


def test_additive_and_helpers_and_parameters():
	k = sympy.Symbol("k")
	m = sympy.Symbol("m")
	SDE = jitcsde([-k * y(0) + m, -0.5 * (y(1) - 1) + m], [0.3, 0.2], helpers=[(m, 0.1 * (y(0) + y(1)))], control_pars=[k], ensemble=8, verbose=False)
	assert SDE.additive
	SDE.set_seed(5)
	SDE.set_initial_value([0.2, 0.4], 0.0)
	SDE.set_parameters(1.5)
	out = SDE.integrate(1.0)
	P = port.PortCore(SDE.spec, [0.2, 0.4], 0.0, 5 + 3)
	P.set_parameters(1.5)
	assert_bits_equal(out[3], P.integrate(1.0), "trajectory 3")


def test_unsuccessful_integration_is_raised():
	"""reference tests/test_jitcsde.py:221-240"""
	for kw in (dict(min_step=10.0), dict(rtol=1e-10, atol=0.0, min_step=1e-3), dict(rtol=0.0, atol=1e-10, min_step=1e-3)):
		SDE = jitcsde(f, g, ensemble=4, verbose=False)
		SDE.set_seed(42)
		SDE.set_initial_value(y0, 0.0)
		import warnings
		with warnings.catch_warnings():
			warnings.simplefilter("ignore")
			SDE.set_integration_parameters(**kw)
		with pytest.raises(UnsuccessfulIntegration) as info:
			SDE.integrate(100.0)
		assert len(info.value.indices) >= 1


def test_pin_noise_through_the_front_end():
	SDE = jitcsde(f, g, verbose=False)
	SDE.set_seed(3)
	SDE.set_initial_value(y0, 0.0)
	SDE.pin_noise(5, 1e-3)
	out = SDE.integrate(0.01)
	P = port.PortCore(SDE.spec, y0, 0.0, 3)
	P.pin_noise(5, 1e-3)
	assert_bits_equal(out, P.integrate(0.01), "pinned run")


def test_jumps_like_the_reference_example():
	"""examples/noisy_and_jumpy_lorenz.py with the callbacks replaced by a tape (amplitude |y2|*xi on component 2)"""
	B, L = 16, 32
	taus, xis = jump_tape(B, L)
	SDE = jitcsde_jump(taus, [0, 0, 0.0 + sympy.Abs(y(2)) * xi], f, g, xis=xis, ensemble=B, verbose=False)
	SDE.set_seed(0)
	SDE.set_initial_value(np.random.default_rng(42).random((B, 3)), 0.0)
	out = SDE.integrate(3.0)
	assert SDE.last_stats["jumps"] > 0
	y0s = np.random.default_rng(42).random((B, 3))
	for k in (0, 5, 15):
		P = port.PortCore(SDE.spec, y0s[k], 0.0, k)
		assert_bits_equal(out[k], P.integrate_jumps(3.0, taus[k], xis[k]), f"trajectory {k}")


def test_callable_waiting_times():
	rng = np.random.default_rng(1)
	SDE = jitcsde_jump(lambda time, state: rng.exponential(0.5), [0.5 * xi], [-y(0)], [0.1], rng=np.random.default_rng(2), ensemble=64, tape_length=64, verbose=False)
	SDE.set_seed(9)
	SDE.set_initial_value([1.0], 0.0)
	out = SDE.integrate(2.0)
	assert out.shape == (64, 1) and np.isfinite(out).all() and SDE.last_stats["jumps"] > 64


def test_generated_jumps_with_a_rate_through_the_public_api():
	"""jitcsde_jump(IJI="generated", waiting_time=tau/rate): no tape; deterministic given the rng that seeds the jump
	generator; the jump count follows the rate."""
	from jitcsde_b200 import tau

	def run(seed):
		SDE = jitcsde_jump("generated", [0.5 * xi], [-y(0)], [0.1], waiting_time=tau / 4.0, rng=np.random.default_rng(seed), ensemble=256, verbose=False)
		SDE.set_seed(7)
		SDE.set_initial_value(np.zeros(1), 0.0)
		out = SDE.integrate(5.0)
		return out, SDE.last_stats["jumps"]
	a, ja = run(3)
	b, jb = run(3)
	c, _ = run(4)
	assert_bits_equal(a, b, "same jump seed")
	assert not np.array_equal(a, c)
	assert ja == jb and abs(ja - 256 * 5.0 * 4.0) < 5 * np.sqrt(256 * 5.0 * 4.0)  # Poisson(rate 4 per unit time)


def test_compile_options_cse_and_simplify_on_the_device():
	"""compile_C(do_cse=True) / (simplify=True) regenerate the model; the device runs exactly the regenerated expressions
	(bit-identical to the oracle's build of the same body)"""
	fs = [sympy.sin(y(0) * y(1)) * 0 + y(1) * y(0) - y(0), -(y(0) * y(1)) - 0.5 * y(1)]
	gs = [0.1 * y(0) * y(1), 0.2 + 0.1 * y(0) * y(1)]
	y0 = np.random.default_rng(4).random((32, 2))
	for kwargs in ({"do_cse": True}, {"simplify": True}):
		SDE = jitcsde(fs, gs, ensemble=32, verbose=False)
		SDE.compile_C(**kwargs)
		SDE.set_seed(0)
		SDE.set_initial_value(y0, 0.0)
		out = SDE.integrate(0.5)
		if "do_cse" in kwargs:
			assert "cse_d0" in SDE.spec.f_body
		ref = port.run_ensemble(SDE.spec, y0, np.arange(32, dtype=np.uint32), 0.0, 0.5)
		assert_bits_equal(out, ref["y"], str(kwargs))
