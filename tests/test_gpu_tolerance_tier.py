"""GPU: the TOLERANCE tier of the parity ladder (BASELINE.json: per-trajectory states within 1e-9 relative wherever the
accepted-step sequence matches; SURVEY.md §8 c rungs 2 and 4) and the reference's own statistical validation
(tests/validation.py, tests/kmc.py: Kramers–Moyal coefficients of the scalar SDE with G = 5 exp(-y^2+y-1.5) + 3).

Which libm calls in generated code are bit-exact on the device and which are tolerance-only:
  exp ........ glibc's algorithm restated (csrc/glibc_math.h: jsde_exp), bit-exact for 2^-54 <= |x| < 512
  log, pow ... (the integrator's own uses: polar method, controller) bit-exact
  + - * / sqrt fabs ... IEEE-exact on both sides
  sin, cos, tan, tanh, ... the CUDA math library's (<= 1-2 ulp from glibc's): tolerance tier, tested here.
"""

import numpy as np
import pytest
import sympy

from jitcsde_b200 import models
from jitcsde_b200._abi import ControllerParameters, Ensemble
from oracle import port
from util import ensemble_inputs

pytestmark = pytest.mark.gpu


def test_single_step_with_cuda_libm_calls_within_1e13():
	"""rung 2: one get_next_step from identical (state, t, h, noise): <= 1e-13 relative"""
	spec, B = models.pendulum, 64
	rng = np.random.default_rng(8)
	y0 = rng.uniform(-3, 3, (B, 2))
	h = 2e-3
	dw = rng.standard_normal((B, 1, 2)) * np.sqrt(h)
	dz = rng.standard_normal((B, 1, 2)) * np.sqrt(h)
	E = Ensemble(spec, B)
	E.seed(np.arange(B, dtype=np.uint32))
	E.set_state(y0, 0.3)
	E.pin_noise_path([h], dw, dz)
	E.get_next_step(h)
	p = E.get_p(1e-10, 1e-5)
	E.accept_step()
	y = E.get_state()
	for k in range(B):
		P = port.PortCore(spec, y0[k], 0.3, k)
		P.pin_path([h], dw[k], dz[k])
		P.get_next_step(h)
		pk = P.get_p(1e-10, 1e-5)
		P.accept_step()
		np.testing.assert_allclose(y[k], P.get_state(), rtol=1e-13, atol=0)
		np.testing.assert_allclose(p[k], pk, rtol=1e-7)  # a ratio of small differences: a last-bit change of sin/cos shows up here


def test_trajectories_with_cuda_libm_calls_within_1e9_where_the_step_sequence_matches():
	"""rung 4.  "Where the accepted-step sequence matches" means the same step SIZES: with the adaptive controller a
	last-bit difference in sin/cos changes the error estimate, hence the next step size by a relative 1e-9 or so, hence every
	later Wiener increment (DW = x*sqrt(h)) — the two sides then integrate different discretisations of the path.  A
	controller that never rejects and never changes the step pins the sequence; 500 steps then agree to 1e-9."""
	spec, B, T = models.pendulum, 256, 1.0
	pinned = dict(first_step=2e-3, max_step=2e-3, atol=1e-2, rtol=1e-2)
	y0, seeds = ensemble_inputs(spec.n, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**pinned))
	E.integrate(T)
	y = E.get_state()
	_, acc, rej = E.controller_state()
	ref = port.run_ensemble(spec, y0, seeds, 0.0, T, **pinned)
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej == 0).all() and (ref["rejected"] == 0).all()
	np.testing.assert_allclose(y, ref["y"], rtol=1e-9, atol=1e-12)


def test_adaptive_trajectories_with_cuda_libm_calls_agree_statistically():
	"""rung 5 for the same model with the default controller: ensemble moments within 3 standard errors of the oracle's"""
	spec, B, T = models.pendulum, 4096, 1.0
	y0, seeds = ensemble_inputs(spec.n, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.integrate(T)
	y = E.get_state()
	Bo = 1024
	ref = port.run_ensemble(spec, y0[:Bo], seeds[:Bo] + np.uint32(500_000), 0.0, T)["y"]
	for i in range(2):
		se = np.sqrt(y[:, i].var() / B + ref[:, i].var() / Bo)
		assert abs(y[:, i].mean() - ref[:, i].mean()) <= 3 * se
	assert np.isfinite(y).all() and (E.t == T).all()


def test_exp_model_is_bit_exact_at_a_long_horizon():
	from util import assert_bits_equal
	spec, B, T = models.validation_exp, 64, 5.0
	y0 = np.zeros((B, 1))
	seeds = np.arange(B, dtype=np.uint32)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.integrate(T)
	ref = port.run_ensemble(spec, y0, seeds, 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], "validation scenario with exp in g")
	_, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()


def kramers_moyal(x0, x1, h, nbins=60, kmax=2):
	"""Conditional moments of the increments per bin of the starting point, with standard errors — the estimator of the
	reference's tests/kmc.py:9-84 (bins over the sample's range, mean(increment^k)/h and its standard error), vectorised."""
	lo, hi = x0.min(), x0.max()
	pad = (hi - lo) / nbins * 0.5
	edges = np.linspace(lo - pad, hi + pad, nbins + 1)
	which = np.clip(((x0 - edges[0]) / (edges[-1] - edges[0]) * nbins).astype(int), 0, nbins - 1)
	centres = (edges[1:] + edges[:-1]) / 2
	inc = x1 - x0
	out = []
	count = np.bincount(which, minlength=nbins).astype(float)
	for k in range(1, kmax + 1):
		v = inc ** k
		s1 = np.bincount(which, weights=v, minlength=nbins)
		s2 = np.bincount(which, weights=v * v, minlength=nbins)
		with np.errstate(invalid="ignore", divide="ignore"):
			mean = s1 / count
			var = (s2 - count * mean ** 2) / (count - 1)
			sem = np.sqrt(var / count)
		out.append((mean / h, sem / h))
	return centres, count, out


@pytest.mark.parametrize("name", ["validation_exp", "cubic"])
def test_kramers_moyal_coefficients_match_theory(name):
	"""tests/validation.py:99-126 on an ensemble: the first two Kramers–Moyal coefficients estimated from increments over
	dt = 1e-3 agree with F + O(dt) and G^2 + O(dt) (theory incl. the first-order corrections of validation.py:109-110)
	in at least 60 % of the populated bins (the reference's thresholds, validation.py:18)."""
	spec = models.ALL[name]
	B, dt, nsamp = 1 << 15, 1e-3, 17
	E = Ensemble(spec, B)
	E.seed(np.arange(B, dtype=np.uint32) + 12345)
	E.set_state(np.zeros((B, 1)), 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.integrate(1.0)  # let the ensemble spread over the stationary density first
	xs = E.integrate_grid(1.0 + dt * np.arange(nsamp))[:, :, 0]  # [nsamp][B]
	x0, x1 = xs[:-1].ravel(), xs[1:].ravel()
	Y = sympy.Symbol("Y")
	F = -Y ** 3 + 4 * Y + Y ** 2
	G = 5 * sympy.exp(-Y ** 2 + Y - sympy.Rational(3, 2)) + 3 if name == "validation_exp" else sympy.Integer(3)
	Fd, Gd = F.diff(Y), G.diff(Y)
	Fdd, Gdd = Fd.diff(Y), Gd.diff(Y)
	M = [sympy.lambdify(Y, F + (F * Fd + G ** 2 * Fdd / 2) * dt / 2, "numpy"),
		sympy.lambdify(Y, G ** 2 + (2 * F * (F + G * Gd) + G ** 2 * (2 * Fd + Gd ** 2 + G * Gdd)) * dt / 2, "numpy")]
	centres, count, kmc = kramers_moyal(x0, x1, dt)
	populated = count >= 200
	assert populated.sum() >= 30
	for k, threshold in enumerate((0.6, 0.6)):
		value, error = kmc[k]
		theory = M[k](centres) * np.ones_like(centres)
		good = (np.abs(theory - value) < error) & populated
		assert good.sum() >= threshold * populated.sum(), f"KM coefficient {k + 1}: theory inside the error bar in {good.sum()} of {populated.sum()} bins"
