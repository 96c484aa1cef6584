"""GPU: the statistical rung of the parity ladder (BASELINE.json: ensemble mean and variance within 3 standard errors),
the analogue of the reference's Kramers–Moyal validation (tests/validation.py), and the full-size run of the headline
configuration checked through size-independent properties."""

import numpy as np
import pytest

from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
from oracle import port
from util import assert_bits_equal

pytestmark = pytest.mark.gpu


def run(spec, B, T, y0=None, **ctrl):
	y0 = sharding.global_initial_states(0, B, spec.n) if y0 is None else y0
	E = Ensemble(spec, B)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**ctrl))
	stats = E.integrate(T)
	return E, stats, y0


def within(value, expected, se, what, k=3.0):
	assert abs(value - expected) <= k * se, f"{what}: {value} vs {expected} (|diff| = {abs(value - expected):.3g} > {k} SE = {k * se:.3g})"


def test_gbm_moments_match_closed_form():
	"""dy = mu y dt + sigma y dW, y0 = 1:  E = exp(mu T), Var = exp(2 mu T)(exp(sigma^2 T) - 1)   (SURVEY §8 d config 3 i)"""
	B, T, mu, sig = 1 << 18, 2.0, 0.05, 0.2
	E, stats, _ = run(models.gbm, B, T, y0=np.ones((B, 1)))
	y = E.get_state()[:, 0]
	mean, var = np.exp(mu * T), np.exp(2 * mu * T) * (np.exp(sig ** 2 * T) - 1)
	within(y.mean(), mean, np.sqrt(var / B), "GBM mean")
	m4 = ((y - y.mean()) ** 4).mean()
	within(y.var(ddof=1), var, np.sqrt((m4 - var ** 2) / B), "GBM variance", k=4.0)
	# log-scale: ln y ~ N((mu - sigma^2/2) T, sigma^2 T) exactly
	ly = np.log(y)
	within(ly.mean(), (mu - sig ** 2 / 2) * T, np.sqrt(sig ** 2 * T / B), "GBM log-mean")
	within(ly.var(ddof=1), sig ** 2 * T, sig ** 2 * T * np.sqrt(2 / B), "GBM log-variance", k=4.0)


def test_ou_moments_match_closed_form():
	"""dy = -theta (y - m) dt + s dW: mean m + (y0-m) e^{-theta T}, var s^2/(2 theta) (1 - e^{-2 theta T})"""
	B, T = 1 << 18, 3.0
	y0 = np.tile([0.5, 0.5], (B, 1))
	E, stats, _ = run(models.ou, B, T, y0=y0)
	y = E.get_state()
	for i, (theta, m, s) in enumerate([(1.0, 0.0, 0.3), (0.5, 1.0, 0.2)]):
		mean = m + (0.5 - m) * np.exp(-theta * T)
		var = s ** 2 / (2 * theta) * (1 - np.exp(-2 * theta * T))
		within(y[:, i].mean(), mean, np.sqrt(var / B), f"OU mean {i}")
		within(y[:, i].var(ddof=1), var, var * np.sqrt(2 / B), f"OU variance {i}", k=4.0)


def test_lorenz_ensemble_moments_vs_oracle_ensemble():
	"""Independent seeds on both sides (GPU: seeds k, oracle: seeds 10^6+k): moments must agree within 3 combined SE."""
	spec, T, Bg, Bo = models.lorenz_additive, 2.0, 1 << 16, 2048
	E, stats, y0 = run(spec, Bg, T)
	yg = E.get_state()
	yo = port.run_ensemble(spec, sharding.global_initial_states(0, Bo, 3), np.arange(Bo, dtype=np.uint32) + 1_000_000, 0.0, T)["y"]
	for i in range(3):
		se = np.sqrt(yg[:, i].var() / Bg + yo[:, i].var() / Bo)
		within(yg[:, i].mean(), yo[:, i].mean(), se, f"Lorenz mean {i}")
		vg, vo = yg[:, i].var(ddof=1), yo[:, i].var(ddof=1)
		m4g = ((yg[:, i] - yg[:, i].mean()) ** 4).mean()
		m4o = ((yo[:, i] - yo[:, i].mean()) ** 4).mean()
		se_v = np.sqrt((m4g - vg ** 2) / Bg + (m4o - vo ** 2) / Bo)
		within(vg, vo, se_v, f"Lorenz variance {i}", k=4.0)


def test_headline_configuration_full_size_properties():
	"""2^20-trajectory noisy-Lorenz ensemble (BASELINE.json configs[1]) on a shortened horizon: every trajectory reaches
	the target, totals add up, a strided sample is bit-identical to the oracle, and a rerun reproduces every bit."""
	spec, B, T = models.lorenz_additive, 1 << 20, 2.0
	E, stats, y0 = run(spec, B, T)
	y, t, status = E.get_state(), E.t, E.status
	dt, acc, rej = E.controller_state()
	assert (status == 0).all() and (t == T).all() and np.isfinite(y).all()
	assert int(acc.sum()) == stats["accepted"] and int(rej.sum()) == stats["rejected"]
	assert stats["n_overflow"] == 0 and stats["n_min_step"] == 0
	idx = np.arange(0, B, B // 128)
	ref = port.run_ensemble(spec, y0[idx], sharding.global_seeds(0, B)[idx], 0.0, T)
	assert_bits_equal(y[idx], ref["y"], "sampled trajectories")
	assert (acc[idx].astype(np.int64) == ref["accepted"]).all()
	# determinism + independence of the ensemble size: the first 2^12 trajectories alone give the same bits
	E2, _, _ = run(spec, 1 << 12, T)
	assert_bits_equal(E2.get_state(), y[: 1 << 12], "independence of ensemble size")
	# checksum of the whole ensemble state is reproducible
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	E.integrate(T)
	assert_bits_equal(E.get_state(), y, "rerun")


def test_multiplicative_full_size_sample():
	spec, B, T = models.lorenz_multiplicative, 1 << 18, 1.0
	E, stats, y0 = run(spec, B, T)
	y = E.get_state()
	idx = np.arange(0, B, B // 64)
	ref = port.run_ensemble(spec, y0[idx], sharding.global_seeds(0, B)[idx], 0.0, T)
	assert_bits_equal(y[idx], ref["y"], "sampled trajectories")
	assert (E.status == 0).all()


def test_config3_duffing_full_size_sample():
	"""BASELINE.json configs[2]: stochastic Duffing ensemble, multiplicative noise (SRI), 2^20 trajectories"""
	spec, B, T = models.duffing, 1 << 20, 1.0
	E, stats, y0 = run(spec, B, T)
	y = E.get_state()
	idx = np.arange(0, B, B // 64)
	ref = port.run_ensemble(spec, y0[idx], sharding.global_seeds(0, B)[idx], 0.0, T)
	assert_bits_equal(y[idx], ref["y"], "sampled trajectories")
	assert (E.status == 0).all() and (E.t == T).all()


def test_config4_jumpy_lorenz_full_size_sample():
	"""BASELINE.json configs[3]: examples/noisy_and_jumpy_lorenz.py physics, 2^18 trajectories, per-trajectory jump tape"""
	from util import jump_tape
	spec, B, T, L = models.lorenz_jumpy, 1 << 18, 2.0, 24
	rng = np.random.default_rng(5)
	taus = rng.exponential(1.0, (B, L))
	xis = rng.standard_normal((B, L))
	y0 = sharding.global_initial_states(0, B, 3)
	E = Ensemble(spec, B)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	stats = E.integrate_jumps(T, taus, xis)
	y = E.get_state()
	assert (E.status == 0).all() and (E.t == T).all() and np.isfinite(y).all()
	expected_jumps = int(((np.cumsum(taus, axis=1) < T).sum(axis=1)).sum())
	assert stats["jumps"] == expected_jumps
	for k in range(0, B, B // 32):
		P = port.PortCore(spec, y0[k], 0.0, int(k))
		assert_bits_equal(y[k], P.integrate_jumps(T, taus[k], xis[k]), f"trajectory {k}")
