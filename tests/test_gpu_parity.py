"""
GPU parity tests: the CUDA path, called through the C ABI, against (a) the committed fixtures generated from the
unmodified reference C core and (b) the C restatement (oracle/) on the same seeded inputs.  Integer work and — because
the device reproduces glibc's log/pow and never contracts — all FP64 results are required to be BIT-EXACT, which is
stricter than the 1e-9 relative tolerance BASELINE.json asks for; ensemble moments are additionally checked at 3
standard errors in test_gpu_statistics.py.
"""

import ctypes
import os

import numpy as np
import pytest

from jitcsde_b200 import _abi, models
from jitcsde_b200._abi import ControllerParameters, Ensemble, JsdeError
from oracle import port
from util import assert_bits_equal, ensemble_inputs, golden, jump_tape, unhex

pytestmark = pytest.mark.gpu
G = golden()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make(spec, B, **ctrl):
	y0, seeds = ensemble_inputs(spec.n, B)
	E = Ensemble(spec, B)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters(**ctrl))
	return E, y0, seeds


# ---- libm restatements on the device --------------------------------------------------------------------------------
def test_device_log_pow_equal_host_libm():
	import subprocess
	so = os.path.join(ROOT, "oracle", "_build", "libm_check.so")
	os.makedirs(os.path.dirname(so), exist_ok=True)
	subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", os.path.join(ROOT, "oracle", "libm_check.c"), "-o", so, "-lm"], check=True)
	host = ctypes.CDLL(so)
	DP = ctypes.POINTER(ctypes.c_double)
	lib = _abi.load_library(models.lorenz_additive)
	rng = np.random.default_rng(3)
	N = 1 << 20
	for x in (rng.random(N), 1 - rng.random(N) * 0.07, 1 + rng.random(N) * 0.07, np.exp(-rng.random(N) * 72)):
		x = np.ascontiguousarray(x[x > 0])
		ref = np.empty_like(x)
		host.libm_log_v(x.ctypes.data_as(DP), ref.ctypes.data_as(DP), ctypes.c_size_t(x.size))
		assert_bits_equal(_abi.device_log(lib, x), ref, "device log vs libm")
	for yexp in (-1 / 1.5, -1 / 2.5):
		for x in (np.exp(rng.uniform(np.log(1.1), np.log(1e30), N)), np.exp(rng.uniform(np.log(1e-300), np.log(0.5), N)), rng.random(N) * 2.2e-308 + 5e-324):
			x = np.ascontiguousarray(x)
			ref = np.empty_like(x)
			host.libm_pow_v(x.ctypes.data_as(DP), ctypes.c_double(yexp), ref.ctypes.data_as(DP), ctypes.c_size_t(x.size))
			assert_bits_equal(_abi.device_pow(lib, x, yexp), ref, "device pow vs libm")


def test_shared_divisor_division_is_exact():
	"""exact_div.cuh: five fused operations + a pre-rounded reciprocal must give the IEEE quotient bit for bit whenever
	the path calls itself tame, and must decline (not guess) outside its proven range."""
	lib = _abi.load_library(models.lorenz_additive)
	rng = np.random.default_rng(11)
	N = 1 << 24
	def mant(n):
		return 1.0 + rng.integers(0, 1 << 52, n, dtype=np.uint64).astype(np.float64) / 2.0 ** 52
	cases = []
	for _ in range(4):  # random mantissas, exponents well inside and around the tame range
		cases.append((mant(N) * 2.0 ** rng.integers(-390, 390, N) * rng.choice([-1.0, 1.0], N), mant(N) * 2.0 ** rng.integers(-390, 390, N)))
	# divisors the stages really use: step sizes, their square roots, 6; numerators of every magnitude incl. zeros
	h = 10.0 ** rng.uniform(-10, 1, N)
	cases.append((rng.standard_normal(N) * 10.0 ** rng.uniform(-12, 3, N), h))
	cases.append((rng.standard_normal(N), np.sqrt(h)))
	cases.append((rng.standard_normal(N) * 10.0 ** rng.uniform(-300, 300, N), np.full(N, 6.0)))
	z = rng.standard_normal(N); z[::3] = 0.0; z[1::3] = -0.0
	cases.append((z, h))
	# hard patterns: divisor mantissa all ones / nearly one, quotients near rounding boundaries (a = q*b rounded)
	b = np.nextafter(2.0, 0.0) * 2.0 ** rng.integers(-50, 50, N)
	cases.append((mant(N), b))
	q = mant(N); b2 = mant(N)
	cases.append((q * b2, b2))
	cases.append((np.nextafter(q * b2, np.inf), b2))
	# outside the proven range: must be flagged, whatever the value
	cases.append((mant(N) * 2.0 ** rng.integers(-1000, -500, N), mant(N)))
	for a, b in cases:
		got, tame = _abi.device_shared_div(lib, a, b)
		with np.errstate(all="ignore"):
			want = a / b
		assert_bits_equal(got[tame], want[tame], "shared-divisor quotient")
	assert not tame.any()  # the last case lies outside the tame range entirely
	got, tame = _abi.device_shared_div(lib, np.array([1.0, np.inf, np.nan, 1e300, 0.0]), np.array([1e-320, 2.0, 2.0, 1e-300, 3.0]))
	assert list(tame) == [False, False, False, False, True] and got[4] == 0.0


def test_inline_division_and_sqrt_are_exact():
	"""exact_div.cuh: the straight-line restatements of nvcc's FP64 division and square root give the operators' bits
	wherever they vouch for the result, vouch for everything the integrator feeds them, and decline at the ends of
	the exponent range."""
	lib = _abi.load_library(models.lorenz_additive)
	rng = np.random.default_rng(12)
	N = 1 << 24
	def mant(n):
		return 1.0 + rng.integers(0, 1 << 52, n, dtype=np.uint64).astype(np.float64) / 2.0 ** 52
	cases = []
	for _ in range(4):
		cases.append((mant(N) * 2.0 ** rng.integers(-900, 900, N) * rng.choice([-1.0, 1.0], N), mant(N) * 2.0 ** rng.integers(-900, 900, N) * rng.choice([-1.0, 1.0], N)))
	# what the Gaussian factor really divides: -2 log(r2) / r2 for radii in (0, 1), and error/tolerance ratios
	r2 = rng.random(N)
	cases.append((-2.0 * np.log(r2), r2))
	r2 = 2.0 ** -rng.uniform(0, 104, N)
	cases.append((-2.0 * np.log(r2), r2))
	cases.append((np.abs(rng.standard_normal(N)) * 10.0 ** rng.uniform(-30, 3, N), 1e-10 + 1e-5 * np.abs(rng.standard_normal(N)) * 10.0 ** rng.uniform(-3, 3, N)))
	# hard patterns: divisor mantissa all ones, quotients at rounding boundaries, perfect squares and their neighbours
	cases.append((mant(N), np.nextafter(2.0, 0.0) * 2.0 ** rng.integers(-50, 50, N)))
	q = mant(N); b2 = mant(N)
	cases.append((q * b2, b2))
	cases.append((np.nextafter(q * b2, np.inf), b2))
	s = mant(N) * 2.0 ** rng.integers(-200, 200, N)
	cases.append((s * s, mant(N)))
	cases.append((np.nextafter(s * s, 0.0), mant(N)))
	vouched_div = vouched_sqrt = 0
	for i, (a, b) in enumerate(cases):
		div, okd, root, oks = _abi.device_inline_math(lib, a, b)
		with np.errstate(all="ignore"):
			assert_bits_equal(div[okd], (a / b)[okd], "inline quotient")
			assert_bits_equal(root[oks], np.sqrt(a)[oks], "inline square root")
		if 4 <= i <= 6:  # the integrator's own operand ranges: no fallbacks
			assert okd.all()
		if 4 <= i <= 5 or i >= 10:
			assert oks.all()
		vouched_div += okd.sum(); vouched_sqrt += oks.sum()
	assert vouched_div > 8 * N and vouched_sqrt > 5 * N
	# specials are declined, never guessed
	a = np.array([0.0, 1.0, np.inf, np.nan, 1.0, 1.0, 1e-320, -1.0, 1e308])
	b = np.array([3.0, 0.0, 2.0, 2.0, np.inf, np.nan, 1.0, 1.0, 1e-308])
	div, okd, root, oks = _abi.device_inline_math(lib, a, b)
	assert list(okd) == [False, False, False, False, False, False, False, True, False] and div[7] == -1.0
	assert not oks[[0, 2, 3, 6, 7]].any() and oks[[1, 4, 5, 8]].all()  # zero, inf, NaN, subnormal, negative are declined
	assert_bits_equal(root[oks], np.sqrt(a[oks]), "inline square root")


# ---- Gaussian stream --------------------------------------------------------------------------------------------------
def test_gauss_stream_bit_exact_many_seeds():
	"""rk_gauss pairs per trajectory == oracle, across several regenerations of the 624-word block."""
	spec = models.ou
	seeds = np.array([0, 1, 42, 12345, 4294967295, 7, 99, 2**31, 2**31 - 1, 1000003] + list(range(100, 154)), dtype=np.uint64)
	B = len(seeds)
	E = Ensemble(spec, B)
	E.seed(seeds)
	pairs = 2000  # ≈ 10 200 words ≈ 16 blocks
	for scale in (1.0, 0.1):
		E.seed(seeds)
		got = E.draw_gauss(scale, pairs)
		for b, s in enumerate(seeds):
			P = port.PortCore(spec, [0, 0], 0.0, int(s))
			assert_bits_equal(got[b], P.draw_gauss(scale, pairs), f"seed {s} scale {scale}")


def test_gauss_stream_golden():
	spec = models.ou
	seeds = np.array([int(s) for s in G["rng"]], dtype=np.uint64)
	E = Ensemble(spec, len(seeds))
	E.seed(seeds)
	got = E.draw_gauss(1.0, 2000)
	for b, s in enumerate(seeds):
		rec = G["rng"][str(int(s))]
		assert_bits_equal(got[b, :8, 0], unhex(rec["dw_scale1_first8"]), f"seed {s}")
		assert_bits_equal(got[b, 1992:, 0], unhex(rec["dw_scale1_1992_2000"]), f"seed {s}")


def test_gauss_stream_continues_across_calls():
	spec = models.ou
	E = Ensemble(spec, 4)
	E.seed([5, 6, 7, 8])
	a = np.concatenate([E.draw_gauss(1.0, 333), E.draw_gauss(1.0, 1), E.draw_gauss(1.0, 700)], axis=1)
	E.seed([5, 6, 7, 8])
	assert_bits_equal(a, E.draw_gauss(1.0, 1034), "split draws")


# ---- single-step surface vs the reference core's golden scripts ------------------------------------------------------
@pytest.mark.parametrize("case", range(len(G["scripted"])))
def test_scripted_golden(case):
	rec = G["scripted"][case]
	spec = models.ALL[rec["model"]]
	B = 3  # same trajectory three times: lanes must not interfere
	E = Ensemble(spec, B)
	E.seed([rec["seed"]] * B)
	E.set_state(np.tile(unhex(rec["y0"]), (B, 1)), 0.0)
	for step in rec["script"]:
		if step["do"] == "step":
			E.get_next_step(unhex(step["h"]))
			assert_bits_equal(E.get_p(1e-10, 1e-5), np.full(B, unhex(step["p"])), "p")
		elif step["do"] == "accept":
			E.accept_step()
			assert_bits_equal(E.get_state(), np.tile(unhex(step["y"]), (B, 1)), "state")
			assert_bits_equal(E.t, np.full(B, unhex(step["t"])), "t")
		else:
			E.pin_noise(step["number"], unhex(step["step"]))


# ---- fused integrate vs golden trajectories ----------------------------------------------------------------------------
@pytest.mark.parametrize("case", range(len(G["trajectories"])))
def test_trajectory_golden(case):
	rec = G["trajectories"][case]
	spec = models.ALL[rec["model"]]
	ctrl = {k: unhex(v) for k, v in rec["ctrl"].items()}
	E, y0, seeds = make(spec, rec["B"], **ctrl)
	stats = E.integrate(unhex(rec["T"]))
	y, t, status = E.get_state(), E.t, E.status
	dt, acc, rej = E.controller_state()
	for k, row in enumerate(rec["rows"]):
		assert status[k] == row["status"]
		assert (int(acc[k]), int(rej[k])) == (row["accepted"], row["rejected"]), f"trajectory {k} step counts"
		assert_bits_equal(y[k], unhex(row["y"]), f"trajectory {k} state")
		assert t[k] == unhex(row["t"])
		assert dt[k] == unhex(row["dt"])
	assert stats["accepted"] == sum(r["accepted"] for r in rec["rows"])
	assert stats["rejected"] == sum(r["rejected"] for r in rec["rows"])
	assert stats["n_min_step"] == sum(r["status"] for r in rec["rows"])


@pytest.mark.parametrize("case", range(len(G["sampled"])))
def test_sampled_golden(case):
	"""10^2 consecutive integrate() calls on a grid (examples/noisy_lorenz.py:95-96): truncated last steps + bridges."""
	rec = G["sampled"][case]
	spec = models.ALL[rec["model"]]
	E, y0, seeds = make(spec, rec["B"])
	times = unhex(rec["times"])
	ys = np.array([(E.integrate(T), E.get_state())[1] for T in times])  # [time][B][n]
	for k, row in enumerate(rec["rows"]):
		assert_bits_equal(ys[:, k, :], unhex(row["ys"]), f"trajectory {k}")
	_, acc, rej = E.controller_state()
	assert [int(a) for a in acc] == [r["accepted"] for r in rec["rows"]]
	assert [int(a) for a in rej] == [r["rejected"] for r in rec["rows"]]


@pytest.mark.parametrize("case", range(len(G["sampled"])))
def test_sampling_grid_in_one_launch_golden(case):
	"""jsde_integrate_grid == the same sequence of integrate() calls (golden from the reference core), one launch."""
	rec = G["sampled"][case]
	spec = models.ALL[rec["model"]]
	E, y0, seeds = make(spec, rec["B"])
	times = unhex(rec["times"])
	ys = E.integrate_grid(times)  # [time][B][n]
	for k, row in enumerate(rec["rows"]):
		assert_bits_equal(ys[:, k, :], unhex(row["ys"]), f"trajectory {k}")
	_, acc, rej = E.controller_state()
	assert [int(a) for a in acc] == [r["accepted"] for r in rec["rows"]]
	assert E.last_stats["kernel_launches"] == 1
	assert_bits_equal(E.get_state(), ys[-1], "final state")


def test_sampling_grid_with_jumps_equals_repeated_calls():
	spec = models.lorenz_jumpy
	B = 64
	taus, xis = jump_tape(B, 16)
	times = np.linspace(0.0, 2.0, 21)
	E, y0, seeds = make(spec, B)
	grid = E.integrate_grid(times, taus, xis)
	E2, _, _ = make(spec, B)
	loop = np.array([(E2.integrate_jumps(T, taus, xis), E2.get_state())[1] for T in times])
	assert_bits_equal(grid, loop, "grid vs loop of calls")


def test_jump_golden():
	rec = G["jumps"][0]
	spec = models.ALL[rec["model"]]
	E, y0, seeds = make(spec, rec["B"])
	taus, xis = jump_tape(rec["B"], rec["tape_len"])
	stats = E.integrate_jumps(unhex(rec["T"]), taus, xis)
	y, t = E.get_state(), E.t
	_, acc, rej = E.controller_state()
	for k, row in enumerate(rec["rows"]):
		assert_bits_equal(y[k], unhex(row["y"]), f"trajectory {k}")
		assert (int(acc[k]), int(rej[k])) == (row["accepted"], row["rejected"])
		assert t[k] == unhex(row["t"])
	assert stats["jumps"] > 0


def test_device_exp_equals_host_libm():
	"""exp() in generated code is glibc's algorithm on the device (csrc/glibc_math.h: jsde_exp)"""
	import math
	lib = _abi.load_library(models.validation_exp)
	rng = np.random.default_rng(11)
	for x in (rng.uniform(-5, 5, 400_000), rng.uniform(-500, 500, 300_000), rng.uniform(-1, 1, 100_000) * 2.0 ** rng.uniform(-70, -10, 100_000),
			-(rng.uniform(-4, 4, 200_000)) ** 2 + rng.uniform(-4, 4, 200_000) - 1.5):
		host = np.array([math.exp(v) for v in x])
		assert_bits_equal(_abi.device_exp(lib, x), host, "exp")


# ---- larger ensembles vs the oracle port (same seeded inputs) ----------------------------------------------------------
@pytest.mark.parametrize("name,B,T", [
	("lorenz_additive", 512, 1.0), ("lorenz_multiplicative", 256, 1.0), ("gbm", 512, 10.0), ("duffing", 256, 5.0),
	("ou", 256, 5.0), ("ou_timedep", 128, 3.0), ("lorenz_additive_fixture", 128, 2.0), ("cubic", 128, 2.0),
	("ring6", 200, 2.0), ("ring6_mult", 200, 1.0), ("validation_exp", 256, 1.0),
])
def test_ensemble_vs_oracle(name, B, T):
	spec = models.ALL[name]
	E, y0, seeds = make(spec, B)
	stats = E.integrate(T)
	ref = port.run_ensemble(spec, y0, seeds, 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], name)
	assert_bits_equal(E.t, ref["t"], name + " t")
	_, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()
	assert stats["accepted"] == int(ref["accepted"].sum())
	assert stats["n_overflow"] == 0


@pytest.mark.parametrize("slice_rounds", [0, 5, 64, 4096])
@pytest.mark.parametrize("name,B,T", [("lorenz_additive", 1000, 0.5), ("lorenz_multiplicative", 700, 0.4), ("gbm", 999, 4.0)])
def test_lane_recycling_small_grid(monkeypatch, name, B, T, slice_rounds):
	"""k_integrate with a grid of two blocks (256 lanes): every lane integrates several trajectories one after the other,
	handing queued generator pairs and stream positions back and forth — and, with time slicing, gives unfinished
	trajectories up for waiting ones every `slice_rounds` rounds (5: a trajectory changes hands dozens of times, between
	warps and blocks).  Results must not depend on that, within one call, across calls, and for the sampling grid."""
	monkeypatch.setenv("JSDE_RECYCLE", "1")
	monkeypatch.setenv("JSDE_RESIDENT_BLOCKS", "2")
	monkeypatch.setenv("JSDE_SLICE", str(slice_rounds))
	spec = models.ALL[name]
	E, y0, seeds = make(spec, B)
	stats = E.integrate(T)
	ref = port.run_ensemble(spec, y0, seeds, 0.0, T)
	assert_bits_equal(E.get_state(), ref["y"], name)
	assert_bits_equal(E.t, ref["t"], name + " t")
	_, acc, rej = E.controller_state()
	assert (acc.astype(np.int64) == ref["accepted"]).all() and (rej.astype(np.int64) == ref["rejected"]).all()
	assert stats["accepted"] == int(ref["accepted"].sum()) and stats["rejected"] == int(ref["rejected"].sum())
	# second call continues every trajectory's stream; then a sampling grid in one launch
	E.integrate(1.5 * T)
	times = np.linspace(1.6 * T, 2.0 * T, 3)
	grid = E.integrate_grid(times)
	for k in range(0, B, 97):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.integrate(T)
		P.integrate(1.5 * T)
		for j, tj in enumerate(times):
			assert_bits_equal(grid[j, k], P.integrate(float(tj)), f"trajectory {k} sample {j}")


@pytest.mark.parametrize("slice_rounds", [0, 7])
def test_lane_recycling_with_jumps(monkeypatch, slice_rounds):
	monkeypatch.setenv("JSDE_RECYCLE", "1")
	monkeypatch.setenv("JSDE_RESIDENT_BLOCKS", "1")
	monkeypatch.setenv("JSDE_SLICE", str(slice_rounds))
	spec = models.lorenz_jumpy
	B = 300
	E, y0, seeds = make(spec, B)
	taus, xis = jump_tape(B, 16)
	E.integrate_jumps(1.0, taus, xis)
	E.integrate_jumps(2.0, taus, xis)
	y = E.get_state()
	for k in range(0, B, 13):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.integrate_jumps(1.0, taus[k], xis[k])
		P.integrate_jumps(2.0, taus[k], xis[k])
		assert_bits_equal(y[k], P.get_state(), f"trajectory {k}")


@pytest.mark.parametrize("name,T", [("lorenz_additive", 1.0), ("lorenz_multiplicative", 0.6)])
def test_time_slicing_across_many_blocks(monkeypatch, name, T):
	"""2.3 waves on a grid of 40 blocks (5120 lanes, 11 800 trajectories — the shape of a 2^20 ensemble split over eight
	GPUs): trajectories migrate between SMs every 24 rounds; every trajectory must end up exactly where the oracle puts it,
	and the call's totals must count every step once."""
	monkeypatch.setenv("JSDE_RECYCLE", "1")
	monkeypatch.setenv("JSDE_RESIDENT_BLOCKS", "40")
	monkeypatch.setenv("JSDE_SLICE", "24")
	spec, B = models.ALL[name], 11800
	E, y0, seeds = make(spec, B)
	stats = E.integrate(T)
	sample = np.arange(0, B, 59)
	ref = port.run_ensemble(spec, y0[sample], seeds[sample], 0.0, T)
	assert_bits_equal(E.get_state()[sample], ref["y"], name)
	assert_bits_equal(E.t[sample], ref["t"], name + " t")
	_, acc, rej = E.controller_state()
	assert (acc[sample].astype(np.int64) == ref["accepted"]).all() and (rej[sample].astype(np.int64) == ref["rejected"]).all()
	assert stats["accepted"] == int(acc.sum()) and stats["rejected"] == int(rej.sum())
	assert (E.t == T).all()
	# the same ensemble without slicing and without recycling: identical bits for all trajectories
	monkeypatch.setenv("JSDE_RECYCLE", "0")
	F, _, _ = make(spec, B)
	F.integrate(T)
	assert_bits_equal(E.get_state(), F.get_state(), "sliced vs plain")


@pytest.mark.parametrize("name", ["lorenz_additive", "lorenz_multiplicative"])
def test_checkpoint_restore_continues_bit_for_bit(name):
	"""save -> integrate on -> restore (into a second handle, too) -> integrate on: the same bits, incl. a pending
	rejected proposal's noise and queued generator pairs."""
	spec = models.ALL[name]
	B = 300
	E, y0, seeds = make(spec, B)
	E.integrate(0.37)
	blob = E.checkpoint()
	E.integrate(0.9)
	want_y, want_t = E.get_state(), E.t.copy()
	want_dt, want_acc, want_rej = E.controller_state()
	E.restore(blob)
	E.integrate(0.9)
	assert_bits_equal(E.get_state(), want_y, "restored in place")
	F = Ensemble(spec, B)  # a fresh handle (never seeded) continues from the blob
	F.restore(blob)
	F.integrate(0.9)
	assert_bits_equal(F.get_state(), want_y, "restored into a fresh handle")
	assert_bits_equal(F.t, want_t, "times")
	dt, acc, rej = F.controller_state()
	assert_bits_equal(dt, want_dt, "step sizes")
	assert (acc == want_acc).all() and (rej == want_rej).all()
	with pytest.raises(JsdeError, match="checkpoint"):
		Ensemble(spec, B + 1).restore(blob)


def test_ragged_ensemble_size_and_single_trajectory():
	"""B not a multiple of the block size, and B = 1 (the reference's own case)."""
	spec = models.lorenz_additive
	for B in (1, 33, 129, 1000):
		E, y0, seeds = make(spec, B)
		E.integrate(0.3)
		ref = port.run_ensemble(spec, y0, seeds, 0.0, 0.3)
		assert_bits_equal(E.get_state(), ref["y"], f"B={B}")


def test_jumps_vs_oracle_many():
	spec = models.lorenz_jumpy
	B = 256
	E, y0, seeds = make(spec, B)
	taus, xis = jump_tape(B, 16)
	E.integrate_jumps(1.0, taus, xis)
	E.integrate_jumps(2.5, taus, xis)  # tape position and pending jump carry over between calls
	y = E.get_state()
	for k in range(0, B, 7):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.integrate_jumps(1.0, taus[k], xis[k])
		P.integrate_jumps(2.5, taus[k], xis[k])
		assert_bits_equal(y[k], P.get_state(), f"trajectory {k}")


def test_state_dependent_jump_intensity_vs_oracle():
	"""JumpSpec.iji (the reference's IJI(time, state)): waiting times computed on the device from the tape's unit
	exponentials and the state at scheduling time."""
	spec = models.lorenz_jumpy_rate
	B = 200
	E, y0, seeds = make(spec, B)
	taus, xis = jump_tape(B, 24)
	st = E.integrate_jumps(1.5, taus, xis)
	E.integrate_jumps(3.0, taus, xis)
	assert st["jumps"] > B // 2
	y = E.get_state()
	for k in range(0, B, 9):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.integrate_jumps(1.5, taus[k], xis[k])
		P.integrate_jumps(3.0, taus[k], xis[k])
		assert_bits_equal(y[k], P.get_state(), f"trajectory {k}")


def test_control_parameters():
	B = 64
	E, y0, seeds = make(models.gbm_pars, B)
	E.set_parameters([0.05, 0.2])
	E.integrate(2.0)
	E2, _, _ = make(models.gbm, B)
	E2.integrate(2.0)
	assert_bits_equal(E.get_state(), E2.get_state(), "shared control parameters")
	# per-trajectory parameters (a parameter scan in one launch)
	mus = np.linspace(0.0, 0.1, B)
	E3, _, _ = make(models.gbm_pars, B)
	E3.set_parameters(np.stack([mus, np.full(B, 0.2)], axis=1))
	E3.integrate(2.0)
	y3 = E3.get_state()
	for k in (0, 17, 63):
		P = port.PortCore(models.gbm_pars, y0[k], 0.0, int(seeds[k]))
		P.set_parameters(mus[k], 0.2)
		assert_bits_equal(y3[k], P.integrate(2.0), f"trajectory {k}")


def test_pin_noise_and_dump():
	spec = models.lorenz_multiplicative
	E, y0, seeds = make(spec, 5)
	E.pin_noise(4, 1e-3)
	E.get_next_step(2.5e-3)
	for k in range(5):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.pin_noise(4, 1e-3)
		P.get_next_step(2.5e-3)
		assert_bits_equal(E.dump_noise(k), P.dump_queue(), f"noise memory {k}")


def test_pin_noise_path_vs_oracle():
	"""Caller-supplied Wiener path: same noise memory as the oracle, same integration along it (whole items, Brownian
	bridges inside items, fresh noise beyond the path's end)."""
	spec = models.lorenz_multiplicative
	B, K = 40, 12
	rng = np.random.default_rng(5)
	steps = rng.uniform(2e-3, 2e-2, K)
	dw = rng.standard_normal((B, K, spec.n)) * np.sqrt(steps)[None, :, None]
	dz = rng.standard_normal((B, K, spec.n)) * np.sqrt(steps)[None, :, None]
	E, y0, seeds = make(spec, B)
	E.pin_noise_path(steps, dw, dz)
	got = E.dump_noise(7)
	assert got.shape == (K, 1 + 2 * spec.n)
	assert_bits_equal(got[:, 0], steps, "item lengths")
	assert_bits_equal(got[:, 1:1 + spec.n], dw[7], "DW")
	T = float(steps.sum() * 1.5)  # half of the run uses fresh noise
	E.integrate(T)
	y = E.get_state()
	for k in range(0, B, 3):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.pin_path(steps, dw[k], dz[k])
		assert_bits_equal(y[k], P.integrate(T), f"trajectory {k}")
	with pytest.raises(JsdeError, match="positive"):
		E.pin_noise_path([0.0], np.zeros((B, 1, spec.n)), np.zeros((B, 1, spec.n)))


def test_unsuccessful_integration_status():
	"""reference tests/test_jitcsde.py:221-240: min_step=10 cannot be met"""
	spec = models.lorenz_multiplicative
	E, y0, seeds = make(spec, 8, min_step=10.0, max_step=10.0, first_step=10.0)
	stats = E.integrate(100.0)
	assert stats["n_min_step"] == 8 and (E.status == 1).all()


def test_noise_overflow_is_reported_not_silent():
	spec = models.lorenz_additive
	y0, seeds = ensemble_inputs(3, 4)
	E = Ensemble(spec, 4, noise_capacity=2)
	E.seed(seeds)
	E.set_state(y0)
	stats = E.integrate(1.0)  # first_step=1.0 needs a dozen consecutive rejections -> deeper than 2 items
	assert stats["n_overflow"] == 4 and (E.status == 2).all()


def test_integrate_to_past_time_is_a_no_op():
	spec = models.lorenz_additive
	E, y0, seeds = make(spec, 4)
	E.integrate(0.1)
	y = E.get_state()
	stats = E.integrate(0.05)
	assert stats["accepted"] == 0 and stats["rejected"] == 0
	assert_bits_equal(E.get_state(), y, "state")


def test_moments():
	import torch
	spec = models.ou
	B = 4096
	E, y0, seeds = make(spec, B)
	E.integrate(1.0)
	acc = torch.zeros(1 + 2 * spec.n, dtype=torch.float64, device="cuda")
	shift = np.array([0.1, 0.9])
	E.moments_into(acc.data_ptr(), shift)
	torch.cuda.synchronize()
	y = E.get_state()
	got = acc.cpu().numpy()
	assert got[0] == B
	np.testing.assert_allclose(got[1:3], (y - shift).sum(axis=0), rtol=1e-12)
	np.testing.assert_allclose(got[3:5], ((y - shift) ** 2).sum(axis=0), rtol=1e-12)
