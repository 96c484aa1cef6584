"""Counter-based jump variates (csrc/philox_jumps.h): the Philox4x32-10 core against the published known-answer vectors
(Random123 kat_vectors), the derived variates' distributions, and — on the GPU — device values == host values bit for
bit and a tape-free run == the same run replayed from the generated tape."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_lib():
	so = os.path.join(ROOT, "oracle", "_build", "philox_tape.so")
	if not os.path.isfile(so):
		os.makedirs(os.path.dirname(so), exist_ok=True)
		subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(ROOT, "oracle", "philox_tape.c"), "-o", so, "-lm"], check=True)
	lib = ctypes.CDLL(so)
	lib.philox_jump_tape.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
	return lib


def host_tape(seed, first, count, length):
	taus, xis = np.empty((count, length)), np.empty((count, length))
	host_lib().philox_jump_tape(seed, first, count, length, taus.ctypes.data, xis.ctypes.data)
	return taus, xis


def test_philox4x32_10_known_answers():
	lib = host_lib()
	kat = [
		([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
		([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
		([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
	]
	for ctr, key, want in kat:
		out = (ctypes.c_uint32 * 4)()
		lib.philox_block((ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), out)
		assert list(out) == want


def test_golden_variates():
	import json
	g = json.load(open(os.path.join(ROOT, "tests", "golden", "philox_jumps.json")))
	taus, xis = host_tape(g["seed"], g["first"], 2, 4)
	assert [[float(v).hex() for v in r] for r in taus] == g["taus"]
	assert [[float(v).hex() for v in r] for r in xis] == g["xis"]


def test_variates_are_what_they_claim():
	taus, xis = host_tape(2024, 0, 64, 4000)
	n = taus.size
	assert (taus >= 0).all() and np.isfinite(taus).all() and np.isfinite(xis).all()
	assert abs(taus.mean() - 1.0) < 4 / np.sqrt(n) and abs(taus.var() - 1.0) < 4 * np.sqrt(8.0 / n)  # Exp(1)
	assert abs(xis.mean()) < 4 / np.sqrt(n) and abs(xis.var() - 1.0) < 4 * np.sqrt(2.0 / n)  # N(0, 1)
	assert abs(np.corrcoef(taus.ravel(), xis.ravel())[0, 1]) < 4 / np.sqrt(n)
	# a function of (seed, trajectory, jump) only: shards see what the whole ensemble sees
	t2, x2 = host_tape(2024, 40, 8, 100)
	assert (t2 == taus[40:48, :100]).all() and (x2 == xis[40:48, :100]).all()
	t3, _ = host_tape(2025, 0, 8, 100)
	assert not (t3 == taus[:8, :100]).any()


@pytest.mark.gpu
def test_device_variates_equal_host_and_tape_free_run_equals_replay():
	from jitcsde_b200 import _abi, models, sharding
	from jitcsde_b200._abi import ControllerParameters, Ensemble
	from util import assert_bits_equal
	spec = models.lorenz_jumpy_rate
	lib = _abi.load_library(spec)
	seed, first, B, L = 0x1234567890ABCDEF, 1000, 300, 600  # a long tape: trajectories in a jump cascade must not exhaust the replay
	taus, xis = _abi.generated_jump_tape(lib, seed, first, B, L)
	ht, hx = host_tape(seed, first, B, L)
	assert_bits_equal(taus, ht, "waiting-time variates")
	assert_bits_equal(xis, hx, "amplitude variates")

	def fresh():
		E = Ensemble(spec, B)
		E.seed(sharding.global_seeds(0, B))
		E.set_state(sharding.global_initial_states(0, B, spec.n), 0.0)
		E.set_integration_parameters(ControllerParameters())
		return E
	A, R = fresh(), fresh()
	sa = A.integrate_jumps_generated(1.0, seed, first)
	A.integrate_jumps_generated(1.8, seed, first)
	R.integrate_jumps(1.0, taus, xis)
	R.integrate_jumps(1.8, taus, xis)
	assert sa["jumps"] > B // 2
	assert_bits_equal(A.get_state(), R.get_state(), "tape-free run vs replay")
	assert_bits_equal(A.t, R.t, "times")
