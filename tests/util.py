"""Shared helpers for the parity tests."""

import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(v):
	"""hex-float string(s) → float64 array/scalar (exact)."""
	if isinstance(v, str):
		return float.fromhex(v)
	return np.array([unhex(x) for x in v], dtype=np.float64)


def golden():
	with open(os.path.join(GOLDEN_DIR, "reference_core.json")) as fh:
		return json.load(fh)


def bits(a):
	return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bits_equal(a, b, what=""):
	a = np.asarray(a, dtype=np.float64)
	b = np.asarray(b, dtype=np.float64)
	assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
	bad = bits(a) != bits(b)
	if bad.any():
		idx = np.argwhere(bad.reshape(a.shape))[:5]
		msg = "; ".join(f"{tuple(i)}: {a[tuple(i)].hex()} vs {b[tuple(i)].hex()}" for i in idx)
		raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} values differ bitwise, e.g. {msg}")


def ensemble_inputs(n, B):
	"""y0 / seed rule of SURVEY.md §8(d); must match oracle/make_golden.py."""
	return np.random.default_rng(42).random((B, n)), np.arange(B, dtype=np.uint32)


def jump_tape(B, length, scale=1.0):
	taus = np.empty((B, length))
	xis = np.empty((B, length))
	for k in range(B):
		r = np.random.default_rng([42, k])
		taus[k] = r.exponential(scale, length)
		xis[k] = r.standard_normal(length)
	return taus, xis
