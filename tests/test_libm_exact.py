"""
The product's restatement of glibc log/pow (jitcsde_b200/csrc/glibc_math.h), compiled for the host, must equal the
running libm bit for bit on the argument ranges the hot path uses.  If the GPU box ever ships
a libm with different tables this test — not a silent 1-ulp drift in the Gaussian stream — is what fails.
"""

import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib():
	out = os.path.join(ROOT, "oracle", "_build")
	os.makedirs(out, exist_ok=True)
	so = os.path.join(out, "libm_check.so")
	subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
		os.path.join(ROOT, "oracle", "libm_check.c"), "-o", so, "-lm"], check=True)
	return ctypes.CDLL(so)


def _run(lib, fn, x, y=None):
	out = np.empty_like(x)
	if y is None:
		getattr(lib, fn)(x.ctypes.data_as(_DP), out.ctypes.data_as(_DP), ctypes.c_size_t(x.size))
	else:
		getattr(lib, fn)(x.ctypes.data_as(_DP), ctypes.c_double(y), out.ctypes.data_as(_DP), ctypes.c_size_t(x.size))
	return out


N = 2_000_000


@pytest.mark.parametrize("name", ["unit", "near_one_below", "near_one_above", "tiny", "wide"])
def test_log_bit_exact(lib, name):
	rng = np.random.default_rng(1)
	x = {
		"unit": lambda: rng.random(N),                       # r2 of the polar method lives here
		"near_one_below": lambda: 1 - rng.random(N) * 0.07,  # the table-free branch
		"near_one_above": lambda: 1 + rng.random(N) * 0.07,
		"tiny": lambda: np.exp(-rng.random(N) * 72),         # down to 2^-104, the smallest possible r2
		"wide": lambda: np.exp(rng.uniform(-700, 700, N)),
	}[name]()
	x = x[x > 0]
	a, b = _run(lib, "restated_log_v", x), _run(lib, "libm_log_v", x)
	assert (a.view(np.uint64) == b.view(np.uint64)).all()


@pytest.mark.parametrize("y", [-1 / 1.5, -1 / 2.5])  # the controller's two exponents, _jitcsde.py:587,592
@pytest.mark.parametrize("name", ["reject", "accept", "subnormal", "near_one", "huge"])
def test_pow_bit_exact(lib, y, name):
	rng = np.random.default_rng(2)
	x = {
		"reject": lambda: np.exp(rng.uniform(np.log(1.1), np.log(1e30), N)),
		"accept": lambda: np.exp(rng.uniform(np.log(1e-300), np.log(0.5), N)),
		"subnormal": lambda: rng.random(N) * 2.2e-308,
		"near_one": lambda: 1 + rng.uniform(-0.1, 0.1, N),
		"huge": lambda: np.exp(rng.uniform(600, 709, N)),
	}[name]()
	x = x[x > 0]
	a, b = _run(lib, "restated_pow_v", x, y), _run(lib, "libm_pow_v", x, y)
	assert (a.view(np.uint64) == b.view(np.uint64)).all()


@pytest.mark.parametrize("name", ["moderate", "wide", "small", "tiny", "validation_argument"])
def test_exp_bit_exact(lib, name):
	"""exp() inside generated drift/diffusion code (e.g. the reference's validation scenario, tests/validation.py:21-32)"""
	rng = np.random.default_rng(3)
	x = {
		"moderate": lambda: rng.uniform(-5, 5, N),
		"wide": lambda: rng.uniform(-511.9, 511.9, N),
		"small": lambda: rng.uniform(-1e-3, 1e-3, N),
		"tiny": lambda: rng.uniform(-1, 1, N) * 2.0 ** rng.uniform(-80, -20, N),
		"validation_argument": lambda: -(rng.uniform(-4, 4, N)) ** 2 + rng.uniform(-4, 4, N) - 1.5,
	}[name]()
	a, b = _run(lib, "restated_exp_v", x), _run(lib, "libm_exp_v", x)
	assert (a.view(np.uint64) == b.view(np.uint64)).all()
