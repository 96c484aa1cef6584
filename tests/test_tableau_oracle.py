"""Secondary oracle (SURVEY.md §8 c, reference tests/test_step_functions.py:16-55): single Rößler SRK steps evaluated by the
reference's own Python code — `_python_core.perform_step` / `perform_SRA_step` and the Butcher-tableau evaluation of
tests/Butcher.py, tests/Butcher_SRA.py — stored in tests/golden/tableau_steps.json by oracle/make_golden_tableau.py.

The port (CPU) and the CUDA single-step surface (GPU) receive the same prescribed Wiener increments through
pin_noise_path and must give the same proposal and the same error estimate.  The tableau evaluation orders its
operations differently, so this rung is a tolerance test: 1e-12 relative on the new state (the reference's own test
uses 1e-5), 1e-6 on the error ratio (a sum of small differences)."""

import json
import os

import numpy as np
import pytest

from jitcsde_b200 import models
from oracle import port
from util import GOLDEN_DIR, unhex

ATOL, RTOL = 1e-10, 1e-5


def cases():
	with open(os.path.join(GOLDEN_DIR, "tableau_steps.json")) as fh:
		return json.load(fh)["cases"]


CASES = cases()


def expected(case, which):
	new_y, E_D, E_N = (unhex(case[which][k]) for k in ("new_y", "E_D", "E_N"))
	p = np.max((np.abs(E_D) + np.abs(E_N)) / (ATOL + RTOL * np.abs(new_y)))
	return new_y, p


def check(case, new_y, p):
	for which in ("python_core", "butcher"):
		ref_y, ref_p = expected(case, which)
		np.testing.assert_allclose(new_y, ref_y, rtol=1e-12, atol=1e-300, err_msg=f"{case['model']} vs {which}: new state")
		np.testing.assert_allclose(p, ref_p, rtol=1e-6, err_msg=f"{case['model']} vs {which}: error ratio")


def test_the_two_reference_evaluations_agree():
	for case in CASES:
		a, pa = expected(case, "python_core")
		b, pb = expected(case, "butcher")
		np.testing.assert_allclose(a, b, rtol=1e-13)
		np.testing.assert_allclose(pa, pb, rtol=1e-7)


@pytest.mark.parametrize("index", range(len(CASES)))
def test_port_step_matches_the_tableau(index):
	case = CASES[index]
	spec = models.ALL[case["model"]]
	h = unhex(case["h"])
	P = port.PortCore(spec, unhex(case["y"]), unhex(case["t"]), 0)
	P.pin_path([h], unhex(case["dw"])[None, :], unhex(case["dz"])[None, :])
	P.get_next_step(h)
	p = P.get_p(ATOL, RTOL)
	P.accept_step()
	check(case, P.get_state(), p)
	assert P.t == unhex(case["t"]) + h


@pytest.mark.gpu
def test_cuda_single_step_surface_matches_the_tableau():
	from jitcsde_b200._abi import Ensemble
	engines = {}
	for case in CASES:
		spec = models.ALL[case["model"]]
		E = engines.get(spec.name) or engines.setdefault(spec.name, Ensemble(spec, 1))
		h = unhex(case["h"])
		E.seed(np.zeros(1, dtype=np.uint32))
		E.set_state(unhex(case["y"])[None, :], unhex(case["t"]))
		E.pin_noise_path([h], unhex(case["dw"])[None, None, :], unhex(case["dz"])[None, None, :])
		E.get_next_step(h)
		p = E.get_p(ATOL, RTOL)[0]
		E.accept_step()
		check(case, E.get_state()[0], p)
		# and bit for bit what the port computes from the same increments
		P = port.PortCore(spec, unhex(case["y"]), unhex(case["t"]), 0)
		P.pin_path([h], unhex(case["dw"])[None, :], unhex(case["dz"])[None, :])
		P.get_next_step(h)
		assert p == P.get_p(ATOL, RTOL)
		P.accept_step()
		assert (E.get_state()[0].view(np.uint64) == P.get_state().view(np.uint64)).all()
