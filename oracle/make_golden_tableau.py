"""
TEST INFRASTRUCTURE.  Secondary oracle (SURVEY.md §8 c "known-answer tests to re-use", §4 row 1): single SRK steps
evaluated by the reference's OWN Python code —

  * `jitcsde/_python_core.py:8-50`  perform_step (SRIW1) / perform_SRA_step (SRA1), the lambdified core's arithmetic, and
  * `tests/Butcher.py:4-80`, `tests/Butcher_SRA.py:4-45`  the human-readable evaluation of the Rößler Butcher tableaus
    that the reference's `tests/test_step_functions.py:16-55` checks the former against —

imported in place from /root/reference (never copied) with stand-ins for the two modules that are not installed here
(`symengine`, `jitcxde_common.symbolic`; `_python_core.py` only needs them for its class, not for the two step
functions).  Inputs: the canned models of jitcsde_b200/models.py (their printed C expressions evaluated as Python),
random states, times and step sizes, and PRESCRIBED Wiener increments (DW, DZ), from which the iterated integrals are
formed exactly as the core does (`jitced_template.c:269-291`).  Output: tests/golden/tableau_steps.json with
(new_y, E_D, E_N) of both evaluations as hex floats.  The port (CPU test) and the CUDA single-step surface (GPU test)
are fed the same increments through pin_noise_path and must reproduce them to 1e-12.

    python -m oracle.make_golden_tableau
"""

import json
import math
import os
import sys
import types

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("JSDE_REFERENCE_ROOT", "/root/reference")

from jitcsde_b200 import models  # noqa: E402


def import_reference_step_functions():
	for name in ("symengine", "jitcxde_common", "jitcxde_common.symbolic"):
		sys.modules.setdefault(name, types.ModuleType(name))
	sys.modules["jitcxde_common.symbolic"].ordered_subs = None
	import importlib.util

	def load(path, name):
		spec = importlib.util.spec_from_file_location(name, path)
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
		return mod
	core = load(os.path.join(REF, "jitcsde", "_python_core.py"), "jsde_ref_python_core")
	sri = load(os.path.join(REF, "tests", "Butcher.py"), "jsde_ref_butcher")
	sra = load(os.path.join(REF, "tests", "Butcher_SRA.py"), "jsde_ref_butcher_sra")
	return core, sri, sra


def python_callables(spec):
	"""f(t, y) and g(t, y) from the model's printed C expressions (polynomial/rational: plain Python arithmetic)"""
	env = {"fabs": abs, "sqrt": math.sqrt, "exp": math.exp, "sin": math.sin, "cos": math.cos}

	def make(exprs):
		code = [compile(e, "<model>", "eval") for e in exprs]

		def fn(t, Y):
			scope = dict(env, t=t, y=lambda i: Y[i])
			return np.array([float(eval(c, scope)) for c in code])
		return fn
	return make(spec.f), make(spec.g)


def hx(a):
	a = np.asarray(a, dtype=np.float64)
	return float(a).hex() if a.ndim == 0 else [hx(x) for x in a]


CASES = ("lorenz_multiplicative", "duffing", "gbm", "ring6_mult", "lorenz_additive", "lorenz_additive_fixture", "ou_timedep", "ring6")


def main():
	core, sri, sra = import_reference_step_functions()
	rng = np.random.default_rng(2024)
	out = []
	for name in CASES:
		spec = models.ALL[name]
		f, g = python_callables(spec)
		for rep in range(6):
			n = spec.n
			y = rng.uniform(0.5, 2.0, n) * rng.choice([-1.0, 1.0], n)
			t = float(rng.uniform(0.0, 2.0))
			h = float(10 ** rng.uniform(-4, -1.5))
			dw = rng.standard_normal(n) * math.sqrt(h)
			dz = rng.standard_normal(n) * math.sqrt(h)
			# get_I (jitced_template.c:269-291)
			I_1 = dw
			I_11 = (dw ** 2 - h) / 2
			I_111 = (dw ** 3 - 3 * h * dw) / 6
			I_10 = (dw + dz / math.sqrt(3)) * h / 2
			if spec.additive:
				ga = lambda tt: g(tt, y)  # additive: g depends on time only  # noqa: E731
				a = core.perform_SRA_step(t, h, f, ga, y, I_1, I_10)
				b = sra.perform_step(t, h, f, ga, y, I_1, I_10)
			else:
				a = core.perform_step(t, h, f, g, y, I_1, I_11, I_111, I_10)
				b = sri.perform_step(t, h, f, g, y, I_1, I_11, I_111, I_10)
			out.append({"model": name, "y": hx(y), "t": hx(t), "h": hx(h), "dw": hx(dw), "dz": hx(dz),
				"python_core": {"new_y": hx(a[0]), "E_D": hx(a[1]), "E_N": hx(a[2])},
				"butcher": {"new_y": hx(b[0]), "E_D": hx(b[1]), "E_N": hx(b[2])}})
	path = os.path.join(ROOT, "tests", "golden", "tableau_steps.json")
	with open(path, "w") as fh:
		json.dump({"source": "reference jitcsde/_python_core.py perform_step/perform_SRA_step and tests/Butcher.py, tests/Butcher_SRA.py",
			"cases": out}, fh, indent=1)
	print("wrote", path, len(out), "steps")


if __name__ == "__main__":
	main()
