"""Front end (CPU part): symbolic input styles, code generation, validation and error behaviour mirror the reference
(tests/test_jitcsde.py:242-263, tests/test_sympy_input.py) as far as they do not need a device."""

import numpy as np
import pytest
import sympy

from jitcsde_b200 import UnsuccessfulIntegration, jitcsde, jitcsde_jump, t, xi, y
from jitcsde_b200._codegen import ccode, spec_from_symbolic

f = [10 * (y(1) - y(0)), y(0) * (28 - y(2)) - y(1), y(0) * y(1) - sympy.Rational(8, 3) * y(2)]
g = [0.1 * y(i) for i in range(3)]


def test_additivity_is_detected():
	assert not jitcsde(f, g, verbose=False).additive
	assert jitcsde(f, [0.1, 0.2, 0.3], verbose=False).additive
	assert jitcsde(f, [0.1 * t, 0.2, 0.3], verbose=False).additive  # time-dependent but state-independent
	assert jitcsde(f, g, additive=True, verbose=False).additive       # the caller's word is taken, like the reference


def test_input_styles_give_the_same_model():
	a = jitcsde(f, g, verbose=False).spec
	b = jitcsde(lambda: (e for e in f), lambda: (e for e in g), n=3, verbose=False).spec
	c = jitcsde({y(i): f[i] for i in range(3)}, {y(i): g[i] for i in range(3)}, verbose=False).spec
	assert a.f == b.f == c.f and a.g == b.g == c.g
	# a user's own sympy.Function("y") is accepted (reference tests/test_sympy_input.py)
	Y = sympy.Function("y")
	d = jitcsde([10 * (Y(1) - Y(0)), Y(0) * (28 - Y(2)) - Y(1), Y(0) * Y(1) - sympy.Rational(8, 3) * Y(2)], [0.1 * Y(i) for i in range(3)], verbose=False).spec
	assert d.f == a.f


def test_helpers_and_control_parameters():
	k, a_ = sympy.symbols("k a_")
	h = sympy.Symbol("coupling")
	S = jitcsde([-k * y(0) + h, -y(1) + h], [a_, a_], helpers=[(h, 0.5 * (y(0) + y(1)))], control_pars=[k, a_], verbose=False)
	assert S.spec.control_pars == ("k", "a_")
	assert "coupling" not in "".join(S.spec.f)
	assert S.additive


def test_integer_powers_are_printed_as_products():
	assert "pow" not in ccode(y(0) ** 3)
	assert ccode(y(0) ** 2) == "((y(0))*(y(0)))"


def test_stratonovich_conversion():
	# dy = -y dt + 0.5 y ∘ dW  ->  Ito drift -y + 0.5*0.5*y/2 ... = -y + g*g'/2
	S = jitcsde([-y(0)], [0.5 * y(0)], ito=False, verbose=False)
	expr = sympy.sympify(S.f_sym[0])
	assert sympy.simplify(expr - (-y(0) + 0.5 * y(0) * 0.5 / 2)) == 0


def test_check_rejects_bad_input():
	with pytest.raises(ValueError):
		jitcsde([y(-1)], [y(0)], verbose=False).check()
	with pytest.raises(ValueError):
		jitcsde([y(1)], [y(0)], verbose=False).check()
	with pytest.raises(ValueError):
		jitcsde([sympy.Symbol("z") * y(0)], [y(0)], verbose=False).check()
	jitcsde(f, g, verbose=False).check()


def test_constructor_errors():
	with pytest.raises(ValueError):
		jitcsde(f, verbose=False)
	with pytest.raises(NotImplementedError):
		jitcsde(f, g, callback_functions=[(sympy.Function("cb"), lambda y: 0.0, 0)], verbose=False)
	with pytest.raises(NotImplementedError):
		jitcsde_jump(np.ones(4), [0, 0, xi], f, g, ito=False, verbose=False)
	with pytest.raises(NotImplementedError):
		jitcsde(f, g, verbose=False).generate_lambdas()
	with pytest.raises(ValueError):
		jitcsde(f, g, verbose=False).set_initial_value([1.0, 2.0])


def test_integration_parameter_validation():
	S = jitcsde(f, g, verbose=False)
	with pytest.warns(UserWarning):
		S.set_integration_parameters(first_step=20.0, max_step=1.0)
	with pytest.raises(AssertionError):
		S.set_integration_parameters(decrease_threshold=0.5)
	with pytest.raises(AssertionError):
		S.set_integration_parameters(atol=-1.0)


def test_jump_spec_is_generated():
	S = jitcsde_jump(np.ones(8), [0, 0, sympy.Abs(y(2)) * xi], f, g, verbose=False)
	assert S.spec.jump is not None and "xi" in S.spec.jump.amp[2] and "fabs" in S.spec.jump.amp[2]
	S.check()
	with pytest.raises(ValueError):
		jitcsde_jump(np.ones(8), [0, xi], f, g, verbose=False)


def test_generated_model_compiles():
	S = jitcsde(f, g, verbose=False)
	assert S.compile_C().endswith(".so")
	assert issubclass(UnsuccessfulIntegration, Exception)


def test_jump_waiting_time_expression_is_printed_into_the_spec():
	from jitcsde_b200 import jitcsde_jump, tau
	f = [10 * (y(1) - y(0)), y(0) * (28 - y(2)) - y(1), y(0) * y(1) - 8 / 3 * y(2)]
	g = [0.1 * y(i) for i in range(3)]
	SDE = jitcsde_jump(np.ones(4), [0, 0, abs(y(2)) * xi], f, g, waiting_time=tau / (1 + y(2) ** 2), verbose=False)
	assert SDE.spec.jump.iji.replace(" ", "") == "tau/(((y(2))*(y(2)))+1)"
	SDE.check()
	plain = jitcsde_jump(np.ones(4), [0, 0, abs(y(2)) * xi], f, g, verbose=False)
	assert plain.spec.jump.iji == "tau" and plain.spec.name != SDE.spec.name


# ---- compile_C options and large symbolic systems -------------------------------------------------------------------
def test_cse_emits_temporaries_and_keeps_the_values():
	"""do_cse (reference `_jitcsde.py:366-375`): same drift, shared subexpressions computed once; the oracle compiles the
	same body, so parity tests see no difference between the two sides."""
	from oracle import port
	fs = [sympy.sin(y(0) * y(1)) + y(0) * y(1), sympy.cos(y(0) * y(1)) - y(1)]
	gs = [0.1 * y(0) * y(1), 0.2]
	plain, _, _ = spec_from_symbolic(fs, gs, name="cse_plain")
	cse, _, _ = spec_from_symbolic(fs, gs, name="cse_on", do_cse=True)
	assert "const double cse_d0" in cse.f_body and cse.f_body.count("y(0)*y(1)") <= 1 and plain.f_body == ""
	y0 = np.array([0.3, -0.7])
	a = port.PortCore(plain, y0, 0.0, 5).integrate(0.2)
	b = port.PortCore(cse, y0, 0.0, 5).integrate(0.2)
	np.testing.assert_allclose(a, b, rtol=1e-11)


def test_compile_options_are_honoured_or_refused():
	S = jitcsde([sympy.sin(y(0)) ** 2 + sympy.cos(y(0)) ** 2 - y(0)], [0.1], verbose=False)
	before = S.spec.f
	S._library = None
	S.spec, S.f_sym, S.g_sym = S._make_spec(simplify=True)
	assert S.spec.f != before and "sin" not in S.spec.f[0]  # sin^2 + cos^2 -> 1
	from jitcsde_b200 import models
	T = jitcsde(spec=models.ou, verbose=False)
	with pytest.raises(NotImplementedError, match="symbolic input"):
		T.compile_C(do_cse=True)
	with pytest.raises(NotImplementedError, match="numpy_rng"):
		T.compile_C(numpy_rng=True)


def symbolic_network(units, seed=7):
	"""the FitzHugh–Nagumo network of models.fhn_network written the way a user of the reference would: Python loops"""
	A = np.random.default_rng(seed).random((units, units))
	fs = [y(i) - y(i) ** 3 / 3 - y(units + i) + 0.5 + sum(float(A[i, j]) / units * (y(j) - y(i)) for j in range(units)) for i in range(units)]
	fs += [0.08 * (y(i) + 0.7 - 0.8 * y(units + i)) for i in range(units)]
	gs = [sympy.Float(0.05)] * units + [sympy.Float(0.0)] * units
	return fs, gs


def test_large_symbolic_systems_become_wide_models():
	"""n above the lane kernels' range: the dense linear coupling is recognised and the components fold into two templates"""
	from oracle import port
	units = 40
	fs, gs = symbolic_network(units)
	spec, fsym, _ = spec_from_symbolic(fs, gs)
	assert spec.wide and spec.coupling_units == units and spec.additive and spec.n == 2 * units
	assert spec.f_component.count("?") == 1 and "coupling" in spec.f_component  # v-template : w-template
	# the generated code evaluates the user's drift: one tiny noise-free step through the oracle's build of the same strings
	y0 = np.random.default_rng(1).random(spec.n)
	P = port.PortCore(spec, y0, 0.0, 3)
	h = 1e-7
	P.pin_path([h], np.zeros((1, spec.n)), np.zeros((1, spec.n)))
	P.get_next_step(h)
	P.accept_step()
	Yb = sympy.IndexedBase("Y")
	drift = sympy.lambdify(Yb, [e.subs({y(k): Yb[k] for k in range(spec.n)}) for e in fsym], "numpy")
	np.testing.assert_allclose((P.get_state() - y0) / h, np.array(drift(y0)), rtol=0, atol=1e-5)


def test_irregular_large_systems_are_refused_with_a_reason():
	n = 80
	rng = np.random.default_rng(0)
	fs = [-(1.0 + float(rng.random())) * y(i) ** (2 + i % 3) - float(rng.random()) * y((i * 7 + 3) % n) ** 2 for i in range(n)]
	with pytest.raises(NotImplementedError, match="structural templates"):
		spec_from_symbolic(fs, [0.1] * n)


def test_jump_randomness_is_renewed_on_reset_unless_given_explicitly():
	"""reference: IJI and amp draw from a live generator, so a reset instance gives an independent sample; here the tape is
	redrawn (callable IJI, implicit xis) and the generated mode takes a new seed — explicit arrays and seeds are replayed"""
	amp = [0.5 * xi]
	rng = np.random.default_rng(0)
	A = jitcsde_jump(lambda time, state: rng.exponential(1.0), amp, [-y(0)], [0.1], rng=np.random.default_rng(1), ensemble=4, tape_length=8, verbose=False)
	A._make_tape()
	first = (A._taus.copy(), A._xis.copy())
	A.set_initial_value([1.0], 0.0)  # -> reset_integrator
	assert A._taus is None and not A._tape_uploaded
	A._make_tape()
	assert not np.array_equal(A._taus, first[0]) and not np.array_equal(A._xis, first[1])
	taus, xis = np.full((4, 3), 0.5), np.zeros((4, 3))
	B = jitcsde_jump(taus, amp, [-y(0)], [0.1], xis=xis, ensemble=4, verbose=False)
	B._make_tape()
	B.set_initial_value([1.0], 0.0)
	assert B._taus is not None and np.array_equal(B._taus, taus)
	C = jitcsde_jump("generated", amp, [-y(0)], [0.1], rng=np.random.default_rng(2), ensemble=4, verbose=False)
	s1 = C._generated_seed()
	C.set_initial_value([1.0], 0.0)
	assert C._generated_seed() != s1
	D = jitcsde_jump("generated", amp, [-y(0)], [0.1], jump_seed=99, ensemble=4, verbose=False)
	D.set_initial_value([1.0], 0.0)
	assert D._generated_seed() == 99
