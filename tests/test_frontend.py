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
