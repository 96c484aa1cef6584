"""The reference's API-level test suite, variant by variant (reference tests/test_jitcsde.py, tests/test_sympy_input.py):
"all kinds of formatting the input actually work and produce the same result" — default, initial value as a dictionary,
helpers (same / prefiltered / auto-filtered), generator and dictionary input, control parameters, `simplify=False`,
`do_cse=True`, `save_compiled` → `module_location`, Stratonovich input (plain and with helpers), additive noise (plain and
with helpers), and symbols imported from elsewhere.  Each variant integrates dy = (−y³+4y+y²)dt + (5·exp(−y²+y−3/2)+3)dW
from y = 1 to t = 0.001 with seed 42 and must agree with the first variant to rtol 1e-3 (the reference's criterion,
test_jitcsde.py:29-39) — here the variants that print the same operations must agree bit for bit, too.

* without a GPU the engine is the oracle built from each variant's printed expressions (what the input handling,
  substitution and Itō conversion produced);
* with a GPU the engine is the CUDA path, compared with that oracle bit for bit (`exp` is restated from glibc).

Not mirrored: `numpy_rng` and the Python core (`generate_lambdas`) are test devices of the reference's C backend and
raise NotImplementedError here, as do Python callbacks inside f and g (they cannot run in a GPU lane)."""

import numpy as np
import pytest
import sympy
from sympy import Rational, exp, symbols

from jitcsde_b200 import UnsuccessfulIntegration, jitcsde, y
from oracle import port
from util import assert_bits_equal

f = [-y(0) ** 3 + 4 * y(0) + y(0) ** 2]
g = [5 * exp(-y(0) ** 2 + y(0) - Rational(3, 2)) + 3]
initial_value = np.array([1.0])
T_END = 0.001

# helpers (test_jitcsde.py:80-93)
y2_m_y, state, exp_term, polynome = symbols("y2_m_y, state, exp_term, polynome")
helpers = [
	(state, y(0)),
	(y2_m_y, state ** 2 - state),
	(polynome, -state ** 3 + 5 * state + y2_m_y),
	(exp_term, exp(-y2_m_y - Rational(3, 2))),
]
f_helpers = [helpers[i] for i in (0, 1, 2)]
g_helpers = [helpers[i] for i in (0, 1, 3)]
f_with_helpers = [polynome]
g_with_helpers = [5 * exp_term + 3]


def f_generator():
	yield from f


def g_generator():
	yield from g


five, four = symbols("five four")
f_control_pars = [-y(0) ** 3 + four * y(0) + y(0) ** 2]
g_control_pars = [five * exp(-y(0) ** 2 + y(0) - Rational(3, 2)) + 3]

# Stratonovich (test_jitcsde.py:166-171): the drift that the Itō conversion turns back into f
f_strat = [
	Rational(5, 2) * exp(-Rational(3, 2) - y(0) ** 2 + y(0)) * (2 * y(0) - 1) * (3 + 5 * exp(-Rational(3, 2) - y(0) ** 2 + y(0)))
	+ y(0) ** 2 - y(0) ** 3 + 4 * y(0)
]
f_strat_with_helpers = [polynome - 5 * exp(-Rational(3, 2) - y2_m_y) * (1 - 2 * state) * (3 + 5 * exp(-Rational(3, 2) - y2_m_y)) / 2]
g_add = [3]

quiet = dict(verbose=False)
MULTIPLICATIVE = {
	"default": lambda: jitcsde(f, g, **quiet),
	"helpers_same": lambda: jitcsde(f_with_helpers, g_with_helpers, helpers=helpers, g_helpers="same", **quiet),
	"helpers_prefiltered": lambda: jitcsde(f_with_helpers, g_with_helpers, helpers=f_helpers, g_helpers=g_helpers, **quiet),
	"helpers_auto": lambda: jitcsde(f_with_helpers, g_with_helpers, helpers=helpers, g_helpers="auto", **quiet),
	"generator": lambda: jitcsde(f_generator, g_generator, **quiet),
	"dictionary": lambda: jitcsde({y(0): f[0]}, {y(0): g[0]}, **quiet),
	"control_pars": lambda: jitcsde(f_control_pars, g_control_pars, control_pars=[five, four], **quiet),
	"stratonovich": lambda: jitcsde(f_strat, g, ito=False, **quiet),
	"stratonovich_helpers_same": lambda: jitcsde(f_strat_with_helpers, g_with_helpers, helpers=helpers, g_helpers="same", ito=False, **quiet),
	"stratonovich_helpers_prefiltered": lambda: jitcsde(f_strat_with_helpers, g_with_helpers, helpers=f_helpers, g_helpers=g_helpers, ito=False, **quiet),
	"stratonovich_helpers_auto": lambda: jitcsde(f_strat_with_helpers, g_with_helpers, helpers=helpers, g_helpers="auto", ito=False, **quiet),
}
ADDITIVE = {
	"additive": lambda: jitcsde(f, g_add, **quiet),
	"additive_helpers_same": lambda: jitcsde(f_with_helpers, g_add, helpers=helpers, g_helpers="same", **quiet),
	"additive_helpers_auto": lambda: jitcsde(f_with_helpers, g_add, helpers=helpers, g_helpers="auto", **quiet),
}
#: variants whose printed operations are those of the first one: bit-equal results, not only rtol 1e-3
SAME_OPERATIONS = {"default", "generator", "dictionary"}


def oracle_result(SDE, pars=None):
	P = port.PortCore(SDE.spec, initial_value, 0.0, 42)
	if pars is not None:
		P.set_parameters(*pars)
	return P.integrate(T_END)


def close_to(base, new):
	np.testing.assert_allclose(new - initial_value, base - initial_value, rtol=1e-3)


# ---- the printed model of every variant, integrated by the oracle (no GPU) ---------------------------------------------------
@pytest.fixture(scope="module")
def oracle_base():
	return {"mult": oracle_result(MULTIPLICATIVE["default"]()), "add": oracle_result(ADDITIVE["additive"]())}


@pytest.mark.parametrize("name", list(MULTIPLICATIVE) + list(ADDITIVE))
def test_variant_prints_the_same_sde(name, oracle_base):
	additive = name in ADDITIVE
	SDE = (ADDITIVE if additive else MULTIPLICATIVE)[name]()
	SDE.check()
	assert SDE.additive == additive and SDE.n == 1
	result = oracle_result(SDE, pars=(5.0, 4.0) if name == "control_pars" else None)
	base = oracle_base["add" if additive else "mult"]
	assert result[0] != initial_value[0]
	close_to(base, result)
	if name in SAME_OPERATIONS:
		assert_bits_equal(result, base, name)


def test_stratonovich_conversion_restores_the_ito_drift():
	S = MULTIPLICATIVE["stratonovich"]()
	assert sympy.simplify(sympy.sympify(S.f_sym[0]) - f[0]) == 0
	for name in ("stratonovich_helpers_same", "stratonovich_helpers_prefiltered", "stratonovich_helpers_auto"):
		assert sympy.simplify(sympy.sympify(MULTIPLICATIVE[name]().f_sym[0]) - f[0]) == 0, name


def test_symbols_from_elsewhere_give_one_model():
	"""tests/test_sympy_input.py: t, y and cos from SymPy by hand, from the package, from its sympy_symbols module, mixed"""
	import jitcsde_b200
	import jitcsde_b200.sympy_symbols
	sources = [
		(sympy.Symbol("t", real=True), sympy.Function("y", real=True), sympy.cos),
		(jitcsde_b200.t, jitcsde_b200.y, sympy.cos),
		(jitcsde_b200.sympy_symbols.t, jitcsde_b200.sympy_symbols.y, sympy.cos),
		(jitcsde_b200.sympy_symbols.t, sympy.Function("y"), sympy.cos),
	]
	printed = set()
	for t_, y_, cos_ in sources:
		SDE = jitcsde([cos_(t_) * y_(0)], [cos_(t_) * y_(0) / 10], **quiet)
		SDE.check()
		printed.add((SDE.spec.f, SDE.spec.g, SDE.spec.digest()))
	assert len(printed) == 1, printed


def test_save_and_load_round_trip_of_the_printed_model(tmp_path):
	"""test_jitcsde.py:65-68 without a device: what `module_location` restores is the model that was saved"""
	from jitcsde_b200.modelspec import ModelSpec
	for factory in (MULTIPLICATIVE["helpers_auto"], ADDITIVE["additive"]):
		spec = factory().spec
		again = ModelSpec.from_json(spec.to_json())
		assert again == spec and again.digest() == spec.digest()
	from jitcsde_b200 import models
	wide = models.fhn_small_jumpy
	again = ModelSpec.from_json(wide.to_json())
	assert again == wide and again.digest() == wide.digest() and again.wide and again.jump == wide.jump


def test_refused_features_say_so():
	with pytest.raises(NotImplementedError):
		jitcsde(f, g, callback_functions=[(sympy.Function("call"), lambda y: 3, 0)], **quiet)
	SDE = jitcsde(f, g, **quiet)
	with pytest.raises(NotImplementedError):
		SDE.compile_C(numpy_rng=True)
	with pytest.raises(NotImplementedError):
		SDE.generate_lambdas()


# ---- the same on the device --------------------------------------------------------------------------------------------------
def device_result(SDE, as_dict=False, pars=None):
	SDE.set_seed(42)
	SDE.set_initial_value({y(0): initial_value[0]} if as_dict else initial_value, 0.0)
	if pars is not None:
		SDE.set_parameters(pars)
	SDE.check()
	result = SDE.integrate(T_END)
	assert SDE.y_dict[y(0)] == result[0]  # test_jitcsde.py:72
	assert SDE.t == T_END
	return result


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MULTIPLICATIVE) + list(ADDITIVE))
def test_variant_on_the_device(name, oracle_base):
	additive = name in ADDITIVE
	SDE = (ADDITIVE if additive else MULTIPLICATIVE)[name]()
	pars = [5, 4] if name == "control_pars" else None
	result = device_result(SDE, pars=pars)
	close_to(oracle_base["add" if additive else "mult"], result)
	assert_bits_equal(result, oracle_result(SDE, pars=pars), f"{name}: CUDA vs the oracle built from the same printed expressions")
	# reproducibility (test_jitcsde.py:50-51): setting seed and initial value twice changes nothing; a second instance agrees
	again = (ADDITIVE if additive else MULTIPLICATIVE)[name]()
	again.set_seed(42)
	again.set_initial_value(initial_value, 0.0)
	assert_bits_equal(device_result(again, pars=pars), result, "second instance")
	# like on the reference's Python object, the adapted step size survives set_initial_value: a rerun on the same
	# instance reproduces the first one once the integration parameters are set again
	SDE.set_integration_parameters()
	assert_bits_equal(device_result(SDE, pars=pars), result, "same instance, parameters reset")


@pytest.mark.gpu
def test_initial_value_as_a_dictionary(oracle_base):
	result = device_result(MULTIPLICATIVE["default"](), as_dict=True)
	assert_bits_equal(result, oracle_base["mult"], "dictionary initial value")


@pytest.mark.gpu
@pytest.mark.parametrize("options", [dict(simplify=False), dict(do_cse=True), dict(simplify=True)])
def test_compile_options(options, oracle_base):
	"""test_jitcsde.py:56-63; simplify=True may re-associate (rtol), the other two keep the operations"""
	SDE = MULTIPLICATIVE["default"]()
	SDE.compile_C(**options)
	result = device_result(SDE)
	close_to(oracle_base["mult"], result)
	assert_bits_equal(result, oracle_result(SDE), "vs the oracle built from the regenerated model")
	if not options.get("simplify"):
		assert_bits_equal(result, oracle_base["mult"], str(options))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["default", "helpers_auto", "control_pars", "additive"])
def test_save_and_load(name, tmp_path, oracle_base):
	"""test_jitcsde.py:65-68: save_compiled → a new object from module_location integrates to the same result"""
	additive = name in ADDITIVE
	first = (ADDITIVE if additive else MULTIPLICATIVE)[name]()
	pars = [5, 4] if name == "control_pars" else None
	expected = device_result(first, pars=pars)
	filename = first.save_compiled(str(tmp_path) + "/", overwrite=True)
	assert filename.endswith(first.MODULE_SUFFIX)
	with pytest.raises(OSError):
		first.save_compiled(filename)
	assert first.save_compiled(filename, overwrite=True) == filename
	loaded = jitcsde(module_location=filename, n=1, control_pars=[five, four] if pars else (), **quiet)
	assert loaded.compile_attempt, "the saved library was not taken"
	assert loaded.additive == additive
	assert_bits_equal(device_result(loaded, pars=pars), expected, "loaded module")
	with pytest.raises(ValueError):
		jitcsde(module_location=filename, n=2, **quiet)


@pytest.mark.gpu
class TestIntegrationParameters:
	"""test_jitcsde.py:221-240"""

	def setup_method(self):
		self.SDE = jitcsde(f, g, **quiet)
		self.SDE.set_initial_value(initial_value, 0.0)
		self.SDE.compile_C()

	def test_min_step_error(self):
		self.SDE.set_integration_parameters(min_step=10.0)
		with pytest.raises(UnsuccessfulIntegration):
			self.SDE.integrate(1000)

	def test_rtol_error(self):
		self.SDE.set_integration_parameters(min_step=1e-3, rtol=1e-10, atol=0)
		with pytest.raises(UnsuccessfulIntegration):
			self.SDE.integrate(1000)

	def test_atol_error(self):
		self.SDE.set_integration_parameters(min_step=1e-3, rtol=0, atol=1e-10)
		with pytest.raises(UnsuccessfulIntegration):
			self.SDE.integrate(1000)


@pytest.mark.gpu
def test_the_package_self_test():
	"""test_jitcsde.py:259-263"""
	from jitcsde_b200 import test
	for sympy_flag in (False, True):
		for omp in (True, False):
			test(omp=omp, sympy=sympy_flag)


# ---- reference tests/test_jumps.py -------------------------------------------------------------------------------------------
RATE = 1000  # λ (test_jumps.py:20)


def jumpy(seed=42, **kwargs):
	"""test_jumps.py:40-44: IJI draws exponential(1/λ) waiting times from a seeded generator; the amplitude
	normal(0, 1 + 1/(1+y²)) becomes the symbolic (1 + 1/(1+y²))·ξ with ξ a standard normal variate from the same generator"""
	from jitcsde_b200 import jitcsde_jump, xi
	R = np.random.RandomState(seed)
	return jitcsde_jump(lambda time, state: R.exponential(1 / RATE), [(1 + 1 / (1 + y(0) ** 2)) * xi], f, g,
		rng=np.random.default_rng(seed), tape_length=16, **quiet, **kwargs)


def test_jump_checks():
	"""test_jumps.py:63-98"""
	from jitcsde_b200 import jitcsde_jump, xi
	with pytest.raises(NotImplementedError):
		jitcsde_jump(np.ones(4), [xi], f, g, ito=False, **quiet)
	for bad_f, bad_g in (([y(0)], [y(-1)]), ([y(0)], [y(1)]), ([y(0)], [symbols("x")])):
		with pytest.raises(ValueError):
			jitcsde_jump(np.ones(4), [xi], bad_f, bad_g, **quiet).check()
	with pytest.raises(NotImplementedError):  # "wrong amp output type": a Python callable is refused outright here
		jitcsde_jump(np.ones(4), lambda time, state: 1, f, g, **quiet)
	with pytest.raises(ValueError):  # wrong amp output size
		jitcsde_jump(np.ones(4), [xi, xi], f, g, **quiet).check()
	jumpy().check()


@pytest.mark.gpu
def test_jumps_default_and_reproducibility():
	"""test_jumps.py:27-59: two instances built the same way give the same result; it is the oracle's with the same tape"""
	results = []
	for _ in range(2):
		SDE = jumpy()
		SDE.set_seed(42)
		SDE.set_initial_value(initial_value, 0.0)
		SDE.check()
		results.append(SDE.integrate(T_END))
	assert_bits_equal(results[0], results[1], "reproducibility")
	P = port.PortCore(SDE.spec, initial_value, 0.0, 42)
	assert_bits_equal(results[0], P.integrate_jumps(T_END, SDE._taus[0], SDE._xis[0]), "vs the oracle with the same tape")
	# with λ = 1000 about one jump happens before t = 0.001: make sure the tape was actually consulted
	plain = oracle_result(MULTIPLICATIVE["default"]())
	assert SDE.last_stats["jumps"] == int(np.searchsorted(np.cumsum(SDE._taus[0]), T_END, side="right")) or not np.array_equal(results[0], plain)
