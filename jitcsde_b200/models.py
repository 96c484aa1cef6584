"""
Ready-printed `ModelSpec`s for the workloads BASELINE.json names (SURVEY.md §8 d) and for the parity tests.
The expression strings are written the way the reference's examples write them
(`examples/noisy_lorenz.py:75-87`, `tests/compare_C_and_Python_core.py:36-46`) so that the oracle, which
compiles the very same strings into the reference C core, sees the same operation order.
"""

from .modelspec import JumpSpec, ModelSpec

_LORENZ_F = (
	"10*(y(1)-y(0))",
	"y(0)*(28-y(2))-y(1)",
	"y(0)*y(1)-(8.0/3.0)*y(2)",
)

#: headline workload (BASELINE.json configs[1]): additive noise ⇒ SRA
lorenz_additive = ModelSpec("lorenz_additive", 3, True, _LORENZ_F, ("0.1", "0.1", "0.1"))

#: the reference's own additive fixture, `tests/compare_C_and_Python_core.py:44` (G = i*0.05)
lorenz_additive_fixture = ModelSpec("lorenz_additive_fixture", 3, True, _LORENZ_F, ("0.0", "0.05", "0.1"))

#: physics of `examples/noisy_lorenz.py:75-87` (multiplicative ⇒ SRI)
lorenz_multiplicative = ModelSpec("lorenz_multiplicative", 3, False, _LORENZ_F, ("0.1*y(0)", "0.1*y(1)", "0.1*y(2)"))

#: `examples/noisy_and_jumpy_lorenz.py:36-58`: same SDE plus jumps on component 2 with amplitude |y2|·ξ
lorenz_jumpy = ModelSpec(
	"lorenz_jumpy", 3, False, _LORENZ_F, ("0.1*y(0)", "0.1*y(1)", "0.1*y(2)"),
	jump=JumpSpec(("0.0", "0.0", "0.0 + fabs(y(2))*xi")),
)

#: the same with a state-dependent jump intensity: the tape holds unit exponentials, the waiting time is E/(1 + y2^2/400)
#: (reference `IJI(time, state)`, `_jitcsde.py:682-687`, evaluated when the jump is scheduled)
lorenz_jumpy_rate = ModelSpec(
	"lorenz_jumpy_rate", 3, False, _LORENZ_F, ("0.1*y(0)", "0.1*y(1)", "0.1*y(2)"),
	jump=JumpSpec(("0.0", "0.0", "0.0 + fabs(y(2))*xi"), iji="tau/(1.0+y(2)*y(2)/400.0)+0.001*t"),
)

#: geometric Brownian motion, closed-form moments (SURVEY §8 d config 3 i)
gbm = ModelSpec("gbm", 1, False, ("0.05*y(0)",), ("0.2*y(0)",))

#: GBM with control parameters (exercises set_parameters, reference `_jitcsde.py:465-484`)
gbm_pars = ModelSpec("gbm_pars", 1, False, ("mu*y(0)",), ("sigma*y(0)",), control_pars=("mu", "sigma"))

#: stochastic Duffing oscillator (SURVEY §8 d config 3 ii): δ=0.3 α=-1 β=1 σ=0.3
duffing = ModelSpec(
	"duffing", 2, False,
	("y(1)", "-0.3*y(1)-(-1.0)*y(0)-1.0*(y(0)*y(0)*y(0))"),
	("0.0", "0.3*y(0)"),
)

#: Ornstein–Uhlenbeck pair: contracting, so per-trajectory parity is meaningful at long horizons
ou = ModelSpec("ou", 2, True, ("-1.0*y(0)", "-0.5*(y(1)-1.0)"), ("0.3", "0.2"))

#: time-dependent additive noise (exercises the g(t+h) / g(t) evaluation order of `jitced_template.c:476,485`)
ou_timedep = ModelSpec("ou_timedep", 2, True, ("-1.0*y(0)+t", "-0.5*(y(1)-1.0)"), ("0.3*(1.0+t)", "0.2/(1.0+t*t)"))

#: six coupled components, additive: the lane kernels beyond the 1-3 components of the benchmark configurations (register
#: and shared-memory budgets scale with n)
ring6 = ModelSpec("ring6", 6, True, tuple(f"-1.0*y({i})+0.5*(y({(i + 1) % 6})-y({(i + 5) % 6}))" for i in range(6)),
	("0.1", "0.2", "0.3", "0.1", "0.2", "0.3"))

#: the same ring with multiplicative noise (SRI with six components)
ring6_mult = ModelSpec("ring6_mult", 6, False, tuple(f"-1.0*y({i})+0.5*(y({(i + 1) % 6})-y({(i + 5) % 6}))" for i in range(6)),
	tuple(f"0.1*y({i})+0.05" for i in range(6)))

#: scalar validation SDE of `tests/validation.py:21-32` without the exponential (polynomial part only)
cubic = ModelSpec("cubic", 1, True, ("-(y(0)*y(0)*y(0))+4*y(0)+y(0)*y(0)",), ("3.0",))

#: the reference's own validation scenario in full (`tests/validation.py:21-32`): F = -y^3 + 4y + y^2,
#: G = 5 exp(-y^2 + y - 1.5) + 3 — multiplicative noise through a transcendental function (exp is evaluated with glibc's
#: algorithm on the device, csrc/glibc_math.h, so this model is bit-exact too)
validation_exp = ModelSpec("validation_exp", 1, False, ("-(y(0)*y(0)*y(0))+4*y(0)+y(0)*y(0)",), ("5*exp(-(y(0)*y(0))+y(0)-1.5)+3.0",))

#: noisy damped pendulum: sin/cos come from the CUDA math library on the device and from glibc in the reference's C, which
#: differ by an ulp now and then — the TOLERANCE tier of the parity ladder (SURVEY.md §8 c rungs 2 and 4)
pendulum = ModelSpec("pendulum", 2, False, ("y(1)", "-sin(y(0))-0.1*y(1)"), ("0.0", "0.2+0.1*cos(y(0))"))


def fhn_network(units: int, name: str, seed: int = 7, multiplicative: bool = False, jumpy: bool = False, pars: bool = False) -> ModelSpec:
	"""
	Noisy FitzHugh–Nagumo network with dense coupling (SURVEY §8 d config 5; units=500 gives n=1000):
	    dv_i = v_i - v_i^3/3 - w_i + I + (K/N) sum_j A_ij (v_j - v_i),    dw_i = eps (v_i + a - b w_i),
	additive noise sigma on v only.  y(i) = v_i for i < units, y(units+i) = w_i.  A = default_rng(seed).random, stored
	transposed so that threads owning neighbouring components read neighbouring table entries; the sum runs over j in
	ascending order on both the device and the oracle (same rounding).
	"""
	import numpy as np
	A = np.random.default_rng(seed).random((units, units))
	table = ",".join(float(v).hex() for v in A.T.ravel())
	preamble = f"""
#ifndef JSDE_MODEL_FN
#define JSDE_MODEL_FN static inline
#define JSDE_MODEL_CONST static const
#endif
#define FHN_UNITS {units}
JSDE_MODEL_CONST double FHN_AT[FHN_UNITS * FHN_UNITS] = {{{table}}};
"""
	f_component = ("(i < FHN_UNITS ? y(i) - y(i) * y(i) * y(i) / 3.0 - y(FHN_UNITS + i) + 0.5 + (1.0 / FHN_UNITS) * coupling"
		" : 0.08 * (y(i - FHN_UNITS) + 0.7 - 0.8 * y(i)))")
	if pars:  # control parameters (shared or one set per trajectory): input current, coupling strength, noise amplitude
		f_component = f_component.replace("+ 0.5 +", "+ I_ext +").replace("(1.0 / FHN_UNITS)", "(K_c / FHN_UNITS)")
	# multiplicative variant (exercises the wide SRI path): state-dependent amplitudes, one of them reaching across
	# components — the diffusion of w_i depends on v_i
	g_component = ("(i < FHN_UNITS ? 0.05 * y(i) + 0.02 : 0.01 * y(i - FHN_UNITS) + 0.01 * y(i))" if multiplicative
		else "(i < FHN_UNITS ? 0.05 : 0.0)")
	if pars:
		g_component = g_component.replace("0.05", "sigma")
	# jumpy variant (exercises jumps on the wide path): kicks on the v components, intensity growing with v_0^2
	jump = JumpSpec(("(i < FHN_UNITS ? 0.2 * xi * (1.0 + y(i)) : 0.0)",), iji="tau/(2.0+y(0)*y(0))") if jumpy else None
	return ModelSpec(name, 2 * units, not multiplicative, (), (), preamble=preamble,
		f_component=f_component, g_component=g_component, jump=jump, control_pars=("I_ext", "K_c", "sigma") if pars else (),
		coupling_units=units, coupling_table="FHN_AT")


#: small network for the parity tests of the wide (block-per-trajectory) kernel
fhn_small = fhn_network(24, "fhn_small")
fhn_small_mult = fhn_network(24, "fhn_small_mult", multiplicative=True)
fhn_small_jumpy = fhn_network(24, "fhn_small_jumpy", jumpy=True)
fhn_small_pars = fhn_network(24, "fhn_small_pars", pars=True)
_FHN_1000 = None


def fhn_1000() -> ModelSpec:
	"""n = 1000 (BASELINE.json configs[4]); built lazily: its coupling table is a 6 MB literal"""
	global _FHN_1000
	if _FHN_1000 is None:
		_FHN_1000 = fhn_network(500, "fhn_1000")
	return _FHN_1000


ALL = {m.name: m for m in (
	lorenz_additive, lorenz_additive_fixture, lorenz_multiplicative, lorenz_jumpy, lorenz_jumpy_rate,
	gbm, gbm_pars, duffing, ou, ou_timedep, cubic, ring6, ring6_mult, validation_exp, pendulum, fhn_small, fhn_small_mult, fhn_small_jumpy, fhn_small_pars,
)}
