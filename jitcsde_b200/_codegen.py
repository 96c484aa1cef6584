"""
Symbolic front end of the code generator: SymPy expressions in `y(i)`, `t`, control parameters and helpers → the C
expression strings of a `ModelSpec`.

Counterpart of the symbolic half of the reference's `compile_C` (reference `jitcsde/_jitcsde.py:355-421`) and of
`_determine_additivity` (`:117-124`) and `_stratonovich_to_ito` (`:126-150`).  The reference prints with SymEngine's C
printer; SymEngine is not available here, so SymPy's `ccode` is used with two adjustments that matter for parity:
small integer powers are printed as products (gcc calls libm `pow` for `pow(x, 3)`, nvcc inlines its own — products
are exact on both sides), and `y(i)` stays a function call so the reference's macro vocabulary applies unchanged.
Helpers are substituted into the expressions (the reference evaluates them into local arrays first,
`jitced_template.c:373-376`; same values, different association only if a helper is itself re-associated by SymPy).
"""

from __future__ import annotations

import sympy
from sympy.printing.c import C99CodePrinter

from .modelspec import JumpSpec, ModelSpec

#: the SymPy flavour of the reference's dynamical variable and time (reference `jitcsde/sympy_symbols.py:5,8`)
y = sympy.Function("y", real=True)
t = sympy.Symbol("t", real=True)
#: the per-jump standard-normal variate available to jump-amplitude expressions (see `JumpSpec`)
xi = sympy.Symbol("xi", real=True)
#: the tape's value for the jump being scheduled, available to waiting-time expressions (see `JumpSpec.iji`)
tau = sympy.Symbol("tau", real=True)


class _Printer(C99CodePrinter):
	def _print_Pow(self, expr):
		base, exp = expr.as_base_exp()
		if exp.is_Integer and 2 <= int(exp) <= 4:
			b = self._print(base)
			return "(" + "*".join([f"({b})"] * int(exp)) + ")"
		if exp.is_Integer and -4 <= int(exp) <= -1:
			b = self._print(base)
			return "(1.0/(" + "*".join([f"({b})"] * (-int(exp))) + "))"
		return super()._print_Pow(expr)

	def _print_Function(self, expr):
		if expr.func.__name__ == "y":
			return f"y({self._print(expr.args[0])})"
		return super()._print_Function(expr)

	def _print_Integer(self, expr):
		return str(int(expr))


_printer = _Printer()


def ccode(expr) -> str:
	return _printer.doprint(sympy.sympify(expr))


def _handle_input(source, n_hint=None):
	"""Accept the input styles of the reference (`jitcxde._handle_input`): iterable, generator function, dict."""
	if callable(source) and not isinstance(source, sympy.Basic):
		source = list(source())
	if isinstance(source, dict):
		n = len(source)
		out = [None] * n
		for key, value in source.items():
			key = sympy.sympify(key)
			if not (key.func.__name__ == "y" and len(key.args) == 1 and key.args[0].is_Integer):
				raise ValueError("Dictionary keys must be dynamical variables y(0), y(1), …")
			i = int(key.args[0])
			if not 0 <= i < n:
				raise ValueError("Dictionary keys must be y(0) … y(n-1).")
			out[i] = value
		if any(v is None for v in out):
			raise ValueError("Dictionary does not define every component.")
		source = out
	return [_to_sympy(e) for e in source]


def _to_sympy(expr):
	"""sympify, mapping any function called `y` onto our `y` (so symbols from elsewhere — e.g. a user's own
	sympy.Function("y") — work, like reference tests/test_sympy_input.py)."""
	e = sympy.sympify(expr)
	repl = {f: y(*f.args) for f in e.atoms(sympy.Function) if f.func.__name__ == "y" and f.func != y}
	return e.xreplace(repl) if repl else e


def _substitute_helpers(exprs, helpers):
	if not helpers:
		return exprs
	helpers = [(sympy.sympify(sym), _to_sympy(val)) for sym, val in helpers]
	out = []
	for e in exprs:
		# helpers may depend on earlier helpers: substitute until nothing changes (bounded by the helper count)
		for _ in range(len(helpers) + 1):
			new = e
			for sym, val in reversed(helpers):
				new = new.subs(sym, val)
			if new == e:
				break
			e = new
		out.append(e)
	return out


def has_state(expr) -> bool:
	return any(f.func.__name__ == "y" for f in sympy.sympify(expr).atoms(sympy.Function))


def check_expressions(f, g, n, control_pars, extra_symbols=()):
	"""The reference's `check()` validators (`_jitcsde.py:211-240`): component indices and foreign symbols."""
	problems = []
	valid = {t, *map(sympy.sympify, control_pars), *extra_symbols}
	for name, exprs in (("f_sym", f), ("g_sym", g)):
		if not exprs:
			problems.append(f"{name} is empty.")
		for i, e in enumerate(exprs):
			for fn in e.atoms(sympy.Function):
				if fn.func.__name__ == "y":
					arg = fn.args[0]
					if not arg.is_Integer:
						problems.append(f"y is called with a non-integer argument ({arg}) in component {i} of {name}.")
					elif arg < 0:
						problems.append(f"y is called with a negative argument ({arg}) in component {i} of {name}.")
					elif arg >= n:
						problems.append(f"y is called with an argument ({arg}) higher than the system’s dimension ({n}) in component {i} of {name}.")
			for sym in e.atoms(sympy.Symbol):
				if sym not in valid and sym.name not in {s.name for s in valid}:
					problems.append(f"Invalid symbol ({sym.name}) in component {i} of {name}.")
	return problems


def stratonovich_to_ito(f, g):
	"""f_i += g_i * dg_i/dy_i / 2 (diagonal noise; reference `_jitcsde.py:142-147`, helpers already substituted)."""
	return [fi + gi * sympy.diff(gi, y(i)) / 2 for i, (fi, gi) in enumerate(zip(f, g))]


def _cse_body(kind, exprs):
	"""Common-subexpression elimination (the reference's `do_cse`, `_jitcsde.py:366-375`): temporaries first, then one
	`set_<kind>(i, expr);` per component — the same text goes to nvcc and to the oracle's gcc, so parity is unaffected."""
	replacements, reduced = sympy.cse(list(exprs), symbols=sympy.numbered_symbols(f"cse_{kind[0]}"), order="none")
	lines = [f"const double {sym} = {ccode(val)};\n" for sym, val in replacements]
	lines += [f"set_{kind}({i}, {ccode(e)});\n" for i, e in enumerate(reduced)]
	return "".join("\t\t" + line for line in lines)


def spec_from_symbolic(f_sym, g_sym, *, n=None, additive=None, ito=True, control_pars=(), helpers=None, g_helpers="auto",
		jump_amp=None, jump_iji=None, name=None, simplify=None, do_cse=False, wide=None) -> ModelSpec:
	f = _handle_input(f_sym)
	g = _handle_input(g_sym)
	if n is None:
		n = len(f)
	if len(f) != n or len(g) != n:
		raise ValueError("f_sym and g_sym must both have n components.")
	f_helpers = list(helpers or [])
	if g_helpers in ("auto", "same"):
		gh = f_helpers
	else:
		gh = list(g_helpers or [])
	f = _substitute_helpers(f, f_helpers)
	g = _substitute_helpers(g, gh)
	if additive is None:
		additive = all(not has_state(e) for e in g)
	if not ito and not additive:
		f = stratonovich_to_ito(f, g)
	if simplify is None:
		simplify = False  # the reference simplifies iff n<=10 (:350); skipped by default to keep the user's operation order
	if simplify:
		f = [sympy.simplify(e) for e in f]
		g = [sympy.simplify(e) for e in g]
	pars = tuple(str(p) for p in control_pars)
	for p in pars:
		if p in ("t", "h", "y", "Y", "out", "par", "xi", "tau") or not p.isidentifier():
			raise ValueError(f"control parameter name {p!r} is not usable in generated code")
	jump = None
	if jump_amp is not None:
		amp = [_to_sympy(e) for e in jump_amp]
		if len(amp) != n:
			raise ValueError("Output of amp function has the wrong dimension")
		jump = JumpSpec(tuple(ccode(e) for e in amp), iji="tau" if jump_iji is None else ccode(_to_sympy(jump_iji)))
	if wide is None:
		from ._codegen_wide import WIDE_THRESHOLD
		wide = n > WIDE_THRESHOLD
	if wide:
		# large systems: one expression in the component index, evaluated by a thread block per group of trajectories
		from ._codegen_wide import wide_spec
		amp = None if jump_amp is None else [_to_sympy(e) for e in jump_amp]
		return wide_spec(name, n, additive, f, g, pars, jump_amp=amp, jump_iji=None if jump_iji is None else _to_sympy(jump_iji)), f, g
	fs = tuple(ccode(e) for e in f)
	gs = tuple(ccode(e) for e in g)
	f_body = _cse_body("drift", f) if do_cse else ""
	g_body = _cse_body("diffusion", g) if do_cse else ""
	if name is None:
		import hashlib
		name = "sde_" + hashlib.sha256("|".join((*fs, "#", *gs, "#", *pars, "#", *(jump.amp if jump else ()), "#", (jump.iji if jump else ""),
			"#cse" if do_cse else "")).encode()).hexdigest()[:12]
	return ModelSpec(name, n, bool(additive), fs, gs, control_pars=pars, jump=jump, f_body=f_body, g_body=g_body), f, g
