"""
Symbolic front end for LARGE systems: n SymPy expressions → a *wide* `ModelSpec` (one expression in the component index
`i`, evaluated with the components spread over a thread block, csrc/wide.cuh) instead of n printed expressions that a
single GPU lane would have to keep in registers.

The reference prints every component's expression separately, whatever n is (`jitcsde/_jitcsde.py:408-421`; its
examples build networks with Python loops and rely on chunking + CSE in jitcxde_common for compile time).  Here the n
expressions are folded back into a few *templates*:

1. **dense linear coupling** (north star: "a drift term that is a dense linear coupling matrix applied across the
   ensemble batch").  Every f_i is expanded and its terms c_ij·y(j) with numeric c_ij are collected.  If the coupled
   components are the leading block i, j < units and the block is dense enough, the terms are removed from f_i and become
   the kernel's batched coupling sum  Σ_j C_ij (y_j − y_i)  (ascending j, three roundings per term) plus the local term
   (Σ_j C_ij)·y_i — the same value as Σ_j C_ij y_j, in the association the kernel and the oracle's loop both use.
2. **structural templates**.  In what is left, y(k) is rewritten as y(i + (k − i)); components whose expressions then
   coincide share one template, selected by index range or by a group table.  A FitzHugh–Nagumo network has two (the
   v- and the w-equations), a ring of identical units one.

Anything that does not fold (more than `MAX_TEMPLATES` distinct templates, symbolic coupling weights, a dense block that
is not leading) raises NotImplementedError with the reason — there is no silent fallback to a kernel shape that cannot
hold the model.
"""

from __future__ import annotations

import hashlib

import numpy as np
import sympy

from ._codegen import _Printer, t, y  # noqa: F401
from .modelspec import JumpSpec, ModelSpec

#: above this many components the lane-per-trajectory kernels stop making sense (state + stages in registers)
WIDE_THRESHOLD = 32
MAX_TEMPLATES = 48
#: a leading block counts as densely coupled when its rows hold at least this fraction of the possible linear terms
DENSE_FRACTION = 0.25

_i = sympy.Symbol("i", integer=True)


class _WidePrinter(_Printer):
	def _print_Function(self, expr):
		if expr.func.__name__ == "y":
			return f"y({self._print(expr.args[0])})"
		if expr.func.__name__ == "jsde_tab":  # jsde_tab(k) -> per-component constant table k, entry i
			return f"JSDE_SYM_TAB{int(expr.args[0])}[i]"
		return super()._print_Function(expr)


_printer = _WidePrinter()
_tab = sympy.Function("jsde_tab")


def _is_state(e) -> bool:
	return isinstance(e, sympy.Function) and e.func.__name__ == "y" and len(e.args) == 1 and e.args[0].is_Integer


def extract_linear_coupling(f, n):
	"""-> (units, C [units][units] float64, remainders) or (0, None, f).  See the module docstring, item 1."""
	rows = []
	for e in f:
		terms = sympy.expand(e).as_coefficients_dict()
		lin = {int(k.args[0]): v for k, v in terms.items() if _is_state(k) and v.is_number}
		rows.append((terms, lin))
	involved = [i for i, (_, lin) in enumerate(rows) if len(lin) >= 4]
	if not involved:
		return 0, None, list(f)
	# columns that many rows refer to form the dense block; linear terms outside it (a unit's own recovery variable, say)
	# stay local
	from collections import Counter
	col_count = Counter(j for i in involved for j in rows[i][1])
	dense_cols = {j for j, c in col_count.items() if c >= max(4, len(involved) // 4)}
	dense_rows = [i for i in involved if len(dense_cols.intersection(rows[i][1])) >= 4]
	if not dense_rows:
		return 0, None, list(f)
	units = max(max(dense_rows), max(dense_cols)) + 1
	filled = sum(len(dense_cols.intersection(rows[i][1])) for i in dense_rows)
	if filled < DENSE_FRACTION * units * units:
		return 0, None, list(f)
	C = np.zeros((units, units))
	rest = list(f)
	for i in dense_rows:
		terms, lin = rows[i]
		for j, v in lin.items():
			if j in dense_cols:
				C[i, j] = float(v)
		rest[i] = sympy.Add(*[v * k for k, v in terms.items() if not (_is_state(k) and v.is_number and int(k.args[0]) in dense_cols)])
	return units, C, rest


def fold_templates(exprs, what):
	"""-> (C expression in `i`, preamble text with the group table if one is needed).  Module docstring, item 2."""
	n = len(exprs)
	groups: dict[str, list[int]] = {}
	templates: dict[str, sympy.Expr] = {}
	for i, e in enumerate(exprs):
		e = sympy.sympify(e)
		rel = e.replace(_is_state, lambda s, i=i: y(_i + (int(s.args[0]) - i)))
		key = sympy.srepr(rel)
		groups.setdefault(key, []).append(i)
		templates[key] = rel
	if len(groups) > MAX_TEMPLATES:
		raise NotImplementedError(f"{what}: the {n} components fold into {len(groups)} structural templates (limit {MAX_TEMPLATES}); "
			"wide models need components that share their expressions up to an index shift")
	keys = sorted(groups, key=lambda k: groups[k][0])
	code = {k: _printer.doprint(templates[k]) for k in keys}
	if len(keys) == 1:
		return f"({code[keys[0]]})", ""
	contiguous = all(groups[k] == list(range(groups[k][0], groups[k][-1] + 1)) for k in keys)
	if contiguous:
		out = f"({code[keys[-1]]})"
		for k in reversed(keys[:-1]):
			out = f"(i < {groups[k][-1] + 1} ? ({code[k]}) : {out})"
		return out, ""
	table = [0] * n
	for g, k in enumerate(keys):
		for i in groups[k]:
			table[i] = g
	tag = hashlib.sha256((what + "".join(map(str, table))).encode()).hexdigest()[:8]
	name = f"JSDE_SYM_GROUP_{tag}"
	pre = f"JSDE_MODEL_CONST unsigned char {name}[{n}] = {{{','.join(map(str, table))}}};\n"
	out = f"({code[keys[-1]]})"
	for g, k in reversed(list(enumerate(keys[:-1]))):
		out = f"({name}[i] == {g} ? ({code[k]}) : {out})"
	return out, pre


def wide_spec(name, n, additive, f, g, pars, jump_amp=None, jump_iji=None) -> ModelSpec:
	"""n SymPy expressions each for drift and diffusion (helpers substituted, Itō form) → wide ModelSpec"""
	preamble = """
#ifndef JSDE_MODEL_FN
#define JSDE_MODEL_FN static inline
#define JSDE_MODEL_CONST static const
#endif
"""
	units, C, rest = extract_linear_coupling(f, n)
	coupling_table = ""
	if units:
		rowsum = np.zeros(n)
		rowsum[:units] = C.sum(axis=1)
		# transposed, like models.fhn_network: threads owning neighbouring components read neighbouring entries
		preamble += f"JSDE_MODEL_CONST double JSDE_SYM_COUPLING[{units * units}] = {{{','.join(float(v).hex() for v in C.T.ravel())}}};\n"
		preamble += f"JSDE_MODEL_CONST double JSDE_SYM_TAB0[{n}] = {{{','.join(float(v).hex() for v in rowsum)}}};\n"
		coupling_table = "JSDE_SYM_COUPLING"
		rest = [e + _tab(0) * y(i) if i < units else e for i, e in enumerate(rest)]
	f_code, pre_f = fold_templates(rest, "f_sym")
	if units:
		f_code = f"({f_code} + coupling)"
	g_code, pre_g = fold_templates(g, "g_sym")
	jump = None
	if jump_amp is not None:
		amp_code, pre_j = fold_templates(jump_amp, "amp")
		preamble += pre_j
		jump = JumpSpec((amp_code,), iji="tau" if jump_iji is None else _printer.doprint(sympy.sympify(jump_iji)))
	preamble += pre_f + pre_g
	if name is None:
		name = "wide_" + hashlib.sha256("|".join((f_code, g_code, preamble, *pars, jump.amp[0] if jump else "", jump.iji if jump else "")).encode()).hexdigest()[:12]
	return ModelSpec(name, n, bool(additive), (), (), control_pars=tuple(pars), jump=jump, preamble=preamble,
		f_component=f_code, g_component=g_code, coupling_units=units, coupling_table=coupling_table)
