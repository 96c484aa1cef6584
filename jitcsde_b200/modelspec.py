"""
Model description shared by the CUDA code generator and (as plain data) by the test oracle.

A `ModelSpec` is the *printed* form of an SDE  dy = f(t,y) dt + g(t,y) ⊙ dW : one C expression string per
component, written against the same tiny macro vocabulary the reference's generated `f.c`/`g.c` use
(reference `jitcsde/jitced_template.c:351-359`):

* ``y(i)``            – i-th state component,
* ``t``               – time,
* ``par(name)`` is **not** used; control parameters appear by bare name (``sigma``) and are bound by the
  consumer (CUDA: per-model constant/array; oracle: ``self->parameter_sigma`` via ``#define``).

Feeding the *same strings* to nvcc (with ``-fmad=false``) and to gcc (``-ffp-contract=off``) is what makes
bit-for-bit parity possible for polynomial/rational models.
"""

from __future__ import annotations

import dataclasses
import hashlib
import json
from dataclasses import dataclass, field


@dataclass(frozen=True)
class JumpSpec:
	"""
	Declarative replacement for the reference's Python callbacks ``IJI(time,state)`` / ``amp(time,state)``
	(reference `jitcsde/_jitcsde.py:656-700`).  Waiting times and one standard-normal variate per jump come from a
	per-trajectory *tape*; ``amp`` holds one C expression per component in ``y(i)``, ``t`` and ``xi`` (the tape's
	variate for this jump).  The expression is evaluated on the *pre-jump* state, like `_jitcsde.py:697`.

	``iji`` is the reference's ``IJI(time, state)`` (`_jitcsde.py:682-687`: evaluated when the next jump is scheduled, on
	the state right after the previous one): a C expression in ``y(i)``, ``t`` and ``tau``, the tape's value for this
	jump.  The default ``"tau"`` makes the tape hold the waiting times themselves; ``"tau/(1.0+y(2)*y(2))"`` with unit
	exponentials on the tape gives a state-dependent intensity (SURVEY.md §8 f item 3).
	"""
	amp: tuple[str, ...]
	iji: str = "tau"


@dataclass(frozen=True)
class ModelSpec:
	name: str
	n: int
	additive: bool
	f: tuple[str, ...]            # drift components (NOT yet multiplied by h; consumers apply `*h` like template:351)
	g: tuple[str, ...]            # diffusion components
	control_pars: tuple[str, ...] = ()
	jump: JumpSpec | None = None
	# Optional verbatim C/CUDA code placed before the component expressions (e.g. static coupling tables and
	# loop-style drift for the large-n network).  When `f_body`/`g_body` is given it replaces the per-component
	# statements and must call set_drift(i, expr) / set_diffusion(i, expr) itself.
	preamble: str = ""
	f_body: str = ""
	g_body: str = ""
	# "Wide" models (large n, e.g. the 1000-dimensional network of SURVEY §8 d config 5): instead of n printed
	# expressions, ONE expression in the component index `i` (and y(j), t, parameters, helpers from `preamble`).
	# The GPU runs one thread block per trajectory with the components spread over the threads; the oracle loops
	# over i.  `preamble` may use JSDE_MODEL_FN (function qualifier) and JSDE_MODEL_CONST (table qualifier).
	f_component: str = ""
	g_component: str = ""
	# Dense linear coupling applied across the ensemble batch: `f_component` may use the word `coupling`, which stands
	# for  sum_j TABLE[j*units + i] * (y(j) - y(i))  (j ascending, separately rounded) when i < coupling_units and for
	# 0.0 otherwise; TABLE (the transposed coupling matrix) is defined in `preamble`.  The GPU evaluates the sums of
	# all trajectories of a thread block together (csrc/wide.cuh); the oracle's C gets a plain loop.
	coupling_units: int = 0
	coupling_table: str = ""

	@property
	def wide(self) -> bool:
		return bool(self.f_component)

	def __post_init__(self):
		if self.wide:
			assert self.g_component, "wide models need g_component too"
			assert self.jump is None or len(self.jump.amp) == 1, "wide models: ONE jump-amplitude expression in the component index i"
			if self.coupling_units:
				assert self.coupling_table and 0 < self.coupling_units <= self.n
				u, tab = self.coupling_units, self.coupling_table
				object.__setattr__(self, "preamble", self.preamble + f"""
#ifndef JSDE_DEVICE_COUPLING  /* host/oracle: the coupling term as a plain loop; the CUDA path batches it over trajectories */
JSDE_MODEL_FN double jsde_coupling_sum(int i, const double *Y)
{{
	if (i >= {u}) return 0.0;
	double s = 0.0;
	for (int j = 0; j < {u}; j++)
		s += {tab}[j * {u} + i] * (Y[j] - Y[i]);
	return s;
}}
#define coupling jsde_coupling_sum(i, Y)
#endif
""")
			object.__setattr__(self, "f_body", f"for (int i = 0; i < {self.n}; i++) set_drift(i, {self.f_component});\n")
			object.__setattr__(self, "g_body", f"for (int i = 0; i < {self.n}; i++) set_diffusion(i, {self.g_component});\n")
		if not self.f_body:
			assert len(self.f) == self.n, "len(f) != n"
		if not self.g_body:
			assert len(self.g) == self.n, "len(g) != n"

	def to_json(self) -> str:
		"""every field as stored (for wide models: after __post_init__ has completed preamble and bodies)"""
		return json.dumps(dataclasses.asdict(self))

	@classmethod
	def from_json(cls, text: str) -> "ModelSpec":
		"""inverse of to_json; does not run __post_init__ again (it would append the wide-model helpers twice)"""
		d = json.loads(text)
		jump = d.pop("jump")
		self = object.__new__(cls)
		for key, value in d.items():
			object.__setattr__(self, key, tuple(value) if isinstance(value, list) else value)
		object.__setattr__(self, "jump", JumpSpec(tuple(jump["amp"]), jump["iji"]) if jump else None)
		return self

	def digest(self) -> str:
		h = hashlib.sha256()
		for part in (self.name, str(self.n), str(self.additive), *self.f, "|", *self.g, "|", *self.control_pars,
				"|", *(self.jump.amp if self.jump else ()), "|", (self.jump.iji if self.jump and self.jump.iji != "tau" else ""), "|", self.preamble, self.f_body, self.g_body, self.f_component, self.g_component,
				str(self.coupling_units), self.coupling_table):
			h.update(part.encode())
			h.update(b"\0")
		return h.hexdigest()[:16]
