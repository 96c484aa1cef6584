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

import hashlib
from dataclasses import dataclass, field


@dataclass(frozen=True)
class JumpSpec:
	"""
	Declarative replacement for the reference's Python callbacks ``IJI(time,state)`` / ``amp(time,state)``
	(reference `jitcsde/_jitcsde.py:656-700`).  Waiting times and one standard-normal variate per jump come from a
	per-trajectory *tape*; ``amp`` holds one C expression per component in ``y(i)``, ``t`` and ``xi`` (the tape's
	variate for this jump).  The expression is evaluated on the *pre-jump* state, like `_jitcsde.py:697`.
	"""
	amp: tuple[str, ...]


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

	def __post_init__(self):
		if not self.f_body:
			assert len(self.f) == self.n, "len(f) != n"
		if not self.g_body:
			assert len(self.g) == self.n, "len(g) != n"

	def digest(self) -> str:
		h = hashlib.sha256()
		for part in (self.name, str(self.n), str(self.additive), *self.f, "|", *self.g, "|", *self.control_pars,
				"|", *(self.jump.amp if self.jump else ()), "|", self.preamble, self.f_body, self.g_body):
			h.update(part.encode())
			h.update(b"\0")
		return h.hexdigest()[:16]
