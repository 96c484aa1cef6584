"""
TEST INFRASTRUCTURE — not part of the product (see oracle/build_ref.py header).

The reference's adaptive controller and jump scheduler live in a Python class that cannot be imported here
(`jitcxde_common` and `symengine` are absent), so they are restated for the oracle.  Each method names the reference
lines it follows.  `core` is any object with the reference's core interface (reference `jitced_template.c:630-646`):
`t`, `get_next_step(h)`, `get_p(atol, rtol)`, `accept_step()`, `get_state()`, `apply_jump(change)`,
`pin_noise(number, step)`, `set_parameters(*floats)` — i.e. the compiled reference core from build_ref.load().
"""

from __future__ import annotations

import numpy as np


class UnsuccessfulIntegration(Exception):
	pass


DEFAULTS = dict(
	atol=1e-10, rtol=1e-5, first_step=1.0, min_step=1e-10, max_step=10.0, decrease_threshold=1.1,
	increase_threshold=0.5, safety_factor=0.9, max_factor=5.0, min_factor=0.2,
)


class RefIntegrator:
	"""One trajectory driven exactly the way `jitcsde.integrate` drives its core."""

	def __init__(self, module, y0, t0=0.0, seed=0, **pars):
		self.core = module.sde_integrator(float(t0), np.array(y0, dtype=np.float64), int(seed))
		self.accepted = 0
		self.rejected = 0
		self.set_integration_parameters(**pars)

	# reference `_jitcsde.py:491-566`
	def set_integration_parameters(self, **kw):
		p = dict(DEFAULTS)
		p.update(kw)
		if p["first_step"] > p["max_step"]:
			p["first_step"] = p["max_step"]
		if p["min_step"] > p["first_step"]:
			p["min_step"] = p["first_step"]
		self.atol = p["atol"]
		self.rtol = p["rtol"]
		self.dt = p["first_step"]
		self.min_step = p["min_step"]
		self.max_step = p["max_step"]
		self.decrease_threshold = p["decrease_threshold"]
		self.increase_threshold = p["increase_threshold"]
		self.safety_factor = p["safety_factor"]
		self.max_factor = p["max_factor"]
		self.min_factor = p["min_factor"]
		self.q = 1.5

	@property
	def t(self):
		return self.core.t

	# reference `_jitcsde.py:581-595` (+ `:569-579` for the min_step check)
	def _adjust_step_size(self, actual_dt):
		p = self.core.get_p(self.atol, self.rtol)
		if p > self.decrease_threshold:
			self.dt = actual_dt * max(self.safety_factor * p ** (-1 / self.q), self.min_factor)
			if self.dt < self.min_step:
				self.rejected += 1  # the failing attempt is a rejected one (bookkeeping only; the reference keeps no counters)
				raise UnsuccessfulIntegration("step size below min_step")
			return False
		if p <= self.increase_threshold:
			factor = self.safety_factor * p ** (-1 / (self.q + 1)) if p else self.max_factor
			sug_step = actual_dt * min(factor, self.max_factor)
			self.dt = max(self.dt, min(sug_step, self.max_step))
		return True

	# reference `_jitcsde.py:597-630`
	def integrate(self, target_time):
		core = self.core
		last_step = core.t >= target_time
		while not last_step:
			if core.t + self.dt < target_time:
				actual_dt = self.dt
			else:
				actual_dt = target_time - core.t
				last_step = True
			core.get_next_step(actual_dt)
			if self._adjust_step_size(actual_dt):
				core.accept_step()
				self.accepted += 1
			else:
				last_step = False
				self.rejected += 1
		return core.get_state()


class RefJumpIntegrator(RefIntegrator):
	"""reference `_jitcsde.py:656-700`; `IJI(time,state)` and `amp(time,state)` are caller-supplied closures."""

	def __init__(self, IJI, amp, *args, **kwargs):
		super().__init__(*args, **kwargs)
		self.IJI = IJI
		self.amp = amp
		self._next_jump = None

	@property
	def next_jump(self):
		if self._next_jump is None:
			self._next_jump = self.core.t + self.IJI(self.core.t, self.core.get_state())
		return self._next_jump

	def integrate(self, target_time):
		while self.next_jump < target_time:
			state = super().integrate(self.next_jump)
			time = self.core.t
			self.core.apply_jump(self.amp(time, state))
			self._next_jump = self.core.t + self.IJI(time, self.core.get_state())
		return super().integrate(target_time)


def tape_closures(taus, xis, amp_fn, iji_fn=None):
	"""
	Feed a per-trajectory jump tape (waiting times `taus`, standard-normal variates `xis`) to the reference's
	callback interface: the k-th call of IJI returns taus[k]; the k-th call of amp evaluates `amp_fn(t, y, xis[k])`.
	Past the end of the tape the waiting time is +inf (no further jumps).  `iji_fn(t, y, tau)` turns the tape's value into
	the waiting time (default: the value itself) — the state-dependent intensities of `JumpSpec.iji`.
	"""
	state = {"i": 0, "j": 0}

	def IJI(time, y):
		k = state["i"]
		state["i"] += 1
		if k >= len(taus):
			return float("inf")
		return float(taus[k]) if iji_fn is None else float(iji_fn(time, y, float(taus[k])))

	def amp(time, y):
		k = state["j"]
		state["j"] += 1
		return np.asarray(amp_fn(time, y, float(xis[k])), dtype=np.float64)

	return IJI, amp
