"""
Python front end: the reference's public names (`jitcsde`, `jitcsde_jump`, `y`, `t`, `UnsuccessfulIntegration`;
reference `jitcsde/__init__.py:1`, `jitcsde/_jitcsde.py:27-722`) over the B200 ensemble engine.

What is different from the reference, and why:
* every instance integrates an **ensemble** of `ensemble` independent trajectories (default 1).  With `ensemble=1`
  results have the reference's shapes (`integrate` returns an array of length n); otherwise a leading axis of length B
  is added.  Trajectory k uses seed `seed + k` (mod 2³²) unless an explicit array of seeds is given, and the initial
  value may be one state (broadcast) or one per trajectory.
* the adaptive loop (`_jitcsde.py:597-630`) does not run in Python: `integrate` is one kernel launch in which every
  lane runs the same controller to `target_time` (csrc/kernels.cuh).
* `compile_C` builds a CUDA library with nvcc instead of a C extension; there is no Python/CPU fallback core
  (`generate_lambdas` raises), and Python callbacks inside f/g cannot run on a GPU lane (`callback_functions` raises).
* jumps (`jitcsde_jump`) are declarative: waiting times and one normal variate per jump come from a per-trajectory
  tape, and the amplitude is a symbolic expression in `y(i)`, `t` and `xi` — see `jitcsde_jump`.
"""

from __future__ import annotations

import os
import random
from warnings import warn

import numpy as np

from . import _abi, _build
from ._abi import ControllerParameters, Ensemble
from ._codegen import check_expressions, spec_from_symbolic, t, tau, xi, y  # noqa: F401  (re-exported)
from .modelspec import ModelSpec


class UnsuccessfulIntegration(Exception):
	"""
	Raised when the integrator cannot meet the accuracy and step-size requirements (reference `_jitcsde.py:21-25`).
	`indices` lists the failing trajectories; all others have been integrated to the target time.
	"""

	def __init__(self, message, indices=()):
		super().__init__(message)
		self.indices = np.asarray(indices)


class jitcsde:
	"""
	Parameters follow reference `jitcsde/_jitcsde.py:29-87`.  Additional keyword arguments:

	ensemble : int — number of independent trajectories B integrated at once (one per GPU lane).
	device : int — CUDA device index.
	noise_capacity : int — bound on the noise memory per trajectory (items); overflow is reported, never silent.
	spec : ModelSpec — bypass the symbolic front end with pre-printed expressions (see `jitcsde_b200.models`).
	coupling : "exact" (default) or "dmma" — large systems with a dense linear coupling term only: evaluate that term on
		the FP64 tensor cores.  Faster, but the sums are fused and re-associated, so results match the reference within
		tolerances (1e-9 per trajectory, 3 standard errors on moments) instead of bit for bit.
	"""

	dynvar = y

	def __init__(self, f_sym=(), g_sym=(), *, helpers=None, g_helpers="auto", n=None, additive=None, ito=True,
			control_pars=(), callback_functions=(), verbose=True, module_location=None,
			ensemble=1, device=0, noise_capacity=0, spec: ModelSpec | None = None, coupling="exact", _jump_amp=None, _jump_iji=None):
		if callback_functions:
			raise NotImplementedError("Python callbacks cannot run inside a GPU kernel; express the term symbolically.")
		saved_library = None
		if module_location is not None and spec is None:
			spec, saved_library = self._load_compiled(module_location)
			if n is not None and int(n) != spec.n:
				raise ValueError(f"the saved module has {spec.n} components, not n = {n}")
		self.verbose = verbose
		self.B = int(ensemble)
		if self.B < 1:
			raise ValueError("ensemble must be a positive integer")
		self.device = device
		self.noise_capacity = noise_capacity
		if coupling not in ("exact", "dmma"):
			raise ValueError("coupling must be 'exact' or 'dmma'")
		self.coupling = coupling
		self.control_pars = tuple(control_pars)
		if spec is not None:
			self.spec = spec
			self.f_sym = self.g_sym = None
		else:
			if f_sym and not g_sym:
				raise ValueError("You gave f_sym as an argument but not g_sym. JiTCSDE cannot properly work with this.")
			self._symbolic = dict(f_sym=f_sym, g_sym=g_sym, n=n, additive=additive, ito=ito, control_pars=control_pars, helpers=helpers,
				g_helpers=g_helpers, jump_amp=_jump_amp, jump_iji=_jump_iji)
			self.spec, self.f_sym, self.g_sym = self._make_spec()
		self.n = self.spec.n
		self.additive = self.spec.additive
		self.integration_parameters_set = False
		self._pars = ControllerParameters()
		self.SDE: Ensemble | None = None     # the compiled ensemble (reference: the core object, `_jitcsde.py:455`)
		self._engine: Ensemble | None = None  # survives reset_integrator (device buffers are reused)
		self._initiated = False
		self.seed = None
		self._y = None
		self._t = None
		self._parameters = None
		self.compile_attempt = None
		self.last_stats = None
		if saved_library is not None:
			self._library = saved_library
			self.compile_attempt = True

	def _make_spec(self, simplify=None, do_cse=False):
		a = self._symbolic
		return spec_from_symbolic(a["f_sym"], a["g_sym"], n=a["n"], additive=a["additive"], ito=a["ito"], control_pars=a["control_pars"],
			helpers=a["helpers"], g_helpers=a["g_helpers"], jump_amp=a["jump_amp"], jump_iji=a["jump_iji"], simplify=simplify, do_cse=do_cse)

	# ---- reporting / checks --------------------------------------------------------------------------------------
	def report(self, message):
		if self.verbose:
			print(message)

	def check(self, fail_fast=True):
		"""Basic validation of the symbolic input (reference `_jitcsde.py:211-240`); raises ValueError."""
		if self.f_sym is None:
			return
		problems = check_expressions(self.f_sym, self.g_sym, self.n, self.control_pars,
			extra_symbols=(xi, tau) if self.spec.jump else ())
		if problems:
			raise ValueError(problems[0] if fail_fast else "\n".join(problems))

	# ---- state -----------------------------------------------------------------------------------------------------
	def _shape_out(self, a):
		return a[0] if self.B == 1 else a

	@property
	def y(self):
		if self._initiated:
			return self._shape_out(self.SDE.get_state())
		return None if self._y is None else self._shape_out(self._y)

	@y.setter
	def y(self, value):
		self._y = value
		self.reset_integrator()

	@property
	def y_dict(self):
		return {self.dynvar(i): self.y[..., i] for i in range(self.n)}

	@property
	def t(self):
		if self._initiated:
			tt = self.SDE.t
			return float(tt[0]) if self.B == 1 else tt
		return self._t

	@t.setter
	def t(self, value):
		self._t = value
		self.reset_integrator()

	def reset_integrator(self):
		"""Forget all stored noise and force re-initiation when needed next (reference `_jitcsde.py:242-246`)."""
		self._initiated = False
		self.SDE = None

	@property
	def is_initiated(self):
		return self._initiated

	def set_initial_value(self, initial_value, time=0.0):
		"""reference `_jitcsde.py:252-268`; `initial_value` is one state (length n) or one per trajectory ([B][n])."""
		if isinstance(initial_value, dict):
			vals = [None] * self.n
			for key, v in initial_value.items():
				vals[int(key.args[0])] = v
			initial_value = vals
		arr = np.array(initial_value, copy=True, dtype=float)
		if arr.ndim == 1:
			if arr.shape[0] != self.n:
				raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
			arr = np.tile(arr, (self.B, 1))
		elif arr.shape != (self.B, self.n):
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		self.y = arr
		self.t = time
		return self

	def set_seed(self, seed=None):
		"""reference `_jitcsde.py:270-274`.  An integer seeds trajectory k with seed+k; an array gives one seed each."""
		self.seed = seed

	def _seeds(self):
		if self.seed is None:
			base = int(random.getrandbits(32))  # reference :454
			return (base + np.arange(self.B, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
		if np.ndim(self.seed) == 0:
			return (np.uint64(int(self.seed) & 0xFFFFFFFF) + np.arange(self.B, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
		seeds = np.asarray(self.seed, dtype=np.uint64)
		if seeds.shape != (self.B,):
			raise ValueError("seed array must have one entry per trajectory")
		return seeds

	# ---- compilation ------------------------------------------------------------------------------------------------
	def compile_C(self, simplify=None, do_cse=False, numpy_rng=False, chunk_size=100, extra_compile_args=None,
			extra_link_args=None, verbose=False, modulename=None, omp=False):
		"""
		Name kept from the reference (`_jitcsde.py:302-442`); builds the CUDA library for this model with nvcc for
		sm_100a.  `simplify` (SymPy's simplify on every component; the reference's default "iff n <= 10" is NOT applied —
		None keeps the operation order the user wrote, which is what the parity oracle compiles too) and `do_cse`
		(common-subexpression elimination, reference `:366-375`: temporaries emitted ahead of the component statements) are
		honoured for symbolic input and regenerate the model; with `spec=` input they are refused.  The C-compiler-specific
		arguments (`chunk_size`, `extra_compile_args`, `extra_link_args`, `omp`) have no meaning for nvcc and are ignored;
		`numpy_rng=True` is refused (that path exists in the reference for testing only, `_jitcsde.py:327-328`).
		"""
		if numpy_rng:
			raise NotImplementedError("numpy_rng is a test device of the reference's C backend; the GPU stream follows random_numbers.c")
		if simplify or do_cse:
			if self.f_sym is None:
				raise NotImplementedError("simplify / do_cse need the symbolic input; this instance was built from a printed ModelSpec")
			if self.spec.wide and do_cse:
				raise NotImplementedError("do_cse is not available for wide models: their components already share one expression per template")
			self.spec, self.f_sym, self.g_sym = self._make_spec(simplify=bool(simplify), do_cse=bool(do_cse))
			self._engine = None  # a model library built from the old expressions must not be reused
			self.reset_integrator()
		self._library = _build.build(self.spec, verbose=verbose)
		self.compile_attempt = True
		return self._library

	compile_cuda = compile_C

	#: suffix of saved modules (the reference enforces the platform's extension-module suffix, jitcxde_common)
	MODULE_SUFFIX = ".jsde"

	def save_compiled(self, destination="", overwrite=False):
		"""
		Saves the compiled model for later use with `module_location=` (reference: `jitcxde.save_compiled`, used by
		tests/test_jitcsde.py:65-68).  The file is an archive of the printed model (ModelSpec, JSON) and the sm_100a
		library built from it; a directory or "" as destination gets the default file name.  Returns the path.
		"""
		import zipfile
		if self.compile_attempt is None:
			self.compile_C()
		default = f"jitced_{self.spec.name}{self.MODULE_SUFFIX}"
		if destination == "":
			destination = default
		elif os.path.isdir(destination) or destination.endswith(os.sep):
			destination = os.path.join(destination, default)
		elif not destination.endswith(self.MODULE_SUFFIX):
			destination = os.path.splitext(destination)[0] + self.MODULE_SUFFIX
		if os.path.exists(destination) and not overwrite:
			raise OSError("Target File already exists and \"overwrite\" is set to False")
		os.makedirs(os.path.dirname(os.path.abspath(destination)), exist_ok=True)
		with zipfile.ZipFile(destination, "w", zipfile.ZIP_DEFLATED) as z:
			z.writestr("spec.json", self.spec.to_json())
			z.writestr("sources_digest", _build._sources_digest())
			z.write(self._library, "library.so")
		return destination

	@staticmethod
	def _load_compiled(location):
		"""-> (ModelSpec, path of the installed library or None).  The saved binary is used when it was built from the
		kernel sources of this installation; otherwise the model is rebuilt from its printed form on first use."""
		import zipfile
		with zipfile.ZipFile(location) as z:
			spec = ModelSpec.from_json(z.read("spec.json").decode())
			if z.read("sources_digest").decode() != _build._sources_digest():
				return spec, None
			target = _build.library_path(spec)
			if not os.path.isfile(target):
				os.makedirs(os.path.dirname(target), exist_ok=True)
				tmp = f"{target}.{os.getpid()}.tmp"
				with open(tmp, "wb") as fh:
					fh.write(z.read("library.so"))
				os.replace(tmp, target)
		return spec, target

	def generate_lambdas(self):
		raise NotImplementedError("There is no Python/CPU core in this implementation (no fallback by design).")

	def _initiate(self):
		if self.compile_attempt is None:
			self.compile_C()
		if not self._initiated:
			assert self._y is not None, "You need to set an initial value first."
			assert self._t is not None, "You need to set an initial time first."
			if self._engine is None:
				self._engine = Ensemble(self.spec, self.B, device=self.device, noise_capacity=self.noise_capacity,
					library=self._library)
				self._engine.set_integration_parameters(self._pars)
				if self.coupling != "exact":
					self._engine.set_option("coupling", self.coupling)
			self._engine.seed(self._seeds())
			self._engine.set_state(self._y, self._t)
			self.SDE = self._engine
			self._initiated = True
			# a fresh reference core starts with zeroed control parameters; we keep the last values instead
			if self._parameters is not None:
				self.SDE.set_parameters(self._parameters)
		self._set_integration_parameters()

	# ---- parameters ---------------------------------------------------------------------------------------------------
	def set_parameters(self, *parameters):
		"""reference `_jitcsde.py:465-484`; additionally accepts a [B][n_par] array (one parameter set per trajectory)."""
		if len(parameters) == 1 and np.ndim(parameters[0]) >= 1:
			values = np.asarray(parameters[0], dtype=float)
		else:
			values = np.asarray(parameters, dtype=float)
		self._parameters = values
		self._initiate()
		self.SDE.set_parameters(values)

	def _set_integration_parameters(self):
		if not self.integration_parameters_set:
			self.report("Using default integration parameters.")
			self.set_integration_parameters()

	def set_integration_parameters(self, atol=1e-10, rtol=1e-5, first_step=1.0, min_step=1e-10, max_step=10.0,
			decrease_threshold=1.1, increase_threshold=0.5, safety_factor=0.9, max_factor=5.0, min_factor=0.2):
		"""reference `_jitcsde.py:491-566` (same defaults, clamps, assertions and warnings)."""
		if first_step > max_step:
			first_step = max_step
			warn("Decreasing first_step to match max_step", stacklevel=2)
		if min_step > first_step:
			min_step = first_step
			warn("Decreasing min_step to match first_step", stacklevel=2)
		assert decrease_threshold >= 1.0, "decrease_threshold smaller than 1"
		assert increase_threshold <= 1.0, "increase_threshold larger than 1"
		assert max_factor >= 1.0, "max_factor smaller than 1"
		assert min_factor <= 1.0, "min_factor larger than 1"
		assert safety_factor <= 1.0, "safety_factor larger than 1"
		assert atol >= 0.0, "negative atol"
		assert rtol >= 0.0, "negative rtol"
		if atol == 0 and rtol == 0:
			warn("atol and rtol are both 0. You probably do not want this.", stacklevel=2)
		self._pars = ControllerParameters(atol, rtol, first_step, min_step, max_step, decrease_threshold,
			increase_threshold, safety_factor, max_factor, min_factor, 1.5)
		self.atol, self.rtol, self.min_step, self.max_step = atol, rtol, min_step, max_step
		self.integration_parameters_set = True
		if self._engine is not None:
			self._engine.set_integration_parameters(self._pars)

	@property
	def dt(self):
		"""current step size of every trajectory (the reference's `self.dt`)"""
		self._initiate()
		d = self.SDE.controller_state()[0]
		return float(d[0]) if self.B == 1 else d

	# ---- integration ---------------------------------------------------------------------------------------------------
	def _raise_on_failure(self, stats):
		if stats["n_min_step"] or stats["n_overflow"]:
			status = self.SDE.status
			bad = np.flatnonzero(status == 1)
			if bad.size:
				raise UnsuccessfulIntegration(
					"\n"
					"Could not integrate with the given tolerance parameters:\n\n"
					f"atol: {self.atol:e}\n"
					f"rtol: {self.rtol:e}\n"
					f"min_step: {self.min_step:e}\n\n"
					f"({bad.size} of {self.B} trajectories, first indices {bad[:8].tolist()})\n"
					"The most likely reasons for this are:\n"
					"• The SDE is ill-posed or stiff.\n"
					"• You did not allow for an absolute error tolerance (atol) though your SDE calls for it. Even a very small absolute tolerance (1e-16) may sometimes help.",
					bad)
			over = np.flatnonzero(status == 2)
			raise MemoryError(f"noise memory exceeded noise_capacity on {over.size} trajectories (first {over[:8].tolist()}); "
				"construct with a larger noise_capacity")

	def integrate(self, target_time):
		"""reference `_jitcsde.py:597-630`: evolve every trajectory to `target_time`; returns the state(s)."""
		self._initiate()
		stats = self.SDE.integrate(target_time)
		self.last_stats = stats
		self._raise_on_failure(stats)
		return self._shape_out(self.SDE.get_state())

	def integrate_grid(self, times):
		"""
		`[self.integrate(T) for T in times]` in one kernel launch: ndarray [len(times)][B][n] (or [len(times)][n] for an
		ensemble of one).  Same results bit for bit as the loop of calls the reference's examples use
		(`examples/noisy_lorenz.py:95-96`), without one launch and one host round trip per sample.
		"""
		self._initiate()
		out = self.SDE.integrate_grid(times)
		self.last_stats = self.SDE.last_stats
		self._raise_on_failure(self.last_stats)
		return out[:, 0, :] if self.B == 1 else out

	def pin_noise(self, number, step_size):
		"""reference `_jitcsde.py:632-654`"""
		self._initiate()
		assert number >= 0, "Number must be non-negative"
		assert step_size > 0, "Step size must be positive"
		if not isinstance(number, int):
			warn("`number` does not appear to be a integer. This is very likely cause an error immediately.", stacklevel=2)
		self.SDE.pin_noise(number, step_size)

	def pin_noise_path(self, steps, dw, dz):
		"""Pre-fill the noise memory with caller-supplied Wiener increments (steps [K]; dw, dz [B][K][n] or [K][n]) instead
		of fresh draws: integrate along a prescribed Brownian path (no counterpart in the reference)."""
		self._initiate()
		self.SDE.pin_noise_path(steps, dw, dz)

	def checkpoint(self):
		"""The whole integrator state (states, times, step sizes, noise memory, every trajectory's generator state,
		counters, jump scheduling, controller) as a uint8 array — for long campaigns (SURVEY.md §8 f item 4; the reference
		has no counterpart).  `restore(blob)` on an instance with the same model and ensemble size continues bit for bit."""
		self._initiate()
		return self.SDE.checkpoint()

	def restore(self, blob):
		self._initiate()
		self.SDE.restore(blob)

	def ensemble_moments(self, process_group=None, shift=None):
		"""
		Mean and variance of the current state over the whole ensemble.  With `torch.distributed` initialised the
		{count, Σ(y−c), Σ(y−c)²} accumulators are all-reduced (NCCL over NVLink) so every rank gets the moments of the
		union of all shards (SURVEY.md §8 e).  `shift` (c, length n) must be the same on every rank; default zeros.
		"""
		import torch

		from . import sharding

		self._initiate()
		dev = torch.device("cuda", self.device)
		acc = torch.zeros(1 + 2 * self.n, dtype=torch.float64, device=dev)
		shift = np.zeros(self.n) if shift is None else np.asarray(shift, dtype=np.float64)
		self.SDE.moments_into(acc.data_ptr(), shift)
		sharding.all_reduce_sum(acc, process_group)
		return sharding.moments_from_accumulators(acc.cpu().numpy(), shift)


class jitcsde_jump(jitcsde):
	"""
	`jitcsde` with random jumps (reference `_jitcsde.py:656-722`).  The reference takes two Python callables; a GPU lane
	cannot call Python, so both are given declaratively and consumed from a per-trajectory *tape*:

	IJI : waiting times.  The string "generated" (unit exponentials drawn on the device by a counter-based generator —
		no tape, no length limit; combine with `waiting_time` for a rate), or an array of shape (L,) / (B, L) — the k-th jump of a trajectory happens that long
		after the previous one — or a callable `IJI(time, state)` that does not depend on its arguments (like the
		reference's example, `examples/noisy_and_jumpy_lorenz.py:48-49`); it is called L times per trajectory up front.
	amp : iterable of n symbolic expressions in `y(i)`, `t` and `xi`, where `xi` is the tape's standard-normal variate
		for this jump; evaluated on the pre-jump state and added to it (reference `:696-698`).
	xis : optional array (L,) / (B, L) of variates; drawn from `rng.standard_normal` if omitted.
	tape_length : L when IJI is a callable (default 1024).  Past the tape's end no further jumps happen: the attribute
		`tape_exhausted` then holds the number of trajectories whose jump process stopped early and a RuntimeWarning is
		issued (the reference draws a new waiting time after every jump and never runs out, `_jitcsde.py:698-699`); use
		a longer tape or `IJI="generated"`.
	jump_seed : seed of the counter-based jump variates for `IJI="generated"` (default: drawn from `rng`).

	Randomness after a reset: `set_initial_value` / `reset_integrator` restart the Wiener noise from the seeds; the jump
	process is renewed as well — a callable `IJI` is called again, variates that were not supplied explicitly (`xis=None`)
	are redrawn from `rng`, and the generated mode takes a new seed from `rng` unless `jump_seed` was given — so repeated
	runs on one instance are independent samples, like in the reference, where IJI and amp draw from a live generator.
	Explicit arrays (and an explicit `jump_seed`) are replayed as given.
	waiting_time : optional symbolic expression in `y(i)`, `t` and `tau` — the reference's state- and time-dependent
		`IJI(time, state)` (`_jitcsde.py:682-687`), evaluated on the device when a jump is scheduled; `tau` is the
		tape's value for that jump (e.g. unit exponentials from `IJI`, and `tau/rate(y)` here).  Default: `tau`.
	"""

	def __init__(self, IJI, amp, *args, rng=None, ito=True, xis=None, tape_length=1024, waiting_time=None, jump_seed=None, **kwargs):
		if not ito:
			raise NotImplementedError("I don’t know how to convert jumpy Stratonovich SDEs to Itō SDEs – nobody does.")
		if rng is None:
			rng = np.random.default_rng()
		if "spec" not in kwargs:
			kwargs["_jump_amp"] = list(amp) if not callable(amp) else self._refuse_callable_amp()
			kwargs["_jump_iji"] = waiting_time
		super().__init__(*args, **kwargs)
		if self.spec.jump is None:
			raise ValueError("the model has no jump amplitude")
		self.rng = rng
		self.IJI = IJI  # array, callable, or the string "generated" (device-side counter-based variates, see integrate)
		self._first_index = 0  # global index of trajectory 0 (set it on every rank of a sharded ensemble)
		self._xis_arg = xis
		self._tape_length = tape_length
		self._taus = self._xis = None
		self._tape_uploaded = False
		self._jump_seed_arg = None if jump_seed is None else int(jump_seed)
		self._jump_seed = self._jump_seed_arg
		self.tape_exhausted = 0  # trajectories whose tape ran out before the target time (sticky until the next reset)

	@property
	def _generated(self):
		return isinstance(self.IJI, str) and self.IJI == "generated"

	def _generated_seed(self):
		if self._jump_seed is None:
			self._jump_seed = int(self.rng.integers(0, 2 ** 63))
		return self._jump_seed

	@staticmethod
	def _refuse_callable_amp():
		raise NotImplementedError("amp must be given as n symbolic expressions in y(i), t and xi (see the class docstring); "
			"a Python callable cannot run inside a GPU kernel.")

	def _make_tape(self):
		if self._taus is not None:
			return
		if callable(self.IJI):
			taus = np.array([[self.IJI(0.0, None) for _ in range(self._tape_length)] for _ in range(self.B)], dtype=float)
		else:
			taus = np.asarray(self.IJI, dtype=float)
			if taus.ndim == 1:
				taus = np.tile(taus, (self.B, 1))
		if taus.shape[0] != self.B or taus.ndim != 2:
			raise ValueError("IJI must have shape (L,) or (B, L)")
		if (taus < 0).any():
			raise ValueError("waiting times must be non-negative")
		if self._xis_arg is None:
			xis = self.rng.standard_normal(taus.shape)
		else:
			xis = np.asarray(self._xis_arg, dtype=float)
			if xis.ndim == 1:
				xis = np.tile(xis, (self.B, 1))
		if xis.shape != taus.shape:
			raise ValueError("xis must have the same shape as the waiting times")
		self._taus = np.ascontiguousarray(taus)
		self._xis = np.ascontiguousarray(xis)
		self._tape_uploaded = False

	def reset_integrator(self):
		"""The engine rewinds every trajectory's jump counter together with the noise (jsde_set_state); the jump variates
		are renewed unless they were given explicitly (see the class docstring)."""
		super().reset_integrator()
		if not hasattr(self, "_taus"):
			return  # called from the base constructor
		self.tape_exhausted = 0
		if callable(self.IJI) or self._xis_arg is None:
			self._taus = self._xis = None
			self._tape_uploaded = False
		self._jump_seed = self._jump_seed_arg

	def _prepare_jumps(self):
		"""-> (mode, seed): uploads the tape if it is new"""
		if self._generated:
			return "generated", self._generated_seed()
		self._make_tape()
		if not self._tape_uploaded:
			self.SDE.set_jump_tape(self._taus, self._xis)
			self._tape_uploaded = True
		return "tape", 0

	def _after_jumps(self, stats):
		self.last_stats = stats
		if stats.get("n_tape_exhausted"):
			self.tape_exhausted += stats["n_tape_exhausted"]
			warn(f"the jump tape (length {self._taus.shape[1]}) ran out on {stats['n_tape_exhausted']} of {self.B} trajectories before the "
				"target time: their jump process stopped early.  Use a larger tape_length or IJI=\"generated\".", RuntimeWarning, stacklevel=3)
		self._raise_on_failure(stats)

	def integrate(self, target_time):
		"""reference `_jitcsde.py:693-700`, all trajectories in one launch"""
		self._initiate()
		mode, seed = self._prepare_jumps()
		if mode == "generated":
			# variates generated on the device (counter-based): unit exponentials feed `waiting_time`, no tape
			stats = self.SDE.integrate_jumps_generated(target_time, seed, self._first_index)
		else:
			stats = self.SDE.integrate_jumps(target_time)
		self._after_jumps(stats)
		return self._shape_out(self.SDE.get_state())

	def integrate_grid(self, times):
		"""`[self.integrate(T) for T in times]` in one launch, jumps included (tape or generated variates)."""
		self._initiate()
		mode, seed = self._prepare_jumps()
		out = self.SDE.integrate_grid(times, jumps=mode, seed=seed, first_index=self._first_index)
		self._after_jumps(self.SDE.last_stats)
		return out[:, 0, :] if self.B == 1 else out

	def checkpoint(self):
		"""The engine's blob (see `jitcsde.checkpoint`) followed by the jump bookkeeping a later `restore` needs: the seed of
		the generated variates.  A tape is an input the caller supplies again (explicit arrays), exactly like
		per-trajectory control parameters."""
		blob = super().checkpoint()
		seed = self._generated_seed() if self._generated else 0
		return np.concatenate([blob, np.frombuffer(np.array([seed], dtype=np.uint64).tobytes(), dtype=np.uint8)])

	def restore(self, blob):
		blob = np.ascontiguousarray(blob, dtype=np.uint8)
		super().restore(blob[:-8])
		seed = int(np.frombuffer(blob[-8:].tobytes(), dtype=np.uint64)[0])
		if self._generated:
			self._jump_seed = seed

	def check(self, fail_fast=True):
		super().check(fail_fast)
		if not self.spec.wide and len(self.spec.jump.amp) != self.n:  # wide models: one expression in the component index
			raise ValueError("Output of amp function has the wrong dimension")


def test(omp=True, sympy=True):
	"""
	Quick self-test (reference `_jitcsde.py:724-…`): builds a small model with nvcc and, if a GPU is present, integrates it.
	"""
	import torch

	SDE = jitcsde([-y(0)], [0.1 * y(0) / (1 + y(0) ** 2)], verbose=False)
	SDE.compile_C()
	if torch.cuda.is_available():
		SDE.set_initial_value([1.0])
		SDE.integrate(0.1)
