"""
ctypes binding of the C ABI (`include/jsde.h`) and a thin object wrapper, `Ensemble`: B copies of the reference's core
object `sde_integrator` (reference `jitcsde/jitced_template.c:630-646`) living on one GPU.

The wrapper adds nothing numerical — every method is one ABI call — and raises instead of returning status codes.
Loading fails loudly (`BuildError` / `OSError` / `JsdeError`) when the CUDA library cannot be built or no device is
present: there is no CPU path in the product.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _build
from .modelspec import ModelSpec

c_double_p = ctypes.POINTER(ctypes.c_double)
c_u32_p = ctypes.POINTER(ctypes.c_uint32)
c_u64_p = ctypes.POINTER(ctypes.c_uint64)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)


class JsdeError(RuntimeError):
	pass


class Ctrl(ctypes.Structure):
	_fields_ = [(name, ctypes.c_double) for name in (
		"atol", "rtol", "first_step", "min_step", "max_step", "decrease_threshold", "increase_threshold",
		"safety_factor", "max_factor", "min_factor", "q")]


class Stats(ctypes.Structure):
	_fields_ = [
		("accepted", ctypes.c_ulonglong), ("rejected", ctypes.c_ulonglong), ("fresh_draws", ctypes.c_ulonglong),
		("jumps", ctypes.c_ulonglong), ("n_min_step", ctypes.c_longlong), ("n_overflow", ctypes.c_longlong),
		("max_noise_depth", ctypes.c_int), ("kernel_launches", ctypes.c_int), ("kernel_ms", ctypes.c_float),
		("n_tape_exhausted", ctypes.c_longlong),
	]

	def as_dict(self):
		return {name: getattr(self, name) for name, _ in self._fields_}


class JumpTape(ctypes.Structure):
	_fields_ = [("taus_host", c_double_p), ("xis_host", c_double_p), ("length", ctypes.c_int64)]


#: every symbol include/jsde.h declares (tests/test_abi_symbols.py checks the built library exports them all)
ABI_SYMBOLS = (
	"jsde_model_info", "jsde_model_name", "jsde_create", "jsde_destroy", "jsde_set_option", "jsde_set_state", "jsde_seed",
	"jsde_set_integration_parameters", "jsde_set_parameters", "jsde_pin_noise", "jsde_integrate", "jsde_set_jump_tape", "jsde_integrate_jumps",
	"jsde_integrate_grid", "jsde_integrate_grid_jumps",
	"jsde_get_state", "jsde_get_controller_state", "jsde_moments", "jsde_get_next_step", "jsde_get_p", "jsde_accept_step",
	"jsde_apply_jump", "jsde_set_time", "jsde_draw_gauss_debug", "jsde_dump_noise", "jsde_measure_fp64_peak",
	"jsde_debug_log", "jsde_debug_pow", "jsde_debug_exp", "jsde_debug_shared_div", "jsde_debug_inline_math", "jsde_pin_noise_path", "jsde_integrate_jumps_generated", "jsde_generated_jump_tape", "jsde_checkpoint_bytes", "jsde_checkpoint_save", "jsde_checkpoint_restore", "jsde_last_error", "jsde_version",
)


def _declare(lib):
	H = ctypes.c_void_p
	lib.jsde_last_error.restype = ctypes.c_char_p
	lib.jsde_version.restype = ctypes.c_char_p
	lib.jsde_model_name.restype = ctypes.c_char_p
	lib.jsde_model_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 4
	lib.jsde_create.argtypes = [ctypes.POINTER(H), ctypes.c_int, ctypes.c_int64, ctypes.c_int]
	lib.jsde_destroy.argtypes = [H]
	lib.jsde_set_option.argtypes = [H, ctypes.c_char_p, ctypes.c_char_p]
	lib.jsde_set_state.argtypes = [H, c_double_p, c_double_p, ctypes.c_double]
	lib.jsde_seed.argtypes = [H, c_u32_p]
	lib.jsde_set_integration_parameters.argtypes = [H, ctypes.POINTER(Ctrl)]
	lib.jsde_set_parameters.argtypes = [H, c_double_p, ctypes.c_int]
	lib.jsde_pin_noise.argtypes = [H, ctypes.c_uint, ctypes.c_double]
	lib.jsde_integrate.argtypes = [H, ctypes.c_double, ctypes.POINTER(Stats)]
	lib.jsde_integrate_jumps.argtypes = [H, ctypes.c_double, ctypes.POINTER(JumpTape), ctypes.POINTER(Stats)]
	lib.jsde_integrate_grid.argtypes = [H, c_double_p, ctypes.c_int64, ctypes.POINTER(JumpTape), c_double_p, ctypes.POINTER(Stats)]
	lib.jsde_set_jump_tape.argtypes = [H, ctypes.POINTER(JumpTape)]
	lib.jsde_integrate_grid_jumps.argtypes = [H, c_double_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, c_double_p, ctypes.POINTER(Stats)]
	lib.jsde_integrate_jumps_generated.argtypes = [H, ctypes.c_double, ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(Stats)]
	lib.jsde_pin_noise_path.argtypes = [H, ctypes.c_uint, c_double_p, c_double_p, c_double_p]
	lib.jsde_generated_jump_tape.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_double_p, c_double_p]
	lib.jsde_get_state.argtypes = [H, c_double_p, c_double_p, c_i32_p]
	lib.jsde_get_controller_state.argtypes = [H, c_double_p, c_u64_p, c_u64_p]
	lib.jsde_moments.argtypes = [H, ctypes.c_void_p, c_double_p]
	lib.jsde_get_next_step.argtypes = [H, c_double_p]
	lib.jsde_get_p.argtypes = [H, ctypes.c_double, ctypes.c_double, c_double_p]
	lib.jsde_accept_step.argtypes = [H, c_u8_p]
	lib.jsde_apply_jump.argtypes = [H, c_double_p]
	lib.jsde_set_time.argtypes = [H, c_double_p]
	lib.jsde_draw_gauss_debug.argtypes = [H, ctypes.c_double, ctypes.c_int, c_double_p]
	lib.jsde_dump_noise.argtypes = [H, ctypes.c_int64, c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
	lib.jsde_measure_fp64_peak.argtypes = [ctypes.c_int, c_double_p]
	lib.jsde_debug_log.argtypes = [ctypes.c_int, c_double_p, c_double_p, ctypes.c_int64]
	lib.jsde_debug_pow.argtypes = [ctypes.c_int, c_double_p, ctypes.c_double, c_double_p, ctypes.c_int64]
	lib.jsde_debug_exp.argtypes = [ctypes.c_int, c_double_p, c_double_p, ctypes.c_int64]
	lib.jsde_debug_shared_div.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, c_u8_p, ctypes.c_int64]
	lib.jsde_checkpoint_bytes.argtypes = [ctypes.c_void_p]
	lib.jsde_checkpoint_bytes.restype = ctypes.c_int64
	lib.jsde_checkpoint_save.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
	lib.jsde_checkpoint_restore.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
	lib.jsde_debug_inline_math.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, c_u8_p, c_double_p, c_u8_p, ctypes.c_int64]
	for name in ABI_SYMBOLS:
		fn = getattr(lib, name)
		if fn.restype is ctypes.c_int and name not in ("jsde_last_error", "jsde_version", "jsde_model_name"):
			fn.restype = ctypes.c_int
	return lib


_LOADED: dict = {}


def load_library(spec: ModelSpec, path: str | None = None):
	"""Build (if needed) and load the model's shared library."""
	path = path or _build.build(spec)
	if path not in _LOADED:
		_LOADED[path] = _declare(ctypes.CDLL(path))
	return _LOADED[path]


def _dp(a):
	return a.ctypes.data_as(c_double_p)


@dataclass
class ControllerParameters:
	"""Defaults of `set_integration_parameters` (reference `jitcsde/_jitcsde.py:491-502`); q at `:564`."""
	atol: float = 1e-10
	rtol: float = 1e-5
	first_step: float = 1.0
	min_step: float = 1e-10
	max_step: float = 10.0
	decrease_threshold: float = 1.1
	increase_threshold: float = 0.5
	safety_factor: float = 0.9
	max_factor: float = 5.0
	min_factor: float = 0.2
	q: float = 1.5


class Ensemble:
	"""B trajectories of one model on one GPU, behind the reference core's method names."""

	def __init__(self, spec: ModelSpec, B: int, device: int = 0, noise_capacity: int = 0, library: str | None = None):
		self.spec = spec
		self.B = int(B)
		self.n = spec.n
		self.device = device
		self.lib = load_library(spec, library)
		n, additive, n_par, has_jump = (ctypes.c_int() for _ in range(4))
		self.lib.jsde_model_info(n, additive, n_par, has_jump)
		if (n.value, bool(additive.value), n_par.value) != (spec.n, spec.additive, len(spec.control_pars)):
			raise JsdeError("library does not match the model description")
		self._h = ctypes.c_void_p()
		self._check(self.lib.jsde_create(ctypes.byref(self._h), device, self.B, noise_capacity))
		self.has_tape = False
		self.last_stats = None

	def _check(self, rc):
		if rc != 0:
			raise JsdeError(f"[{rc}] {self.lib.jsde_last_error().decode()}")

	def close(self):
		if getattr(self, "_h", None):
			self.lib.jsde_destroy(self._h)
			self._h = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	def set_option(self, name: str, value) -> None:
		"""`jsde_set_option`: "coupling" = "exact" | "dmma" (wide models), "recycle" = 0 | 1 and "slice" = rounds between yields (lane models)"""
		self._check(self.lib.jsde_set_option(self._h, name.encode(), str(value).encode()))

	# ---- state / seed / parameters --------------------------------------------------------------------------------
	def set_state(self, y0, t0=0.0):
		y0 = np.ascontiguousarray(y0, dtype=np.float64)
		if y0.shape != (self.B, self.n):
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		if np.ndim(t0) == 0:
			self._check(self.lib.jsde_set_state(self._h, _dp(y0), None, float(t0)))
		else:
			t0 = np.ascontiguousarray(t0, dtype=np.float64)
			assert t0.shape == (self.B,)
			self._check(self.lib.jsde_set_state(self._h, _dp(y0), _dp(t0), 0.0))

	def seed(self, seeds):
		seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64) & np.uint64(0xFFFFFFFF), dtype=np.uint32)
		assert seeds.shape == (self.B,)
		self._check(self.lib.jsde_seed(self._h, seeds.ctypes.data_as(c_u32_p)))

	def set_integration_parameters(self, pars: ControllerParameters):
		c = Ctrl(*(getattr(pars, name) for name, _ in Ctrl._fields_))
		self._check(self.lib.jsde_set_integration_parameters(self._h, ctypes.byref(c)))

	def set_parameters(self, values):
		v = np.ascontiguousarray(values, dtype=np.float64)
		npar = len(self.spec.control_pars)
		if v.shape == (npar,):
			self._check(self.lib.jsde_set_parameters(self._h, _dp(v), 0))
		elif v.shape == (self.B, npar):
			self._check(self.lib.jsde_set_parameters(self._h, _dp(v), 1))
		else:
			raise ValueError("Wrong input.")

	def pin_noise(self, number, step):
		self._check(self.lib.jsde_pin_noise(self._h, int(number), float(step)))

	def pin_noise_path(self, steps, dw, dz):
		"""Append caller-supplied Wiener increments to the noise memory: steps [K], dw and dz [B][K][n] (or [K][n],
		broadcast).  Integration then runs along this path for sum(steps) time units (Brownian bridges in between)."""
		steps = np.ascontiguousarray(steps, dtype=np.float64)
		dw = np.asarray(dw, dtype=np.float64)
		dz = np.asarray(dz, dtype=np.float64)
		if dw.ndim == 2:
			dw = np.broadcast_to(dw, (self.B,) + dw.shape)
			dz = np.broadcast_to(dz, (self.B,) + dz.shape)
		if dw.shape != (self.B, steps.size, self.n) or dz.shape != dw.shape:
			raise ValueError("dw and dz must have shape [B][len(steps)][n]")
		dw, dz = np.ascontiguousarray(dw), np.ascontiguousarray(dz)
		self._check(self.lib.jsde_pin_noise_path(self._h, steps.size, _dp(steps), _dp(dw), _dp(dz)))

	# ---- the hot path ---------------------------------------------------------------------------------------------------
	def integrate(self, target_time) -> dict:
		st = Stats()
		self._check(self.lib.jsde_integrate(self._h, float(target_time), ctypes.byref(st)))
		self.last_stats = st.as_dict()
		return self.last_stats

	def set_jump_tape(self, taus, xis):
		"""Copy a jump tape ([B][L] waiting-time values and standard-normal variates) to the device.  Always uploads — the
		arrays are not referenced afterwards, and nothing is cached by identity; call it again after editing a tape."""
		t = np.ascontiguousarray(taus, dtype=np.float64)
		x = np.ascontiguousarray(xis, dtype=np.float64)
		if t.ndim != 2 or t.shape != x.shape or t.shape[0] != self.B:
			raise ValueError("taus and xis must both have shape [B][L]")
		self._check(self.lib.jsde_set_jump_tape(self._h, ctypes.byref(JumpTape(_dp(t), _dp(x), t.shape[1]))))
		self.has_tape = True

	def integrate_jumps(self, target_time, taus=None, xis=None) -> dict:
		"""`taus`/`xis` given: upload them first (set_jump_tape); omitted: continue on the tape set before."""
		if taus is not None:
			self.set_jump_tape(taus, xis)
		st = Stats()
		self._check(self.lib.jsde_integrate_jumps(self._h, float(target_time), None, ctypes.byref(st)))
		self.last_stats = st.as_dict()
		return self.last_stats

	def integrate_jumps_generated(self, target_time, seed, first_index=0) -> dict:
		"""Jumps whose variates are generated on the device (counter-based, csrc/philox_jumps.h): no tape."""
		st = Stats()
		self._check(self.lib.jsde_integrate_jumps_generated(self._h, float(target_time), int(seed), int(first_index), ctypes.byref(st)))
		self.last_stats = st.as_dict()
		return self.last_stats

	def integrate_grid(self, times, taus=None, xis=None, jumps=None, seed=0, first_index=0):
		"""States at every sample time, [n_times][B][n], from ONE launch (semantics of consecutive integrate() calls).
		Jump models: pass `taus`/`xis` (uploaded first), or `jumps="tape"` (the tape set before) or `jumps="generated"`
		with `seed`/`first_index` (counter-based variates)."""
		times = np.ascontiguousarray(times, dtype=np.float64)
		out = np.zeros((times.size, self.B, self.n))
		if taus is not None:
			self.set_jump_tape(taus, xis)
			jumps = "tape"
		mode = {None: 0, "tape": 1, "generated": 2}[jumps]
		st = Stats()
		self._check(self.lib.jsde_integrate_grid_jumps(self._h, _dp(times), times.size, mode, int(seed), int(first_index), _dp(out), ctypes.byref(st)))
		self.last_stats = st.as_dict()
		return out

	# ---- results -------------------------------------------------------------------------------------------------------
	def get_state(self, out=None):
		y = out if out is not None else np.empty((self.B, self.n))
		self._check(self.lib.jsde_get_state(self._h, _dp(y), None, None))
		return y

	@property
	def t(self):
		t = np.empty(self.B)
		self._check(self.lib.jsde_get_state(self._h, None, _dp(t), None))
		return t

	@t.setter
	def t(self, value):
		v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.B,)))
		self._check(self.lib.jsde_set_time(self._h, _dp(v)))

	@property
	def status(self):
		s = np.empty(self.B, dtype=np.int32)
		self._check(self.lib.jsde_get_state(self._h, None, None, s.ctypes.data_as(c_i32_p)))
		return s

	def controller_state(self):
		dt = np.empty(self.B)
		acc = np.empty(self.B, dtype=np.uint64)
		rej = np.empty(self.B, dtype=np.uint64)
		self._check(self.lib.jsde_get_controller_state(self._h, _dp(dt), acc.ctypes.data_as(c_u64_p), rej.ctypes.data_as(c_u64_p)))
		return dt, acc, rej

	def moments_into(self, acc_dev_ptr: int, shift=None):
		"""Accumulate {count, Σ(y-c), Σ(y-c)²} into device memory at `acc_dev_ptr` (1+2n doubles)."""
		s = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
		self._check(self.lib.jsde_moments(self._h, ctypes.c_void_p(acc_dev_ptr), None if s is None else _dp(s)))

	# ---- single-step surface (reference core methods) ------------------------------------------------------------------
	def get_next_step(self, h):
		v = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.float64), (self.B,)))
		self._check(self.lib.jsde_get_next_step(self._h, _dp(v)))

	def get_p(self, atol, rtol):
		p = np.empty(self.B)
		self._check(self.lib.jsde_get_p(self._h, float(atol), float(rtol), _dp(p)))
		return p

	def accept_step(self, mask=None):
		if mask is None:
			self._check(self.lib.jsde_accept_step(self._h, None))
		else:
			m = np.ascontiguousarray(mask, dtype=np.uint8)
			self._check(self.lib.jsde_accept_step(self._h, m.ctypes.data_as(c_u8_p)))

	def apply_jump(self, change):
		c = np.ascontiguousarray(change, dtype=np.float64)
		if c.shape != (self.B, self.n):
			raise ValueError("Wrong input. Note that the function returning jump amplitudes must return a NumPy array.")
		self._check(self.lib.jsde_apply_jump(self._h, _dp(c)))

	def draw_gauss(self, scale, pairs):
		out = np.empty((self.B, pairs, 2))
		self._check(self.lib.jsde_draw_gauss_debug(self._h, float(scale), int(pairs), _dp(out)))
		return out

	def checkpoint(self) -> np.ndarray:
		"""Everything needed to continue this ensemble bit for bit later (states, noise memory, generator states, controller,
		counters, jump scheduling) as one uint8 array; see `jsde_checkpoint_save`."""
		size = self.lib.jsde_checkpoint_bytes(self._h)
		if size < 0:
			raise JsdeError(self.lib.jsde_last_error().decode())
		blob = np.empty(size, dtype=np.uint8)
		self._check(self.lib.jsde_checkpoint_save(self._h, blob.ctypes.data_as(ctypes.c_void_p), size))
		return blob

	def restore(self, blob: np.ndarray) -> None:
		"""Inverse of `checkpoint()` for an ensemble of the same model, size and noise capacity."""
		blob = np.ascontiguousarray(blob, dtype=np.uint8)
		self._check(self.lib.jsde_checkpoint_restore(self._h, blob.ctypes.data_as(ctypes.c_void_p), blob.size))

	def dump_noise(self, b, max_items=64):
		out = np.zeros((max_items, 1 + 2 * self.n))
		count = ctypes.c_int()
		self._check(self.lib.jsde_dump_noise(self._h, int(b), _dp(out), max_items, ctypes.byref(count)))
		return out[:min(count.value, max_items)]


def measure_fp64_peak(lib, device=0) -> float:
	tf = ctypes.c_double()
	if lib.jsde_measure_fp64_peak(device, ctypes.byref(tf)) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return tf.value


def device_log(lib, x, device=0):
	x = np.ascontiguousarray(x, dtype=np.float64)
	out = np.empty_like(x)
	if lib.jsde_debug_log(device, _dp(x), _dp(out), x.size) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return out


def device_exp(lib, x, device=0):
	x = np.ascontiguousarray(x, dtype=np.float64)
	out = np.empty_like(x)
	if lib.jsde_debug_exp(device, _dp(x), _dp(out), x.size) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return out


def device_pow(lib, x, y, device=0):
	x = np.ascontiguousarray(x, dtype=np.float64)
	out = np.empty_like(x)
	if lib.jsde_debug_pow(device, _dp(x), float(y), _dp(out), x.size) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return out


def device_shared_div(lib, a, b, device=0):
	"""(a/b through the stages' shared-divisor path, tame mask)"""
	a = np.ascontiguousarray(a, dtype=np.float64)
	b = np.ascontiguousarray(b, dtype=np.float64)
	out = np.empty_like(a)
	tame = np.empty(a.size, dtype=np.uint8)
	if lib.jsde_debug_shared_div(device, _dp(a), _dp(b), _dp(out), tame.ctypes.data_as(c_u8_p), a.size) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return out, tame.astype(bool)


def device_inline_math(lib, a, b, device=0):
	"""(a/b, ok mask, sqrt(a), ok mask) through the straight-line division and square root of csrc/exact_div.cuh"""
	a = np.ascontiguousarray(a, dtype=np.float64)
	b = np.ascontiguousarray(b, dtype=np.float64)
	div, root = np.empty_like(a), np.empty_like(a)
	okd, oks = np.empty(a.size, dtype=np.uint8), np.empty(a.size, dtype=np.uint8)
	if lib.jsde_debug_inline_math(device, _dp(a), _dp(b), _dp(div), okd.ctypes.data_as(c_u8_p), _dp(root), oks.ctypes.data_as(c_u8_p), a.size) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return div, okd.astype(bool), root, oks.astype(bool)


def generated_jump_tape(lib, seed, first_index, count, length, device=0):
	"""(taus, xis), each [count][length]: the counter-based jump variates `integrate_jumps_generated` uses"""
	taus = np.empty((count, length))
	xis = np.empty((count, length))
	if lib.jsde_generated_jump_tape(device, int(seed), int(first_index), int(count), int(length), _dp(taus), _dp(xis)) != 0:
		raise JsdeError(lib.jsde_last_error().decode())
	return taus, xis
