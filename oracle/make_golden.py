"""
TEST INFRASTRUCTURE.  Generates tests/golden/*.json from the UNMODIFIED reference C core (strict-IEEE build,
oracle/build_ref.py) driven by the transcribed reference controller (oracle/ref_driver.py).  Needs /root/reference,
i.e. runs only in the build container; the fixtures it writes are committed and travel to the GPU box.

    python -m oracle.make_golden

All floating-point values are stored as C99 hex-float strings (exact).
"""

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from jitcsde_b200 import models  # noqa: E402  (plain data: expression strings)
from oracle import build_ref, ref_driver  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def hx(a):
	a = np.asarray(a, dtype=np.float64)
	if a.ndim == 0:
		return float(a).hex()
	return [hx(x) for x in a]


def ensemble_inputs(spec, B):
	"""The y0 / seed rule of SURVEY.md §8(d): y0_k = default_rng(42).random((B,n))[k], seed_k = k."""
	return np.random.default_rng(42).random((B, spec.n)), np.arange(B, dtype=np.uint32)


def jump_tape(B, length, scale=1.0):
	"""Per-trajectory tape: waiting times Exp(scale), variates N(0,1), from default_rng([42,k])."""
	taus = np.empty((B, length))
	xis = np.empty((B, length))
	for k in range(B):
		r = np.random.default_rng([42, k])
		taus[k] = r.exponential(scale, length)
		xis[k] = r.standard_normal(length)
	return taus, xis


def rng_goldens():
	# the RNG is model-independent; any model's core exposes it through pin_noise only, so draw through a 1-D
	# additive model with f=0, g=1: one get_next_step(h) from empty memory consumes exactly one pair with scale sqrt(h)
	out = {}
	spec = models.ModelSpec("unit_noise", 1, True, ("0.0",), ("1.0",))
	mod = build_ref.load(spec, "strict")
	for seed in (0, 1, 42, 12345, 4294967295):
		core = mod.sde_integrator(0.0, np.zeros(1), seed)
		# new_state = 0 + E_N(=0) + (1/3)*(0+2*0) + I_1*g1  = DW exactly; time does not matter
		dws = []
		for _ in range(2000):
			core.get_next_step(1.0)   # scale = sqrt(1.0) = 1
			core.accept_step()
			dws.append(core.get_state()[0])
			core.apply_jump(-core.get_state())  # back to 0 exactly
		out[str(seed)] = {"dw_scale1_first8": hx(dws[:8]), "dw_scale1_1992_2000": hx(dws[1992:])}
	return out


def scripted(spec, y0, seed, script):
	"""Run a script of core calls; record state/t/p after each action."""
	mod = build_ref.load(spec, "strict")
	core = mod.sde_integrator(0.0, np.array(y0, dtype=np.float64), seed)
	rec = []
	for action in script:
		kind = action[0]
		if kind == "step":
			core.get_next_step(action[1])
			rec.append({"do": "step", "h": hx(action[1]), "p": hx(core.get_p(1e-10, 1e-5))})
		elif kind == "accept":
			core.accept_step()
			rec.append({"do": "accept", "y": hx(core.get_state()), "t": hx(core.t)})
		elif kind == "pin":
			core.pin_noise(action[1], action[2])
			rec.append({"do": "pin", "number": action[1], "step": hx(action[2])})
	return {"model": spec.name, "y0": hx(y0), "seed": seed, "script": rec}


def trajectories(spec, B, T, **ctrl):
	mod = build_ref.load(spec, "strict")
	y0, seeds = ensemble_inputs(spec, B)
	rows = []
	for k in range(B):
		R = ref_driver.RefIntegrator(mod, y0[k], 0.0, int(seeds[k]), **ctrl)
		status = 0
		try:
			R.integrate(T)
		except ref_driver.UnsuccessfulIntegration:
			status = 1
		rows.append({"y": hx(R.core.get_state()), "t": hx(R.core.t), "accepted": R.accepted, "rejected": R.rejected,
			"dt": hx(R.dt), "status": status})
	return {"model": spec.name, "B": B, "T": hx(T), "ctrl": {k: hx(v) for k, v in ctrl.items()}, "rows": rows}


def sampled(spec, B, times):
	"""Repeated integrate() calls on a time grid (the way examples/noisy_lorenz.py:95-96 uses the integrator)."""
	mod = build_ref.load(spec, "strict")
	y0, seeds = ensemble_inputs(spec, B)
	rows = []
	for k in range(B):
		R = ref_driver.RefIntegrator(mod, y0[k], 0.0, int(seeds[k]))
		ys = [hx(R.integrate(float(T))) for T in times]
		rows.append({"ys": ys, "accepted": R.accepted, "rejected": R.rejected})
	return {"model": spec.name, "B": B, "times": hx(times), "rows": rows}


def jumpy(spec, B, T, tape_len):
	mod = build_ref.load(spec, "strict")
	y0, seeds = ensemble_inputs(spec, B)
	taus, xis = jump_tape(B, tape_len)
	rows = []
	for k in range(B):
		# amplitude law of examples/noisy_and_jumpy_lorenz.py:51-56 with the normal variate taken from the tape
		IJI, amp = ref_driver.tape_closures(taus[k], xis[k], lambda t, y, xi: np.array([0.0, 0.0, 0.0 + abs(y[2]) * xi]))
		R = ref_driver.RefJumpIntegrator(IJI, amp, mod, y0[k], 0.0, int(seeds[k]))
		R.integrate(T)
		rows.append({"y": hx(R.core.get_state()), "t": hx(R.core.t), "accepted": R.accepted, "rejected": R.rejected})
	return {"model": spec.name, "B": B, "T": hx(T), "tape_len": tape_len, "rows": rows}


def long_horizon():
	"""
	The configured horizons ("parity at the configured horizon is untested"): whole trajectories to T = 100
	for the additive (SRA, BASELINE.json configs[1]) and the multiplicative (SRI) Lorenz system, and configs[0] itself —
	examples/noisy_lorenz.py: y0 = default_rng(42).random(3), 10 000 integrate() calls on arange(0, 100, 0.01) — of
	which every 250th sample, the last one and the step counts are kept.
	"""
	g = {"trajectories": [trajectories(models.lorenz_additive, 8, 100.0), trajectories(models.lorenz_multiplicative, 8, 100.0)]}
	spec = models.lorenz_multiplicative
	mod = build_ref.load(spec, "strict")
	y0 = np.random.default_rng(seed=42).random(3)  # examples/noisy_lorenz.py:75,91
	times = np.arange(0.0, 100.0, 0.01)            # :95
	R = ref_driver.RefIntegrator(mod, y0, 0.0, 42)
	ys = np.array([R.integrate(float(T)) for T in times])
	keep = sorted(set(range(0, len(times), 250)) | {len(times) - 1})
	g["config1"] = {"model": spec.name, "y0": hx(y0), "seed": 42, "n_times": len(times), "kept": keep, "ys": hx(ys[keep]),
		"accepted": R.accepted, "rejected": R.rejected, "checksum_xor": format(int(np.bitwise_xor.reduce(ys.view(np.uint64).ravel())), "016x")}
	path = os.path.join(OUT, "reference_core_T100.json")
	with open(path, "w") as fh:
		json.dump(g, fh, indent=1)
	print("wrote", path, os.path.getsize(path), "bytes")


def main():
	os.makedirs(OUT, exist_ok=True)
	if len(sys.argv) > 1 and sys.argv[1] == "long":
		return long_horizon()
	g = {}
	g["rng"] = rng_goldens()
	lor = [1.0, 1.0, 20.0]
	bridge_script = [("step", 1e-3), ("step", 4e-4), ("accept",), ("step", 1e-3), ("accept",)]
	deep_script = [("pin", 3, 2e-4), ("step", 1e-3), ("step", 5e-4), ("step", 1e-4), ("accept",), ("step", 3e-4), ("accept",),
		("step", 2e-3), ("accept",), ("pin", 2, 0.5), ("step", 0.25), ("step", 1e-3), ("accept",)]
	g["scripted"] = [
		scripted(models.lorenz_additive, lor, 42, [("step", 1e-3), ("accept",)]),
		scripted(models.lorenz_additive, lor, 42, bridge_script),
		scripted(models.lorenz_multiplicative, lor, 42, [("step", 1e-3), ("accept",)]),
		scripted(models.lorenz_multiplicative, lor, 42, bridge_script),
		scripted(models.lorenz_additive_fixture, lor, 7, deep_script),
		scripted(models.lorenz_multiplicative, lor, 7, deep_script),
		scripted(models.duffing, [0.3, 0.7], 11, deep_script),
		scripted(models.ou_timedep, [0.3, 0.7], 11, deep_script),
	]
	g["trajectories"] = [
		trajectories(models.lorenz_additive, 16, 0.5),
		trajectories(models.lorenz_additive, 8, 10.0),
		trajectories(models.lorenz_additive_fixture, 8, 2.0),
		trajectories(models.lorenz_multiplicative, 16, 0.5),
		trajectories(models.lorenz_multiplicative, 8, 5.0),
		trajectories(models.gbm, 16, 10.0),
		trajectories(models.duffing, 8, 20.0),
		trajectories(models.ou, 8, 20.0),
		trajectories(models.ou_timedep, 8, 5.0),
		trajectories(models.cubic, 8, 2.0),
		# non-default controller parameters (incl. the clamps of _jitcsde.py:536-541)
		trajectories(models.lorenz_additive, 4, 1.0, atol=1e-6, rtol=1e-3, first_step=0.01, max_step=0.05),
		trajectories(models.lorenz_multiplicative, 4, 1.0, first_step=20.0, max_step=0.5, min_step=1.0),
		# failure path (tests/test_jitcsde.py:221-240): min_step too large
		trajectories(models.lorenz_multiplicative, 4, 1.0, min_step=1e-3, rtol=1e-10, atol=0.0),
	]
	g["sampled"] = [
		sampled(models.lorenz_additive, 4, np.arange(0.0, 1.0, 0.01)),
		sampled(models.lorenz_multiplicative, 4, np.arange(0.0, 0.5, 0.01)),
	]
	g["jumps"] = [jumpy(models.lorenz_jumpy, 8, 5.0, 32)]
	path = os.path.join(OUT, "reference_core.json")
	with open(path, "w") as fh:
		json.dump(g, fh, indent=1)
	print("wrote", path, os.path.getsize(path), "bytes")
	long_horizon()


if __name__ == "__main__":
	main()
