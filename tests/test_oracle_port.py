"""
Pins the oracle.  The C restatement (oracle/sde_port.c) must reproduce, bit for bit,
  (1) the golden vectors of SURVEY.md §8(c) (strict build of the reference core, recorded at survey time),
  (2) the committed fixtures in tests/golden/reference_core.json (generated from the unmodified reference C core by
      oracle/make_golden.py), and
  (3) where the reference core itself is available (build container: /root/reference; GPU box: prebuilt oracle/_ref),
      whole trajectories of that core driven by the transcribed controller.
"""

import numpy as np
import pytest

from jitcsde_b200 import models
from oracle import build_ref, port, ref_driver
from util import assert_bits_equal, ensemble_inputs, golden, jump_tape, unhex

G = golden()


# ---- (1) survey goldens ---------------------------------------------------------------------------------------
def test_survey_mt19937_words():
	for seed, words in {0: [2357136044, 2546248239, 3071714933, 3626093760], 42: [1608637542, 3421126067, 4083286876, 787846414],
			4294967295: [419326371, 479346978, 3918654476, 2416749639]}.items():
		P = port.PortCore(models.ou, [0, 0], 0.0, seed)
		assert list(P.random_words(4)) == words


def test_survey_doubles_and_pairs():
	P = port.PortCore(models.ou, [0, 0], 0.0, 0)
	assert [x.hex() for x in P.random_doubles(2)] == ["0x1.18fe1565f12a8p-1", "0x1.6e2d4cf608733p-1"]
	P = port.PortCore(models.ou, [0, 0], 0.0, 42)
	assert [x.hex() for x in P.random_doubles(2)] == ["0x1.7f8771e5f51ecp-2", "0x1.e6c4068bbd654p-1"]
	for seed, (dw, dz, pos) in {0: ("0x1.c398ef3e5cfa5p+0", "0x1.99c2cfacc8953p-2", 4), 42: ("0x1.fca2a28a9307cp-2", "-0x1.1b2a505de052ap-3", 4),
			4294967295: ("0x1.4bfc38c486d8ap-1", "0x1.56b192e507eeep-1", 8)}.items():
		P = port.PortCore(models.ou, [0, 0], 0.0, seed)
		pair = P.draw_gauss(1.0, 1)[0]
		assert (pair[0].hex(), pair[1].hex(), P.rng_pos) == (dw, dz, pos)
	P = port.PortCore(models.ou, [0, 0], 0.0, 42)
	pairs = P.draw_gauss(0.1, 3)
	assert [(a.hex(), b.hex()) for a, b in pairs] == [("0x1.96e88208759fcp-5", "-0x1.c510809633b76p-7"),
		("0x1.094b10ce9e520p-4", "0x1.37eaa0b348822p-3"), ("-0x1.7fa30b2ac8500p-6", "-0x1.7f9c2852715c9p-6")]
	P = port.PortCore(models.ou, [0, 0], 0.0, 42)
	last = P.draw_gauss(1.0, 1000)[-1]
	assert (last[0].hex(), last[1].hex(), P.rng_pos) == ("-0x1.4df60d6992629p-3", "-0x1.7d63e1244b92ap-1", 104)


def test_survey_lorenz_steps():
	P = port.PortCore(models.lorenz_additive, [1, 1, 20], 0.0, 42)
	P.get_next_step(1e-3)
	assert P.get_p(1e-10, 1e-5).hex() == "0x1.0505b1bfdca22p+0"
	P.accept_step()
	assert [x.hex() for x in P.get_state()] == ["0x1.00699e24b4928p+0", "0x1.02528acd53d8dp+0", "0x1.3f26ea8db7e3bp+4"]
	assert P.t.hex() == "0x1.0624dd2f1a9fcp-10"
	P = port.PortCore(models.lorenz_multiplicative, [1, 1, 20], 0.0, 42)
	P.get_next_step(1e-3)
	assert P.get_p(1e-10, 1e-5).hex() == "0x1.4ce7ca0a807d5p+1"
	P.accept_step()
	assert [x.hex() for x in P.get_state()] == ["0x1.006a2327cfb8bp+0", "0x1.02530fad03a7ap+0", "0x1.3eed12319924ep+4"]


# ---- (2) committed fixtures -----------------------------------------------------------------------------------
def test_golden_rng_stream():
	spec = models.ModelSpec("unit_noise", 1, True, ("0.0",), ("1.0",))
	for seed, rec in G["rng"].items():
		P = port.PortCore(spec, [0.0], 0.0, int(seed))
		dws = P.draw_gauss(1.0, 2000)[:, 0]
		assert_bits_equal(dws[:8], unhex(rec["dw_scale1_first8"]), f"seed {seed} first 8")
		assert_bits_equal(dws[1992:], unhex(rec["dw_scale1_1992_2000"]), f"seed {seed} last 8")


@pytest.mark.parametrize("case", range(len(G["scripted"])))
def test_golden_scripted(case):
	rec = G["scripted"][case]
	spec = models.ALL[rec["model"]]
	P = port.PortCore(spec, unhex(rec["y0"]), 0.0, rec["seed"])
	for step in rec["script"]:
		if step["do"] == "step":
			P.get_next_step(unhex(step["h"]))
			assert P.get_p(1e-10, 1e-5).hex() == float.fromhex(step["p"]).hex()
		elif step["do"] == "accept":
			P.accept_step()
			assert_bits_equal(P.get_state(), unhex(step["y"]), "state")
			assert P.t == unhex(step["t"])
		else:
			P.pin_noise(step["number"], unhex(step["step"]))


@pytest.mark.parametrize("case", range(len(G["trajectories"])))
def test_golden_trajectories(case):
	rec = G["trajectories"][case]
	spec = models.ALL[rec["model"]]
	y0, seeds = ensemble_inputs(spec.n, rec["B"])
	ctrl = {k: unhex(v) for k, v in rec["ctrl"].items()}
	res = port.run_ensemble(spec, y0, seeds, 0.0, unhex(rec["T"]), **ctrl)
	for k, row in enumerate(rec["rows"]):
		assert res["status"][k] == row["status"]
		assert (res["accepted"][k], res["rejected"][k]) == (row["accepted"], row["rejected"]), f"trajectory {k} step counts"
		assert_bits_equal(res["y"][k], unhex(row["y"]), f"trajectory {k}")
		assert res["t"][k] == unhex(row["t"])


@pytest.mark.parametrize("case", range(len(G["sampled"])))
def test_golden_sampled(case):
	rec = G["sampled"][case]
	spec = models.ALL[rec["model"]]
	y0, seeds = ensemble_inputs(spec.n, rec["B"])
	times = unhex(rec["times"])
	for k, row in enumerate(rec["rows"]):
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		ys = np.array([P.integrate(T) for T in times])
		assert_bits_equal(ys, unhex(row["ys"]), f"trajectory {k}")
		c = P.counters()
		assert (c["accepted"], c["rejected"]) == (row["accepted"], row["rejected"])


def test_golden_jumps():
	rec = G["jumps"][0]
	spec = models.ALL[rec["model"]]
	y0, seeds = ensemble_inputs(spec.n, rec["B"])
	taus, xis = jump_tape(rec["B"], rec["tape_len"])
	res = port.run_ensemble(spec, y0, seeds, 0.0, unhex(rec["T"]), taus=taus, xis=xis)
	for k, row in enumerate(rec["rows"]):
		assert_bits_equal(res["y"][k], unhex(row["y"]), f"trajectory {k}")
		assert (res["accepted"][k], res["rejected"][k]) == (row["accepted"], row["rejected"])
		assert res["t"][k] == unhex(row["t"])


# ---- (3) the reference core itself, when it can be loaded -------------------------------------------------------
@pytest.mark.parametrize("name", ["lorenz_additive", "lorenz_multiplicative", "duffing"])
def test_against_live_reference_core(name):
	spec = models.ALL[name]
	if not build_ref.available(spec, "strict"):
		pytest.skip("reference core neither buildable nor prebuilt here")
	mod = build_ref.load(spec, "strict")
	y0, seeds = ensemble_inputs(spec.n, 3)
	for k in range(3):
		R = ref_driver.RefIntegrator(mod, y0[k], 0.0, int(seeds[k]) + 1000)
		yr = R.integrate(3.0)
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]) + 1000)
		yp = P.integrate(3.0)
		assert_bits_equal(yp, yr, f"{name} trajectory {k}")
		c = P.counters()
		assert (c["accepted"], c["rejected"]) == (R.accepted, R.rejected)
		assert P.dt == R.dt


def test_state_dependent_jump_intensity_against_live_reference_core():
	"""JumpSpec.iji: the port's compiled waiting-time expression == the same law fed to the reference's IJI(time, state)
	callback (`_jitcsde.py:682-700`), bit for bit."""
	spec = models.lorenz_jumpy_rate
	base = models.lorenz_jumpy  # same SDE: the reference core has no notion of jumps, its Python driver has
	if not build_ref.available(base, "strict"):
		pytest.skip("reference core neither buildable nor prebuilt here")
	mod = build_ref.load(base, "strict")
	y0, seeds = ensemble_inputs(spec.n, 3)
	taus, xis = jump_tape(3, 24)
	for k in range(3):
		IJI, amp = ref_driver.tape_closures(taus[k], xis[k], lambda t, y, xi: np.array([0.0, 0.0, 0.0 + abs(y[2]) * xi]),
			iji_fn=lambda t, y, tau: tau / (1.0 + y[2] * y[2] / 400.0) + 0.001 * t)
		R = ref_driver.RefJumpIntegrator(IJI, amp, mod, y0[k], 0.0, int(seeds[k]))
		R.integrate(2.0)
		yr = R.integrate(4.0)
		P = port.PortCore(spec, y0[k], 0.0, int(seeds[k]))
		P.integrate_jumps(2.0, taus[k], xis[k])
		P.integrate_jumps(4.0, taus[k], xis[k])
		assert_bits_equal(P.get_state(), yr, f"trajectory {k}")
		c = P.counters()
		assert (c["accepted"], c["rejected"]) == (R.accepted, R.rejected)


def test_pinned_path_is_consumed_exactly():
	"""port.pin_path (the checker of jsde_pin_noise_path): steps that coincide with the pinned items consume them whole,
	so a noise-driven-only SDE (f = 0, g = 1) reproduces the path's partial sums exactly."""
	spec = models.ModelSpec("unit_noise3", 1, True, ("0.0",), ("1.0",))
	rng = np.random.default_rng(3)
	steps = rng.uniform(0.01, 0.1, 6)
	dw = rng.standard_normal((6, 1)) * np.sqrt(steps)[:, None]
	dz = rng.standard_normal((6, 1)) * np.sqrt(steps)[:, None]
	P = port.PortCore(spec, [0.0], 0.0, 1)
	P.pin_path(steps, dw, dz)
	q = P.dump_queue()
	assert_bits_equal(q[:, 0], steps, "lengths")
	assert_bits_equal(q[:, 1], dw[:, 0], "DW")
	total = 0.0
	for j, h in enumerate(steps):
		P.get_next_step(float(h))
		P.accept_step()
		total = total + dw[j, 0]
		assert P.get_state()[0] == total
	assert len(P.dump_queue()) == 0


# ---- noise-memory invariants (reference tests/test_python_core.py:31-61, restated for the port's queue) -------
def test_noise_memory_invariants():
	spec = models.ModelSpec("unit_noise2", 1, True, ("0.0",), ("1.0",))
	rng = np.random.default_rng(42)
	P = port.PortCore(spec, [0.0], 0.0, 5)
	hs = rng.exponential(size=10)
	for h in hs:
		P.pin_noise(1, h)
	before = P.dump_queue()
	total_dw = before[:, 1].sum()
	# a request inside the pinned range inserts a bridge but leaves cumulative sums (almost) unchanged
	P.get_next_step(0.37 * hs.sum())
	first = P.get_p(1.0, 0.0)
	P.get_next_step(0.37 * hs.sum())
	assert P.get_p(1.0, 0.0) == first  # idempotent: same noise served twice
	after = P.dump_queue()
	assert len(after) == len(before) + 1
	np.testing.assert_allclose(after[:, 0].sum(), hs.sum(), rtol=1e-14)
	np.testing.assert_allclose(after[:, 1].sum(), total_dw, rtol=1e-12, atol=1e-12)
	# accepting drops exactly the consumed length
	P.accept_step()
	rest = P.dump_queue()
	np.testing.assert_allclose(rest[:, 0].sum(), (1 - 0.37) * hs.sum(), rtol=1e-12)
