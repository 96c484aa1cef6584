"""Development aid: stress the yield protocol of the recycling integrate kernel (kernels.cuh) — tiny slices, odd grid sizes,
repeated runs — and compare every run with the plain grid bit for bit.  python tools/slice_stress.py [repeats]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models
from jitcsde_b200._abi import ControllerParameters, Ensemble

repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bad = runs = 0
for name, B, T in (("lorenz_multiplicative", 6007, 0.3), ("lorenz_additive", 9001, 0.4), ("lorenz_jumpy", 3001, 0.5), ("gbm", 8191, 2.0), ("duffing", 5003, 1.0)):
	spec = models.ALL[name]
	y0 = np.random.default_rng(1).random((B, spec.n))
	seeds = np.arange(B, dtype=np.uint32)
	want = None
	for resident in (0, 3, 17, 40):
		if resident:
			os.environ["JSDE_RESIDENT_BLOCKS"] = str(resident)
		else:
			os.environ.pop("JSDE_RESIDENT_BLOCKS", None)
		E = Ensemble(spec, B)
		for slice_rounds in ((0,) if not resident else (1, 2, 3, 5, 8, 13, 100)):
			for rep in range(1 if not resident else repeats):
				E.set_option("recycle", "1" if resident else "0")
				E.set_option("slice", str(slice_rounds))
				E.seed(seeds)
				E.set_state(y0, 0.0)
				E.set_integration_parameters(ControllerParameters())
				if spec.jump is not None:
					st = E.integrate_jumps_generated(T, 5, 0)
				else:
					st = E.integrate(T)
				got = (E.get_state().copy(), E.t.copy(), st["accepted"], st["rejected"], st.get("jumps", 0))
				if want is None:
					want = got
				same = all(np.array_equal(np.asarray(a).view(np.uint64) if isinstance(a, np.ndarray) else a, np.asarray(b).view(np.uint64) if isinstance(b, np.ndarray) else b) for a, b in zip(got, want))
				runs += 1
				if not same:
					bad += 1
					print(f"MISMATCH {name} resident={resident} slice={slice_rounds} rep={rep}", flush=True)
		E.close()
	print(f"{name}: done", flush=True)
print(f"{runs} runs, {bad} mismatches")
