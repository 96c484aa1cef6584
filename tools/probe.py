"""Quick throughput probe (development aid, not the benchmark): python tools/probe.py [model] [log2B] [T]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import _abi, models
from jitcsde_b200._abi import ControllerParameters, Ensemble

name = sys.argv[1] if len(sys.argv) > 1 else "lorenz_additive"
B = 1 << int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
T = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
T0 = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0   # untimed lead-in (skips the first_step=1.0 transient)
spec = models.fhn_1000() if name == 'fhn_1000' else models.ALL[name]
lib = _abi.load_library(spec)
print("fp64 peak TFLOP/s:", _abi.measure_fp64_peak(lib) if not spec.wide else "n/a (wide library)")
y0 = np.random.default_rng(42).random((B, spec.n))
E = Ensemble(spec, B)
E.seed(np.arange(B, dtype=np.uint32))
for rep in range(2):
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	if T0 > 0:
		E.integrate(T0)
	t0 = time.time()
	st = E.integrate(T0 + T)
	wall = time.time() - t0
	print(f"{name} B={B} T={T}: {st['accepted']} acc {st['rejected']} rej, kernel {st['kernel_ms']:.1f} ms (wall {wall*1e3:.1f}), "
		f"{st['accepted']/st['kernel_ms']*1e3:.3e} accepted steps/s, {(st['accepted']+st['rejected'])/st['kernel_ms']*1e3:.3e} attempts/s, depth {st['max_noise_depth']}")
