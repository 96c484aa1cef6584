"""Development aid: throughput of the jumpy-Lorenz configuration (BASELINE.json configs[3]) with tape-free jumps."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
spec = models.lorenz_jumpy
E = Ensemble(spec, B)
E.seed(sharding.global_seeds(0, B))
for rep in range(2):
	E.set_state(sharding.global_initial_states(0, B, 3), 0.0)
	E.set_integration_parameters(ControllerParameters())
	st = E.integrate_jumps_generated(T, 42)
	print(f"lorenz_jumpy B={B} T={T}: {st['accepted']} acc {st['rejected']} rej {st['jumps']} jumps, kernel {st['kernel_ms']:.1f} ms, "
		f"{st['accepted']/st['kernel_ms']*1e3:.3e} accepted steps/s, failed {st['n_min_step']}+{st['n_overflow']}, depth {st['max_noise_depth']}")
