"""Small end-to-end run for compute-sanitizer (development aid)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from util import jump_tape  # noqa
for name, B, T in (("lorenz_additive", 200, 0.3), ("lorenz_multiplicative", 100, 0.2), ("duffing", 70, 0.5), ("gbm", 33, 1.0)):
	spec = models.ALL[name]
	E = Ensemble(spec, B)
	E.seed(sharding.global_seeds(0, B))
	E.set_state(sharding.global_initial_states(0, B, spec.n), 0.0)
	E.set_integration_parameters(ControllerParameters())
	print(name, E.integrate(T)["accepted"], E.integrate_grid(np.linspace(T, 2 * T, 5)).shape)
	E.pin_noise(2, 1e-3); E.get_next_step(1e-3); E.accept_step(); E.draw_gauss(1.0, 5)
spec = models.lorenz_jumpy
E = Ensemble(spec, 64); E.seed(np.arange(64)); E.set_state(sharding.global_initial_states(0, 64, 3), 0.0)
taus, xis = jump_tape(64, 8)
print("jumpy", E.integrate_jumps(1.0, taus, xis)["jumps"])
