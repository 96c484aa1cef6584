"""Development aid: results must not depend on lane recycling nor on what the handle did before (python tools/recycle_check.py [log2B] [T])"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 19)
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
spec = models.lorenz_additive
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # global index of trajectory 0
y0 = sharding.global_initial_states(first, first + B, 3)
seeds = sharding.global_seeds(first, first + B)
def run(E, recycle):
	E.set_option("recycle", recycle)
	E.seed(seeds)
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	st = E.integrate(T)
	return st["accepted"], st["rejected"], E.get_state().copy(), E.controller_state()[1].copy()
E = Ensemble(spec, B)
res = {}
for tag, rec in (("same handle, recycle=1", "1"), ("same handle, recycle=0", "0")):
	res[tag] = run(E, rec)
E.close()
for tag, rec in (("fresh handle, recycle=0", "0"),):
	F = Ensemble(spec, B)
	res[tag] = run(F, rec)
	F.close()
ref = res["fresh handle, recycle=0"]
for tag, (a, r, y, acc) in res.items():
	diff = np.flatnonzero((y.view(np.uint64) != ref[2].view(np.uint64)).any(axis=1))
	print(f"{tag:32s} accepted {a} rejected {r}  trajectories differing from 'fresh, recycle=0': {diff.size} {diff[:8].tolist()}")
