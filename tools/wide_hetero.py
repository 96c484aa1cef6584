"""Development aid: distribution of attempts per trajectory for the n=1000 network (python tools/wide_hetero.py T)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models, sharding
from jitcsde_b200._abi import ControllerParameters, Ensemble
T = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
B = 2048
spec = models.fhn_1000()
E = Ensemble(spec, B)
E.seed(sharding.global_seeds(0, B))
E.set_state(sharding.global_initial_states(0, B, spec.n), 0.0)
E.set_integration_parameters(ControllerParameters())
st = E.integrate(T)
dt, acc, rej = E.controller_state()
att = (acc + rej).astype(float)
print(f"T={T}: kernel {st['kernel_ms']:.0f} ms, attempts/s {att.sum()/st['kernel_ms']*1e3:.3e}; attempts per trajectory mean {att.mean():.0f} std {att.std():.0f} min {att.min():.0f} max {att.max():.0f}")
for tpb in (2, 4, 7, 8):
	n = B // tpb * tpb
	g = att[:n].reshape(-1, tpb)
	print(f"  blocks of {tpb}: mean of block-max / mean = {g.max(axis=1).mean()/att.mean():.3f}; max/mean {att.max()/att.mean():.3f}")
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
if os.path.isdir(out):
	np.save(os.path.join(out, "wide_attempts.npy"), att)
