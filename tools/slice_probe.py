"""Development aid: the strong-scaling shape on one GPU — B trajectories (default 2^17 = a 2^20 ensemble split over eight
GPUs, 2.3 waves of resident lanes) to T, with recycling off / on / on with time slicing.
python tools/slice_probe.py [model] [log2B] [T]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from jitcsde_b200 import models
from jitcsde_b200._abi import ControllerParameters, Ensemble

name = sys.argv[1] if len(sys.argv) > 1 else "lorenz_additive"
B = 1 << int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
T = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
spec = models.ALL[name]
y0 = np.random.default_rng(42).random((B, spec.n))
E = Ensemble(spec, B)
reference = None
for recycle, slice_rounds in (("0", 0), ("1", 0), ("1", 16384), ("1", 4096), ("1", 1024), ("0", 0)):
	E.set_option("recycle", recycle)
	E.set_option("slice", str(slice_rounds))
	E.seed(np.arange(B, dtype=np.uint32))
	E.set_state(y0, 0.0)
	E.set_integration_parameters(ControllerParameters())
	st = E.integrate(T)
	y = E.get_state()
	if reference is None:
		reference = y
	same = (y.view(np.uint64) == reference.view(np.uint64)).all()
	print(f"{name} B={B} T={T} recycle={recycle} slice={slice_rounds}: kernel {st['kernel_ms']:.1f} ms, "
		f"{st['accepted'] / st['kernel_ms'] * 1e3:.4e} accepted steps/s, bits equal to the first run: {same}", flush=True)
