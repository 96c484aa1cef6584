"""Development aid: compile, here (nvcc cross-compiles without a GPU), the models that the GPU tests build from symbolic
input, so that the libraries travel with the tree and the tests on the GPU box find them in jitcsde_b200/_cache."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_reference_suite as S

for name, factory in {**S.MULTIPLICATIVE, **S.ADDITIVE}.items():
	print(name, os.path.relpath(factory().compile_C(), ROOT))
for options in (dict(simplify=False), dict(do_cse=True), dict(simplify=True)):
	print(options, os.path.relpath(S.MULTIPLICATIVE["default"]().compile_C(**options), ROOT))
print("jumpy", os.path.relpath(S.jumpy().compile_C(), ROOT))
import sympy
import test_gpu_wide as W
print("symbolic network", os.path.relpath(W.symbolic_network(48, 11).compile_C(), ROOT))
I_ext, sigma = sympy.symbols("I_ext sigma")
print("symbolic network with parameters", os.path.relpath(W.symbolic_network(20, 6, I_ext, sigma, control_pars=[I_ext, sigma]).compile_C(), ROOT))
