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
