"""CPU-side checks of the boundary: the CUDA library builds with nvcc for sm_100a, loads, and exports every symbol that
include/jsde.h declares.  No compute call is made (there is no GPU here); creating a handle must fail loudly."""

import ctypes
import os
import re

import pytest

from jitcsde_b200 import _abi, _build, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
	text = open(os.path.join(ROOT, "include", "jsde.h")).read()
	text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
	return sorted(set(re.findall(r"\b(jsde_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
	assert sorted(_abi.ABI_SYMBOLS) == declared_symbols()


@pytest.mark.parametrize("name", ["lorenz_additive", "lorenz_jumpy", "gbm_pars"])
def test_library_exports_every_declared_symbol(name):
	spec = models.ALL[name]
	path = _build.build(spec)
	lib = ctypes.CDLL(path)
	for sym in declared_symbols():
		assert hasattr(lib, sym), f"{sym} missing from {path}"
	lib.jsde_model_name.restype = ctypes.c_char_p
	assert lib.jsde_model_name().decode() == name
	n, add, npar, jump = (ctypes.c_int() for _ in range(4))
	lib.jsde_model_info(ctypes.byref(n), ctypes.byref(add), ctypes.byref(npar), ctypes.byref(jump))
	assert (n.value, bool(add.value), npar.value, bool(jump.value)) == (spec.n, spec.additive, len(spec.control_pars), spec.jump is not None)


def test_no_cpu_fallback():
	import torch
	if torch.cuda.is_available():
		pytest.skip("a GPU is present")
	with pytest.raises(_abi.JsdeError, match="no CUDA device|CUDA"):
		_abi.Ensemble(models.lorenz_additive, 4)


def test_built_for_sm100a_without_contraction():
	assert "arch=compute_100a,code=sm_100a" in _build.NVCC_FLAGS
	assert "-fmad=false" in _build.NVCC_FLAGS
