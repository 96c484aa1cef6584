"""
TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke()/build() and bench.py's
CPU-baseline legs may use anything under oracle/.

Builds the *unmodified* reference C core for one model: renders `/root/reference/jitcsde/jitced_template.c` with
jinja2 (the template variables of reference `_jitcsde.py:427-436`), writes `f.c`/`g.c` from the model's expression
strings in the form the reference's generator emits (`set_drift(i, expr);`, `_jitcsde.py:415-421`), puts
`random_numbers.c` beside it (`_jitcsde.py:438-440`) and compiles the lot with gcc into a CPython extension under
`oracle/_ref/`.  Reference sources are read where they lie and never copied into the repository; only the
compiled `.so` is kept (git-ignored, but shipped to the GPU box by gpurun).

Two flag sets (SURVEY.md §0.4, §8 c):
  strict  = -std=c11 -O2 -ffp-contract=off     parity target (machine-independent IEEE semantics)
  default = -std=c11 -Ofast -g0 -march=native -mtune=native -Wno-unknown-pragmas   the reference's shipping flags
            (jitcxde_common.DEFAULT_COMPILE_ARGS); speed baseline only.

/root/reference exists only in the build container.  On the GPU box the prebuilt modules are loaded if present.
"""

from __future__ import annotations

import importlib.util
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("JSDE_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")

FLAGS = {
	"strict": ["-std=c11", "-O2", "-ffp-contract=off"],
	"default": ["-std=c11", "-Ofast", "-g0", "-march=native", "-mtune=native", "-Wno-unknown-pragmas"],
	# the reference's omp=True path (`_jitcsde.py:331-334`: -fopenmp added to compile and link arguments): the template's
	# `#pragma omp parallel for` loops over components become active (jitced_template.c:132-557)
	"default_omp": ["-std=c11", "-Ofast", "-g0", "-march=native", "-mtune=native", "-fopenmp"],
}


def cpu_signature() -> str:
	"""Digest of the host CPU's feature flags: a -march=native build is only valid on a CPU with the same features."""
	import hashlib
	try:
		for line in open("/proc/cpuinfo"):
			if line.startswith("flags"):
				return hashlib.sha256(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()[:12]
	except OSError:
		pass
	return "unknown"


def module_name(spec, flavour: str) -> str:
	return f"jsderef_{spec.name}_{flavour}"


def reference_available() -> bool:
	return os.path.isfile(os.path.join(REF_ROOT, "jitcsde", "jitced_template.c"))


def _statements(kind: str, exprs, body: str) -> str:
	if body:
		return body
	return "".join(f"set_{kind}({i}, {e});\n" for i, e in enumerate(exprs))


def build(spec, flavour: str = "strict", force: bool = False) -> str:
	"""Compile the reference core for `spec`; returns the path of the extension module."""
	name = module_name(spec, flavour)
	target = os.path.join(OUT_DIR, name + ".so")
	stamp = target + ".digest"
	if not force and os.path.isfile(target) and os.path.isfile(stamp) and open(stamp).read() == spec.digest():
		return target
	if not reference_available():
		raise FileNotFoundError(f"reference sources not found under {REF_ROOT} and no prebuilt {target}")
	import jinja2
	import numpy

	os.makedirs(OUT_DIR, exist_ok=True)
	src_dir = os.path.join(REF_ROOT, "jitcsde")
	env = jinja2.Environment(loader=jinja2.FileSystemLoader(src_dir))
	rendered = env.get_template("jitced_template.c").render(
		n=spec.n,
		number_of_f_helpers=0,
		number_of_g_helpers=0,
		control_pars=list(spec.control_pars),
		additive=spec.additive,
		numpy_rng=False,
		chunk_size=100,
		callbacks=[],
		module_name=name,
	)
	with tempfile.TemporaryDirectory(prefix="jsderef_") as tmp:
		with open(os.path.join(tmp, name + ".c"), "w") as fh:
			fh.write(rendered)
		shutil.copy(os.path.join(src_dir, "random_numbers.c"), tmp)
		# control parameters: the reference substitutes `self->parameter_X` (`_jitcsde.py:388-391`)
		# (as scoped macros inside the two generated bodies, so that a parameter called e.g. `k` cannot touch the template)
		par_defs = "".join(f"#define {p} (self->parameter_{p})\n" for p in spec.control_pars)
		par_undefs = "".join(f"#undef {p}\n" for p in spec.control_pars)
		with open(os.path.join(tmp, "f_definitions.c"), "w") as fh:
			fh.write(spec.preamble)
		with open(os.path.join(tmp, "g_definitions.c"), "w") as fh:
			fh.write("")
		with open(os.path.join(tmp, "f.c"), "w") as fh:
			fh.write(par_defs + _statements("drift", spec.f, spec.f_body) + par_undefs)
		with open(os.path.join(tmp, "g.c"), "w") as fh:
			fh.write(par_defs + _statements("diffusion", spec.g, spec.g_body) + par_undefs)
		cmd = [
			os.environ.get("CC", "gcc"), "-shared", "-fPIC", *FLAGS[flavour],
			"-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
			os.path.join(tmp, name + ".c"), "-o", target, "-lm",
		]
		res = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
		if res.returncode != 0 and cmd[0] != "gcc" and shutil.which("gcc"):  # e.g. a $CC wrapper that lacks libgomp.spec
			cmd[0] = shutil.which("gcc")
			res = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
		if res.returncode != 0:
			raise RuntimeError("reference core failed to compile:\n" + res.stderr[-4000:])
	with open(stamp, "w") as fh:
		fh.write(spec.digest())
	with open(target + ".cpu", "w") as fh:
		fh.write(cpu_signature())
	return target


def load(spec, flavour: str = "strict"):
	"""Import the (pre)built reference core for `spec`; builds it when the reference sources are present."""
	name = module_name(spec, flavour)
	target = os.path.join(OUT_DIR, name + ".so")
	if reference_available():
		target = build(spec, flavour)
	elif not os.path.isfile(target):
		raise FileNotFoundError(f"no prebuilt reference core {target}")
	if name in sys.modules:
		return sys.modules[name]
	loader = importlib.machinery.ExtensionFileLoader(name, target)
	spec_ = importlib.util.spec_from_loader(name, loader)
	module = importlib.util.module_from_spec(spec_)
	loader.exec_module(module)
	sys.modules[name] = module
	return module


def available(spec, flavour: str = "strict") -> bool:
	if reference_available():
		return True
	target = os.path.join(OUT_DIR, module_name(spec, flavour) + ".so")
	if not os.path.isfile(target):
		return False
	if flavour.startswith("default"):  # built with -march=native: usable only on a CPU with the same feature set
		try:
			return open(target + ".cpu").read() == cpu_signature()
		except OSError:
			return False
	return True
