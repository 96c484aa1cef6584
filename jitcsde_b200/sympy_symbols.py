"""`t` and `y` as SymPy objects (reference: `jitcsde/sympy_symbols.py`).  This implementation is SymPy-based throughout, so
these are the package's own symbols; the module exists so that `import jitcsde.sympy_symbols`-style code keeps working."""
from ._codegen import t, y  # noqa: F401
