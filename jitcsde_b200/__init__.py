"""
B200-native ensemble SDE integrator behind the JiTCSDE API (reference `jitcsde/__init__.py:1`).

    from jitcsde_b200 import jitcsde, y, t
    SDE = jitcsde(f_sym, g_sym, ensemble=2**20)      # same constructor as the reference + an ensemble dimension
    SDE.set_seed(0); SDE.set_initial_value(y0, 0.0)
    states = SDE.integrate(100.0)                    # ndarray [B][n]

The CUDA library is compiled at run time with nvcc for sm_100a; there is no CPU fallback.
"""

from ._jitcsde import UnsuccessfulIntegration, jitcsde, jitcsde_jump, t, tau, test, xi, y  # noqa: F401
from .modelspec import JumpSpec, ModelSpec  # noqa: F401

__all__ = ["jitcsde", "jitcsde_jump", "y", "t", "xi", "tau", "UnsuccessfulIntegration", "test", "ModelSpec", "JumpSpec"]
