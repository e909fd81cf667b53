"""rlsolver_b200 -- B200-native (sm_100a) hot path of RLSolver's parallel max-cut / QUBO environments.

Only the environment hot path lives here (graph store, cut evaluation, flip / local-search
kernels, select ops and their Python mirrors of the reference classes).  See DESIGN.md.
"""
from . import _lib

__all__ = ["build", "lib"]
build = _lib.build
lib = _lib.lib
