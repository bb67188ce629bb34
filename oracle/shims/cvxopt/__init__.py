"""Stand-in for the `cvxopt` package, ORACLE ONLY (test infrastructure).

cvxopt is a third-party dependency of robotarium_python_simulator (rps) and is NOT in
/root/reference nor installable here (no network).  This package restates, from the published
algorithm (Andersen, Dahl, Vandenberghe, "CVXOPT: cone programming" - coneprog.coneqp, misc.py,
kktsolver 'chol2'), exactly the slice rps uses: solvers.qp(P, q, G, h) with only a
componentwise ('l') cone, no equality constraints, default initial point, refinement = 0.

PARITY UNPINNED: this restatement could not be executed against real cvxopt in this environment.
"""
import numpy as np
from . import solvers  # noqa: F401
from . import blas  # noqa: F401


def matrix(x, size=None, tc="d"):
    """Dense column-major matrix -> float64 ndarray; 1-D input becomes a column vector (cvxopt rule)."""
    a = np.array(x, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if size is not None:
        a = a.reshape(size, order="F")
    return a


def sparse(x, tc="d"):
    """rps only wraps 2*I in sparse(); keep it dense, the algebra is identical."""
    return np.array(x, dtype=np.float64)
