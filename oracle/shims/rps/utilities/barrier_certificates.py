"""rps.utilities.barrier_certificates stand-in (oracle only).  SURVEY.md App. A.8.

Assembles the pairwise single-integrator barrier QP exactly as rps does (pair order i<j
lexicographic, rows -2e / +2e, b = gain*h^3, pre-clip of dxi IN PLACE, H = 2I, f = -2 vec_F(dxi))
and hands it to cvxopt.solvers.qp.  Importing this module sets cvxopt's global options the way rps
does (show_progress False, reltol 1e-2, feastol 1e-2, maxiters 50).
"""
from cvxopt import matrix, sparse
from cvxopt.solvers import qp, options
import numpy as np
from scipy.special import comb

from rps.utilities.transformations import *  # noqa: F401,F403

options['show_progress'] = False
options['reltol'] = 1e-2
options['feastol'] = 1e-2
options['maxiters'] = 50


def _solve(dxi, x, safety_radius, gain_of_h, magnitude_limit):
    N = dxi.shape[1]
    num_constraints = int(comb(N, 2))
    A = np.zeros((num_constraints, 2 * N))
    b = np.zeros(num_constraints)
    H = sparse(matrix(2 * np.identity(2 * N)))

    count = 0
    for i in range(N - 1):
        for j in range(i + 1, N):
            error = x[:, i] - x[:, j]
            h = (error[0] * error[0] + error[1] * error[1]) - np.power(safety_radius, 2)
            A[count, (2 * i, (2 * i + 1))] = -2 * error
            A[count, (2 * j, (2 * j + 1))] = 2 * error
            b[count] = gain_of_h(h) * np.power(h, 3)
            count += 1

    norms = np.linalg.norm(dxi, 2, 0)
    idxs_to_normalize = (norms > magnitude_limit)
    dxi[:, idxs_to_normalize] *= magnitude_limit / norms[idxs_to_normalize]

    f = -2 * np.reshape(dxi, 2 * N, order='F')
    result = qp(H, matrix(f), matrix(A), matrix(b))['x']
    return np.reshape(result, (2, -1), order='F')


def create_single_integrator_barrier_certificate(barrier_gain=100, safety_radius=0.17, magnitude_limit=0.2):
    def f(dxi, x):
        return _solve(dxi, x, safety_radius, lambda h: barrier_gain, magnitude_limit)
    return f


def create_single_integrator_barrier_certificate2(barrier_gain=100, unsafe_barrier_gain=1e6,
                                                  safety_radius=0.17, magnitude_limit=0.2):
    def f(dxi, x):
        return _solve(dxi, x, safety_radius,
                      lambda h: barrier_gain if h >= 0 else unsafe_barrier_gain, magnitude_limit)
    return f
