"""cvxopt.solvers stand-in: `qp` -> `coneqp` for the pure-inequality ('l' cone) case.  ORACLE ONLY.

Follows cvxopt's coneprog.coneqp step by step in its Nesterov-Todd scaled variables
(W['d'] = sqrt(s/z), lmbda = sqrt(s*z)) with the 'chol2' KKT solver
(S = P + G' diag(d)^-2 G, dense Cholesky), Mehrotra predictor/corrector with
STEP = 0.99, EXPON = 3, default starting point, and cvxopt's stopping rule.  SURVEY.md App. A.9.

Module-level `options` mirrors cvxopt.solvers.options (rps mutates it at import:
show_progress False, reltol 1e-2, feastol 1e-2, maxiters 50; abstol keeps its default 1e-7).
"""
import math
import numpy as np

options = {}

STEP = 0.99
EXPON = 3

# statistics of the most recent call (not part of cvxopt; used by the oracle tests)
last = {"iterations": 0, "status": None}


def _chol_solve(S, b):
    L = np.linalg.cholesky(S)          # raises LinAlgError when not PD (cvxopt: ArithmeticError)
    y = np.linalg.solve(L, b)
    return np.linalg.solve(L.T, y)


def coneqp_l(P, q, G, h, opts=None):
    opts = options if opts is None else opts
    MAXITERS = opts.get("maxiters", 100)
    ABSTOL = opts.get("abstol", 1e-7)
    RELTOL = opts.get("reltol", 1e-6)
    FEASTOL = opts.get("feastol", 1e-7)

    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    n = q.size
    m = h.size
    if m == 0:
        x = _chol_solve(P, -q)
        return {"x": x.reshape(-1, 1), "status": "optimal", "iterations": 0}

    resx0 = max(1.0, math.sqrt(float(q @ q)))
    resz0 = max(1.0, math.sqrt(float(h @ h)))

    # ---- default starting point: [P G'; G -I][x; z] = [-q; h], s = -z, shifted into the cone
    try:
        x = _chol_solve(P + G.T @ G, -q + G.T @ h)
    except np.linalg.LinAlgError:
        raise ValueError("Rank(A) < p or Rank([P; A; G]) < n")
    z = G @ x - h
    s = -z
    nrms = math.sqrt(float(s @ s))
    ts = float(np.max(-s))
    if ts >= -1e-8 * max(nrms, 1.0):
        s = s + (1.0 + ts)
    nrmz = math.sqrt(float(z @ z))
    tz = float(np.max(-z))
    if tz >= -1e-8 * max(nrmz, 1.0):
        z = z + (1.0 + tz)

    gap = float(s @ z)
    d = None
    lmbda = None

    for iters in range(MAXITERS + 1):
        # residuals and costs
        rx = P @ x + q
        f0 = 0.5 * (float(x @ rx) + float(x @ q))
        rx = rx + G.T @ z
        resx = math.sqrt(float(rx @ rx))
        rz = s + G @ x - h
        resz = math.sqrt(float(rz @ rz))

        pcost = f0
        dcost = f0 + float(z @ rz) - gap
        if pcost < 0.0:
            relgap = gap / -pcost
        elif dcost > 0.0:
            relgap = gap / dcost
        else:
            relgap = None
        pres = resz / resz0
        dres = resx / resx0

        if (pres <= FEASTOL and dres <= FEASTOL and
                (gap <= ABSTOL or (relgap is not None and relgap <= RELTOL))) or iters == MAXITERS:
            status = "unknown" if iters == MAXITERS else "optimal"   # cvxopt tests MAXITERS first
            last["iterations"], last["status"] = iters, status
            return {"x": x.reshape(-1, 1), "s": s.reshape(-1, 1), "z": z.reshape(-1, 1),
                    "status": status, "iterations": iters, "gap": gap}

        if iters == 0:
            d = np.sqrt(s / z)           # misc.compute_scaling, 'l' block
            lmbda = np.sqrt(s * z)
        lmbdasq = lmbda * lmbda

        # kkt_chol2 factor: S = P + Gs'Gs, Gs = diag(1/d) G
        Gs = G / d[:, None]
        try:
            L = np.linalg.cholesky(P + Gs.T @ Gs)
        except np.linalg.LinAlgError:
            if iters == 0:
                raise ValueError("Rank(A) < p or Rank([P; A; G]) < n")
            last["iterations"], last["status"] = iters, "unknown"
            return {"x": x.reshape(-1, 1), "s": s.reshape(-1, 1), "z": z.reshape(-1, 1),
                    "status": "unknown", "iterations": iters, "gap": gap}

        def f4(bx, bz, bs):
            # f4_no_ir + kkt_chol2.solve; returns (ux, W*uz, W^-T*us) i.e. scaled dz, ds
            bs = bs / lmbda                      # sinv
            bz = bz - d * bs                     # z := z - W'*(lmbda o\ bs)
            bz = bz / d                          # scale(z, W, trans='T', inverse='I')
            bx = bx + Gs.T @ bz
            ux = np.linalg.solve(L.T, np.linalg.solve(L, bx))
            uz = Gs @ ux - bz
            us = bs - uz
            return ux, uz, us

        mu = gap / m
        sigma = 0.0
        dsdz_aff = None
        for i in (0, 1):
            bs = -lmbdasq + sigma * mu
            if i == 1:
                bs = bs - dsdz_aff
            dx, dz, ds = f4(-rx, -rz, bs)
            if i == 0:
                dsdz = float(ds @ dz)
                dsdz_aff = ds * dz               # saved for the Mehrotra correction
            ds = ds / lmbda                      # scale2
            dz = dz / lmbda
            ts = float(np.max(-ds))
            tz = float(np.max(-dz))
            t = max(0.0, ts, tz)
            if t == 0:
                step = 1.0
            elif i == 0:
                step = min(1.0, 1.0 / t)
            else:
                step = min(1.0, STEP / t)
            if i == 0:
                sigma = min(1.0, max(0.0, 1.0 - step + dsdz / gap * step ** 2)) ** EXPON

        x = x + step * dx
        # updated iterates in the current scaling, then misc.update_scaling ('l' block)
        ds = (1.0 + step * ds) * lmbda
        dz = (1.0 + step * dz) * lmbda
        ds = np.sqrt(ds)
        dz = np.sqrt(dz)
        d = d * ds / dz
        lmbda = ds * dz
        s = d * lmbda
        z = lmbda / d
        gap = float(lmbda @ lmbda)

    raise AssertionError("unreachable")


def qp(P, q, G=None, h=None, A=None, b=None, solver=None, kktsolver=None, initvals=None, **kwargs):
    if A is not None or b is not None or initvals is not None or solver is not None:
        raise NotImplementedError("cvxopt stand-in: only the inequality-constrained qp(P,q,G,h) is restated")
    return coneqp_l(P, q, G, h, kwargs.get("options", options))
