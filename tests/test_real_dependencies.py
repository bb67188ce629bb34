"""The pin this repo cannot have in its build container: the REAL rps / cvxopt packages.

Every fixture under tests/golden was produced by the unmodified reference running on the restated stand-ins in
oracle/shims, so every parity statement reads "versus the restated rps / cvxopt".  These tests close that gap wherever
the real packages can be imported (a developer machine with `pip install cvxopt` and
robotarium_python_simulator @ 6bb184e on the path, reference README.md:10-11): they re-solve the committed QP vectors
with the real solver / the real barrier certificate and compare with what the fixtures hold.  They skip -- loudly --
when the packages are absent, which is the case in the build container and on the GPU box (no network, nothing in
/opt/wheelhouse)."""
import importlib
import os
import sys

import numpy as np
import pytest

import golden_util as gu

SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")


def _real(name):
    """Import `name` from outside oracle/shims, or None."""
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == name or k.startswith(name + ".")}
    try:
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != SHIMS]
        for k in saved_mods:
            del sys.modules[k]
        mod = importlib.import_module(name)
        if os.path.abspath(getattr(mod, "__file__", "") or "").startswith(SHIMS):
            return None
        return mod
    except Exception:
        return None
    finally:
        sys.path = saved_path
        if saved_mods:
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
            sys.modules.update(saved_mods)


def _qp_problem(dxi, xi, default):
    """The QP rps hands to cvxopt: pair order i < j, rows -2e / +2e, b = gain h^3, H = 2I, f = -2 dxi."""
    N = xi.shape[1]
    r2, lim = (0.17 ** 2, 0.2) if default else (0.2 ** 2, 0.2)
    d = dxi.copy()
    nrm = np.linalg.norm(d, 2, 0)
    big = nrm > lim
    d[:, big] *= lim / nrm[big]
    rows, b = [], []
    for i in range(N - 1):
        for j in range(i + 1, N):
            e = xi[:, i] - xi[:, j]
            h = e @ e - r2
            a = np.zeros(2 * N)
            a[2 * i:2 * i + 2] = -2 * e
            a[2 * j:2 * j + 2] = 2 * e
            rows.append(a)
            b.append((100.0 if (default or h >= 0) else 1e6) * h ** 3)
    return 2 * np.eye(2 * N), -2 * d.reshape(-1, order="F"), np.array(rows), np.array(b)


def test_fixture_qp_vectors_against_real_cvxopt():
    cvxopt = _real("cvxopt")
    if cvxopt is None:
        pytest.skip("real cvxopt is not installed here: parity stays 'versus the restated cvxopt' (DESIGN.md 2)")
    from cvxopt import matrix, solvers
    opts = {"show_progress": False, "reltol": 1e-2, "feastol": 1e-2, "maxiters": 50}
    worst, flips, total = 0.0, 0, 0
    for N, v in sorted(gu.qp_vectors().items()):
        for k in range(min(64, len(v["iters"]))):
            if v["iters"][k] >= 25:
                continue                                   # limit cycles of the loose stopping rule: not reproducible
            P, q, G, h = _qp_problem(v["dxi"][k], v["xi"][k], bool(v["default"][k]))
            sol = solvers.qp(matrix(P), matrix(q), matrix(G), matrix(h), options=opts)
            u = np.array(sol["x"]).reshape(2, N, order="F")
            total += 1
            flips += int(sol["iterations"] != v["iters"][k])
            worst = max(worst, np.abs(u - v["u"][k]).max())
    print("real cvxopt vs fixtures: %d problems, max |du| %.2e, iteration-count differences %d" % (total, worst, flips))
    assert worst < 1e-4                                    # north_star: QP velocities within 1e-4 of cvxopt's solution
    assert flips <= 0.01 * total


def test_fixture_qp_vectors_against_real_rps_certificate():
    if _real("cvxopt") is None or _real("rps") is None:
        pytest.skip("real rps (robotarium_python_simulator @ 6bb184e) / cvxopt are not installed here")
    saved = list(sys.path)
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != SHIMS]
    for k in [k for k in sys.modules if k == "rps" or k.startswith("rps.") or k == "cvxopt" or k.startswith("cvxopt.")]:
        del sys.modules[k]
    try:
        from rps.utilities.barrier_certificates import create_single_integrator_barrier_certificate, \
            create_single_integrator_barrier_certificate2
        safe, default = create_single_integrator_barrier_certificate2(safety_radius=0.2), create_single_integrator_barrier_certificate()
        worst = 0.0
        for N, v in sorted(gu.qp_vectors().items()):
            for k in range(min(32, len(v["iters"]))):
                if v["iters"][k] >= 25:
                    continue
                u = (default if v["default"][k] else safe)(v["dxi"][k].copy(), v["xi"][k].copy())
                worst = max(worst, np.abs(np.asarray(u) - v["u"][k]).max())
        print("real rps certificate vs fixtures: max |du| %.2e" % worst)
        assert worst < 1e-4
    finally:
        sys.path = saved
        for k in [k for k in sys.modules if k == "rps" or k.startswith("rps.") or k == "cvxopt" or k.startswith("cvxopt.")]:
            del sys.modules[k]
