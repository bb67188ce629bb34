"""Pin the C restatement (oracle/marbler_oracle.c) to the reference's own code.

The fixtures hold outputs of /root/reference/robotarium_gym run unmodified on oracle/shims
(oracle/gen_golden.py).  The reference has no tests or golden vectors of its own,
and rps / cvxopt are restated, so this is "parity unpinned" at that boundary by construction.
"""
import numpy as np
import pytest

import golden_util as gu


@pytest.mark.parametrize("name", gu.fixture_names())
def test_step_matches_reference(oracle_lib, name):
    g = gu.Golden(name)
    o = oracle_lib.COracle(g.scenario, g.cfg)
    out, s1 = o.step(g.s0, g.actions)
    gu.compare_step(g, out, s1, pose_tol=1e-9, obs_tol=1e-9, rew_tol=1e-9, dist_tol=1e-9)
    assert np.array_equal(out["qp_iters"], g.qp_iters)
    assert out["obs"].shape[1:] == g.out["obs"].shape[1:]


def test_barrier_qp_matches_restated_cvxopt(oracle_lib):
    for N, v in gu.qp_vectors().items():
        for i in range(v["dxi"].shape[0]):
            u, it = oracle_lib.barrier_qp(v["dxi"][i], v["xi"][i], bool(v["default"][i]))
            assert it == v["iters"][i], (N, i, it, v["iters"][i])
            assert np.abs(u - v["u"][i]).max() < (1e-6 if it == 50 else 1e-9), (N, i)


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert oracle_lib.philox([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_lib.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_lib.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("scenario", ["PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple"])
def test_reset_distribution(oracle_lib, scenario):
    """Reset parity is distributional: spawn cells are the reference's grid,
    distinct, uniformly used; flags cleared."""
    g = gu.Golden(scenario + "_rollout")
    o = oracle_lib.COracle(scenario, g.cfg)
    st = o.reset(4096, seed=3)
    N = o.N
    assert (st["episode_steps"] == 0).all() and (st["prev_valid"] == 0).all()
    x, y = st["poses"][:, 0], st["poses"][:, 1]
    # every spawn pose the reference itself produced must be on our grid and vice versa
    starts = g.s0["episode_steps"] == 0
    ref_xy = {(round(a, 9), round(b, 9)) for a, b in zip(g.s0["poses"][starts][:, 0].ravel(), g.s0["poses"][starts][:, 1].ravel())}
    our_xy = {(round(a, 9), round(b, 9)) for a, b in zip(x.ravel(), y.ravel())}
    assert ref_xy <= our_xy
    d = np.hypot(x[:, :, None] - x[:, None, :], y[:, :, None] - y[:, None, :]) + 9 * np.eye(N)
    assert d.min() > 0.19
    if scenario == "ArcticTransport":
        grid = st["grid"]
        assert (grid[:, 7, 1:11] == 0).all() and ((grid == 3).sum(axis=(1, 2)) == 4).all()
        gc = st["goal_col"]
        assert gc.min() == 1 and gc.max() == 11
        assert all((grid[b, 0:2, gc[b] - 1:gc[b] + 1] == 3).all() for b in range(64))
        frac = [(grid[:, 2:7] == v).mean() for v in range(3)]
        assert max(abs(f - 1 / 3) for f in frac) < 0.01
    else:
        cells, counts = np.unique(np.round(x * 1000) * 10000 + np.round(y * 1000), return_counts=True)
        sp = o.c.spawn_robots
        assert len(cells) == sp.xr * sp.yr
        assert counts.min() > 0.8 * counts.mean() and counts.max() < 1.2 * counts.mean()
    if scenario == "PredatorCapturePrey":
        px = st["prey_loc"][:, :, 0]
        assert set(np.round(px.ravel(), 6)) == {0.5, 0.7, 0.9, 1.1}
        assert (st["prey_sensed"] == 0).all() and (st["prey_captured"] == 0).all()
    if scenario == "MaterialTransport":
        z = st["zone_load"]
        assert abs(z[:, 0].mean() - 99.5) < 0.6 and abs(z[:, 1].mean() - 19.5) < 0.3   # int() truncation: mean - 0.5
        assert abs(z[:, 0].std() - 10) < 0.5 and abs(z[:, 1].std() - 4) < 0.25
    if scenario == "Warehouse":
        th = st["poses"][:, 2]
        assert th.min() < -3.0 and th.max() > 3.0 and abs(th.mean()) < 0.1
