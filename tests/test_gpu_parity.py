"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(marbler_b200._lib -> libmarbler_b200.so); the oracle (oracle/) is only the checker.

Bars (BASELINE.json north_star): discrete outputs (message, done, step counters, event flags, loads,
cells) bit-exact; poses within 1e-5; barrier-QP velocities within 1e-4 of the (restated) cvxopt iterate.
The measured agreement is ~1e-12, so the tests assert much tighter bounds than the bar where that is
robust, and the bar itself where float32 outputs are involved."""
import os

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-5          # north_star
QP_TOL = 1e-4            # north_star
F32_TOL = 2e-6           # obs / reward / dist are emitted as float32 (|values| <= 5)


def _vec(scenario, cfg, B, **kw):
    from marbler_b200.vec_env import VecEnv
    return VecEnv(scenario, cfg, num_envs=B, device="cuda:0", **kw)


def _step_from(env, s0, actions):
    env.set_state(s0)
    a = torch.as_tensor(np.asarray(actions), dtype=torch.int32, device=env.device)
    env.step(a)
    torch.cuda.synchronize()
    out = {"obs": env.obs.cpu().numpy(), "reward": env.reward.cpu().numpy(), "done": env.done.cpu().numpy(),
           "message": env.message.cpu().numpy(), "dist": env.dist.cpu().numpy(),
           "remaining": env.remaining.cpu().numpy()}
    return out, env.get_state()


def test_barrier_qp_matches_restated_cvxopt():
    from marbler_b200.vec_env import barrier_qp
    for N, v in sorted(gu.qp_vectors().items()):
        for kind in (0, 1):
            sel = v["default"] == kind
            if not sel.any():
                continue
            dxi = torch.tensor(v["dxi"][sel], device="cuda:0")
            xi = torch.tensor(v["xi"][sel], device="cuda:0")
            u, it = barrier_qp(dxi, xi, barrier_default=bool(kind))
            it = it.cpu().numpy()
            err = np.abs(u.cpu().numpy() - v["u"][sel]).reshape(len(it), -1).max(axis=1)
            conv = v["iters"][sel] < 50              # non-converged problems (cap 50) are compared loosely
            assert np.array_equal(it[conv], v["iters"][sel][conv]), (N, kind)
            assert err[conv].max() < QP_TOL, (N, kind, err[conv].max())
            assert err[conv].max() < 1e-8, (N, kind, err[conv].max())


def test_tensor_core_qp_equals_the_block_solver_on_hard_layouts(monkeypatch):
    """20 robots: the FP64 tensor-core solver (DMMA tile factorisation, inverted diagonal tiles, tile solves) against the
    2 x 2-block solver of the run-time team sizes on 16,384 random layouts that include what the fixtures have few of:
    pairs inside the safety radius (gain 1e6, KKT condition ~1e10), near-coincident robots and exactly symmetric rows."""
    from marbler_b200.vec_env import barrier_qp
    rng = np.random.RandomState(11)
    B, N = 16384, 20
    xi = np.stack([rng.uniform(-1.5, 1.5, (B, N)), rng.uniform(-0.9, 0.9, (B, N))], axis=1)
    close = rng.rand(B) < 0.5                              # pull a few robots onto their neighbours
    for b in np.nonzero(close)[0]:
        k = rng.randint(1, 6)
        src, dst = rng.randint(0, N, k), rng.randint(0, N, k)
        ok = src != dst
        xi[b][:, dst[ok]] = xi[b][:, src[ok]] + rng.uniform(-0.12, 0.12, (2, ok.sum())) * rng.choice([1.0, 0.05], ok.sum())
    grid = rng.rand(B) < 0.1                               # spawn-grid symmetry: one column, equal spacing
    xi[grid, 0, :] = -1.4
    xi[grid, 1, :] = np.linspace(-0.9, 0.9, N)
    dxi = rng.uniform(-0.3, 0.3, (B, 2, N))
    d, x = torch.tensor(dxi, device="cuda:0"), torch.tensor(xi, device="cuda:0")
    for kind in (False, True):
        u_t, it_t = barrier_qp(d, x, barrier_default=kind)
        monkeypatch.setenv("MRB_WARP_GENERIC", "1")
        u_g, it_g = barrier_qp(d, x, barrier_default=kind)
        monkeypatch.delenv("MRB_WARP_GENERIC")
        it_t, it_g = it_t.cpu().numpy(), it_g.cpu().numpy()
        err = np.abs((u_t - u_g).cpu().numpy()).reshape(B, -1).max(axis=1)
        assert np.isfinite(u_t.cpu().numpy()).all()
        same = it_t == it_g
        settled = same & (it_g < 25)                       # equal iteration counts outside cvxopt's limit cycles
        assert same.mean() > 0.995, (kind, same.mean())
        # two orderings of the same arithmetic on KKT systems of condition ~1e10: agreement degrades to ~1e-7 on the
        # worst layouts (north_star bar for QP velocities: 1e-4), and stays at rounding level on the rest
        q50, q99 = np.percentile(err[settled], [50, 99])
        assert err[settled].max() < 0.1 * QP_TOL, (kind, err[settled].max())
        assert q99 < 1e-8 and q50 < 1e-10, (kind, q50, q99)
        print("tensor-core vs block solver, default=%s: equal iteration counts %.4f, |du| max %.2e  99%% %.2e  median %.2e, mean iterations %.2f" % (
            kind, same.mean(), err[settled].max(), q99, q50, it_t.mean()))


def test_every_team_size_matches_the_block_solver(monkeypatch):
    """mrb_barrier_qp for every team size of 7..32 robots: the kernel with the compile-time size of the Newton system
    (FP64 tensor cores; N rounded up to a multiple of four, phantom robots when it is not one) against the run-time team
    size kernel (2 x 2 blocks on the FP64 pipe) on 512 random layouts each, a third of them with close pairs."""
    from marbler_b200.vec_env import barrier_qp
    rng = np.random.RandomState(5)
    B = 512
    for N in range(7, 33):
        xi = np.stack([rng.uniform(-1.5, 1.5, (B, N)), rng.uniform(-0.9, 0.9, (B, N))], axis=1)
        for b in range(0, B, 3):
            a, c = rng.choice(N, 2, replace=False)
            xi[b][:, c] = xi[b][:, a] + rng.uniform(-0.15, 0.15, 2)
        dxi = rng.uniform(-0.3, 0.3, (B, 2, N))
        d, x = torch.tensor(dxi, device="cuda:0"), torch.tensor(xi, device="cuda:0")
        u_t, it_t = barrier_qp(d, x)
        monkeypatch.setenv("MRB_WARP_GENERIC", "1")
        u_g, it_g = barrier_qp(d, x)
        monkeypatch.delenv("MRB_WARP_GENERIC")
        it_t, it_g = it_t.cpu().numpy(), it_g.cpu().numpy()
        same = it_t == it_g
        err = np.abs((u_t - u_g).cpu().numpy()).reshape(B, -1).max(axis=1)
        assert same.mean() >= 0.99, (N, same.mean())
        assert err[same & (it_g < 25)].max() < 0.1 * QP_TOL, (N, err[same & (it_g < 25)].max())
        assert np.median(err[same]) < 1e-10, (N, np.median(err[same]))


@pytest.mark.parametrize("name", gu.fixture_names())
def test_step_matches_reference_fixture(name):
    g = gu.Golden(name)
    env = _vec(g.scenario, g.cfg, g.B)
    out, s1 = _step_from(env, g.s0, g.actions)
    err = gu.compare_step(g, out, s1, pose_tol=POSE_TOL, obs_tol=F32_TOL, rew_tol=F32_TOL, dist_tol=F32_TOL)
    assert err["poses"] < 1e-9, err
    assert np.array_equal(s1["prev_pose"][:, :2].shape, g.s1["prev_pose"][:, :2].shape)
    assert np.abs(s1["prev_pose"][:, :2] - g.s1["prev_pose"][:, :2]).max() < 1e-9
    if "remaining" in out and g.scenario == "PredatorCapturePrey":
        assert np.array_equal(out["remaining"], g.cfg["num_prey"] - g.s1["prey_captured"].sum(axis=1))


SCN_B = [("PredatorCapturePrey", 8192), ("Warehouse", 4096), ("MaterialTransport", 4096),
         ("ArcticTransport", 4096), ("Simple", 4096)]


STALL_ITERS = 25


def _resync(env, orc, sf, si, idx):
    """Copy the oracle's state of envs `idx` into the CUDA env (after an unreproducible solve) through
    mrb_set_state; the env keeps its own episode return (a field the oracle does not carry)."""
    env.set_state(orc.unpack(sf[idx], si[idx]), envs=idx)


LOCKSTEP_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "lockstep_counts.jsonl")


def _log_counts(rec):
    """Exclusion counts of every lock-step run, one JSON line each (copied to profiles/ from the GPU run)."""
    import json
    print("lockstep", json.dumps(rec))
    try:
        os.makedirs(os.path.dirname(LOCKSTEP_LOG), exist_ok=True)
        with open(LOCKSTEP_LOG, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def _lockstep(oracle_lib, scenario, B, T, overrides=None, stall_frac=1e-4, label=None):
    """Same reset (same Philox draws), same random actions, T steps: the CUDA env and the C oracle must
    agree at every step (discrete bit-exact, poses 1e-5), including across auto-resets.

    Known, documented exception (DESIGN.md "limit cycles"): on exactly symmetric layouts (robots in one
    spawn column, all headings 0) cvxopt's Mehrotra iteration with rps' loose tolerances falls into a
    period-4 limit cycle and only leaves it through rounding noise, so the exit iteration (25..50) and
    the returned iterate are not reproducible between ANY two implementations (the C oracle and the
    Python restatement disagree with each other there too).  Such env-steps (oracle reports a solve of
    >= 25 iterations; ~2e-6 of solves) are excluded and the env is re-synchronised."""
    cfg = dict(gu.Golden(scenario + "_rollout").cfg)
    cfg.update(overrides or {})
    env = _vec(scenario, cfg, B, seed=5, auto_reset=True)
    orc = oracle_lib.COracle(scenario, cfg)
    env.reset()
    threads = os.cpu_count() or 8
    sf, si = orc.reset_flat(B, seed=5, threads=threads)
    rng = np.random.RandomState(0)
    st = env.get_state()
    ost = orc.unpack(sf, si)
    assert np.abs(st["poses"] - ost["poses"]).max() < 1e-12
    n_done = n_msg = n_stalled = n_loose = n_tie = 0
    knn = scenario in ("PredatorCapturePrey", "Warehouse") and cfg.get("num_neighbors", 0) < orc.N - 1
    for t in range(T):
        a = rng.randint(0, orc.n_actions, size=(B, orc.N)).astype(np.int32)
        env.step(torch.as_tensor(a, device=env.device))
        obs, rew, dist, out_i = orc.step_flat(sf, si, a, auto_reset=False, seed=5, threads=threads)
        final_poses = orc.unpack(sf, si)["poses"] if knn else None       # the poses the observations were built from
        orc.reset_envs(sf, si, np.where(out_i[:, 1] != 0)[0], seed=5)    # auto-reset, as the step kernel does
        ok = out_i[:, 5] < STALL_ITERS
        n_stalled += int((~ok).sum())
        assert np.array_equal(env.message.cpu().numpy()[ok], out_i[ok, 0]), t
        assert np.array_equal(env.done.cpu().numpy()[ok], out_i[ok, 1]), t
        st, ost = env.get_state(), orc.unpack(sf, si)
        dth = st["poses"][:, 2] - ost["poses"][:, 2]
        perr = np.maximum(np.abs(np.arctan2(np.sin(dth), np.cos(dth))).max(axis=1),
                          np.abs(st["poses"][:, :2] - ost["poses"][:, :2]).max(axis=(1, 2)))
        # ill-conditioned solves (a pair inside the safety radius: gain 1e6, KKT condition ~1e10) agree to
        # ~1e-6 only, which heading dynamics (1/0.05 projection) amplify to ~1e-5; they are rare, counted,
        # bounded by 1e-3 and re-synchronised.  Everything else must meet the 1e-5 bar (it is ~1e-12).
        loose = ok & (perr >= POSE_TOL)
        n_loose += int(loose.sum())
        assert perr[ok].max() < 1e-3, (t, perr[ok].max())
        tight = ok & ~loose
        obs_ok = tight
        if knn:
            # K-nearest neighbour blocks (utilities/misc.py:20-25): on mirror-symmetric layouts (common on the
            # spawn grid) two neighbours are EXACTLY equidistant and the order is decided by the last bit of
            # the integrated poses (the reference's own argpartition order is unspecified there): skip the obs
            # check for envs with such a tie, count them
            P = final_poses
            d = np.hypot(P[:, 0, :, None] - P[:, 0, None, :], P[:, 1, :, None] - P[:, 1, None, :])
            ds = np.sort(d, axis=2)
            tie = (np.diff(ds[:, :, 1:], axis=2) < 1e-9).any(axis=(1, 2))
            n_tie += int(tie.sum())
            obs_ok = tight & ~tie
        assert np.abs(env.obs.cpu().numpy() - obs)[obs_ok].max() < POSE_TOL + F32_TOL, t
        assert np.abs(env.reward.cpu().numpy() - rew)[tight].max() < POSE_TOL + F32_TOL, t
        assert np.abs(env.dist.cpu().numpy() - dist)[tight].max() < POSE_TOL + F32_TOL, t
        for k in gu.DISCRETE_STATE + ("episode_count",):
            if k in ost:
                a_, b_ = np.asarray(st[k]).astype(np.int64), np.asarray(ost[k]).astype(np.int64)
                assert np.array_equal(a_[ok], b_[ok]), (t, k)
        if not tight.all():
            _resync(env, orc, sf, si, np.where(~tight)[0])
        n_done += int(out_i[ok, 1].sum())
        n_msg += int((out_i[ok, 0] != 0).sum())
    stats = env.read_stats()
    _log_counts({"case": label or scenario, "scenario": scenario, "overrides": overrides or {}, "envs": B, "steps": T,
                 "env_steps": B * T, "excluded_stalled_solves": n_stalled, "loose_pose_env_steps": n_loose,
                 "knn_tie_env_steps": n_tie, "episodes_finished": n_done, "violations": n_msg,
                 "qp_iterations": stats["qp_iterations"], "qp_iterations_warp": stats["qp_iterations_warp"]})
    assert stats["substeps"] <= B * T * cfg["update_frequency"] and stats["qp_iterations_warp"] >= stats["qp_iterations"]
    assert n_stalled <= max(2, int(stall_frac * B * T)), n_stalled
    # measured on the B200 (profiles/r02_lockstep_counts.jsonl): loose <= 1.2e-5, exact ties <= 4.3e-4 of env-steps
    assert n_loose <= max(2, int(5e-5 * B * T)), n_loose
    assert n_tie <= max(4, int(2e-3 * B * T)), n_tie
    assert abs(stats["episodes"] - n_done) <= n_stalled and stats["env_steps"] == B * T
    assert stats["collisions"] + stats["boundary_exits"] >= n_msg - n_stalled
    assert stats["qp_stalls"] <= 4 * max(n_stalled, 1)
    return n_stalled, n_loose


@pytest.mark.parametrize("scenario,B", SCN_B)
def test_rollout_lockstep_with_oracle(oracle_lib, scenario, B):
    # MaterialTransport spawns every robot in ONE column (1 x 6 grid, heading 0), i.e. exactly the symmetric
    # layout of the limit cycle, so it sees far more of them than the other scenarios
    # measured shares of excluded env-steps: MaterialTransport 9.0e-4, ArcticTransport 8.5e-5, the others <= 1.2e-5
    _lockstep(oracle_lib, scenario, B, 40, stall_frac={"MaterialTransport": 3e-3, "ArcticTransport": 3e-4}.get(scenario, 1e-4))


TEAM_CASES = [
    ("PredatorCapturePrey", dict(predator=1, capture=1, num_neighbors=1)),                       # N=2 thread kernel
    ("PredatorCapturePrey", dict(predator=2, capture=1, num_neighbors=1, num_prey=9)),           # N=3, K nearest
    ("PredatorCapturePrey", dict(predator=3, capture=2, num_neighbors=4, capability_aware=True)),  # N=5 (primal QP)
    ("PredatorCapturePrey", dict(predator=4, capture=4, num_neighbors=2, ROBOT_INIT_RIGHT_THRESH=0.1)),  # N=8 warp kernel
    ("Simple", dict(n_agents=3)),
    ("Simple", dict(n_agents=10, ROBOT_INIT_RIGHT_THRESH=0.1)),                                  # warp kernel
    ("Warehouse", dict(n_agents=4, num_neighbors=3)),
    ("Warehouse", dict(n_agents=9, num_neighbors=3)),                                            # warp kernel, K nearest
    ("MaterialTransport", dict(n_agents=5, n_fast_agents=2, n_slow_agents=3, capability_aware=True)),
    ("PredatorCapturePrey", dict(robotarium=True, update_frequency=10)),                         # controller every sub-step
    ("PredatorCapturePrey", dict(barrier_certificate="default")),
    ("Warehouse", dict(penalize_violations=False)),
]


@pytest.mark.parametrize("scenario,overrides", TEAM_CASES, ids=lambda v: v if isinstance(v, str) else "-".join("%s=%s" % kv for kv in sorted(v.items()))[:60])
def test_other_team_sizes_and_options_lockstep(oracle_lib, scenario, overrides):
    """Team sizes / options the reference fixtures do not cover, against the C oracle (which is pinned to the
    reference for N = 4, 6, 20): every kernel dispatch path (thread N = 2..6, warp N = 7..32)."""
    _lockstep(oracle_lib, scenario, 1024, 25, overrides=overrides, stall_frac=1e-3)      # measured: none in any of these


@pytest.mark.parametrize("predator,capture", [(12, 11), (13, 13)], ids=["N=23", "N=26"])
def test_large_teams_lockstep(oracle_lib, predator, capture):
    """The widest instantiations of the run-time team size kernel (8 and 16 pair slots per lane; an odd team size takes
    the single-column tail of the factorisation): 23 and 26 robots on the 5 x 6 spawn grid, K = 3 nearest neighbours."""
    _lockstep(oracle_lib, "PredatorCapturePrey", 256, 12, stall_frac=1e-2,
              overrides=dict(predator=predator, capture=capture, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))


TENSOR_TEAMS = [("PredatorCapturePrey", dict(predator=6, capture=6, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)),
                ("PredatorCapturePrey", dict(predator=8, capture=8, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=15)),
                ("PredatorCapturePrey", dict(predator=12, capture=12, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)),
                ("PredatorCapturePrey", dict(predator=14, capture=14, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)),
                ("PredatorCapturePrey", dict(predator=15, capture=14, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)),   # 29 on 32
                ("PredatorCapturePrey", dict(predator=9, capture=9, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=17)),    # 18 on 20
                ("Warehouse", dict(n_agents=11, num_neighbors=10)),                                                    # 11 on 12
                ("Warehouse", dict(n_agents=8, num_neighbors=7)),
                ("Simple", dict(n_agents=8, ROBOT_INIT_RIGHT_THRESH=0.1)),
                ("Simple", dict(n_agents=12, ROBOT_INIT_RIGHT_THRESH=0.1)),
                ("Simple", dict(n_agents=16, ROBOT_INIT_RIGHT_THRESH=0.1))]


@pytest.mark.parametrize("scenario,overrides", TENSOR_TEAMS, ids=["PCP-12", "PCP-16", "PCP-24", "PCP-28", "PCP-29on32", "PCP-18on20", "Warehouse-11on12", "Warehouse-8", "Simple-8", "Simple-12", "Simple-16"])
def test_tensor_core_team_sizes_lockstep(oracle_lib, scenario, overrides):
    """Every (scenario, size) that has a kernel of its own with the FP64 tensor-core solver (csrc/kern_team_*.cu), with
    the team size folded into the code and padded with phantom robots (PCP-8, PCP-10 on 12 ... are in TEAM_CASES and
    test_large_teams_lockstep, PCP-20 in test_lockstep_20_robots), against the C oracle."""
    _lockstep(oracle_lib, scenario, 512, 14, overrides=overrides, stall_frac=1e-2)


PCP20 = dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)


def test_lockstep_baseline_config2_full_size(oracle_lib):
    """BASELINE.json configs[1] literally: PredatorCapturePrey, 65,536 envs, barrier certificates on, random actions,
    20 steps against the oracle with the bit-exact event / done / message / flag / counter check at every step."""
    _lockstep(oracle_lib, "PredatorCapturePrey", 65536, 20, label="config2 PCP 65536")


@pytest.mark.parametrize("generic", [False, True], ids=["specialised", "generic"])
def test_lockstep_20_robots(oracle_lib, generic, monkeypatch):
    """BASELINE.json configs[4]'s team (10 predators + 10 capture agents, K = 3 nearest neighbours): 4,096 envs x 25
    steps on the compile-time-specialised warp kernel and on the generic one (MRB_WARP_GENERIC is read per launch)."""
    if generic:
        monkeypatch.setenv("MRB_WARP_GENERIC", "1")
    else:
        monkeypatch.delenv("MRB_WARP_GENERIC", raising=False)
    _lockstep(oracle_lib, "PredatorCapturePrey", 4096, 25, overrides=PCP20, stall_frac=2e-3,
              label="config5 PCP20 " + ("generic" if generic else "specialised"))


@pytest.mark.parametrize("scenario", ["Warehouse", "MaterialTransport", "ArcticTransport"])
def test_lockstep_full_size_other_configs(oracle_lib, scenario):
    """BASELINE.json configs[2], [3] at 65,536 envs, a few steps (the CPU oracle needs ~0.3 s per step here)."""
    # the first steps after a reset are the symmetric layouts of the limit cycle (MaterialTransport: one spawn column;
    # ArcticTransport: four robots in a row, same heading) - measured 4.9e-3 / 3.1e-4 / 2.5e-6 of env-steps here
    budget = {"MaterialTransport": 1e-2, "ArcticTransport": 1e-3}.get(scenario, 1e-4)
    _lockstep(oracle_lib, scenario, 65536, 6, stall_frac=budget, label="full size " + scenario)


def test_projected_collision_form_lockstep(oracle_lib):
    """The other published form of rps' collision test (heading-projected points, offset 0.025) in closed loop."""
    _lockstep(oracle_lib, "PredatorCapturePrey", 4096, 40, overrides=dict(rps_collision_offset=0.025), label="PCP projected collision")
    _lockstep(oracle_lib, "Warehouse", 2048, 40, overrides=dict(rps_collision_offset=0.025), label="Warehouse projected collision")


def test_state_access_through_the_c_abi():
    """mrb_get_state / mrb_set_state against the raw SoA rows (marbler_b200/layout.py is the independent description
    of the row layout): every field of every scenario, env ranges, and fields left NULL keep their values."""
    from marbler_b200 import layout
    for scenario, B in SCN_B:
        g = gu.Golden(scenario + "_rollout")
        env = _vec(scenario, g.cfg, 300, seed=3, auto_reset=True)
        env.reset()
        gen = torch.Generator(device="cuda:0").manual_seed(1)
        for _ in range(5):
            env.step(torch.randint(0, env.n_actions, (300, env.N), generator=gen, device="cuda:0", dtype=torch.int32))
        torch.cuda.synchronize()
        raw = layout.unpack(scenario, env.N, env.P, env.state_f64.cpu().numpy(), env.state_i32.cpu().numpy())
        st = env.get_state()
        assert set(st) == set(raw), (set(st) ^ set(raw))
        for k in raw:
            assert np.array_equal(np.asarray(st[k]).astype(np.float64), np.asarray(raw[k]).astype(np.float64)), (scenario, k)
        part = env.get_state(env_lo=37, count=100)
        for k in raw:
            assert np.array_equal(part[k], st[k][37:137]), (scenario, k)
        # write a permuted copy back through the C ABI, compare the raw rows with the independent packer
        perm = np.random.RandomState(0).permutation(300)
        env.set_state({k: v[perm] for k, v in st.items()})
        pf, pi = layout.pack(scenario, env.N, env.P, {k: v[perm] for k, v in st.items()}, 300)
        assert np.array_equal(env.state_f64.cpu().numpy(), pf) and np.array_equal(env.state_i32.cpu().numpy(), pi)
        # partial update: only poses of envs [10, 20); everything else untouched
        before = env.get_state()
        new = before["poses"][10:20] + 0.25
        env.set_state({"poses": new}, env_lo=10, count=10)
        after = env.get_state()
        assert np.array_equal(after["poses"][10:20], new)
        after["poses"][10:20] = before["poses"][10:20]
        for k in before:
            assert np.array_equal(after[k], before[k]), (scenario, k)
        env.set_state({k: v[[5, 200]] for k, v in st.items()}, envs=[7, 250])
        again = env.get_state()
        for k in st:
            assert np.array_equal(again[k][7], st[k][5]) and np.array_equal(again[k][250], st[k][200]), (scenario, k)


def test_single_env_returns_float64_like_the_reference():
    """num_envs == 1: observations are float64 rows and rewards Python floats at full precision (the reference:
    PredatorCapturePrey.py:176), not values rounded through the float32 buffers of the batched path.  States and
    expected outputs are cases of the reference's own rollout fixtures (the ones without a limit-cycle solve)."""
    import marbler_b200
    for scenario in ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple"):
        g = gu.Golden(scenario + "_rollout")
        env = marbler_b200.make("robotarium_gym:%s-v0" % scenario, seed=5)
        env.reset()
        scn = env.env
        checked = 0
        solves = -(-g.cfg["update_frequency"] // 15)
        for i in range(20, g.B):
            # the fixture records the TOTAL iterations of the step's solves; a solve takes at least 4, so below this
            # bound none of them ran into the limit cycle (>= 25)
            if g.qp_iters[i] >= STALL_ITERS + 4 * (solves - 1) or checked == 8:
                continue
            scn.set_state({k: v[i:i + 1] for k, v in g.s0.items()})
            obs, rew, done, info = env.step([int(a) for a in g.actions[i]])
            assert all(o.dtype == np.float64 for o in obs) and all(isinstance(r, float) for r in rew)
            assert np.abs(np.asarray(obs) - g.out["obs"][i]).max() < 1e-9, (scenario, i)
            assert np.abs(np.asarray(rew) - g.out["reward"][i]).max() < 1e-9, (scenario, i)
            assert np.abs(np.asarray(scn.get_observations()) - g.out["obs"][i]).max() < 1e-9
            assert done == [bool(g.out["done"][i][0])] * scn.num_robots
            if scenario == "PredatorCapturePrey" and int(g.out["message"][i]) == 0 and abs(g.out["reward"][i][0] + 0.05) < 1e-12:
                assert rew[0] == -0.05                           # not float32(-0.05) = -0.0500000007
            checked += 1
        assert checked == 8, (scenario, checked)


def test_fp64_peak_and_solver_statistics():
    from marbler_b200.vec_env import fp64_peak
    tf = fp64_peak(0, 100.0)
    assert 15.0 < tf < 60.0, tf                                  # B200: ~37-40 TFLOP/s nominal
    g = gu.Golden("PredatorCapturePrey_rollout")
    env = _vec("PredatorCapturePrey", g.cfg, 8192, seed=1, auto_reset=True)
    env.reset()
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    for _ in range(10):
        env.step(torch.randint(0, 5, (8192, 4), generator=gen, device="cuda:0", dtype=torch.int32))
    s = env.read_stats()
    assert s["env_steps"] == 81920 and 0 < s["substeps"] <= 81920 * g.cfg["update_frequency"]
    assert 1.0 <= s["qp_iterations_warp"] / s["qp_iterations"] < 2.0     # divergence overhead of the solver loop


@pytest.mark.parametrize("scenario", [s for s, _ in SCN_B])
def test_reset_matches_oracle_draws(oracle_lib, scenario):
    g = gu.Golden(scenario + "_rollout")
    B = 4096
    env = _vec(scenario, g.cfg, B, seed=11, env_id0=1000)
    env.reset()
    assert float(env.obs.abs().max()) == 0.0
    st = env.get_state()
    ost = oracle_lib.COracle(scenario, g.cfg).reset(B, seed=11, env_id0=1000)
    assert np.abs(st["poses"] - ost["poses"]).max() < 1e-12
    for k in gu.DISCRETE_STATE + ("episode_count",):
        if k in ost:
            same = np.asarray(st[k]).astype(np.int64) == np.asarray(ost[k]).astype(np.int64)
            assert same.mean() > 0.9999 if k == "zone_load" else same.all(), k
    for k in ("prey_loc", "goal"):
        if k in ost:
            assert np.array_equal(st[k], ost[k])


def test_masked_reset_touches_only_masked_envs():
    g = gu.Golden("PredatorCapturePrey_rollout")
    env = _vec("PredatorCapturePrey", g.cfg, 512, seed=2)
    env.reset()
    before = env.get_state()
    mask = np.zeros(512, dtype=np.uint8)
    mask[::3] = 1
    env.reset(mask=mask)
    after = env.get_state()
    keep = mask == 0
    assert np.array_equal(before["poses"][keep], after["poses"][keep])
    assert (after["episode_count"][keep] == 1).all() and (after["episode_count"][~keep] == 2).all()
    assert not np.array_equal(before["poses"][~keep], after["poses"][~keep])


def test_sharded_envs_reproduce_single_device_run():
    """Two handles owning env ids [0,B/2) and [B/2,B) give the same trajectories as one handle over
    [0,B): sharding needs no data-path collective."""
    g = gu.Golden("PredatorCapturePrey_rollout")
    B = 2048
    full = _vec("PredatorCapturePrey", g.cfg, B, seed=9, auto_reset=True)
    halves = [_vec("PredatorCapturePrey", g.cfg, B // 2, seed=9, env_id0=k * B // 2, auto_reset=True) for k in (0, 1)]
    for e in [full] + halves:
        e.reset()
    rng = np.random.RandomState(1)
    for _ in range(90):
        a = torch.as_tensor(rng.randint(0, 5, size=(B, 4)).astype(np.int32), device="cuda:0")
        full.step(a)
        halves[0].step(a[:B // 2].contiguous())
        halves[1].step(a[B // 2:].contiguous())
    torch.cuda.synchronize()
    cat = torch.cat([halves[0].state_f64, halves[1].state_f64], dim=1)
    assert torch.equal(full.state_f64, cat)
    assert torch.equal(full.state_i32, torch.cat([halves[0].state_i32, halves[1].state_i32], dim=1))
    assert torch.equal(full.obs, torch.cat([halves[0].obs, halves[1].obs], dim=0))
    s = [e.read_stats() for e in [full] + halves]
    assert s[0]["episodes"] == s[1]["episodes"] + s[2]["episodes"] > 0


def test_host_path_equals_device_path():
    g = gu.Golden("Warehouse_inject")
    a_env, b_env = _vec(g.scenario, g.cfg, g.B), _vec(g.scenario, g.cfg, g.B)
    a_env.set_state(g.s0)
    b_env.set_state(g.s0)
    a_env.step(torch.as_tensor(g.actions, device="cuda:0"))
    obs, rew, done, msg = b_env.step_host(g.actions)
    torch.cuda.synchronize()
    assert torch.equal(a_env.obs.cpu(), obs) and torch.equal(a_env.reward.cpu(), rew)
    assert torch.equal(a_env.done.cpu(), done) and torch.equal(a_env.message.cpu(), msg)


@pytest.mark.parametrize("mode", ["0", "1"])
def test_host_path_output_modes(mode):
    """mrb_step_host writes the host buffers three ways (include/marbler_b200.h): small outputs stored by the
    kernel (default, covered by the tests around this one), everything downloaded (MRB_HOST_DIRECT=0) and
    everything stored by the kernel (=1).  The switch is read once per process, hence the subprocesses."""
    import subprocess
    import sys
    env = dict(os.environ, MRB_HOST_DIRECT=mode)
    res = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k",
                          "test_host_path_equals_device_path or test_chunked_host_path_equals_device_path"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "3 passed" in res.stdout


@pytest.mark.parametrize("B", [40000, 70001])
def test_chunked_host_path_equals_device_path(B):
    """Batches >= 32,768 envs take the multi-stream chunked path of mrb_step_host (ragged last chunk included):
    same results as one device-path launch, step after step, with auto-reset on."""
    g = gu.Golden("PredatorCapturePrey_rollout")
    a_env = _vec(g.scenario, g.cfg, B, seed=9, auto_reset=True)
    b_env = _vec(g.scenario, g.cfg, B, seed=9, auto_reset=True)
    a_env.reset(), b_env.reset()
    rng = np.random.RandomState(1)
    for _ in range(3):
        a = rng.randint(0, 5, size=(B, 4)).astype(np.int32)
        a_env.step(torch.as_tensor(a, device="cuda:0"))
        obs, rew, done, msg = b_env.step_host(a)
        torch.cuda.synchronize()
        assert torch.equal(a_env.obs.cpu(), obs) and torch.equal(a_env.reward.cpu(), rew)
        assert torch.equal(a_env.done.cpu(), done) and torch.equal(a_env.message.cpu(), msg)
    assert torch.equal(a_env.state_f64, b_env.state_f64) and torch.equal(a_env.state_i32, b_env.state_i32)
    sa, sb = a_env.read_stats(), b_env.read_stats()
    assert sa["env_steps"] == sb["env_steps"] == 3 * B and sa["episodes"] == sb["episodes"]


def test_wrapper_single_env_keeps_reference_types():
    import marbler_b200
    for key, n, d, na in (("PredatorCapturePrey-v0", 4, 16, 5), ("Warehouse-v0", 6, 18, 5), ("MaterialTransport-v0", 4, 9, 20),
                          ("ArcticTransport-v0", 4, 30, 5), ("Simple-v0", 4, 10, 5)):
        env = marbler_b200.make("robotarium_gym:" + key, seed=3)
        assert env.n_agents == n and len(env.action_space) == n and env.action_space[0].n == na
        assert env.observation_space[0].shape == (d,)
        obs = env.reset()
        assert len(obs) == n and all(len(o) == d and not any(o) for o in obs)
        obs, rew, done, info = env.step([1] * n)
        assert isinstance(obs, tuple) and len(obs) == n and obs[0].shape == (d,)
        assert isinstance(rew, list) and isinstance(rew[0], float) and len(rew) == n
        assert isinstance(done, list) and isinstance(done[0], bool) and len(done) == n
        assert isinstance(info, dict) and info["dist_travelled"].shape == (n,)
        goals = env.env._generate_step_goal_positions([0] * n)
        assert goals.shape == (3, n)


def test_full_size_properties():
    """BASELINE config 2 size (65,536 PCP envs): determinism, physical bounds, flag consistency."""
    g = gu.Golden("PredatorCapturePrey_rollout")
    B = 65536
    runs = []
    for _ in range(2):
        env = _vec("PredatorCapturePrey", g.cfg, B, seed=123, auto_reset=False)
        env.reset()
        gen = torch.Generator(device="cuda:0").manual_seed(7)
        p0 = env.agent_poses.clone()
        for _t in range(3):
            a = torch.randint(0, 5, (B, 4), generator=gen, device="cuda:0", dtype=torch.int32)
            env.step(a)
        torch.cuda.synchronize()
        runs.append((env.state_f64.clone(), env.state_i32.clone(), env.obs.clone(), env.message.clone(), env.done.clone()))
    for x, y in zip(*runs):
        assert torch.equal(x, y)
    p1 = env.agent_poses
    moved = (p1[:, :2] - p0[:, :2]).norm(dim=1)
    assert float(moved.max()) <= 3 * 29 * 0.033 * 0.2 + 1e-9            # |v| <= 0.2 m/s
    assert float(p1[:, 2].abs().max()) <= np.pi + 1e-12
    msg, done = runs[0][3], runs[0][4]
    assert bool(((msg != 0) <= (done != 0)).all())
    steps = env.state_i32[0]
    assert int(steps.max()) == 3 and int(steps.min()) >= 1
    assert bool((env.obs[:, :, 2:4] >= -5).all())


FULL_SIZE = [("Warehouse", 262144, {}), ("MaterialTransport", 262144, {}), ("ArcticTransport", 262144, {}),
             ("PredatorCapturePrey", 131072, dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))]


@pytest.mark.parametrize("scenario,B,over", FULL_SIZE, ids=[s + "-" + str(b) for s, b, _ in FULL_SIZE])
def test_full_size_properties_other_configs(scenario, B, over):
    """BASELINE configs 3-5 at their full per-GPU sizes: size-independent properties of the env step -
    run-to-run determinism (bit-exact state, obs, flags), |v| <= 0.2 m/s displacement bound, wrapped headings,
    violation => done, step counters advance by one, observations finite and inside the declared Box."""
    cfg = dict(gu.Golden(scenario + "_rollout").cfg)
    cfg.update(over)
    runs = []
    for _ in range(2):
        env = _vec(scenario, cfg, B, seed=77, auto_reset=False)
        env.reset()
        gen = torch.Generator(device="cuda:0").manual_seed(9)
        p0 = env.agent_poses.clone()
        for _t in range(2):
            a = torch.randint(0, env.n_actions, (B, env.N), generator=gen, device="cuda:0", dtype=torch.int32)
            env.step(a)
        torch.cuda.synchronize()
        runs.append((env.state_f64.clone(), env.state_i32.clone(), env.obs.clone(), env.message.clone(),
                     env.done.clone(), env.reward.clone()))
    for x, y in zip(*runs):
        assert torch.equal(x, y)
    p1 = env.agent_poses
    uf = cfg["update_frequency"]
    assert float((p1[:, :2] - p0[:, :2]).norm(dim=1).max()) <= 2 * uf * 0.033 * 0.2 + 1e-9
    assert float(p1[:, 2].abs().max()) <= np.pi + 1e-12
    msg, done = runs[0][3], runs[0][4]
    assert bool(((msg != 0) <= (done != 0)).all())
    assert bool((env.state_i32[0] == 2).all())
    assert bool(torch.isfinite(env.obs).all()) and bool(torch.isfinite(env.reward).all())
    assert float(env.obs.min()) >= -5.0 and float(env.obs.max()) <= 150.0
    stats = env.read_stats()
    assert stats["env_steps"] == 2 * B and stats["qp_solves"] > 0
