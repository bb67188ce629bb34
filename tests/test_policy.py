"""On-device policy: mrb_policy_act against the reference's own agent classes.

Fixtures (tests/golden/policy, made by oracle/gen_policy_golden.py): the reference's shipped checkpoints driven
through utilities/rnn_agent.py RNNAgent / utilities/rnn_ns_agent.py RNNNSAgent the way utilities/misc.py:155-170
run_env drives them.  The kernels multiply FP16 operands on the tensor cores with FP32 accumulation (FP16 has the
same 11-bit significand as TF32, which the first version of the kernel used and the fixture keys are named after),
so two bars:
  * against the evaluation of the same network with operands rounded to 11 significant bits (`q_tf32`): agreement to
    FP32 rounding noise - this pins the kernel's logic (fragment layouts, gate order, bias placement, in-place
    hidden update);
  * against the reference's float32 output (`q`): within 1e-3 of the largest |q|, and the SAME greedy action
    wherever the reference's top-2 gap exceeds twice that."""
import glob
import os

import numpy as np
import pytest
import torch

import golden_util as gu

POLICY = os.path.join(gu.GOLDEN, "policy")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(POLICY, "*.npz")))


def _load(name):
    z = np.load(os.path.join(POLICY, name + ".npz"))
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    return z, sd


def test_fixtures_present_and_flatten():
    """Host logic only: state_dict -> flat weight order of mrb_policy_create."""
    from marbler_b200.policy import flatten_state_dict
    assert len(NAMES) >= 3
    for name in NAMES:
        z, sd = _load(name)
        flat, f = flatten_state_dict(sd, int(z["n_agents"]))
        assert flat.dtype == np.float32 and flat.size == sum(v.size for v in sd.values())
        assert f["input_dim"] == int(z["obs_dim"]) + (int(z["n_agents"]) if int(z["obs_agent_id"]) else 0)
        assert f["n_actions"] == z["q"].shape[-1] and f["hidden_dim"] == z["h"].shape[-1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_policy_matches_reference_agents(name):
    from marbler_b200.policy import Policy
    z, sd = _load(name)
    N, D = int(z["n_agents"]), int(z["obs_dim"])
    pol = Policy(sd, N, D, obs_agent_id=bool(z["obs_agent_id"]), device="cuda:0")
    T, B = z["obs"].shape[:2]
    hidden = pol.init_hidden(B)
    q = torch.zeros((B, N, pol.n_actions), device="cuda:0")
    scale = float(np.abs(z["q"]).max())
    agree = []
    for t in range(T):
        obs = torch.tensor(z["obs"][t], device="cuda:0")
        a = pol.act(obs, hidden, q=q)
        torch.cuda.synchronize()
        qg, ag = q.cpu().numpy(), a.cpu().numpy()
        assert np.abs(qg - z["q_tf32"][t]).max() < 1e-4 * scale + 2e-4, (t, np.abs(qg - z["q_tf32"][t]).max())
        assert np.abs(qg - z["q"][t]).max() < 1e-3 * scale, (t, np.abs(qg - z["q"][t]).max())
        assert np.array_equal(ag, qg.argmax(axis=2)), t                      # the kernel's own argmax (first maximum)
        top = np.sort(z["q"][t], axis=2)
        clear = (top[..., -1] - top[..., -2]) > 2e-3 * scale
        assert np.array_equal(ag[clear], z["actions"][t][clear]), t
        agree.append((ag == z["actions"][t]).mean())
    assert np.mean(agree) >= 0.99
    hg = hidden.cpu().numpy()
    hs = max(1.0, float(np.abs(z["h"]).max()))                    # GRU states are in (-1, 1); the Linear+ReLU variant is unbounded
    assert np.abs(hg - z["h_tf32"]).max() < 1e-3 * hs and np.abs(hg - z["h"]).max() < 1e-2 * hs


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_accurate_policy_reproduces_the_reference_actions(name):
    """Policy(accurate=True): float32 operands and accumulation (csrc/policy_f32.cuh) against the float32 output of the
    reference's own agent classes on the shipped checkpoints (hidden 128 GRU shared, hidden 64 per-agent sets, 20
    actions): q to float32 rounding (only the summation order differs) and EVERY greedy action the reference's."""
    from marbler_b200.policy import Policy
    z, sd = _load(name)
    N, D = int(z["n_agents"]), int(z["obs_dim"])
    pol = Policy(sd, N, D, obs_agent_id=bool(z["obs_agent_id"]), device="cuda:0", accurate=True)
    T, B = z["obs"].shape[:2]
    hidden = pol.init_hidden(B)
    q = torch.zeros((B, N, pol.n_actions), device="cuda:0")
    scale = float(np.abs(z["q"]).max())
    worst = 0.0
    for t in range(T):
        a = pol.act(torch.tensor(z["obs"][t], device="cuda:0"), hidden, q=q)
        torch.cuda.synchronize()
        qg, ag = q.cpu().numpy(), a.cpu().numpy()
        worst = max(worst, float(np.abs(qg - z["q"][t]).max()))
        assert np.abs(qg - z["q"][t]).max() < 2e-5 * scale, (t, np.abs(qg - z["q"][t]).max(), scale)
        assert np.array_equal(ag, z["actions"][t]), (t, int((ag != z["actions"][t]).sum()))
    hs = max(1.0, float(np.abs(z["h"]).max()))
    assert np.abs(hidden.cpu().numpy() - z["h"]).max() < 2e-5 * hs
    print("accurate policy %s: max |q - q_ref| = %.3g (|q|max %.3g)" % (name, worst, scale))


@pytest.mark.gpu
def test_accurate_rollout_follows_the_host_loop_exactly(oracle_lib):
    """The closed loop with the float32 policy: every env follows run_env's loop restated on the host (float32 torch
    agent + C oracle env) action for action - no near-tie flips (the FP16 kernels track 94.5 % of the envs)."""
    from marbler_b200.policy import Policy, Rollout
    from marbler_b200.vec_env import VecEnv
    z, sd = _load("PredatorCapturePrey_vdn")
    cfg = dict(gu.Golden("PredatorCapturePrey_rollout").cfg)
    B, T, N, D = 1024, 12, 4, 16
    env = VecEnv("PredatorCapturePrey", cfg, num_envs=B, device="cuda:0", seed=21, auto_reset=True)
    ro = Rollout(env, Policy(sd, N, D, device="cuda:0", accurate=True), use_graph=False)
    ro.reset()
    orc = oracle_lib.COracle("PredatorCapturePrey", cfg)
    sf, si = orc.reset_flat(B, seed=21, threads=8)
    h = torch.zeros(B * N, 128)
    obs = np.zeros((B, N, D))
    fresh = np.ones(B, dtype=bool)
    follow = np.ones(B, dtype=bool)
    eye = torch.eye(N).repeat(B, 1)
    for t in range(T):
        ro.run(1)
        torch.cuda.synchronize()
        o = torch.tensor(obs, dtype=torch.float32)
        o[torch.tensor(fresh)] = 0
        h = h.reshape(B, N, 128)
        h[torch.tensor(fresh)] = 0
        q, h = _cpu_agent(sd, torch.cat([o.reshape(B * N, D), eye], dim=1), h.reshape(B * N, 128))
        a = q.argmax(dim=1).reshape(B, N).numpy().astype(np.int32)
        follow &= (ro.actions.cpu().numpy() == a).all(axis=1)
        obs, rew, dist, out_i = orc.step_flat(sf, si, a, auto_reset=True, seed=21, threads=8)
        fresh = out_i[:, 1].astype(bool)
    assert follow.mean() >= 0.999, follow.mean()


@pytest.mark.gpu
def test_fresh_mask_means_zero_hidden_and_zero_obs():
    from marbler_b200.policy import Policy
    z, sd = _load("PredatorCapturePrey_vdn")
    N, D = int(z["n_agents"]), int(z["obs_dim"])
    pol = Policy(sd, N, D, device="cuda:0")
    B = 200                                                               # not a multiple of the 64-env CTA tile
    g = torch.Generator(device="cuda:0").manual_seed(3)
    obs = torch.rand((B, N, D), generator=g, device="cuda:0") * 2 - 1
    hid = torch.rand((B, N, pol.hidden_dim), generator=g, device="cuda:0") - 0.5
    fresh = (torch.arange(B, device="cuda:0") % 3 == 0).to(torch.uint8)
    h1, q1 = hid.clone(), torch.zeros((B, N, 5), device="cuda:0")
    a1 = pol.act(obs, h1, q=q1, fresh=fresh)
    m = fresh.bool()
    obs2, h2, q2 = obs.clone(), hid.clone(), torch.zeros((B, N, 5), device="cuda:0")
    obs2[m] = 0
    h2[m] = 0
    a2 = pol.act(obs2, h2, q=q2)
    torch.cuda.synchronize()
    assert torch.equal(a1, a2) and torch.equal(q1, q2) and torch.equal(h1, h2)


def _cpu_agent(sd, obs, h):
    """utilities/rnn_agent.py:21-29 forward (shared weights, GRU), float32 on the host - the checker."""
    F = torch.nn.functional
    t = {k: torch.tensor(v) for k, v in sd.items()}
    x = F.relu(F.linear(obs, t["fc1.weight"], t["fc1.bias"]))
    gi, gh = F.linear(x, t["rnn.weight_ih"], t["rnn.bias_ih"]), F.linear(h, t["rnn.weight_hh"], t["rnn.bias_hh"])
    H = h.shape[-1]
    r, zg = torch.sigmoid(gi[:, :H] + gh[:, :H]), torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    hn = (1 - zg) * n + zg * h
    return F.linear(hn, t["fc2.weight"], t["fc2.bias"]), hn


@pytest.mark.gpu
def test_device_rollout_tracks_host_loop(oracle_lib):
    """Rollout (policy kernel + step kernel, CUDA-graph replay) against run_env's loop restated on the host:
    float32 torch agent + C oracle env, same Philox resets.  Reduced-precision operands may flip a near-tie argmax, after which that
    env's trajectory is simply a different valid one, so: graph == eager bit for bit and >= 90 % of envs follow
    the host loop action for action over the first 12 steps (48 agent decisions each)."""
    from marbler_b200.policy import Policy, Rollout
    from marbler_b200.vec_env import VecEnv
    z, sd = _load("PredatorCapturePrey_vdn")
    cfg = dict(gu.Golden("PredatorCapturePrey_rollout").cfg)
    B, T, N, D = 1024, 12, 4, 16

    def make():
        env = VecEnv("PredatorCapturePrey", cfg, num_envs=B, device="cuda:0", seed=21, auto_reset=True)
        return Rollout(env, Policy(sd, N, D, device="cuda:0"), use_graph=True, steps_per_graph=4)
    ro = make()
    ro.reset()
    orc = oracle_lib.COracle("PredatorCapturePrey", cfg)
    sf, si = orc.reset_flat(B, seed=21, threads=8)
    h = torch.zeros(B * N, 128)
    obs = np.zeros((B, N, D))
    fresh = np.ones(B, dtype=bool)
    follow = np.ones(B, dtype=bool)
    eye = torch.eye(N).repeat(B, 1)
    for t in range(T):
        ro.use_graph = False
        ro.run(1)
        torch.cuda.synchronize()
        o = torch.tensor(obs, dtype=torch.float32)
        o[torch.tensor(fresh)] = 0
        h = h.reshape(B, N, 128)
        h[torch.tensor(fresh)] = 0
        q, h = _cpu_agent(sd, torch.cat([o.reshape(B * N, D), eye], dim=1), h.reshape(B * N, 128))
        a = q.argmax(dim=1).reshape(B, N).numpy().astype(np.int32)
        follow &= (ro.actions.cpu().numpy() == a).all(axis=1)
        obs, rew, dist, out_i = orc.step_flat(sf, si, a, auto_reset=True, seed=21, threads=8)
        fresh = out_i[:, 1].astype(bool)
    assert follow.mean() >= 0.90, follow.mean()      # measured 0.945: ~1e-3 of agent-steps flip a near-tie under 11-bit operands
    # graph replay == eager launches
    r1, r2 = make(), make()
    r1.reset(), r2.reset()
    r2.use_graph = False
    r1.run(13), r2.run(13)
    torch.cuda.synchronize()
    assert r1._graph is not None
    assert torch.equal(r1.env.state_f64, r2.env.state_f64) and torch.equal(r1.hidden, r2.hidden)
    assert torch.equal(r1.actions, r2.actions) and torch.equal(r1.env.stats, r2.env.stats)


@pytest.mark.gpu
def test_mma_sync_policy_kernel_meets_the_same_bars():
    """Models with hidden 128 + GRUCell and <= 8 actions run on the persistent tcgen05 / TMEM kernel
    (csrc/policy_tc2.cuh) in the tests above; MRB_POLICY_TC=0 sends them to the mma.sync kernel that serves every
    other model shape, which must pass the same tests.  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    env = dict(os.environ, MRB_POLICY_TC="0")
    res = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k",
                          "test_policy_matches_reference_agents or test_fresh_mask or test_device_rollout or test_hidden128"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def _emulated_forward(sd, prefix, obs, h):
    """RNNAgent.forward (utilities/rnn_agent.py:21-29) in float32 torch with the matmul operands rounded to FP16 -
    the arithmetic contract of both policy kernels."""
    r16 = lambda t: t.half().float()
    g = lambda k: torch.as_tensor(sd[prefix + k], device=obs.device)
    H = g("fc1.weight").shape[0]
    x = torch.relu(r16(obs) @ r16(g("fc1.weight")).T + g("fc1.bias"))
    gi = r16(x) @ r16(g("rnn.weight_ih")).T + g("rnn.bias_ih")
    gh = r16(h) @ r16(g("rnn.weight_hh")).T + g("rnn.bias_hh")
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    hn = (1 - z) * n + z * h
    return r16(hn) @ r16(g("fc2.weight")).T + g("fc2.bias"), hn


@pytest.mark.gpu
@pytest.mark.parametrize("non_shared,B,D,agent_id", [(True, 1000, 11, False), (False, 777, 11, False), (False, 300, 30, True)])
def test_hidden128_gru_models_per_agent_and_shared(non_shared, B, D, agent_id):
    """The reference ships no RNNNSAgent checkpoint with hidden 128 + GRUCell, the shape the persistent tcgen05 kernel
    serves with per-agent weight sets (one agent index per CTA): random weights, ragged last tile, three steps of
    recurrence, against the emulated float32 forward above.  The 30-wide observation + one-hot id is ArcticTransport's
    input (fc1 k extent 48: the kernel then runs with a two-deep weight ring)."""
    from marbler_b200.policy import Policy
    rng = np.random.RandomState(5)
    N, A, H = (4 if agent_id else 3), 6, 128
    Din = D + (N if agent_id else 0)
    shapes = {"fc1.weight": (H, Din), "fc1.bias": (H,), "rnn.weight_ih": (3 * H, H), "rnn.weight_hh": (3 * H, H),
              "rnn.bias_ih": (3 * H,), "rnn.bias_hh": (3 * H,), "fc2.weight": (A, H), "fc2.bias": (A,)}
    prefixes = ["agents.%d." % i for i in range(N)] if non_shared else [""]
    sd = {p + k: (rng.standard_normal(s) / np.sqrt(s[-1] if len(s) > 1 else 16)).astype(np.float32) for p in prefixes for k, s in shapes.items()}
    pol = Policy(sd, N, D, obs_agent_id=agent_id, device="cuda:0")
    hidden = pol.init_hidden(B)
    h_ref = torch.zeros((B, N, H), device="cuda:0")
    q = torch.zeros((B, N, A), device="cuda:0")
    for t in range(3):
        obs = torch.as_tensor(rng.standard_normal((B, N, D)).astype(np.float32), device="cuda:0")
        a = pol.act(obs, hidden, q=q)
        torch.cuda.synchronize()
        for i in range(N):
            x = obs[:, i]
            if agent_id:                                   # utilities/misc.py:161-162
                onehot = torch.zeros((B, N), device="cuda:0"); onehot[:, i] = 1.0
                x = torch.cat([x, onehot], dim=1)
            qr, hr = _emulated_forward(sd, prefixes[i if non_shared else 0], x, h_ref[:, i])
            h_ref[:, i] = hr
            scale = float(qr.abs().max())
            assert float((q[:, i] - qr).abs().max()) < 2e-4 * scale + 2e-4, (t, i)
            top = torch.sort(qr, dim=1).values
            clear = (top[:, -1] - top[:, -2]) > 1e-3 * scale
            assert torch.equal(a[:, i][clear].long(), qr.argmax(dim=1)[clear]), (t, i)
        assert float((hidden - h_ref).abs().max()) < 1e-3
