"""BipedalWalker-v3 on the GPU against the CPU oracle (oracle/walker_oracle.c).  `pytest -m gpu`.

PARITY UNPINNED w.r.t. the real reference (no Box2D here).  Checked: the generic C implementation
and the specialised CUDA one agree bit for bit (observations incl. lidar, rewards, flags, terrain,
RNG stream incl. numpy's buffered 32-bit integers) over long roll-outs; invariants; API shape.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bit_exact_against_oracle():
    import gym_b200
    import torch
    from oracle.oracle import OracleWalker
    N, T, seed = 512, 300, 21
    env = gym_b200.vector.make("BipedalWalker-v3", N)
    orc = OracleWalker(N)
    obs, _ = env.reset(seed=seed)
    ref = orc.reset(seed=seed)
    assert tuple(obs.shape) == (N, 24) and obs.dtype == torch.float32
    assert np.array_equal(obs.cpu().numpy(), ref), "reset observations (terrain, lidar, embedded step)"
    rng = np.random.default_rng(5)
    n_term = 0
    for t in range(T):
        # a mix of random, saturated and zero actions (motor torque clipping, sign(0) = 0)
        a = rng.uniform(-1.5, 1.5, size=(N, 4)).astype(np.float32)
        a[rng.random((N, 4)) < 0.1] = 0.0
        o, r, te, tr, info = env.step(torch.as_tensor(a, device=env.device))
        ro, rr, rte, rtr, rfo = orc.step(a)
        assert np.array_equal(te.cpu().numpy(), rte), f"step {t}: terminated"
        assert np.array_equal(tr.cpu().numpy(), rtr), f"step {t}: truncated"
        o_h = o.cpu().numpy()
        if not np.array_equal(o_h, ro):
            bad = np.argwhere(o_h != ro)
            raise AssertionError(f"step {t}: {len(bad)} observation values differ, first {bad[0]} "
                                 f"got {o_h[tuple(bad[0])]!r} want {ro[tuple(bad[0])]!r}")
        assert np.array_equal(r.cpu().numpy(), rr), f"step {t}: reward"
        done = rte | rtr
        if done.any():
            assert np.array_equal(info["final_observation"].cpu().numpy()[done], rfo[done])
        n_term += int(rte.sum())
    assert n_term > N  # random walkers fall within ~100 steps, several resets (fresh terrain) per env
    bodies, flags = env.walker_bodies()
    for i in (0, 7, N - 1):
        ob, of = orc.bodies(i)
        assert np.array_equal(bodies[i].cpu().numpy(), ob)
        assert flags[i].tolist() == of.tolist()
    env.close()
    orc.close()


def test_invariants_and_api():
    import gym_b200
    import torch
    from gym_b200 import spaces
    N = 2048
    env = gym_b200.vector.make("BipedalWalker-v3", N)
    assert isinstance(env.single_action_space, spaces.Box) and env.single_action_space.shape == (4,)
    assert env.action_space.shape == (N, 4) and env.single_observation_space.shape == (24,)
    assert env.max_episode_steps == 1600
    obs, _ = env.reset(seed=0)
    assert bool((obs[:, 14:] > 0.3).all()) and bool((obs[:, 14:] <= 1.0).all())   # lidar sees the ground
    assert bool((obs[:, [8, 13]] == 0).all())                                      # legs not yet on the ground
    gen = torch.Generator(device="cuda").manual_seed(0)
    total_term = 0
    for t in range(200):
        a = torch.rand((N, 4), device="cuda", generator=gen) * 2 - 1
        o, r, te, tr, info = env.step(a)
        assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(r).all())
        assert bool(((o[:, [8, 13]] == 0) | (o[:, [8, 13]] == 1)).all())
        assert bool((o[:, 14:] >= 0).all()) and bool((o[:, 14:] <= 1).all())
        # an episode under random torques ends with the hull on the ground (-100).  Without continuous
        # collision (TOI) the solver very rarely (~1 per 10^6 env-steps) blows up in an over-constrained
        # pose and launches the walker off the far end instead, which terminates without the penalty.
        assert int((r[te] != -100).sum()) <= max(1, int(te.sum()) // 500)
        assert bool((r[~te] > -30).all())       # shaping deltas: a few points per step at most
        total_term += int(te.sum())
    assert total_term > N
    with pytest.raises(NotImplementedError):
        gym_b200.vector.make("BipedalWalker-v3", 4, hardcore=True)
    env.close()
    single = gym_b200.make("BipedalWalker-v3")
    s, _ = single.reset(seed=3)
    assert s.shape == (24,) and s in single.observation_space
    s2, r, term, trunc, _ = single.step(np.array([0.5, -0.5, 1.0, 0.0], dtype=np.float32))
    assert s2.shape == (24,) and isinstance(r, float) and term is False
    single.close()
