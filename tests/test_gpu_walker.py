"""BipedalWalker-v3 / BipedalWalkerHardcore-v3 on the GPU against the CPU oracle (oracle/walker_oracle.c).  `pytest -m gpu`.

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
    assert env.box2d_overflows() == 0 and orc.overflows() == 0   # 10 touching pairs per env were always enough
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
    # the feet start 2 leg lengths above flat ground and reach it within reset()'s embedded step(0): with continuous
    # collision the contact listener already fires in that step (the TOI sub-step), so both flags are up
    assert bool(((obs[:, [8, 13]] == 0) | (obs[:, [8, 13]] == 1)).all())
    gen = torch.Generator(device="cuda").manual_seed(0)
    total_term = 0
    for t in range(200):
        a = torch.rand((N, 4), device="cuda", generator=gen) * 2 - 1
        o, r, te, tr, info = env.step(a)
        assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(r).all())
        assert bool(((o[:, [8, 13]] == 0) | (o[:, [8, 13]] == 1)).all())
        assert bool((o[:, 14:] >= 0).all()) and bool((o[:, 14:] <= 1).all())
        # an episode under random torques ends with the hull on the ground (-100); allow the odd walker that
        # leaves the course at the far end instead (terminates without the penalty)
        assert int((r[te] != -100).sum()) <= max(1, int(te.sum()) // 500)
        assert bool((r[~te] > -30).all())       # shaping deltas: a few points per step at most
        total_term += int(te.sum())
    assert total_term > N
    with pytest.raises(TypeError):
        gym_b200.vector.make("BipedalWalker-v3", 4, continuous=True)
    env.close()
    single = gym_b200.make("BipedalWalker-v3")
    s, _ = single.reset(seed=3)
    assert s.shape == (24,) and s in single.observation_space
    s2, r, term, trunc, _ = single.step(np.array([0.5, -0.5, 1.0, 0.0], dtype=np.float32))
    assert s2.shape == (24,) and isinstance(r, float) and term is False
    single.close()


def _compare_step(t, got, want):
    o, r, te, tr, info = got
    ro, rr, rte, rtr, rfo = want
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
    return done


@pytest.mark.parametrize("env_id,hardcore", [("BipedalWalker-v3", False), ("BipedalWalkerHardcore-v3", True)])
def test_demo_gait_bit_exact_against_oracle(env_id, hardcore):
    """The reference's demo controller (bipedal_walker.py:775-854) drives every env: the walkers cross the course
    (plain) or run into the stumps / stairs / pits (hardcore: polygon-polygon manifolds, boxes seen by the lidar),
    and the kernel still agrees with the CPU oracle bit for bit."""
    import gym_b200
    import torch
    from oracle.oracle import OracleWalker, WalkerHeuristic
    N, T, seed = 96, 700, 40
    env = gym_b200.vector.make(env_id, N)
    assert env.max_episode_steps == (2000 if hardcore else 1600)
    assert env.get_attr("hardcore") == (hardcore,) * N
    orc = OracleWalker(N, hardcore=hardcore, max_episode_steps=env.max_episode_steps)
    obs, _ = env.reset(seed=seed)
    ref = orc.reset(seed=seed)
    assert np.array_equal(obs.cpu().numpy(), ref)
    terrain, polys, npoly = env.walker_terrain()
    for i in range(N):
        assert np.array_equal(terrain[i].cpu().numpy(), orc.terrain(i))
        want = orc.polys(i)
        assert int(npoly[i]) == len(want) and np.array_equal(polys[i, :len(want)].cpu().numpy(), want)
    assert bool((npoly > 8).all()) if hardcore else bool((npoly == 0).all())
    gaits = [WalkerHeuristic() for _ in range(N)]
    a = np.zeros((N, 4), dtype=np.float32)
    n_done, min_lidar, far = 0, 1.0, 0.0
    for t in range(T):
        got = env.step(torch.as_tensor(a, device=env.device))
        want = orc.step(a)
        done = _compare_step(t, got, want)
        n_done += int(done.sum())
        min_lidar = min(min_lidar, float(want[0][:, 14:].min()))
        for i in range(N):
            if done[i]:
                gaits[i] = WalkerHeuristic()
                a[i] = 0.0
            else:
                a[i] = gaits[i](want[0][i])
        if t % 50 == 0:
            bodies, _ = env.walker_bodies()
            far = max(far, float(bodies[:, 0, 0].max()))
    if hardcore:
        assert n_done >= N // 2 and min_lidar < 0.4      # the gait falls at the obstacles
        assert far > 9.0                                   # ... having walked up to them (first obstacle at x >= 9.33)
    else:
        assert far > 30.0                                  # a third of the course in 700 steps
    bodies, flags = env.walker_bodies()
    for i in (0, N // 2, N - 1):
        ob, of = orc.bodies(i)
        assert np.array_equal(bodies[i].cpu().numpy(), ob)
        assert flags[i].tolist() == of.tolist()
    env.close()
    orc.close()


def test_hardcore_random_actions_bit_exact_and_kwarg_selects_the_variant():
    """BipedalWalker-v3 with hardcore=True is BipedalWalkerHardcore-v3 minus the registry's TimeLimit
    (gym/envs/__init__.py:70-85); random torques, many resets (fresh obstacle layouts)."""
    import gym_b200
    import torch
    from oracle.oracle import OracleWalker
    N, T, seed = 512, 250, 77
    env = gym_b200.vector.make("BipedalWalker-v3", N, hardcore=True)
    assert env.max_episode_steps == 1600
    orc = OracleWalker(N, hardcore=True, max_episode_steps=1600)
    assert np.array_equal(env.reset(seed=seed)[0].cpu().numpy(), orc.reset(seed=seed))
    rng = np.random.default_rng(6)
    n_term = 0
    for t in range(T):
        a = rng.uniform(-1.2, 1.2, size=(N, 4)).astype(np.float32)
        done = _compare_step(t, env.step(torch.as_tensor(a, device=env.device)), orc.step(a))
        n_term += int(done.sum())
    assert n_term > N
    terrain, polys, npoly = env.walker_terrain()
    for i in (0, 100, N - 1):
        want = orc.polys(i)
        assert int(npoly[i]) == len(want) and np.array_equal(polys[i, :len(want)].cpu().numpy(), want)
    assert env.box2d_overflows() == 0 and orc.overflows() == 0   # fallen walkers among stumps and stairs included
    env.close()
    orc.close()


# ---------------------------------------------------------------------------------------------------------
# frozen roll-outs of the oracle (tests/golden_self): LunarLander and BipedalWalker variants through the public API
# ---------------------------------------------------------------------------------------------------------
import self_fixtures  # noqa: E402


@pytest.mark.parametrize("name", self_fixtures.names())
def test_kernels_reproduce_the_frozen_rollouts(name):
    """Regression pins of the re-derived Box2D physics (not reference data: see oracle/gen_self_fixtures.py): the
    frozen actions replayed on the GPU give the frozen observations, rewards, flags and final observations."""
    import gym_b200
    import torch
    d = self_fixtures.load(name)
    env_id, kw = self_fixtures.env_id(d)
    env = gym_b200.vector.make(env_id, d["n"], max_episode_steps=d["max_episode_steps"], **kw)
    obs, _ = env.reset(seed=d["seed"])

    def step(a):
        o, r, te, tr, info = env.step(torch.as_tensor(a, device=env.device))
        return (o.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy(),
                info["final_observation"].cpu().numpy())

    self_fixtures.check(d, step, obs.cpu().numpy())
    env.close()
