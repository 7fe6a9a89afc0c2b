"""Device-side vector wrappers against numpy restatements of the reference wrappers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _update_moments(mean, var, count, x):
    # gym/wrappers/normalize.py:19-46
    bm, bv, bc = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
    delta = bm - mean
    tot = count + bc
    new_mean = mean + delta * bc / tot
    m2 = var * count + bv * bc + np.square(delta) * count * bc / tot
    return new_mean, m2 / tot, tot


def test_record_episode_statistics():
    import gym_b200
    import torch
    from gym_b200.wrappers import RecordEpisodeStatistics
    N, T = 4096, 120
    env = RecordEpisodeStatistics(gym_b200.vector.make("CartPole-v1", N), deque_size=64)
    env.reset(seed=0)
    ret = np.zeros(N, np.float32)
    length = np.zeros(N, np.int32)
    gen = torch.Generator(device="cuda").manual_seed(1)
    finished = []
    for t in range(T):
        a = torch.randint(0, 2, (N,), device="cuda", generator=gen)
        obs, rew, term, trunc, infos = env.step(a)
        r, done = rew.cpu().numpy(), (term | trunc).cpu().numpy()
        ret += r                      # record_episode_statistics.py:116-117 (float32 accumulator)
        length += 1
        assert np.array_equal(infos["_episode"].cpu().numpy(), done)
        assert np.array_equal(infos["episode"]["r"].cpu().numpy()[done], ret[done])
        assert np.array_equal(infos["episode"]["l"].cpu().numpy()[done], length[done])
        assert infos["episode"]["r"].dtype == torch.float32 and infos["episode"]["l"].dtype == torch.int32
        assert "final_observation" in infos and isinstance(infos["episode"]["t"], float)
        finished += list(zip(ret[done].tolist(), length[done].tolist()))
        ret[done] = 0
        length[done] = 0
    assert env.episode_count == len(finished) > N
    rq, lq = env.return_queue, env.length_queue
    assert len(rq) == 64 and len(lq) == 64
    # CartPole returns equal lengths (reward 1 per step); every ring slot holds ONE finished episode even though
    # ~190 episodes finish per step here and several of them land on the same slot (64-bit packed store)
    assert all(float(a) == float(b) for a, b in zip(rq, lq))
    assert set(zip(rq, lq)) <= set(finished)
    env.close()


def test_fused_episode_statistics_equal_the_standalone_kernel():
    """RecordEpisodeStatistics inside the step kernel (no extra launch) vs the stand-alone launch over the returned
    tensors (what a sharded env's gathered outputs go through): same accumulators, rows, masks and counters, for a
    classic-control kind, a Box2D kind (their own step kernels carry the same fused bookkeeping) and float rewards."""
    import gym_b200
    import torch
    from gym_b200.wrappers import RecordEpisodeStatistics, VectorWrapper
    for env_id, N, T, kw in (("CartPole-v1", 3000, 80, {}), ("Pendulum-v1", 1500, 60, dict(max_episode_steps=17)),
                             ("LunarLander-v2", 512, 140, {})):
        fused = RecordEpisodeStatistics(gym_b200.vector.make(env_id, N, **kw), deque_size=0)
        plain = RecordEpisodeStatistics(VectorWrapper(gym_b200.vector.make(env_id, N, **kw)), deque_size=0)
        assert fused._fused_target() is not None and plain._fused_target() is None
        fused.reset(seed=21)
        plain.reset(seed=21)
        gen = torch.Generator(device="cuda").manual_seed(3)
        inner = fused.env
        for t in range(T):
            if inner.discrete:
                a = torch.randint(0, inner.single_action_space.n, (N,), device="cuda", generator=gen)
            else:
                a = torch.rand((N, inner.act_dim), device="cuda", generator=gen) * 4 - 2
            fa, fb = fused.step(a), plain.step(a)
            assert torch.equal(fa[1], fb[1]) and torch.equal(fa[2], fb[2]) and torch.equal(fa[3], fb[3])
            m = fa[4]["_episode"]
            assert torch.equal(m, fb[4]["_episode"])
            assert torch.equal(fa[4]["episode"]["r"][m], fb[4]["episode"]["r"][m])
            assert torch.equal(fa[4]["episode"]["l"][m], fb[4]["episode"]["l"][m])
            assert torch.equal(fused.episode_returns, plain.episode_returns)
            assert torch.equal(fused.episode_lengths, plain.episode_lengths)
        assert fused.episode_count == plain.episode_count > 0
        fused.close()
        plain.close()


def test_normalize_observation_and_reward():
    import gym_b200
    import torch
    from gym_b200.wrappers import NormalizeObservation, NormalizeReward
    N, T = 2048, 60
    base = gym_b200.vector.make("Pendulum-v1", N, max_episode_steps=25)
    twin = gym_b200.vector.make("Pendulum-v1", N, max_episode_steps=25)
    env = NormalizeReward(NormalizeObservation(base), gamma=0.9)
    o, _ = env.reset(seed=4)
    raw, _ = twin.reset(seed=4)
    mean, var, count = np.zeros(3), np.ones(3), 1e-4
    rmean, rvar, rcount = np.zeros(()), np.ones(()), 1e-4
    returns = np.zeros(N)
    x = raw.cpu().numpy().astype(np.float64)
    mean, var, count = _update_moments(mean, var, count, x)
    np.testing.assert_allclose(o.cpu().numpy(), (x - mean) / np.sqrt(var + 1e-8), rtol=2e-5, atol=2e-6)
    gen = torch.Generator(device="cuda").manual_seed(2)
    for t in range(T):
        a = torch.rand((N, 1), device="cuda", generator=gen) * 4 - 2
        o, r, te, tr, _ = env.step(a)
        raw, rr, rte, rtr, _ = twin.step(a)
        x = raw.cpu().numpy().astype(np.float64)
        mean, var, count = _update_moments(mean, var, count, x)                      # normalize.py:83-95
        np.testing.assert_allclose(o.cpu().numpy(), (x - mean) / np.sqrt(var + 1e-8), rtol=2e-5, atol=2e-6)
        rew = rr.cpu().numpy()
        returns = returns * 0.9 + rew                                                 # :130-144
        rmean, rvar, rcount = _update_moments(rmean, rvar, rcount, returns)
        np.testing.assert_allclose(r.cpu().numpy(), rew / np.sqrt(rvar + 1e-8), rtol=1e-9)
        returns[(rte | rtr).cpu().numpy()] = 0.0
        assert torch.equal(te, rte) and torch.equal(tr, rtr)
    np.testing.assert_allclose(env.env.obs_rms.mean.cpu().numpy(), mean, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(env.env.obs_rms.var.cpu().numpy(), var, rtol=1e-9)
    np.testing.assert_allclose(env.return_rms.var.cpu().numpy()[0], rvar, rtol=1e-9)
    assert abs(float(env.env.obs_rms.count.item()) - count) < 1e-6
    env.close()
    twin.close()
