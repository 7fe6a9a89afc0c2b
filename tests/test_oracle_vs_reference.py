"""Oracle (and the host-side space objects) against the LIVE reference.

Runs only where /root/reference exists (the build container); on the GPU box
the committed fixtures in tests/golden stand in for it.
"""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import REFERENCE

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")


@pytest.fixture(scope="module")
def gym():
    for n, v in (("bool8", np.bool_), ("float_", np.float64)):
        if not hasattr(np, n):
            setattr(np, n, v)
    sys.path.insert(0, REFERENCE)
    warnings.filterwarnings("ignore")
    import gym as ref
    yield ref
    sys.path.remove(REFERENCE)


def _actions(env_id, rng, T, N):
    if env_id.startswith("CartPole"):
        return rng.integers(0, 2, size=(T, N))
    if env_id.startswith(("MountainCar-", "Acrobot")):
        return rng.integers(0, 3, size=(T, N))
    lim = 2.5 if env_id.startswith("Pendulum") else 1.3
    return rng.uniform(-lim, lim, size=(T, N, 1)).astype(np.float32)


@pytest.mark.parametrize("env_id", ["CartPole-v1", "MountainCar-v0", "MountainCarContinuous-v0",
                                    "Pendulum-v1", "Acrobot-v1"])
def test_oracle_bit_exact_against_live_reference(gym, oracle_mod, env_id):
    N, T, seed = 6, 260, 4242
    rng = np.random.default_rng(77)
    acts = _actions(env_id, rng, T, N)
    envs = gym.vector.SyncVectorEnv([lambda: gym.make(env_id, disable_env_checker=True) for _ in range(N)])
    v = oracle_mod.OracleVec(env_id, N)
    ref_obs, _ = envs.reset(seed=seed)
    obs = v.reset(seed=seed)
    if env_id.startswith("Acrobot"):
        np.testing.assert_allclose(obs, ref_obs, rtol=2e-7)
    else:
        assert np.array_equal(obs, ref_obs)
    for t in range(T):
        ro, rr, rte, rtr, info = envs.step(acts[t])
        o, r, te, tr, fo = v.step(acts[t])
        assert np.array_equal(te, rte) and np.array_equal(tr, rtr)
        assert np.array_equal(r, rr)
        done = te | tr
        if env_id.startswith("Acrobot") and done.any():
            np.testing.assert_allclose(o, ro, rtol=2e-7)  # numpy's float32 trig on reset rows
            assert np.array_equal(o[~done], ro[~done])
        else:
            assert np.array_equal(o, ro)
        for i in np.flatnonzero(done):
            assert np.array_equal(info["final_observation"][i], fo[i])
    envs.close()
    v.close()


def test_space_objects_sample_like_the_reference(gym):
    from gym_b200 import spaces
    from gym_b200.envs import KINDS
    ref_ids = {0: "CartPole-v1", 1: "MountainCar-v0", 2: "MountainCarContinuous-v0", 3: "Pendulum-v1",
               4: "Acrobot-v1"}
    for kind, env_id in ref_ids.items():
        ref = gym.make(env_id, disable_env_checker=True)
        obs_space, act_space = KINDS[kind].spaces(None)
        assert obs_space.shape == ref.observation_space.shape
        assert np.array_equal(obs_space.low, ref.observation_space.low)
        assert np.array_equal(obs_space.high, ref.observation_space.high)
        ref.action_space.seed(5)
        act_space.seed(5)
        for _ in range(5):
            assert np.array_equal(np.asarray(act_space.sample()), np.asarray(ref.action_space.sample()))
        rb = gym.vector.utils.batch_space(ref.action_space, 7)
        mb = spaces.batch_space(act_space, 7)
        rb.seed(9)
        mb.seed(9)
        assert np.array_equal(mb.sample(), rb.sample())
        assert type(mb).__name__ == type(rb).__name__ and mb.shape == rb.shape and mb.dtype == rb.dtype
        ref.close()


def test_plugin_registers_engine_ids_in_the_real_registry(gym):
    """gym/envs/registration.py:266-309,434-499: the ids resolve through gym's own registry and gym.make
    reaches the engine's constructor (which refuses to run without a B200: no CPU fallback)."""
    import gym_b200
    from gym_b200 import error, plugin
    ids = plugin.register_all()
    assert "B200/CartPole-v1" in ids and "B200/BipedalWalkerHardcore-v3" in ids
    for env_id in ids:
        s = gym.spec(env_id)
        assert s.namespace == "B200" and s.entry_point == "gym_b200.plugin:make_env"
        assert s.max_episode_steps is None and s.order_enforce is False and s.disable_env_checker is True
        assert s.kwargs == {"env_id": env_id.split("/", 1)[1]}
        assert s.reward_threshold == gym_b200.spec(s.kwargs["env_id"]).reward_threshold
    # the reference's own ids are untouched
    assert gym.spec("CartPole-v1").entry_point == "gym.envs.classic_control.cartpole:CartPoleEnv"
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(error.DependencyNotInstalled):
            gym.make("B200/CartPole-v1")
    else:
        env = gym.make("B200/CartPole-v1")
        obs, info = env.reset(seed=0)
        assert obs.shape == (4,) and env.spec.id == "B200/CartPole-v1"
        env.close()
