"""API contract of the engine-backed VectorEnv / Env (needs a GPU: `pytest -m gpu`).

These follow the reference's own behavioural tests: tests/vector/test_sync_vector_env.py,
tests/vector/test_vector_env_info.py, tests/vector/test_async_vector_env.py (call ordering),
tests/wrappers/test_time_limit.py, tests/wrappers/test_autoreset.py,
tests/envs/test_action_dim_check.py, tests/envs/test_env_implementation.py (reset bounds),
tests/envs/test_envs.py (determinism).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shapes_dtypes_and_spaces():
    import gym_b200
    import torch
    from gym_b200 import spaces
    envs = gym_b200.vector.make("CartPole-v1", num_envs=8)
    assert envs.num_envs == 8 and envs.is_vector_env and not envs.closed
    assert isinstance(envs.single_action_space, spaces.Discrete)
    assert isinstance(envs.action_space, spaces.MultiDiscrete) and envs.action_space.shape == (8,)
    assert envs.observation_space.shape == (8, 4) and envs.single_observation_space.shape == (4,)
    obs, infos = envs.reset(seed=0)
    assert obs.is_cuda and obs.dtype == torch.float32 and tuple(obs.shape) == (8, 4) and infos == {}
    # python lists, numpy arrays and tensors of several integer dtypes are all accepted
    # (tests/vector/test_sync_vector_env.py:49-52 passes a list of ints)
    for actions in ([0, 1] * 4, np.array([1, 0] * 4), torch.tensor([0, 1] * 4, dtype=torch.int32),
                    torch.tensor([0, 1] * 4, dtype=torch.uint8, device="cuda"), envs.action_space.sample()):
        obs, rew, term, trunc, infos = envs.step(actions)
        assert tuple(obs.shape) == (8, 4) and obs.dtype == torch.float32
        assert tuple(rew.shape) == (8,) and rew.dtype == torch.float64
        assert term.dtype == torch.bool and trunc.dtype == torch.bool and tuple(term.shape) == (8,)
        assert "final_observation" in infos and tuple(infos["final_observation"].shape) == (8, 4)
        assert infos["_final_observation"].dtype == torch.bool
    envs.close()
    p = gym_b200.vector.make("Pendulum-v1", num_envs=3)
    p.reset(seed=1)
    obs, rew, *_ = p.step(np.zeros((3, 1), dtype=np.float64))  # float64 input is cast to the Box dtype
    assert tuple(obs.shape) == (3, 3)
    p.close()


def test_order_enforcing_closed_and_async_state_machine():
    import gym_b200
    from gym_b200 import error
    envs = gym_b200.vector.make("MountainCar-v0", num_envs=4)
    with pytest.raises(error.ResetNeeded):          # gym/wrappers/order_enforcing.py:33-37
        envs.step([0, 1, 2, 0])
    with pytest.raises(error.NoAsyncCallError):     # tests/vector/test_async_vector_env.py:184-252
        envs.step_wait()
    with pytest.raises(error.NoAsyncCallError):
        envs.reset_wait()
    envs.reset_async(seed=0)
    with pytest.raises(error.AlreadyPendingCallError):
        envs.reset_async()
    with pytest.raises(error.AlreadyPendingCallError):
        envs.step_async([0, 0, 0, 0])
    envs.reset_wait()
    envs.step_async([0, 1, 2, 0])
    with pytest.raises(error.AlreadyPendingCallError):
        envs.step_async([0, 1, 2, 0])
    envs.step_wait()
    envs.close()
    envs.close()                                     # idempotent (vector_env.py:201-206)
    assert envs.closed
    with pytest.raises(error.ClosedEnvironmentError):  # async_vector_env.py:518-522
        envs.step([0, 0, 0, 0])
    with pytest.raises(error.ClosedEnvironmentError):
        envs.reset()


def test_invalid_actions_raise():
    import gym_b200
    import torch
    from gym_b200 import error
    # numpy backend validates on the host before launching (tests/envs/test_action_dim_check.py:62-78)
    envs = gym_b200.vector.make("CartPole-v1", num_envs=4, backend="numpy")
    envs.reset(seed=0)
    with pytest.raises(Exception):
        envs.step(np.array([0, 1, 2, 0]))
    with pytest.raises(Exception):
        envs.step(np.array([0.5, 1, 0, 0]))
    envs.close()
    # torch backend: device-side sticky counter, checked on demand; the offending env is left untouched
    envs = gym_b200.vector.make("Acrobot-v1", num_envs=4)
    envs.reset(seed=0)
    st0, el0, _ = envs.get_state()
    obs, rew, term, trunc, _ = envs.step(torch.tensor([0, 3, -1, 2], device="cuda"))
    st1, el1, _ = envs.get_state()
    assert torch.equal(st0[1:3], st1[1:3]) and el1.tolist() == [1, 0, 0, 1]
    assert torch.isnan(rew[1:3]).all() and not torch.isnan(rew[[0, 3]]).any()
    with pytest.raises(error.InvalidAction):
        envs.check_actions()
    assert envs.check_actions() == 0
    # ... and without asking: the kernels raise a host-visible flag, so the NEXT step() raises the reference's error
    # (one call late, no synchronisation on the clean path) and the flag is consumed by it
    envs.step(torch.tensor([0, 1, 7, 2], device="cuda"))
    torch.cuda.synchronize()
    with pytest.raises(error.InvalidAction):
        envs.step(torch.tensor([0, 1, 2, 2], device="cuda"))
    envs.step(torch.tensor([0, 1, 2, 2], device="cuda"))
    with pytest.raises(error.InvalidAction):
        envs.step(torch.zeros(4, device="cuda"))     # float actions for a Discrete space
    with pytest.raises(error.InvalidAction):
        envs.step(torch.zeros(5, dtype=torch.int64))  # wrong batch size
    envs.close()
    single = gym_b200.make("CartPole-v1")
    single.reset(seed=0)
    with pytest.raises(Exception):
        single.step(2)
    single.close()


def test_time_limit_truncation_and_both_flags():
    """tests/wrappers/test_time_limit.py:17-57: truncation at N; terminated and truncated may both be set."""
    import gym_b200
    import torch
    envs = gym_b200.vector.make("Pendulum-v1", num_envs=2, max_episode_steps=5)
    envs.reset(seed=0)
    for t in range(1, 12):
        _, _, term, trunc, infos = envs.step(torch.zeros(2, 1, device="cuda"))
        assert trunc.tolist() == [t % 5 == 0] * 2 and not term.any()
        assert infos["_final_observation"].tolist() == [t % 5 == 0] * 2
    envs.close()
    # CartPole pushed one way falls after ~9-10 steps; a TimeLimit on that very step sets both flags
    probe = gym_b200.vector.make("CartPole-v1", num_envs=1)
    probe.reset(seed=0)
    n = 0
    while True:
        n += 1
        _, _, term, _, _ = probe.step([0])
        if term.item():
            break
    probe.close()
    both = gym_b200.vector.make("CartPole-v1", num_envs=1, max_episode_steps=n)
    both.reset(seed=0)
    for _ in range(n):
        _, _, term, trunc, _ = both.step([0])
    assert term.item() and trunc.item()
    both.close()


def test_autoreset_semantics_match_appendix_b():
    """SyncVectorEnv(2 x CartPole-v1).reset(seed=0), constant actions [0, 1], step 9."""
    import gym_b200
    envs = gym_b200.vector.make("CartPole-v1", num_envs=2, backend="numpy")
    envs.reset(seed=0)
    for t in range(9):
        obs, rew, term, trunc, infos = envs.step(np.array([0, 1]))
    assert term.tolist() == [False, True] and trunc.tolist() == [False, False] and rew.tolist() == [1.0, 1.0]
    np.testing.assert_allclose(obs[0], [-0.13062044978141785, -1.7782915830612183, 0.1494298130273819,
                                        2.5877106189727783], rtol=1e-5)
    np.testing.assert_allclose(obs[1], [-0.018816854804754257, -0.007667355239391327, 0.03277026116847992,
                                        -0.009080085903406143], rtol=1e-5)
    assert infos["final_observation"][0] is None
    np.testing.assert_allclose(infos["final_observation"][1], [0.15024752914905548, 1.8084592819213867,
                                                               -0.25012344121932983, -2.820631980895996], rtol=1e-5)
    assert infos["_final_observation"].tolist() == [False, True]
    assert infos["final_info"][1] == {} and infos["_final_info"].tolist() == [False, True]
    envs.close()


def test_reset_options_bounds():
    """tests/envs/test_env_implementation.py:150-215"""
    import gym_b200
    envs = gym_b200.vector.make("CartPole-v1", num_envs=256)
    obs, _ = envs.reset(seed=0, options={"low": 0.1, "high": 0.2})
    assert bool(((obs >= 0.1) & (obs <= 0.2)).all())
    with pytest.raises(ValueError):
        envs.reset(options={"low": 0.3, "high": 0.2})
    with pytest.raises(ValueError):
        envs.reset(options={"low": "a"})
    obs, _ = envs.reset()  # defaults again, stream continues
    assert bool((obs.abs() <= 0.05).all())
    envs.close()
    pend = gym_b200.vector.make("Pendulum-v1", num_envs=256)
    obs, _ = pend.reset(seed=0, options={"x_init": 0.2, "y_init": 0.1})
    st, _, _ = pend.get_state()
    assert bool((st[:, 0].abs() <= 0.2).all()) and bool((st[:, 1].abs() <= 0.1).all())
    pend.close()


def test_output_double_buffering_and_copy():
    import gym_b200
    import torch
    envs = gym_b200.vector.make("CartPole-v1", num_envs=16)
    obs0, _ = envs.reset(seed=0)
    keep = obs0.clone()
    obs1, *_ = envs.step(torch.zeros(16, dtype=torch.int64, device="cuda"))
    assert torch.equal(obs0, keep), "the previous step's tensors stay valid for one more step"
    assert obs1.data_ptr() != obs0.data_ptr()
    envs.close()
    envs = gym_b200.vector.make("CartPole-v1", num_envs=16, copy=True)
    a, _ = envs.reset(seed=0)
    b, *_ = envs.step(torch.zeros(16, dtype=torch.int64, device="cuda"))
    c, *_ = envs.step(torch.zeros(16, dtype=torch.int64, device="cuda"))
    assert len({a.data_ptr(), b.data_ptr(), c.data_ptr()}) == 3
    envs.close()


def test_seeding_variants():
    import gym_b200
    import torch
    from gym_b200 import error
    a = gym_b200.vector.make("CartPole-v1", num_envs=4)
    b = gym_b200.vector.make("CartPole-v1", num_envs=4)
    oa, _ = a.reset(seed=[3, 4, 5, 6])
    ob, _ = b.reset(seed=3)
    assert torch.equal(oa, ob)                        # list seeds == int fan-out
    oa2, _ = a.reset(seed=[None, 4, None, 6])         # None keeps that env's stream
    ob2, _ = b.reset()
    assert torch.equal(oa2[[0, 2]], ob2[[0, 2]]) and torch.equal(oa2[[1, 3]], oa[[1, 3]])
    with pytest.raises(error.Error):
        a.reset(seed=-1)
    u1 = gym_b200.vector.make("CartPole-v1", num_envs=4)
    u2 = gym_b200.vector.make("CartPole-v1", num_envs=4)
    assert not torch.equal(u1.reset()[0], u2.reset()[0])  # unseeded: OS entropy
    for e in (a, b, u1, u2):
        e.close()


def test_call_and_attrs():
    import gym_b200
    envs = gym_b200.vector.make("Pendulum-v1", num_envs=3, g=9.81)
    assert envs.get_attr("g") == (9.81, 9.81, 9.81)
    assert envs.call("max_torque") == (2.0, 2.0, 2.0)
    with pytest.raises(AttributeError):
        envs.get_attr("nope")
    with pytest.raises(AttributeError):
        envs.set_attr("g", 1.0)
    assert "Pendulum-v1" in repr(envs)
    envs.close()


def test_single_env_facade():
    """gym.make-shaped env: numpy in/out, no autoreset, CartPole's reward after termination."""
    import gym_b200
    from gym_b200 import error
    env = gym_b200.make("CartPole-v1")
    with pytest.raises(error.ResetNeeded):
        env.step(0)
    obs, info = env.reset(seed=0)
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float32 and obs.shape == (4,) and info == {}
    np.testing.assert_allclose(obs, [0.013696168549358845, -0.023021329194307327, -0.04590264707803726,
                                     -0.04834723472595215], rtol=1e-6)
    assert obs in env.observation_space
    expected = [[0.013235742226243019, -0.21745604276657104, -0.04686959087848663, 0.2295069843530655],
                [0.008886621333658695, -0.021696746349334717, -0.042279452085494995, -0.07758410274982452]]
    for a, want in zip([0, 1], expected):
        obs, reward, terminated, truncated, info = env.step(a)
        assert type(reward) is float and reward == 1.0 and terminated is False and truncated is False and info == {}
        np.testing.assert_allclose(obs, want, rtol=1e-5)
    # run to termination: the terminating step pays 1.0, later steps 0.0 (cartpole.py:169-184)
    terminated = False
    while not terminated:
        obs, reward, terminated, truncated, _ = env.step(0)
    assert reward == 1.0
    obs2, reward2, terminated2, _, _ = env.step(0)
    assert reward2 == 0.0 and terminated2 and not np.array_equal(obs, obs2)
    obs, _ = env.reset()
    assert np.all(np.abs(obs) <= 0.05)
    env.close()
    # determinism rollout (tests/envs/test_envs.py:60-115)
    e1, e2 = gym_b200.make("Acrobot-v1"), gym_b200.make("Acrobot-v1")
    o1, _ = e1.reset(seed=7)
    o2, _ = e2.reset(seed=7)
    assert np.array_equal(o1, o2)
    e1.action_space.seed(7)
    for _ in range(50):
        a = e1.action_space.sample()
        r1, r2 = e1.step(a), e2.step(a)
        assert np.array_equal(r1[0], r2[0]) and r1[1:4] == r2[1:4]
        assert r1[0] in e1.observation_space
    e1.close()
    e2.close()
    # Box env: out-of-bound actions act like the bound (tests/envs/test_action_dim_check.py:90-136)
    p1, p2 = gym_b200.make("MountainCarContinuous-v0"), gym_b200.make("MountainCarContinuous-v0")
    p1.reset(seed=42)
    p2.reset(seed=42)
    o1 = p1.step(np.array([1.0], dtype=np.float32))[0]
    o2 = p2.step(np.array([101.0], dtype=np.float32))[0]
    assert np.array_equal(o1, o2)
    p1.close()
    p2.close()
    with gym_b200.make("Pendulum-v1", autoreset=True, max_episode_steps=3) as env:
        env.reset(seed=0)
        for t in range(3):
            obs, reward, terminated, truncated, info = env.step(np.array([0.0], dtype=np.float32))
        assert isinstance(reward, np.float64) and truncated and "final_observation" in info
