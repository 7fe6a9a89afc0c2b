"""gym_b200 -- B200-native (sm_100a) vectorised Gym environment engine.

Drop-in for the ``step()/reset()`` hot path of openai/gym 0.26's classic-control
environments behind gym's own ``VectorEnv`` / ``Env`` API.  See DESIGN.md.

    import gym_b200
    envs = gym_b200.vector.make("CartPole-v1", num_envs=1 << 20)     # torch.cuda tensors
    obs, infos = envs.reset(seed=0)
    obs, rewards, terminateds, truncateds, infos = envs.step(actions)

    env = gym_b200.make("CartPole-v1")                                # numpy, single env
"""
from gym_b200 import error, spaces
from gym_b200.registration import EnvSpec, register, registry, spec

__version__ = "0.1.0"

__all__ = ["make", "vector", "spec", "register", "registry", "EnvSpec", "spaces", "error",
           "B200VectorEnv", "B200Env"]


def __getattr__(name):
    # torch is imported lazily so that `import gym_b200` (spaces, registry) stays light
    if name == "B200VectorEnv":
        from gym_b200.vector_env import B200VectorEnv
        return B200VectorEnv
    if name == "B200Env":
        from gym_b200.env import B200Env
        return B200Env
    if name == "vector":
        import importlib
        return importlib.import_module("gym_b200.vector")
    raise AttributeError(name)


def make(id, max_episode_steps=None, autoreset=False, disable_env_checker=None, **kwargs):
    """``gym.make`` (gym/envs/registration.py:502-691) for engine-backed ids -> single env."""
    from gym_b200.env import B200Env
    return B200Env(id, max_episode_steps=max_episode_steps, autoreset=autoreset, **kwargs)
