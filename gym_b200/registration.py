"""Environment registry: the ids, TimeLimit caps and kwargs of the in-scope envs.

Restates the rows of the reference registry table that are on the hot path
(gym/envs/__init__.py:11-60) and the small part of ``gym.make`` / ``gym.spec`` /
``gym.register`` that resolves an id to constructor arguments
(gym/envs/registration.py:117-160, 434-499, 502-691).  The wrapper stack
``make`` would build (PassiveEnvChecker -> OrderEnforcing -> TimeLimit
[-> AutoResetWrapper], registration.py:659-683) is not built as Python objects:
its step/reset semantics are fused into the CUDA step kernel.
"""
import re
from dataclasses import dataclass, field
from typing import Optional

from gym_b200 import _lib, error

# same grammar as gym/envs/registration.py:46-48
ENV_ID_RE = re.compile(r"^(?:(?P<namespace>[\w:-]+)\/)?(?:(?P<name>[\w:.-]+?))(?:-v(?P<version>\d+))?$")


@dataclass
class EnvSpec:
    """The fields of gym's EnvSpec (registration.py:117-147) that the engine uses."""
    id: str
    kind: int                                  # enum b200gym_kind
    entry_point: str = ""                      # the reference class this kind re-implements
    reward_threshold: Optional[float] = None
    nondeterministic: bool = False
    max_episode_steps: Optional[int] = None
    order_enforce: bool = True
    autoreset: bool = False
    kwargs: dict = field(default_factory=dict)

    @property
    def name(self):
        return ENV_ID_RE.fullmatch(self.id).group("name")

    @property
    def version(self):
        v = ENV_ID_RE.fullmatch(self.id).group("version")
        return None if v is None else int(v)


registry = {}


def register(id, kind, entry_point="", reward_threshold=None, max_episode_steps=None, **kwargs):
    """gym.register for engine-backed ids (registration.py:434-499)."""
    if ENV_ID_RE.fullmatch(id) is None:
        raise error.Error(f"Malformed environment ID: {id}.")
    registry[id] = EnvSpec(id=id, kind=kind, entry_point=entry_point, reward_threshold=reward_threshold,
                           max_episode_steps=max_episode_steps, kwargs=dict(kwargs))
    return registry[id]


def spec(env_id):
    """gym.spec (registration.py:694-703) with the reference's error classes for unknown ids."""
    if isinstance(env_id, EnvSpec):
        return env_id
    if env_id in registry:
        return registry[env_id]
    m = ENV_ID_RE.fullmatch(env_id)
    if m is None:
        raise error.Error(f"Malformed environment ID: {env_id}.")
    name = m.group("name")
    versions = sorted(s.version for s in registry.values() if s.name == name and s.version is not None)
    if not versions:
        raise error.NameNotFound(f"Environment {name} doesn't exist in gym_b200 "
                                 f"(available: {sorted(registry)}).")
    if m.group("version") is None:
        # registration.py:548-570: an unversioned id resolves to the latest version
        return registry[f"{name}-v{versions[-1]}"]
    raise error.VersionNotFound(f"Environment version `v{m.group('version')}` for `{name}` doesn't exist. "
                                f"It provides versioned environments: {['v%d' % v for v in versions]}.")


# gym/envs/__init__.py:11-60
register("CartPole-v0", _lib.KIND_CARTPOLE, "gym.envs.classic_control.cartpole:CartPoleEnv",
         reward_threshold=195.0, max_episode_steps=200)
register("CartPole-v1", _lib.KIND_CARTPOLE, "gym.envs.classic_control.cartpole:CartPoleEnv",
         reward_threshold=475.0, max_episode_steps=500)
register("MountainCar-v0", _lib.KIND_MOUNTAINCAR, "gym.envs.classic_control.mountain_car:MountainCarEnv",
         reward_threshold=-110.0, max_episode_steps=200)
register("MountainCarContinuous-v0", _lib.KIND_MOUNTAINCAR_CONT,
         "gym.envs.classic_control.continuous_mountain_car:Continuous_MountainCarEnv",
         reward_threshold=90.0, max_episode_steps=999)
register("Pendulum-v1", _lib.KIND_PENDULUM, "gym.envs.classic_control.pendulum:PendulumEnv",
         max_episode_steps=200)
register("Acrobot-v1", _lib.KIND_ACROBOT, "gym.envs.classic_control.acrobot:AcrobotEnv",
         reward_threshold=-100.0, max_episode_steps=500)
register("LunarLander-v2", _lib.KIND_LUNARLANDER, "gym.envs.box2d.lunar_lander:LunarLander",
         reward_threshold=200, max_episode_steps=1000)
register("LunarLanderContinuous-v2", _lib.KIND_LUNARLANDER_CONT, "gym.envs.box2d.lunar_lander:LunarLander",
         reward_threshold=200, max_episode_steps=1000, continuous=True)
register("BipedalWalker-v3", _lib.KIND_BIPEDALWALKER, "gym.envs.box2d.bipedal_walker:BipedalWalker",
         reward_threshold=300, max_episode_steps=1600)
register("BipedalWalkerHardcore-v3", _lib.KIND_BIPEDALWALKER_HARDCORE, "gym.envs.box2d.bipedal_walker:BipedalWalker",
         reward_threshold=300, max_episode_steps=2000, hardcore=True)
