"""Registration of the engine-backed ids inside a real ``gym`` installation.

gym resolves ``gym.make(id)`` through its registry (gym/envs/registration.py:434-499,
502-691) and fills that registry, at ``import gym``, from entry points of the group
``"gym.envs"`` (registration.py:266-309): for a plugin named ``B200`` it runs the plugin's
function inside ``with namespace("B200")``, so every ``register(id=...)`` the function
makes lands under ``B200/<id>``.  This module is that plugin:

    # pyproject.toml of this repo
    [project.entry-points."gym.envs"]
    B200 = "gym_b200.plugin:register_envs"

    import gym
    env = gym.make("B200/CartPole-v1")            # -> gym_b200.gym_compat.GymEnv (a gym.Env) on cuda:0

or, without installing the entry point, ``gym_b200.plugin.register_all()`` after ``import gym``.

The wrapper stack ``gym.make`` would add is switched off in the spec: TimeLimit
(``max_episode_steps=None``; the limit of the reference id is fused in the step kernel),
OrderEnforcing (``B200Env`` raises ``ResetNeeded`` itself) and the passive checker (it insists on
``gym.spaces`` classes).  Batched use goes through ``gym_b200.vector.make``: ``gym.vector.make``
would build one single-env object per sub-environment.
"""
from gym_b200.registration import registry

NAMESPACE = "B200"


def make_env(env_id, **kwargs):
    """The ``entry_point`` of every registered id: what ``gym.make`` calls as ``env_creator(**kwargs)``
    (registration.py:640)."""
    from gym_b200.gym_compat import GymEnv   # a real gym.Env subclass: this function only runs inside gym
    return GymEnv(env_id, **kwargs)


def register_envs():
    """Entry-point function of group ``"gym.envs"``; gym supplies the namespace."""
    from gym.envs.registration import register
    for spec in registry.values():
        register(id=spec.id, entry_point="gym_b200.plugin:make_env", reward_threshold=spec.reward_threshold,
                 nondeterministic=False, max_episode_steps=None, order_enforce=False, autoreset=False,
                 disable_env_checker=True, kwargs={"env_id": spec.id})


def register_all(namespace=NAMESPACE):
    """Manual registration (no installed entry point): ``B200/<id>`` for every engine-backed id."""
    from gym.envs.registration import namespace as _namespace
    with _namespace(namespace):
        register_envs()
    return [f"{namespace}/{env_id}" for env_id in registry]
