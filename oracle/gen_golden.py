"""Generate tests/golden/*.npz from the REAL reference (openai/gym 0.26.2).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

It imports gym from /root/reference (with the two numpy-2 aliases the frozen
reference needs), rolls `gym.vector.SyncVectorEnv`
forward on seeded random actions and freezes (actions, obs, reward, terminated,
truncated, final_observation) as small fixtures.  It also checks the C oracle
(oracle/gym_oracle.c) against the reference on the same inputs and prints how
many values are bit-identical.  Test infrastructure, not product code.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

for _n, _v in (("bool8", np.bool_), ("float_", np.float64)):
    if not hasattr(np, _n):
        setattr(np, _n, _v)
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import gym  # noqa: E402  (the reference)
from oracle.oracle import OracleVec, seed_sequence  # noqa: E402

assert gym.__version__ == "0.26.2", gym.__version__


def make_actions(env_id, rng, T, N, wild):
    """Seeded random actions; `wild` also draws Box actions beyond the bounds
    (tests/envs/test_action_dim_check.py:90-136 pins the clipping)."""
    if env_id.startswith(("CartPole",)):
        return rng.integers(0, 2, size=(T, N)).astype(np.int64)
    if env_id.startswith(("MountainCar-", "Acrobot")):
        return rng.integers(0, 3, size=(T, N)).astype(np.int64)
    if env_id.startswith("Pendulum"):
        lim = 3.0 if wild else 2.0
        return rng.uniform(-lim, lim, size=(T, N, 1)).astype(np.float32)
    if env_id.startswith("MountainCarContinuous"):
        lim = 1.5 if wild else 1.0
        return rng.uniform(-lim, lim, size=(T, N, 1)).astype(np.float32)
    raise KeyError(env_id)


def rollout_reference(env_id, N, seed, actions, max_episode_steps=None, options=None, make_kwargs=None):
    kw = dict(make_kwargs or {})
    if max_episode_steps is not None:
        kw["max_episode_steps"] = max_episode_steps
    envs = gym.vector.SyncVectorEnv(
        [lambda: gym.make(env_id, disable_env_checker=True, **kw) for _ in range(N)])
    D = envs.single_observation_space.shape[0]
    T = actions.shape[0]
    obs0, _ = envs.reset(seed=seed, options=options)
    obs = np.zeros((T, N, D), np.float32)
    rew = np.zeros((T, N), np.float64)
    term = np.zeros((T, N), bool)
    trunc = np.zeros((T, N), bool)
    fobs = np.zeros((T, N, D), np.float32)
    fmask = np.zeros((T, N), bool)
    for t in range(T):
        o, r, te, tr, info = envs.step(actions[t])
        obs[t], rew[t], term[t], trunc[t] = o, r, te, tr
        if "final_observation" in info:
            for i in range(N):
                if info["_final_observation"][i]:
                    fobs[t, i] = info["final_observation"][i]
                    fmask[t, i] = True
    envs.close()
    return dict(obs0=obs0.astype(np.float32), obs=obs, reward=rew, terminated=term,
                truncated=trunc, final_obs=fobs, final_mask=fmask)


def rollout_oracle(env_id, N, seed, actions, max_episode_steps=None, bounds=None, param0=None):
    v = OracleVec(env_id, N, max_episode_steps=max_episode_steps, param0=param0)
    T = actions.shape[0]
    D = v.obs_dim
    obs0 = v.reset(seed=seed, bounds=bounds)
    out = dict(obs0=obs0, obs=np.zeros((T, N, D), np.float32), reward=np.zeros((T, N)),
               terminated=np.zeros((T, N), bool), truncated=np.zeros((T, N), bool),
               final_obs=np.zeros((T, N, D), np.float32), final_mask=np.zeros((T, N), bool))
    for t in range(T):
        o, r, te, tr, fo = v.step(actions[t])
        done = te | tr
        out["obs"][t], out["reward"][t], out["terminated"][t], out["truncated"][t] = o, r, te, tr
        out["final_obs"][t][done] = fo[done]
        out["final_mask"][t] = done
    v.close()
    return out


def compare(name, ref, orc):
    msgs = []
    for k in ("terminated", "truncated", "final_mask"):
        bad = int((ref[k] != orc[k]).sum())
        msgs.append(f"{k}:{'ok' if bad == 0 else f'{bad} MISMATCH'}")
    for k in ("obs0", "obs", "final_obs"):
        a, b = ref[k], orc[k]
        nb = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        rel = float(np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(a), 1e-30))) if nb else 0.0
        msgs.append(f"{k}: {a.size - nb}/{a.size} bit-exact" + (f" (max rel {rel:.2e})" if nb else ""))
    a, b = ref["reward"], orc["reward"]
    nb = int((a != b).sum())
    msgs.append(f"reward: {a.size - nb}/{a.size} bit-exact" +
                (f" (max abs {np.max(np.abs(a - b)):.2e})" if nb else ""))
    print(f"[{name}] " + "; ".join(msgs))


CASES = [
    # name, env_id, N, T, seed, max_episode_steps, options, bounds, wild, make_kwargs, param0
    ("cartpole_v1", "CartPole-v1", 16, 400, 0, None, None, None, False, None, None),
    ("cartpole_v1_limit10", "CartPole-v1", 8, 64, 123, 10, None, None, False, None, None),
    ("cartpole_v0_bounds", "CartPole-v0", 8, 64, 7, None, {"low": -0.1, "high": 0.2}, (-0.1, 0.2), False, None, None),
    ("mountaincar_v0", "MountainCar-v0", 16, 450, 1, None, None, None, False, None, None),
    ("mountaincar_cont_v0", "MountainCarContinuous-v0", 16, 1100, 2, None, None, None, True, None, None),
    ("mountaincar_cont_v0_limit50", "MountainCarContinuous-v0", 8, 200, 5, 50, None, None, False, None, None),
    ("pendulum_v1", "Pendulum-v1", 16, 450, 3, None, None, None, True, None, None),
    ("pendulum_v1_init", "Pendulum-v1", 8, 64, 11, None, {"x_init": 0.5, "y_init": 0.25}, (0.5, 0.25), False, None, None),
    ("pendulum_v1_g", "Pendulum-v1", 8, 64, 12, None, None, None, False, {"g": 9.81}, 9.81),
    ("acrobot_v1", "Acrobot-v1", 16, 600, 4, None, None, None, False, None, None),
    ("acrobot_v1_limit40", "Acrobot-v1", 8, 128, 99, 40, None, None, False, None, None),
    ("cartpole_v1_bigseed", "CartPole-v1", 4, 64, 2**40 + 7, None, None, None, False, None, None),
]


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    # --- RNG known-answer vectors straight from numpy ----------------------
    seeds = [0, 1, 2, 123, 2**31, 2**32 - 1, 2**32, 2**40 + 7, 2**64 + 5, 10**30]
    ss = np.stack([np.random.SeedSequence(s).generate_state(4, np.uint64) for s in seeds])
    dbl = np.stack([np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))).random(8) for s in seeds])
    for s, w in zip(seeds, ss):
        assert (seed_sequence(s) == w).all(), s
    np.savez(os.path.join(GOLDEN, "rng_kat.npz"),
             seeds_lo=np.array([s & (2**64 - 1) for s in seeds], dtype=np.uint64),
             seeds_hi=np.array([s >> 64 for s in seeds], dtype=np.uint64),
             seed_sequence=ss, doubles=dbl)
    print(f"[rng_kat] {len(seeds)} seeds: SeedSequence words bit-exact")

    for (name, env_id, N, T, seed, mes, options, bounds, wild, mk, p0) in CASES:
        arng = np.random.default_rng(1000 + (seed % 1000))
        actions = make_actions(env_id, arng, T, N, wild)
        ref = rollout_reference(env_id, N, seed, actions, mes, options, mk)
        orc = rollout_oracle(env_id, N, seed, actions, mes, bounds, p0)
        compare(name, ref, orc)
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), env_id=env_id, N=N, T=T,
            seed_lo=np.uint64(seed & (2**64 - 1)), seed_hi=np.uint64(seed >> 64),
            max_episode_steps=-1 if mes is None else mes,
            bounds=np.array(bounds if bounds is not None else [np.nan, np.nan]),
            param0=np.nan if p0 is None else p0, actions=actions, **ref)
    print("numpy", np.__version__, "gym", gym.__version__)


if __name__ == "__main__":
    main()
