"""Freeze roll-outs of the Box2D-task ORACLE as regression fixtures (tests/golden_self/*.npz).

NOT reference data: box2d-py cannot run here, so these come from oracle/lunar_oracle.c / walker_oracle.c themselves
(parity with the real library stays unpinned).  Their job is to hold the re-derived physics still
across rounds: the oracle, the host build of the device source (tests/hostsim) and the CUDA kernels must all keep
reproducing them bit for bit, so a change that moves the oracle AND the kernels together is still caught.

    python oracle/gen_self_fixtures.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden_self")
sys.path.insert(0, ROOT)

from oracle.oracle import OracleLunar, OracleWalker, WalkerHeuristic, lunar_heuristic  # noqa: E402

CASES = [
    # name, family, ctor kwargs, max_episode_steps, policy, N, T, seed
    ("lunarlander_v2_heuristic", "lunar", dict(), 1000, "heuristic", 4, 420, 1),
    ("lunarlander_v2_random", "lunar", dict(), 1000, "random", 6, 260, 7),
    ("lunarlandercontinuous_v2_heuristic", "lunar", dict(continuous=True), 1000, "heuristic", 4, 420, 1),
    ("lunarlander_v2_wind_gravity", "lunar", dict(enable_wind=True, wind_power=12.0, turbulence_power=1.0, gravity=-8.0,
                                                  wind_idx=[5, -40, 999, 1234], torque_idx=[-7, 80, -999, 4321]), 300,
     "heuristic", 4, 320, 11),
    ("bipedalwalker_v3_gait", "walker", dict(), 1600, "gait", 3, 500, 4),
    ("bipedalwalker_v3_random", "walker", dict(), 1600, "random", 4, 220, 5),
    ("bipedalwalkerhardcore_v3_gait", "walker", dict(hardcore=True), 2000, "gait", 3, 420, 0),
]


def rollout(family, kwargs, max_steps, policy, n, t_steps, seed):
    rng = np.random.default_rng(seed + 1000)
    if family == "lunar":
        env = OracleLunar(n, max_episode_steps=max_steps, **kwargs)
        cont = bool(kwargs.get("continuous"))
    else:
        env = OracleWalker(n, max_episode_steps=max_steps, **kwargs)
        gaits = [WalkerHeuristic() for _ in range(n)]
    obs0 = env.reset(seed=seed)
    cur = obs0
    acts, O, R, TE, TR, FO = [], [], [], [], [], []
    a = np.zeros((n, 4), dtype=np.float32)
    for t in range(t_steps):
        if family == "lunar":
            if policy == "random":
                a = rng.uniform(-1.5, 1.5, size=(n, 2)).astype(np.float32) if cont else rng.integers(0, 4, size=n)
            else:
                a = np.stack([lunar_heuristic(s, continuous=cont) for s in cur])
                a = a.astype(np.float32) if cont else a.astype(np.int64)
        elif policy == "random":
            a = rng.uniform(-1.2, 1.2, size=(n, 4)).astype(np.float32)
        o, r, te, tr, fo = env.step(a)
        acts.append(np.array(a, copy=True)); O.append(o); R.append(r); TE.append(te); TR.append(tr); FO.append(fo)
        if family == "walker" and policy == "gait":
            for i in range(n):
                if te[i] or tr[i]:
                    gaits[i] = WalkerHeuristic()
                    a[i] = 0.0
                else:
                    a[i] = gaits[i](o[i])
        cur = o
    return dict(obs0=obs0, actions=np.stack(acts), obs=np.stack(O), reward=np.stack(R), terminated=np.stack(TE),
                truncated=np.stack(TR), final_obs=np.stack(FO))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, family, kwargs, max_steps, policy, n, t_steps, seed in CASES:
        d = rollout(family, kwargs, max_steps, policy, n, t_steps, seed)
        meta = dict(family=family, max_episode_steps=max_steps, n=n, seed=seed, kwargs=json.dumps(kwargs))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d, **{f"meta_{k}": np.asarray(v) for k, v in meta.items()})
        done = d["terminated"] | d["truncated"]
        print(f"{name}: {t_steps} steps x {n} envs, {int(done.sum())} episodes ended, "
              f"returns of first episodes {[round(float(d['reward'][:int(np.argmax(done[:, i])) + 1, i].sum()), 2) if done[:, i].any() else None for i in range(n)]}")


if __name__ == "__main__":
    main()
