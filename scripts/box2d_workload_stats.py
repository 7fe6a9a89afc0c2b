"""Workload characterisation of the Box2D tasks under random actions (CPU, through the C oracle's per-step counters):
how many contact constraints an env has and how many of the <= 60 position iterations it runs -- the two data-dependent
parts of `world.Step(1/50, 180, 60)` that make lanes of one warp diverge.  Output: profiles/r1_box2d_workload_stats.txt

    python scripts/box2d_workload_stats.py > profiles/r1_box2d_workload_stats.txt
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402


def run(env, getter, act, n, steps, warm):
    env.reset(seed=0)
    rng = np.random.default_rng(0)
    hc, hp = np.zeros(12, int), np.zeros(61, int)
    resets = tot = 0
    for t in range(steps):
        o, r, te, tr, fo = env.step(act(rng, n))
        if t >= warm:
            st = np.zeros((n, 2), np.int32)
            getter(env._h, st.ctypes.data)
            hc += np.bincount(np.minimum(st[:, 0], 11), minlength=12)
            hp += np.bincount(st[:, 1], minlength=61)
            resets += int((te | tr).sum())
            tot += n
    return hc / tot, hp / tot, resets / tot


def main():
    L = orc.lib()
    L.orc_lunar_get_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.orc_walker_get_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    n = 4096
    disc = lambda r, k: r.integers(0, 4, k)                                   # noqa: E731
    box4 = lambda r, k: r.uniform(-1, 1, (k, 4)).astype(np.float32)           # noqa: E731
    cases = [("LunarLander-v2", orc.OracleLunar(n), L.orc_lunar_get_stats, disc),
             ("BipedalWalker-v3", orc.OracleWalker(n), L.orc_walker_get_stats, box4),
             ("BipedalWalkerHardcore-v3", orc.OracleWalker(n, hardcore=True, max_episode_steps=2000), L.orc_walker_get_stats, box4)]
    print(f"# {n} envs, uniformly random actions, steps 100..399 after reset(seed=0) (steady state of random play)")
    for name, env, getter, act in cases:
        hc, hp, rs = run(env, getter, act, n, 400, 100)
        k = np.arange(12)
        it = np.arange(61)
        p_contact, p_60 = 1 - hc[0], hp[60]
        print(f"{name}: autoresets {rs:.4f} per env-step")
        print("  touching contacts per env-step: " + "  ".join(f"{i}: {v:.3f}" for i, v in enumerate(hc) if v > 0.0005)
              + f"   (mean {float((hc * k).sum()):.2f}, P[any] = {p_contact:.3f})")
        print(f"  position iterations run (max 60): mean {float((hp * it).sum()):.1f};  1: {hp[1]:.3f}  2-5: {hp[2:6].sum():.3f}  "
              f"6-59: {hp[6:60].sum():.3f}  60: {p_60:.3f}")
        print(f"  => a 32-lane warp contains an env with contacts with probability {1 - (1 - p_contact) ** 32:.2f}, "
              f"one that runs all 60 position iterations with probability {1 - (1 - p_60) ** 32:.2f}")


def gait_case(L, n=128, steps=600, warm=50):
    """BipedalWalker driven by the reference's demo gait instead of random torques."""
    env = orc.OracleWalker(n)
    obs = env.reset(seed=3)
    gaits = [orc.WalkerHeuristic() for _ in range(n)]
    a = np.zeros((n, 4), np.float32)
    hc, hp, tot = np.zeros(12, int), np.zeros(61, int), 0
    for t in range(steps):
        o, r, te, tr, fo = env.step(a)
        if t >= warm:
            st = np.zeros((n, 2), np.int32)
            L.orc_walker_get_stats(env._h, st.ctypes.data)
            hc += np.bincount(np.minimum(st[:, 0], 11), minlength=12)
            hp += np.bincount(st[:, 1], minlength=61)
            tot += n
        for i in range(n):
            if te[i] or tr[i]:
                gaits[i] = orc.WalkerHeuristic()
                a[i] = 0
            else:
                a[i] = gaits[i](o[i])
    hc, hp = hc / tot, hp / tot
    print(f"BipedalWalker-v3 under the reference's demo gait ({n} envs): contacts mean {float((hc * np.arange(12)).sum()):.2f}; "
          f"position iterations mean {float((hp * np.arange(61)).sum()):.1f};  1: {hp[1]:.3f}  2-5: {hp[2:6].sum():.3f}  "
          f"6-59: {hp[6:60].sum():.3f}  60: {hp[60]:.3f}   (the 60-iteration regime belongs to random torques: fallen walkers "
          "with joints pressed against their limits)")


if __name__ == "__main__":
    main()
    gait_case(orc.lib())
