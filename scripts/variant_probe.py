"""Timing + parity probe of the Box2D-task variants (run on the GPU box: PYTHONPATH=. python scripts/variant_probe.py)."""
import time

import numpy as np
import torch

import gym_b200
from oracle.oracle import OracleLunar, OracleWalker, WalkerHeuristic, lunar_heuristic

T0 = time.time()


def lap(msg, t):
    torch.cuda.synchronize()
    print(f"[{time.time() - T0:7.2f}s] {msg}: {time.time() - t:.3f}s", flush=True)
    return time.time()


def lunar(env_id, n, steps, cont, **kw):
    t = time.time()
    env = gym_b200.vector.make(env_id, n, **kw)
    orc = OracleLunar(n, continuous=cont, **{k: v for k, v in kw.items() if k != "continuous"})
    t = lap(f"{env_id} {kw.keys()} make", t)
    o, _ = env.reset(seed=3)
    r = orc.reset(seed=3)
    ok = np.array_equal(o.cpu().numpy(), r)
    t = lap(f"  reset ok={ok}", t)
    rng = np.random.default_rng(0)
    tg = to = 0.0
    for s in range(steps):
        a = rng.uniform(-1.5, 1.5, (n, 2)).astype(np.float32) if cont else rng.integers(0, 4, n)
        t1 = time.time()
        got = env.step(torch.as_tensor(a, device=env.device))
        torch.cuda.synchronize()
        t2 = time.time()
        want = orc.step(a)
        t3 = time.time()
        tg += t2 - t1
        to += t3 - t2
        ok = ok and np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[1].cpu().numpy(), want[1])
    t = lap(f"  {steps} steps ok={ok} gpu {tg:.3f}s oracle {to:.3f}s", t)
    env.close()


lunar("LunarLander-v2", 768, 60, False)
lunar("LunarLanderContinuous-v2", 768, 60, True)
lunar("LunarLander-v2", 768, 60, False, enable_wind=True, wind_idx=5, torque_idx=9)

t = time.time()
env = gym_b200.make("LunarLanderContinuous-v2")
s, _ = env.reset(seed=1)
t = lap("single-env facade make+reset", t)
tot = 0.0
for k in range(200):
    s, r, te, tr, _ = env.step(lunar_heuristic(s, continuous=True).astype(np.float32))
    tot += r
    if te or tr:
        break
t = lap(f"single-env facade {k + 1} steps, return {tot:.1f}", t)
env.close()

for env_id, hc in (("BipedalWalker-v3", False), ("BipedalWalkerHardcore-v3", True)):
    t = time.time()
    n = 128
    env = gym_b200.vector.make(env_id, n)
    orc = OracleWalker(n, hardcore=hc, max_episode_steps=env.max_episode_steps)
    o, _ = env.reset(seed=40)
    r = orc.reset(seed=40)
    ok = np.array_equal(o.cpu().numpy(), r)
    terrain, polys, npoly = env.walker_terrain()
    okp = all(np.array_equal(polys[i, :int(npoly[i])].cpu().numpy(), orc.polys(i)) for i in range(n))
    t = lap(f"{env_id} make+reset ok={ok} polys ok={okp} npoly {int(npoly.min())}..{int(npoly.max())}", t)
    gaits = [WalkerHeuristic() for _ in range(n)]
    a = np.zeros((n, 4), dtype=np.float32)
    tg = to = th = 0.0
    first_bad = None
    ndone = 0
    for s in range(450):
        t1 = time.time()
        got = env.step(torch.as_tensor(a, device=env.device))
        torch.cuda.synchronize()
        t2 = time.time()
        want = orc.step(a)
        t3 = time.time()
        same = (np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[1].cpu().numpy(), want[1])
                and np.array_equal(got[2].cpu().numpy(), want[2]))
        if not same and first_bad is None:
            first_bad = s
            bad = np.argwhere(got[0].cpu().numpy() != want[0])
            print("   first mismatch at step", s, "envs", sorted(set(bad[:, 0].tolist()))[:8], "cols", sorted(set(bad[:, 1].tolist())))
        done = want[2] | want[3]
        ndone += int(done.sum())
        for i in range(n):
            if done[i]:
                gaits[i] = WalkerHeuristic()
                a[i] = 0
            else:
                a[i] = gaits[i](want[0][i])
        t4 = time.time()
        tg += t2 - t1
        to += t3 - t2
        th += t4 - t3
    bodies, _ = env.walker_bodies()
    t = lap(f"  450 gait steps first_bad={first_bad} done={ndone} far={float(bodies[:, 0, 0].max()):.1f} "
            f"gpu {tg:.3f}s oracle {to:.3f}s heuristic {th:.3f}s", t)
    env.close()
print("probe done")
