"""The DEVICE source of the Box2D tasks, compiled for the CPU, against the C oracle.  Runs without a GPU.

tests/hostsim/hostsim.cpp includes gym_b200/csrc/{rng,b2lite,lunar,walker}.cuh and box2d_consts.h through a small
CUDA shim and mirrors the kernel bodies with plain loops.  What this pins on every CPU run: the register-resident
solver (compile-time island topology), the scenes, the record packing (load_world / store_world), the host-computed
shape / mass / joint constants and the numpy-compatible RNG of the device code are bit-identical to the independent
C implementation in oracle/ -- the same statement tests/test_gpu_lunar.py / test_gpu_walker.py make on the GPU for
the nvcc build of this source.  (Not covered here: the kernels' own indexing / launch code and nvcc's code
generation -- that is what the `-m gpu` tests are for.)
"""
import numpy as np
import pytest

from hostsim.sim import HostSim
from oracle import oracle as orc


def _compare(t, got, want):
    for k, name in enumerate(("obs", "reward", "terminated", "truncated")):
        if not np.array_equal(got[k], want[k]):
            bad = np.argwhere(np.asarray(got[k]) != np.asarray(want[k]))
            raise AssertionError(f"step {t}: {name} differs at {bad[0]}: {np.asarray(got[k])[tuple(bad[0])]!r} "
                                 f"vs {np.asarray(want[k])[tuple(bad[0])]!r} ({len(bad)} values)")
    done = want[2] | want[3]
    assert np.array_equal(got[4][done], want[4][done]), f"step {t}: final observations"
    return done


LUNAR = [
    ("LunarLander", dict(), "random"),
    ("LunarLander", dict(), "heuristic"),
    ("LunarLanderContinuous", dict(), "random"),
    ("LunarLanderContinuous", dict(gravity=-6.5), "heuristic"),
    ("LunarLander", dict(enable_wind=True, wind_power=19.0, turbulence_power=1.9, gravity=-11.5), "random"),
    ("LunarLanderContinuous", dict(enable_wind=True), "heuristic"),
]


@pytest.mark.parametrize("name,kwargs,policy", LUNAR)
def test_lunar_device_source_on_cpu_equals_oracle(name, kwargs, policy):
    N, T, seed = 96, 400, 5
    cont = name.endswith("Continuous")
    rng = np.random.default_rng(2)
    extra = {}
    if kwargs.get("enable_wind"):
        extra = dict(wind_idx=rng.integers(-9999, 9999, size=N), torque_idx=rng.integers(-9999, 9999, size=N))
    sim = HostSim(name, N, 1000, **kwargs, **extra)
    ref = orc.OracleLunar(N, max_episode_steps=1000, continuous=cont, **kwargs, **extra)
    cur = ref.reset(seed=seed)
    assert np.array_equal(sim.reset(seed=seed), cur)
    n_done = n_sleep = 0
    for t in range(T):
        if policy == "random":
            a = rng.uniform(-1.6, 1.6, size=(N, 2)).astype(np.float32) if cont else rng.integers(0, 4, size=N)
        else:
            a = np.stack([orc.lunar_heuristic(s, continuous=cont) for s in cur])
        want = ref.step(a)
        done = _compare(t, sim.step(a), want)
        n_done += int(done.sum())
        n_sleep += int((want[1][want[2]] == 100).sum())
        cur = want[0]
    assert n_done > 0
    if policy == "heuristic" and not kwargs.get("enable_wind"):
        assert n_sleep > N // 3
    if kwargs.get("enable_wind"):
        assert all(np.array_equal(x, y) for x, y in zip(sim.wind_idx(), ref.wind_idx()))


@pytest.mark.parametrize("name,hardcore,policy", [("BipedalWalker", False, "gait"), ("BipedalWalkerHardcore", True, "gait"),
                                                  ("BipedalWalker", False, "random"), ("BipedalWalkerHardcore", True, "random")])
def test_walker_device_source_on_cpu_equals_oracle(name, hardcore, policy):
    N, T, seed = 24, 420, 31
    steps = 2000 if hardcore else 1600
    sim = HostSim(name, N, steps)
    ref = orc.OracleWalker(N, hardcore=hardcore, max_episode_steps=steps)
    cur = ref.reset(seed=seed)
    assert np.array_equal(sim.reset(seed=seed), cur)
    for i in range(N):
        t, boxes = sim.terrain(i)
        assert np.array_equal(t, ref.terrain(i)) and np.array_equal(boxes, ref.polys(i))
    rng = np.random.default_rng(8)
    gaits = [orc.WalkerHeuristic() for _ in range(N)]
    a = np.zeros((N, 4), dtype=np.float32)
    n_done = 0
    for t in range(T):
        if policy == "random":
            a = rng.uniform(-1.3, 1.3, size=(N, 4)).astype(np.float32)
        want = ref.step(a)
        done = _compare(t, sim.step(a), want)
        n_done += int(done.sum())
        if policy == "gait":
            for i in range(N):
                if done[i]:
                    gaits[i] = orc.WalkerHeuristic()
                    a[i] = 0.0
                else:
                    a[i] = gaits[i](want[0][i])
    assert n_done > (N // 2 if policy == "random" or hardcore else 0)
