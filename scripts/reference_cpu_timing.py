"""Time the REAL reference (openai/gym 0.26.2, pure Python) on this machine's CPU cores: SURVEY.md 8(d) "CPU path
timing".  Runs only where /root/reference exists (the build container; the GPU box has no copy of the reference, there
bench.py times the C oracle instead).  Output is committed as profiles/r1_reference_python_timing.txt.

    python scripts/reference_cpu_timing.py > profiles/r1_reference_python_timing.txt
"""
import os
import sys
import time
import warnings

import numpy as np

REFERENCE = "/root/reference"


def main():
    for n, v in (("bool8", np.bool_), ("float_", np.float64)):
        if not hasattr(np, n):
            setattr(np, n, v)
    sys.path.insert(0, REFERENCE)
    warnings.filterwarnings("ignore")
    import gym

    def run(make_vec, n, min_seconds=1.0, min_steps=20):
        envs = make_vec(n)
        envs.reset(seed=0)
        envs.action_space.seed(0)
        acts = [envs.action_space.sample() for _ in range(8)]
        for k in range(3):
            envs.step(acts[k])
        steps, t0 = 0, time.perf_counter()
        while True:
            envs.step(acts[steps % 8])
            steps += 1
            el = time.perf_counter() - t0
            if steps >= min_steps and el >= min_seconds:
                break
        envs.close()
        return n * steps / el, steps, el

    print(f"# openai/gym {gym.__version__} (reference tree), numpy {np.__version__}, python {sys.version.split()[0]}, "
          f"{os.cpu_count()} logical cores: {open('/proc/cpuinfo').read().split('model name')[1].split(':')[1].splitlines()[0].strip()}")
    print("# env-steps/s = num_envs * vector steps / wall seconds, random actions from action_space.sample(), autoreset on")
    print(f"{'vector env':28s} {'env id':26s} {'num_envs':>8s} {'env-steps/s':>12s} {'vector steps':>12s} {'seconds':>8s}")
    for env_id in ("CartPole-v1", "Pendulum-v1", "Acrobot-v1", "MountainCar-v0", "MountainCarContinuous-v0"):
        for n in ((4, 1024, 16384) if env_id == "CartPole-v1" else (1024,)):
            v, s, el = run(lambda k: gym.vector.SyncVectorEnv([lambda: gym.make(env_id, disable_env_checker=True)] * k), n)
            print(f"{'SyncVectorEnv (1 core)':28s} {env_id:26s} {n:8d} {v:12.4g} {s:12d} {el:8.2f}")
    workers = os.cpu_count() or 1
    for n in (workers, 8 * workers):
        v, s, el = run(lambda k: gym.vector.AsyncVectorEnv([lambda: gym.make("CartPole-v1", disable_env_checker=True)] * k,
                                                           shared_memory=True), n)
        print(f"{f'AsyncVectorEnv ({n} procs)':28s} {'CartPole-v1':26s} {n:8d} {v:12.4g} {s:12d} {el:8.2f}")
    # BASELINE.json configs[0]: 4 envs, 1000 steps
    envs = gym.vector.SyncVectorEnv([lambda: gym.make("CartPole-v1", disable_env_checker=True)] * 4)
    envs.reset(seed=0)
    envs.action_space.seed(0)
    t0 = time.perf_counter()
    for _ in range(1000):
        envs.step(envs.action_space.sample())
    print(f"# BASELINE.json configs[0] (SyncVectorEnv, 4 x CartPole-v1, 1000 steps incl. action sampling): "
          f"{(time.perf_counter() - t0) * 1e3:.1f} ms")


if __name__ == "__main__":
    if not os.path.isdir(REFERENCE):
        sys.exit("reference tree not present")
    main()
