#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused step kernel (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): CartPole-v1, num_envs = 2^20 per GPU, random
actions, fused step + TimeLimit + autoreset; a "step" is one vector step of the
whole batch (one kernel launch per GPU).  For N > 1 (launched by torchrun, one
rank per GPU) every rank owns 2^20 envs of a global batch and every step ends
with the NCCL all-gather of (obs, reward, terminated, truncated) -- weak scaling.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      algorithmic bytes per launch / average kernel time (CUDA events
                on the launch stream) against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the CPU oracle (C port of the reference path) on the host cores
  e2e           the same metric through the host-buffer C-ABI path (pinned host
                actions in, results to host memory, copies inside the timed region)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ENV_ID = "CartPole-v1"
LOG2_ENVS = 20
# Algorithmic HBM bytes per env-step of the CartPole kernel:
#   read : state 4 x f64 (32) + TimeLimit counter (4) + action int64 (8)            = 44
#   write: state (32) + counter (4) + obs 4 x f32 (16) + reward f64 (8) + 2 flags (2) = 62
# (SURVEY.md 8(d) counts 102 B with a float32 reward; we emit the reference's float64.)
BYTES_PER_ENV_STEP = {"CartPole-v1": 106, "CartPole-v0": 106, "Pendulum-v1": 16 * 2 + 8 + 4 + 12 + 8 + 2,
                      "Acrobot-v1": 32 * 2 + 8 + 8 + 24 + 8 + 2, "MountainCar-v0": 16 * 2 + 8 + 8 + 8 + 8 + 2,
                      "MountainCarContinuous-v0": 16 * 2 + 8 + 4 + 8 + 8 + 2,
                      # LunarLander: 103-word solver record r+w, PCG64 state (2 draws/step), counter, action, outputs
                      "LunarLander-v2": 101 * 4 * 2 + 32 + 16 + 8 + 8 + 32 + 8 + 2,
                      # BipedalWalker: 130-word record r+w (the 200 terrain words are read on demand), action, outputs
                      "BipedalWalker-v3": 130 * 4 * 2 + 8 + 16 + 96 + 8 + 2}
BYTES_PER_ENV_STEP["LunarLanderContinuous-v2"] = BYTES_PER_ENV_STEP["LunarLander-v2"]
BYTES_PER_ENV_STEP["BipedalWalkerHardcore-v3"] = BYTES_PER_ENV_STEP["BipedalWalker-v3"] + 4
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (2^20 envs)
NCU_DRAM_BYTES_PER_LAUNCH = {"CartPole-v1": 56.20e6 + 25.84e6}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=1000)
    p.add_argument("--warmup", type=int, default=100)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--env", default=ENV_ID)
    p.add_argument("--log2-envs", type=int, default=LOG2_ENVS, help="envs per GPU = 2**k")
    p.add_argument("--e2e-steps", type=int, default=200)
    p.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                   help="N>1 exchange: fused NVLink peer stores in the step kernel, or NCCL all-gather")
    return p.parse_args()


# --------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def random_actions_np(env_id, n, rng):
    if env_id.startswith("LunarLanderContinuous"):
        return rng.uniform(-1.0, 1.0, size=(n, 2)).astype(np.float32)
    if env_id.startswith("LunarLander"):
        return rng.integers(0, 4, size=n).astype(np.int64)
    if env_id.startswith("BipedalWalker"):
        return rng.uniform(-1.0, 1.0, size=(n, 4)).astype(np.float32)
    from oracle.oracle import ENV_IDS, KINDS, lib
    kind = KINDS[ENV_IDS[env_id][0]]
    nact = lib().orc_num_actions(kind)
    if nact:
        return rng.integers(0, nact, size=n).astype(np.int64)
    return rng.uniform(-2.0, 2.0, size=(n, 1)).astype(np.float32)


def make_oracle(env_id, n):
    """The C port of the reference path for `env_id` (oracle/): one object stepping n envs on host threads."""
    from oracle.oracle import OracleLunar, OracleVec, OracleWalker
    if env_id.startswith("LunarLander"):
        return OracleLunar(n, continuous="Continuous" in env_id)
    if env_id.startswith("BipedalWalker"):
        hard = "Hardcore" in env_id
        return OracleWalker(n, hardcore=hard, max_episode_steps=2000 if hard else 1600)
    return OracleVec(env_id, n)


def cpu_oracle_throughput(env_id, n, seconds, threads, min_steps=3):
    """Time the C port of the reference path (oracle/) on the host cores: bounded sample."""
    rng = np.random.default_rng(0)
    v = make_oracle(env_id, n)
    v.reset(seed=0)
    pool = [random_actions_np(env_id, n, rng) for _ in range(4)]
    kw = {"nthreads": threads}
    v.step(pool[0], **kw)  # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        v.step(pool[steps % 4], **kw)
        steps += 1
        el = time.perf_counter() - t0
        if steps >= min_steps and el >= seconds:
            break
        if steps >= 2000:
            break
    v.close()
    return n * steps / el, steps, el


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    openai/gym is pure Python and /root/reference does not exist on the GPU box, so this
    arm times the oracle port (oracle/gym_oracle.c: the reference's step/reset/TimeLimit/
    autoreset restated in C, bit-exact against the reference) with all host threads.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1 << args.log2_envs
    threads = host_threads()
    rng = np.random.default_rng(0)
    v = make_oracle(args.env, n)
    v.reset(seed=0)
    pool = [random_actions_np(args.env, n, rng) for _ in range(8)]
    for w in range(max(args.warmup, 1)):
        v.step(pool[w % 8], nthreads=threads)
        if w >= 10:
            break
    steps = min(args.steps, 2000)
    t0 = time.perf_counter()
    for k in range(steps):
        v.step(pool[k % 8], nthreads=threads)
    el = time.perf_counter() - t0
    v.close()
    value = n * steps / el
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.env.startswith(("LunarLander", "BipedalWalker")) else "f64", "data": "synthetic",
        "config": {"workload": f"{args.env} num_envs=2^{args.log2_envs}, random actions, "
                               "step+TimeLimit+autoreset on the host cores (C port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} vector steps of 2^{args.log2_envs} envs"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import gym_b200
    from gym_b200.distributed import ShardedVectorEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = 1 << args.log2_envs
    K, W = args.steps, max(args.warmup, 3)

    gather = args.gather
    if world > 1:
        try:
            env = ShardedVectorEnv(args.env, n * world, gather=gather)
        except Exception as exc:  # e.g. CUDA IPC unavailable: fall back to the NCCL exchange
            if rank == 0:
                print(f"[bench] gather={gather} unavailable ({exc}); using nccl", file=sys.stderr)
            gather = "nccl"
            env = ShardedVectorEnv(args.env, n * world, gather=gather)
        inner = env.env
    else:
        env = gym_b200.vector.make(args.env, n)
        inner = env
    env.reset(seed=0)

    # pre-generated device-resident random actions (pool of 16, cycled)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    if inner.discrete:
        pool = torch.randint(0, inner.single_action_space.n, (16, n), device=dev, dtype=torch.int64, generator=gen)
    else:
        scale = 2.0 if inner.act_dim == 1 else 1.0
        pool = (torch.rand((16, n, inner.act_dim), device=dev, generator=gen) * 2.0 - 1.0) * scale
    # L2 flush between timed steps: write 256 MiB (> 126 MB L2), then stream-read another 256 MiB so
    # that the lines left in L2 are CLEAN (otherwise the timed kernel pays the write-back of the
    # flush's own dirty lines, which is an artefact of the flush, not of the kernel)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if world == 1 else None
    flush_rd = torch.zeros(32 << 20, dtype=torch.int64, device=dev) if world == 1 else None
    flush_sink = torch.zeros(1, dtype=torch.int64, device=dev) if world == 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(W):
        env.step(pool[w % 16])
    barrier()

    sampler = ClockSampler(torch.cuda.current_device() if world == 1 else local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    if world == 1:
        # one kernel per step; L2 flushed (256 MB write) before every timed step
        for k in range(K):
            flush.fill_(k & 0xFF)
            torch.sum(flush_rd, dim=0, keepdim=True, out=flush_sink)
            starts[k].record()
            env.step(pool[k % 16])
            ends[k].record()
        barrier()
        per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        total_ms = float(sum(per_step_ms))
        kernel_ms = total_ms / K
        # the same K steps back to back with a warm L2 (reported, not the headline)
        barrier()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for k in range(K):
            env.step(pool[k % 16])
        e0.record()
        barrier()
        warm_ms = s0.elapsed_time(e0) / K
        l2_note = "flushed before every timed step (256 MiB write, then 256 MiB read so L2 holds clean lines); per-step CUDA events summed"
    else:
        # the gathered outputs (world x 26 MB) exceed L2; K steps back to back incl. the all-gather
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for k in range(K):
            env.step(pool[k % 16])
        e0.record()
        barrier()
        t = torch.tensor([s0.elapsed_time(e0)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        # kernel-only time of this rank's shard (no collective), L2 is overwritten by the gather anyway
        inner_s = [torch.cuda.Event(enable_timing=True) for _ in range(32)]
        inner_e = [torch.cuda.Event(enable_timing=True) for _ in range(32)]
        for k in range(32):
            inner_s[k].record()
            inner.step(pool[k % 16])
            inner_e[k].record()
        barrier()
        kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in zip(inner_s, inner_e)]))
        warm_ms = None
        l2_note = f"per-step working set (all-gather output, {world} x 26 MiB) exceeds L2; steps back to back"
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = total_ms / K
    value = world * n * K / (total_ms * 1e-3)

    # ---- e2e: pinned host actions in, results to host, through the public numpy-backend API
    e2e = None
    if not args.no_e2e:
        host_env = gym_b200.vector.make(args.env, n, backend="numpy", copy=False, dense_infos=True,
                                        first_index=rank * n)
        host_env.reset(seed=0)
        hpool = [pool[k].cpu().pin_memory() for k in range(4)]
        hnp = [t.numpy() for t in hpool]
        for k in range(5):
            host_env.step(hnp[k % 4])
        E = max(10, min(args.e2e_steps, K))
        barrier()
        t0 = time.perf_counter()
        for k in range(E):
            obs, rew, term, trunc, infos = host_env.step(hnp[k % 4])
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = float(tt.item())
        D = inner.obs_dim
        h2d = n * (8 if inner.discrete else 4 * inner.act_dim)
        # obs + reward + 2 flags, plus the compacted final observations (index + row) of the envs that
        # finished in the last step
        n_done = int((term | trunc).sum())
        d2h = n * (4 * D + 8 + 1 + 1) + n_done * (4 + 4 * D) + 8
        e2e = {"value": world * n * E / el, "unit": "env-steps/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "steps": E, "ms_per_step": 1e3 * el / E,
               "api": "B200VectorEnv(backend='numpy').step(pinned numpy actions) -> numpy results "
                      "(b200gym_step_host)"}
        host_env.close()

    # ---- cpu baseline (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        v, cpu_steps, cpu_el = cpu_oracle_throughput(args.env, n, args.cpu_seconds, threads)
        cpu = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
               "sample": f"{cpu_steps} vector steps of 2^{args.log2_envs} envs ({cpu_el:.1f} s) of the same workload, "
                         + ("oracle/lunar_oracle.c / walker_oracle.c" if args.env.startswith(("LunarLander", "BipedalWalker"))
                            else "oracle/gym_oracle.c") + " with one pthread per host core"}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bytes_per_launch = BYTES_PER_ENV_STEP.get(args.env, 0) * n
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.env.startswith(("LunarLander", "BipedalWalker")) else "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.env} num_envs=2^{args.log2_envs} per GPU x {world} GPU(s), random "
                                   f"{'int64' if inner.discrete else 'float32'} "
                                   "actions resident in HBM, fused step+TimeLimit+autoreset"
                                   + ((", all-gather of (obs,reward,terminated,truncated) per step fused into the "
                                       "step kernel (NVLink peer stores + flag exchange)" if gather == "p2p" else
                                       ", NCCL all-gather of (obs,reward,terminated,truncated) per step")
                                      if world > 1 else ""),
                       "state": "float64 (reference-faithful)", "l2": l2_note,
                       "parallelism": f"env-batch data parallel x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.env) if args.log2_envs == 20 else None,
                         "traffic_source": "profiles/r1_cartpole_step_kernel_A_ncu_full.txt: dram__bytes_read.sum + "
                                           "dram__bytes_write.sum of one launch (cold L2; lines still dirty in L2 at "
                                           "kernel end are not in it)", "peak_source": peak_src,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP.get(args.env), "kernel_ms": kernel_ms,
                         "kernel": ("step_kernel_tma" if os.environ.get("B200GYM_KERNEL", "a") == "b" else "step_kernel")
                         + ("<CARTPOLE, int64>" if args.env.startswith("CartPole") else "")},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": K,
            "clocks": clocks,
        }
        if warm_ms is not None:
            line["warm_l2"] = {"ms_per_step": warm_ms, "value": n * 1e3 / warm_ms,
                               "note": "same K steps back to back, state resident in L2 (not the headline)"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
