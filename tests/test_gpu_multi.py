"""Multi-GPU path on real GPUs (needs >= 2 GPUs: `gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_multi.py`).

Rank r owns the global envs [r*n, (r+1)*n); both exchange flavours -- the fused NVLink
peer-store kernel ("p2p") and the NCCL all-gather ("nccl") -- must deliver, on every rank, exactly
the tensors a single GPU computes for the whole batch.
"""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, env_id, n_local, steps, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gym_b200
        from gym_b200.distributed import ShardedVectorEnv
        total = world * n_local
        rng = np.random.default_rng(0)
        full = gym_b200.vector.make(env_id, total, max_episode_steps=25)      # single-GPU reference
        if full.discrete:
            acts = torch.as_tensor(rng.integers(0, full.single_action_space.n, size=(steps, total)), device="cuda")
        else:
            acts = torch.as_tensor(rng.uniform(-2, 2, size=(steps, total, 1)).astype(np.float32), device="cuda")
        envs = {g: ShardedVectorEnv(env_id, total, gather=g, max_episode_steps=25) for g in ("p2p", "nccl")}
        ref_obs, _ = full.reset(seed=77)
        for g, e in envs.items():
            o, _ = e.reset(seed=77)
            assert torch.equal(o, ref_obs), f"{g}: reset obs differ from the single-GPU batch"
        lo, hi = rank * n_local, (rank + 1) * n_local
        n_done = 0
        for t in range(steps):
            ro, rr, rte, rtr, rinfo = full.step(acts[t])
            for g, e in envs.items():
                o, r, te, tr, info = e.step(acts[t, lo:hi])
                torch.cuda.synchronize()
                assert torch.equal(o, ro), f"{g} step {t}: obs"
                assert torch.equal(r, rr) and torch.equal(te, rte) and torch.equal(tr, rtr), f"{g} step {t}"
                m = info["_final_observation"]
                assert torch.equal(m, rinfo["_final_observation"][lo:hi])
                assert torch.equal(info["final_observation"][m], rinfo["final_observation"][lo:hi][m])
            n_done += int((rte | rtr).sum())
        assert n_done > 0
        for e in envs.values():
            e.close()
        full.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("env_id,n_local", [("CartPole-v1", 148 * 256 + 640), ("Pendulum-v1", 5000)])
@pytest.mark.timeout(300)
def test_sharded_step_matches_single_gpu(env_id, n_local):
    import torch.multiprocessing as mp
    world = min(_ngpus(), 8)   # every GPU of the box (the driver's scaling run uses 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, env_id, n_local, 40, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=280) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert all(v == "ok" for v in results.values()), results
