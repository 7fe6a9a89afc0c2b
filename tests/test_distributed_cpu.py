"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_b200.distributed import GatherBuffers, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_local, obs_dim, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard_range(world * n_local, world, rank)
        assert (first, count) == (rank * n_local, n_local)
        bufs = [GatherBuffers(world, rank, n_local, obs_dim, "cpu") for _ in range(2)]
        # emulate what the step kernel does: write this rank's results into its slice; every value
        # is a function of the GLOBAL env index, like the seed fan-out `seed + first_index + i`
        for step in range(3):
            b = bufs[step & 1]
            gi = torch.arange(first, first + count)
            b.local["obs"].copy_((gi[:, None] * 10 + torch.arange(obs_dim)[None, :] + 1000 * step).float())
            b.local["reward"].copy_(gi.double() * 0.5 + step)
            b.local["terminated"].copy_((gi + step) % 3 == 0)
            b.local["truncated"].copy_((gi + step) % 5 == 0)
            g = b.all_gather()
            G = torch.arange(world * n_local)
            assert torch.equal(g["obs"], (G[:, None] * 10 + torch.arange(obs_dim)[None, :] + 1000 * step).float())
            assert torch.equal(g["reward"], G.double() * 0.5 + step)
            assert torch.equal(g["terminated"], (G + step) % 3 == 0)
            assert torch.equal(g["truncated"], (G + step) % 5 == 0)
            assert g["terminated"].dtype == torch.bool and g["obs"].shape == (world * n_local, obs_dim)
            # the local views alias the global tensors (the gather is in place)
            assert b.local["obs"].data_ptr() == g["obs"][first:].data_ptr()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_buffers_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 96, 4, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert results == {0: "ok", 1: "ok"}, results


def test_shard_range_partitions_the_batch():
    total = 1 << 23
    seen = 0
    for r in range(8):
        first, count = shard_range(total, 8, r)
        assert first == seen and count == 1 << 20
        seen += count
    assert seen == total
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)


def test_single_process_gather_is_identity():
    b = GatherBuffers(1, 0, 8, 3, "cpu")
    b.local["obs"].fill_(2.0)
    g = b.all_gather()
    assert torch.equal(g["obs"], torch.full((8, 3), 2.0)) and g["obs"].data_ptr() == b.local["obs"].data_ptr()
