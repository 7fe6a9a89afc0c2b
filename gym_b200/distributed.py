"""Sharding a global batch of environments over the GPUs of one node.

The reference has no distributed code: its only "gather" is AsyncVectorEnv
collecting per-worker results over pipes / shared memory
(gym/vector/async_vector_env.py:319-328, gym/vector/utils/shared_memory.py:164-170).
The B200-native equivalent: one process per GPU, rank r owns the
contiguous global envs [r*n, (r+1)*n); env i is seeded ``seed + i`` with its
GLOBAL index, so results do not depend on the number of GPUs; every step the
ranks exchange (obs, reward, terminated, truncated) with ONE in-place
all-gather per tensor over NVLink (``torch.distributed``, NCCL backend).

The step kernel writes straight into this rank's slice of the gather buffers
(the "pack" is fused into the kernel's stores), so the all-gather is in place
and there is no extra copy pass.
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous block partition of `total` envs: (first_index, count) of `rank`."""
    if total % world_size != 0:
        raise ValueError(f"num_envs={total} must be divisible by world_size={world_size}")
    n = total // world_size
    return rank * n, n


class GatherBuffers:
    """Global (world*n, ...) result tensors plus this rank's writable slices.

    Works on any device/backend (NCCL on GPUs, gloo on CPU in the tests).
    """

    FIELDS = (("obs", torch.float32), ("reward", torch.float64),
              ("terminated", torch.bool), ("truncated", torch.bool))

    def __init__(self, world_size, rank, n_local, obs_dim, device):
        self.world_size, self.rank, self.n_local = world_size, rank, n_local
        total = world_size * n_local
        self.glob = {
            "obs": torch.zeros((total, obs_dim), dtype=torch.float32, device=device),
            "reward": torch.zeros((total,), dtype=torch.float64, device=device),
            "terminated": torch.zeros((total,), dtype=torch.bool, device=device),
            "truncated": torch.zeros((total,), dtype=torch.bool, device=device),
        }
        lo, hi = rank * n_local, (rank + 1) * n_local
        self.local = {k: v[lo:hi] for k, v in self.glob.items()}

    def all_gather(self, group=None):
        """In-place all-gather of the four result tensors (input = this rank's slice)."""
        if self.world_size == 1:
            return self.glob
        for name, _ in self.FIELDS:
            out, inp = self.glob[name], self.local[name]
            if out.dtype == torch.bool:  # collectives want a numeric dtype; same bytes
                out, inp = out.view(torch.uint8), inp.view(torch.uint8)
            dist.all_gather_into_tensor(out, inp, group=group)
        return self.glob


class ShardedVectorEnv:
    """This rank's shard of a global ``B200VectorEnv`` batch + per-step all-gather.

    ``step(actions_local)`` takes the actions of this rank's envs (a replicated
    policy only needs its own slice) and returns the GLOBAL
    (obs, rewards, terminateds, truncateds) tensors, identical on every rank.
    """

    def __init__(self, env_id, total_envs, group=None, gather=True, **kwargs):
        from gym_b200.vector_env import B200VectorEnv
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.first_index, self.n_local = shard_range(total_envs, self.world_size, self.rank)
        self.total_envs = total_envs
        self.gather = gather
        self.env = B200VectorEnv(env_id, self.n_local, first_index=self.first_index, **kwargs)
        # two sets of gather buffers, alternating like the env's own double buffering
        self.buffers = [GatherBuffers(self.world_size, self.rank, self.n_local, self.env.obs_dim, self.env.device)
                        for _ in range(2)]
        for k in range(2):
            self.env._out[k].update(self.buffers[k].local)  # kernel writes into the gather slices
        self.num_envs = total_envs
        self.single_observation_space = self.env.single_observation_space
        self.single_action_space = self.env.single_action_space

    def _finish(self):
        buf = self.buffers[self.env._flip]
        return buf.all_gather(self.group) if self.gather else buf.glob

    def reset(self, *, seed=None, options=None):
        self.env.reset(seed=seed, options=options)
        return self._finish()["obs"], {}

    def step(self, actions_local):
        _, _, _, _, infos = self.env.step(actions_local)
        g = self._finish()
        return g["obs"], g["reward"], g["terminated"], g["truncated"], infos

    def close(self):
        self.env.close()
