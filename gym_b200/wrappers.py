"""Vector-aware wrappers of the reference, re-built for device-resident batches.

* ``RecordEpisodeStatistics``  gym/wrappers/record_episode_statistics.py:79-151
* ``NormalizeObservation`` / ``NormalizeReward``  gym/wrappers/normalize.py:50-144
* ``VectorListInfo``  gym/wrappers/vector_list_info.py:56-111
* ``step_api_compatibility``  gym/utils/step_api_compatibility.py:24-161

The first three run on the device through the C ABI -- ``RecordEpisodeStatistics`` inside the step
kernel itself (no extra launch), the normalisers as two launches per step with one pass over the batch
for the moments and, for sharded envs, an all-reduce of the 2 x dim batch sums in between -- with no
Python loop over the batch and no host synchronisation; the last two are host-side adapters.  They wrap
a ``B200VectorEnv`` with ``backend="torch"`` (or a ``ShardedVectorEnv``).
"""
import ctypes
import time
from collections import deque

import numpy as np

from gym_b200 import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class VectorWrapper:
    """Minimal ``gym.vector.VectorEnvWrapper`` (gym/vector/vector_env.py:277-332): forwards everything."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, actions):
        return self.env.step(actions)

    def close(self, **kwargs):
        return self.env.close(**kwargs)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class RecordEpisodeStatistics(VectorWrapper):
    """Cumulative reward and length of every episode, kept on the device.

    ``infos["episode"] = {"r": float32 (N,), "l": int32 (N,), "t": float}`` with the ``_episode`` mask
    (rows valid where the mask is set), every step and without a host sync.  ``return_queue`` /
    ``length_queue`` (the last ``deque_size`` finished episodes) and ``episode_count`` synchronise on access.
    """

    def __init__(self, env, deque_size=100):
        super().__init__(env)
        import torch
        self._torch = torch
        n, dev = env.num_envs, env.device
        self.num_envs = n
        self.is_vector_env = True
        self.t0 = time.perf_counter()
        self.deque_size = int(deque_size)
        self.episode_returns = torch.zeros(n, dtype=torch.float32, device=dev)
        self.episode_lengths = torch.zeros(n, dtype=torch.int32, device=dev)
        self._ep_r = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)]
        self._ep_l = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(2)]
        self._ep_m = [torch.zeros(n, dtype=torch.bool, device=dev) for _ in range(2)]
        # one int64 word per slot: length << 32 | float32 bits of the return (written with a single store)
        self._ring = torch.zeros(max(self.deque_size, 1), dtype=torch.int64, device=dev)
        self._counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self._flip = 0

    def reset(self, **kwargs):
        out = self.env.reset(**kwargs)
        self.episode_returns.zero_()
        self.episode_lengths.zero_()
        return out

    def _fused_target(self):
        """The engine handle whose step kernel can do the bookkeeping itself: the wrapped env must be the
        B200VectorEnv proper (torch backend) -- not a sharded env returning gathered global tensors, not another
        wrapper that changes rewards."""
        from gym_b200.vector_env import B200VectorEnv
        e = self.env
        return e if (isinstance(e, B200VectorEnv) and e.backend == "torch" and e.num_envs == self.num_envs) else None

    def step(self, actions):
        self._flip ^= 1
        k = self._flip
        fused = self._fused_target()
        if fused is not None:
            # one call that only stores pointers; the step kernel launched next does the accounting
            _lib.check(fused._lib.b200gym_set_episode_stats(
                fused._handle, _p(self.episode_returns), _p(self.episode_lengths), _p(self._ep_r[k]), _p(self._ep_l[k]),
                _p(self._ring), _p(self._counter), self.deque_size), fused._handle)
            try:
                obs, rew, term, trunc, infos = self.env.step(actions)
            finally:
                fused._lib.b200gym_set_episode_stats(fused._handle, None, None, None, None, None, None, 0)
            mask = term | trunc
        else:
            obs, rew, term, trunc, infos = self.env.step(actions)
            env = self.env.unwrapped
            _lib.check(env._lib.b200gym_episode_stats(
                _p(rew), _p(term), _p(trunc), _p(self.episode_returns), _p(self.episode_lengths), _p(self._ep_r[k]),
                _p(self._ep_l[k]), _p(self._ep_m[k]), _p(self._ring), _p(self._counter),
                self.deque_size, self.num_envs, env._stream()))
            mask = self._ep_m[k]
        infos["episode"] = {"r": self._ep_r[k], "l": self._ep_l[k], "t": round(time.perf_counter() - self.t0, 6)}
        infos["_episode"] = mask
        return obs, rew, term, trunc, infos

    @property
    def episode_count(self):
        return int(self._counter.item())

    def _queue(self, returns):
        return unpack_episode_ring(self._ring.cpu().numpy(), self.episode_count, self.deque_size, returns)

    @property
    def return_queue(self):
        return self._queue(True)

    @property
    def length_queue(self):
        return self._queue(False)


def unpack_episode_ring(words, count, deque_size, returns):
    """The ring written by b200gym_episode_stats -> the reference's deque (record_episode_statistics.py:90-91):
    slot (episode number mod size) holds length << 32 | float32 bits of the return; oldest kept episode first."""
    k = min(int(count), int(deque_size))
    words = np.ascontiguousarray(words).view(np.uint64)
    vals = (words & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32) if returns else \
        (words >> np.uint64(32)).astype(np.int32)
    order = [(count - k + j) % max(int(deque_size), 1) for j in range(k)]
    return deque((vals[i].item() for i in order), maxlen=int(deque_size))


class _RunningMeanStd:
    """Device copy of RunningMeanStd's state (normalize.py:8-29): mean 0, var 1, count epsilon.

    ``stats[2][2*d + 1]`` = {mean, var, count} and ``scratch[2][2*d]`` = this step's batch sums, both
    double-buffered by step parity (include/b200gym.h); ``cur`` is the set holding the current statistics."""

    def __init__(self, torch, d, device, epsilon=1e-4):
        self.d = int(d)
        self.stats = torch.zeros((2, 2 * self.d + 1), dtype=torch.float64, device=device)
        self.stats[:, self.d:2 * self.d] = 1.0
        self.stats[:, 2 * self.d] = epsilon
        self.scratch = torch.zeros((2, 2 * self.d), dtype=torch.float64, device=device)
        self.cur = 0

    @property
    def mean(self):
        return self.stats[self.cur, :self.d]

    @property
    def var(self):
        return self.stats[self.cur, self.d:2 * self.d]

    @property
    def count(self):
        return self.stats[self.cur, 2 * self.d:]


def _shard_group(env, group):
    """(process group, world size) when the batch statistics must be combined across GPUs: the wrapped env is this
    rank's shard of a global batch (``first_index``/``ShardedVectorEnv(gather=None)``) and torch.distributed is up."""
    if group is False:
        return None, 1
    try:
        import torch.distributed as dist
    except Exception:
        return None, 1
    if not (dist.is_available() and dist.is_initialized()):
        return None, 1
    if group is None and not getattr(env, "shards_global_batch", False):
        return None, 1
    g = None if group in (None, True) else group
    return (g if g is not None else dist.group.WORLD), dist.get_world_size(g)


class NormalizeObservation(VectorWrapper):
    """obs -> (obs - running_mean) / sqrt(running_var + epsilon), statistics over the whole batch.

    ``group``: a torch.distributed process group (or True for the default one) when ``env`` is ONE RANK'S SHARD of
    a global batch: the per-column batch sums are all-reduced over the group before the running statistics are
    updated (normalize.py:32-46 applied to the global batch), so every rank normalises with the same numbers."""

    def __init__(self, env, epsilon=1e-8, group=None):
        super().__init__(env)
        import torch
        self._torch = torch
        self.num_envs, self.is_vector_env = env.num_envs, True
        self.epsilon = float(epsilon)
        d = env.single_observation_space.shape[0]
        self.obs_rms = _RunningMeanStd(torch, d, env.unwrapped.device)
        self._out = [None, None]
        self._flip = 0
        self._group, self._world = _shard_group(env, group)

    def _normalize(self, obs):
        torch = self._torch
        self._flip ^= 1
        if self._out[self._flip] is None or self._out[self._flip].shape != obs.shape:
            self._out[self._flip] = torch.empty_like(obs)
        out = self._out[self._flip]
        env = self.env.unwrapped
        r, lib = self.obs_rms, env._lib
        n, d = obs.shape
        st, sc = r.stats[r.cur], r.scratch[r.cur]
        with torch.cuda.device(obs.device):
            _lib.check(lib.b200gym_rms_moments(_p(obs), 0, n, d, _p(st), _p(sc), env._stream()))
            if self._group is not None:
                import torch.distributed as dist
                dist.all_reduce(sc, group=self._group)
            _lib.check(lib.b200gym_rms_apply_obs(_p(obs), _p(out), n, d, _p(st), _p(r.stats[r.cur ^ 1]), _p(sc),
                                                 _p(r.scratch[r.cur ^ 1]), float(n * self._world), self.epsilon, 1,
                                                 env._stream()))
        r.cur ^= 1
        return out

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self._normalize(obs), info

    def step(self, actions):
        obs, rew, term, trunc, infos = self.env.step(actions)
        return self._normalize(obs), rew, term, trunc, infos


class NormalizeReward(VectorWrapper):
    """Scale rewards so that the discounted return has unit running variance (normalize.py:98-144)."""

    def __init__(self, env, gamma=0.99, epsilon=1e-8, group=None):
        super().__init__(env)
        import torch
        self._torch = torch
        self.num_envs, self.is_vector_env = env.num_envs, True
        self.gamma, self.epsilon = float(gamma), float(epsilon)
        dev = env.unwrapped.device
        self.return_rms = _RunningMeanStd(torch, 1, dev)
        self.returns = None
        self._out = [None, None]
        self._flip = 0
        self._group, self._world = _shard_group(env, group)

    def step(self, actions):
        torch = self._torch
        obs, rew, term, trunc, infos = self.env.step(actions)
        if self.returns is None or self.returns.shape != rew.shape:
            self.returns = torch.zeros_like(rew)
        self._flip ^= 1
        if self._out[self._flip] is None or self._out[self._flip].shape != rew.shape:
            self._out[self._flip] = torch.empty_like(rew)
        out = self._out[self._flip]
        env = self.env.unwrapped
        r, lib = self.return_rms, env._lib
        n = rew.shape[0]
        st, sc = r.stats[r.cur], r.scratch[r.cur]
        with torch.cuda.device(rew.device):
            _lib.check(lib.b200gym_return_moments(_p(self.returns), _p(rew), self.gamma, n, _p(st), _p(sc), env._stream()))
            if self._group is not None:
                import torch.distributed as dist
                dist.all_reduce(sc, group=self._group)
            _lib.check(lib.b200gym_rms_apply_reward(_p(rew), _p(out), _p(self.returns), _p(term), _p(trunc), n, _p(st),
                                                    _p(r.stats[r.cur ^ 1]), _p(sc), _p(r.scratch[r.cur ^ 1]),
                                                    float(n * self._world), self.epsilon, env._stream()))
        r.cur ^= 1
        return obs, out, term, trunc, infos


class VectorListInfo(VectorWrapper):
    """Dict-of-arrays ``infos`` -> list of per-env dicts (vector_list_info.py:56-111).

    Host-side adapter for agents that expect the pre-0.25 info format; it synchronises and walks the
    batch in Python, so it is meant for small batches."""

    def _to_host(self, v):
        return v.detach().cpu().numpy() if hasattr(v, "detach") else v

    def _convert(self, infos):
        n = self.env.num_envs
        out = [{} for _ in range(n)]
        for k in list(infos.keys()):
            if k.startswith("_"):
                continue
            mask = np.asarray(self._to_host(infos[f"_{k}"])).astype(bool)
            if k == "episode":  # vector_list_info.py:86-111
                ep = {kk: self._to_host(vv) for kk, vv in infos[k].items()}
                for i in np.flatnonzero(mask):
                    out[i]["episode"] = {kk: (vv[i] if np.ndim(vv) else vv) for kk, vv in ep.items()}
                continue
            vals = self._to_host(infos[k])
            for i in np.flatnonzero(mask):
                out[i][k] = vals[i]
        return out

    def reset(self, **kwargs):
        obs, infos = self.env.reset(**kwargs)
        return obs, self._convert(infos)

    def step(self, actions):
        obs, rew, term, trunc, infos = self.env.step(actions)
        return obs, rew, term, trunc, self._convert(infos)


def step_api_compatibility(step_returns, output_truncation_bool=True, is_vector_env=True):
    """5-tuple <-> 4-tuple step API for vector envs (gym/utils/step_api_compatibility.py:24-161).

    Works on numpy arrays and torch tensors alike (``|``, ``&``, ``~`` only)."""
    if output_truncation_bool:
        if len(step_returns) == 5:
            return step_returns
        obs, rew, dones, infos = step_returns
        trunc_key = infos.get("TimeLimit.truncated") if isinstance(infos, dict) else None
        truncated = trunc_key if trunc_key is not None else dones & ~dones
        return obs, rew, dones & ~truncated, truncated, infos
    if len(step_returns) == 4:
        return step_returns
    obs, rew, terminated, truncated, infos = step_returns
    done = terminated | truncated
    # like the reference (:114-118) the key only appears in steps where some env finished (for device tensors this
    # `any` is the one host synchronisation of the adapter)
    if isinstance(infos, dict) and bool(done.any()):
        infos["TimeLimit.truncated"] = truncated & ~terminated
    return obs, rew, done, infos
