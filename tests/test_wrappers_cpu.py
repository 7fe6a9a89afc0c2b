"""Host-side info adapters (no GPU): gym_b200.wrappers.VectorListInfo / step_api_compatibility."""
import numpy as np

from gym_b200 import wrappers


class _FakeVec:
    num_envs = 3
    is_vector_env = True

    def __init__(self):
        self.t = 0

    @property
    def unwrapped(self):
        return self

    def reset(self, **kwargs):
        return np.zeros((3, 2), np.float32), {}

    def step(self, actions):
        self.t += 1
        done = np.array([False, True, False])
        fo = np.zeros((3, 2), np.float32)
        fo[1] = [1.5, -2.5]
        infos = {"final_observation": fo, "_final_observation": done,
                 "episode": {"r": np.array([0, 7.0, 0], np.float32), "l": np.array([0, 9, 0], np.int32), "t": 0.25},
                 "_episode": done}
        return np.ones((3, 2), np.float32), np.ones(3), done, np.zeros(3, bool), infos


def test_vector_list_info_matches_reference_format():
    # gym/wrappers/vector_list_info.py:56-111 and tests/wrappers/test_vector_list_info.py
    env = wrappers.VectorListInfo(_FakeVec())
    obs, infos = env.reset()
    assert infos == [{}, {}, {}]
    obs, rew, term, trunc, infos = env.step(None)
    assert isinstance(infos, list) and len(infos) == 3
    assert infos[0] == {} and infos[2] == {}
    assert np.array_equal(infos[1]["final_observation"], [1.5, -2.5])
    assert infos[1]["episode"]["r"] == 7.0 and infos[1]["episode"]["l"] == 9 and infos[1]["episode"]["t"] == 0.25


def test_step_api_compatibility_roundtrip():
    # gym/utils/step_api_compatibility.py:24-161 (vector env, dict infos)
    term = np.array([True, False, False])
    trunc = np.array([False, True, False])
    five = (np.zeros((3, 1)), np.ones(3), term, trunc, {})
    obs, rew, dones, infos = wrappers.step_api_compatibility(five, output_truncation_bool=False)
    assert dones.tolist() == [True, True, False]
    assert infos["TimeLimit.truncated"].tolist() == [False, True, False]
    back = wrappers.step_api_compatibility((obs, rew, dones, infos), output_truncation_bool=True)
    assert back[2].tolist() == term.tolist() and back[3].tolist() == trunc.tolist()
    assert wrappers.step_api_compatibility(five) is five


def test_episode_ring_unpacks_to_the_reference_deques():
    """return_queue / length_queue (record_episode_statistics.py:90-91,123-144): the packed device ring, oldest kept
    episode first, before and after it wraps."""
    size = 5
    ring = np.zeros(size, dtype=np.int64)
    episodes = [(1.5, 3), (-2.25, 7), (100.0, 1), (0.0, 9), (42.0, 4), (7.0, 8), (-1.0, 2)]
    for count, (ret, length) in enumerate(episodes, start=1):
        word = (np.int64(length) << np.int64(32)) | np.int64(np.float32(ret).view(np.uint32))
        ring[(count - 1) % size] = word
        kept = episodes[max(0, count - size):count]
        assert list(wrappers.unpack_episode_ring(ring, count, size, True)) == [r for r, _ in kept]
        assert list(wrappers.unpack_episode_ring(ring, count, size, False)) == [l for _, l in kept]
    assert wrappers.unpack_episode_ring(ring, 0, size, True).maxlen == size
    assert list(wrappers.unpack_episode_ring(ring, 0, size, True)) == []
