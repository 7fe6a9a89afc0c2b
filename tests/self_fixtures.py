"""Loader of tests/golden_self/*.npz: roll-outs frozen from the Box2D-task ORACLE (oracle/gen_self_fixtures.py).
They are regression pins of the re-derived physics, not reference data."""
import glob
import json
import os

import numpy as np

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_self")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(DIR, name + ".npz"))
    d = {k: z[k] for k in ("obs0", "actions", "obs", "reward", "terminated", "truncated", "final_obs")}
    d["family"] = str(z["meta_family"])
    d["max_episode_steps"] = int(z["meta_max_episode_steps"])
    d["n"] = int(z["meta_n"])
    d["seed"] = int(z["meta_seed"])
    d["kwargs"] = json.loads(str(z["meta_kwargs"]))
    return d


def env_id(d):
    """The registered id + remaining constructor kwargs of a fixture."""
    kw = dict(d["kwargs"])
    if d["family"] == "lunar":
        return ("LunarLanderContinuous-v2" if kw.pop("continuous", False) else "LunarLander-v2"), kw
    return ("BipedalWalkerHardcore-v3" if kw.pop("hardcore", False) else "BipedalWalker-v3"), kw


def check(d, step, reset_obs):
    """Replay the frozen actions through `step(a) -> (obs, reward, terminated, truncated, final_obs)`."""
    assert np.array_equal(np.asarray(reset_obs), d["obs0"]), "reset observations"
    for t in range(len(d["actions"])):
        o, r, te, tr, fo = step(d["actions"][t])
        assert np.array_equal(np.asarray(te), d["terminated"][t]) and np.array_equal(np.asarray(tr), d["truncated"][t]), f"flags at step {t}"
        assert np.array_equal(np.asarray(o), d["obs"][t]), f"observations at step {t}"
        assert np.array_equal(np.asarray(r), d["reward"][t]), f"rewards at step {t}"
        done = d["terminated"][t] | d["truncated"][t]
        assert np.array_equal(np.asarray(fo)[done], d["final_obs"][t][done]), f"final observations at step {t}"
