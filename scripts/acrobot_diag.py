"""Diagnostic (GPU): first step at which the Acrobot float64 state on the device differs from the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, gym_b200
from oracle import oracle as orc
env_id = sys.argv[1] if len(sys.argv) > 1 else "Acrobot-v1"
N, T = 4096, 450
env = gym_b200.vector.make(env_id, N); o = orc.OracleVec(env_id, N)
env.reset(seed=2024); o.reset(seed=2024)
rng = np.random.default_rng(1)
prev = o.get_state()[0].copy()
found = 0
for t in range(T):
    a = rng.integers(0, 3, size=N) if env.discrete else rng.uniform(-2, 2, size=(N, 1)).astype(np.float32)
    env.step(torch.as_tensor(a, device="cuda")); o.step(a)
    gs = env.get_state()[0].cpu().numpy(); os_ = o.get_state()[0]
    bad = np.argwhere(gs.view(np.int64) != os_.view(np.int64))
    if bad.size:
        i = bad[0][0]
        print(f"step {t}: {len(set(bad[:,0]))} envs differ; env {i} action {a[i]}")
        print(" prev  ", [float.hex(v) for v in prev[i]])
        print(" gpu   ", [float.hex(v) for v in gs[i]])
        print(" oracle", [float.hex(v) for v in os_[i]])
        found += 1
        if found >= 3: break
        env.set_state(state=os_, elapsed=o.get_state()[1], rng=o.get_rng())
    prev = os_.copy()
print("done, mismatching steps seen:", found)
