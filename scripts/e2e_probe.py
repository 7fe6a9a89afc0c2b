import time, numpy as np, torch, gym_b200, sys
n=1<<20
for ar in (True, False):
    env=gym_b200.vector.make("CartPole-v1", n, backend="numpy", copy=False, dense_infos=True, autoreset=ar)
    env.reset(seed=0)
    acts=[torch.randint(0,2,(n,),dtype=torch.int64).pin_memory().numpy() for _ in range(4)]
    for k in range(5): env.step(acts[k%4])
    torch.cuda.synchronize(); t0=time.perf_counter()
    for k in range(40): env.step(acts[k%4])
    torch.cuda.synchronize(); el=time.perf_counter()-t0
    print('autoreset',ar,'ms/step',1e3*el/40)
    # breakdown: time only the C call
    import ctypes
    from gym_b200 import _lib
    h=env._hio
    t0=time.perf_counter()
    for k in range(40):
        _lib.check(env._lib.b200gym_step_host(env._handle, acts[k%4].ctypes.data, 0, None,None,None,None, h["final_obs"].ctypes.data if ar else None))
    el=time.perf_counter()-t0
    print('   C call only ms/step',1e3*el/40)
    env.close()
