import sys; sys.path.insert(0,'/root/repo')
import torch, time
from gobblet_rl_b200 import gobblet_v1, ops
dev=torch.device('cuda',0); n=1<<20
logger = gobblet_v1.vec_env(n, device=dev, seed=1)
log = logger.rollout_random(20, emit=False, log_actions=True)["actions"]
h_log = torch.zeros(log.shape, dtype=torch.uint8, pin_memory=True); h_log.copy_(log); torch.cuda.synchronize()
for kw in (dict(), dict(blocking_events=True), dict(host_threads=15), dict(host_threads=15, blocking_events=True), dict(), dict(blocking_events=True), dict(chunks=16), dict(chunks=1)):
    host = gobblet_v1.HostVecEnv(n, device=dev, seed=1, **kw)
    host.reset()
    ts=[]
    for k in range(20):
        t0=time.perf_counter(); host.step(h_log[k]); dt=time.perf_counter()-t0
        if k>=4: ts.append((dt, ops.host_last_timing()))
    ts.sort(key=lambda x:x[0]); dt,m=ts[len(ts)//2]
    print(kw, f"call {dt*1e3:.3f} ms | enqueue {m[0]*1e3:.3f} | published", [round(x*1e3,3) for x in m[1:-1]], f"| finish {m[-1]*1e3:.3f}")
