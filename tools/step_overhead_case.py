#!/usr/bin/env python
"""Host-side cost of one externally driven step at a few thousand envs (where the Python side, not the kernel, is the
step): microseconds per `VecEnv.step(actions)` call from a plain loop, for the pre-marshalled fast path and for the
general path (forced by passing `out=`), next to the kernel's own duration in a CUDA graph."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

for n in (4096, 65536):
    vec = gobblet_v1.vec_env(n, device="cuda:0", seed=0)
    acts = torch.zeros(n, dtype=torch.int64, device="cuda:0")
    res = {}
    for name, call in (("fast", lambda: vec.step(acts)), ("general", lambda: vec.step(acts, out=(vec.obs, vec.mask)))):
        for _ in range(200):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            call()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        res[name] = (t_issue / 2000 * 1e6, (time.perf_counter() - t0) / 2000 * 1e6)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        vec.step(acts)
        with torch.cuda.graph(g, stream=s):
            for _ in range(64):
                vec.step(acts)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"{n} envs: step() from Python  fast {res['fast'][0]:.1f} us issue / {res['fast'][1]:.1f} us total,  "
          f"general {res['general'][0]:.1f} / {res['general'][1]:.1f} us;  in a CUDA graph {a.elapsed_time(b) * 1e3 / 640:.2f} us per step")
