#!/usr/bin/env python
"""Experiment: where is the gap between the rollout kernel (6.9 TB/s) and a pure fill (7.5 TB/s)?  Runs the
fused rollout with both output streams, observation only, mask only and no stores at all
(GBL_MEASURE_SKIP_*_STORES).  Result on B200: observation-only also tops out at 6.9 TB/s (so it is not the two
streams interfering), mask-only is compute-bound (1.04 ms), no stores = 0.86 ms = the kernel's compute floor
(7.8e10 env-steps/s)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1, ops
n, T = 1 << 20, 64
vec = gobblet_v1.vec_env(n, seed=0)
obs, mask = vec._ring_buffers(T)
for name, extra, nbytes in (("both", 0, 171), ("obs only", 2 << 8, 117), ("mask only", 1 << 8, 54), ("none", 3 << 8, 0)):
    def run():
        ops.rollout_random(vec.state, T, 0, 0, 0, obs, mask, None, None, None, None, None, None, vec.stats, vec.flags | extra)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): run()
    b.record(); torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / 100
    print(name, "%.3f ms" % (dt * 1e3), "%.3e steps/s" % (n * T / dt), "%.0f GB/s" % (n * T * nbytes / dt / 1e9), flush=True)
