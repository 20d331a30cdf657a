#!/usr/bin/env python
"""Experiment: HostVecEnv (the `e2e` path) throughput vs number of stream-pipelined chunks."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

n, K = 1 << 20, 16
logger = gobblet_v1.vec_env(n, seed=1)
log = logger.rollout_random(K + 3, emit=False, log_actions=True)["actions"]
h_log = torch.zeros(log.shape, dtype=torch.uint8, pin_memory=True)
h_log.copy_(log)
for chunks in (1, 2, 4, 8, 16, 32):
    host = gobblet_v1.HostVecEnv(n, chunks=chunks, seed=1)
    host.reset()
    for k in range(3):
        host.step(h_log[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(3, 3 + K):
        host.step(h_log[k])
    dt = time.perf_counter() - t0
    print(chunks, "%.4e env-steps/s" % (n * K / dt), "%.1f GB/s D2H" % (n * K * 176 / dt / 1e9), flush=True)
    del host
