#!/usr/bin/env python
"""BASELINE config 2 as a short command (for ncu): 4096 envs, fused rollout of 512 steps, a few launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

n, T = 4096, 512
vec = gobblet_v1.vec_env(n, device="cuda:0", seed=0)
for _ in range(3):
    vec.rollout_random(T, ring=4)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    vec.rollout_random(T, ring=4)
b.record()
torch.cuda.synchronize()
print(f"{a.elapsed_time(b) * 1e3 / 5 / T:.3f} us per lockstep step")
