#!/usr/bin/env python
"""The packed-record step (e2e path) as a short command (for ncu): 2^20 envs, a few launches; prints us per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

n = 1 << 20
vec = gobblet_v1.vec_env(n, device="cuda:0", seed=0)
vec.rollout_random(6, emit=False)
acts = torch.randint(0, 54, (n,), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    vec.step_packed(acts)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    vec.step_packed(acts)
b.record()
torch.cuda.synchronize()
print(f"step_packed: {a.elapsed_time(b) * 1e3 / 20:.2f} us per launch of {n} envs")
