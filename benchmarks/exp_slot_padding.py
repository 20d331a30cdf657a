#!/usr/bin/env python
"""Experiment: does padding the trajectory slot stride change DRAM channel balance?  (ncu shows per-channel DRAM
utilisation between 60 % and 89 % for the fused rollout.)  Result on B200: the unpadded stride (2^20 envs x 117 B)
is the BEST case, 1.058 of the copy peak; any padding lands at 1.00 -- so no padding is applied."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1, ops
n, T = 1 << 20, 64
vec = gobblet_v1.vec_env(n, seed=0)
res = {}
for pad_envs in (0, 16, 272, 1040, 4112, 16400, 65552, 262160):
    pad = n + pad_envs
    obs = torch.zeros((T, pad, 3, 3, 13), dtype=torch.int8, device="cuda")[:, :n]
    mask = torch.zeros((T, pad, 54), dtype=torch.int8, device="cuda")[:, :n]
    def run():
        ops.rollout_random(vec.state, T, 0, 0, 0, obs, mask, None, None, None, None, None, None, vec.stats, vec.flags)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): run()
    b.record(); torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / 100
    res[pad_envs] = n * T / dt
    print(pad_envs, "%.4e" % res[pad_envs], round(res[pad_envs] * 171 / 1e9 / 6552.6, 4), flush=True)
    del obs, mask
    torch.cuda.empty_cache()
