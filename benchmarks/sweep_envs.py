#!/usr/bin/env python
"""Fused-rollout throughput vs number of lockstep envs (trajectory output, every byte once per launch):
shows where the kernel leaves the latency-bound regime and reaches the HBM-write roof."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

peak = 6552.6
p = os.path.join(REPO, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
rows = []
for logn in (10, 12, 14, 16, 18, 20, 22, 24):
    n = 1 << logn
    T = max(8, min(4096, (1 << 33) // (n * 171)))           # ~8 GB of trajectory per launch at most
    T = min(T, 1024)
    vec = gobblet_v1.vec_env(n, seed=0)
    for _ in range(3):
        vec.rollout_random(T, ring=T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        vec.rollout_random(T, ring=T)
    b.record()
    torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / reps
    rate = n * T / dt
    rows.append({"envs": n, "fused_steps": T, "trajectory_gb": n * T * 171 / 1e9, "env_steps_per_s": rate,
                 "emitted_gbs": rate * 171 / 1e9, "frac_of_hbm_peak": rate * 171 / 1e9 / peak})
    del vec
    torch.cuda.empty_cache()
print(json.dumps(rows, indent=1))
