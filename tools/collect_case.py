#!/usr/bin/env python
"""BASELINE config 5 as a short command (for ncu launch lists): 131 072 envs x 16 steps collected into a
TrajectoryBuffer by the fused masked-uniform collector, a few times; prints env-steps/s (CUDA events)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import adapters, gobblet_v1  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--final", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
vec = gobblet_v1.vec_env(a.envs, device=dev, seed=1)
buf = adapters.TrajectoryBuffer(a.steps, a.envs, device=dev, keep_final=a.final)
col = adapters.VecCollector(vec, adapters.RandomLegalPolicy(seed=3), buf)
for _ in range(3):
    col.collect()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    col.collect()
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) * 1e-3 / a.iters
print(f"collect: {a.envs} envs x {a.steps} steps, {dt * 1e6:.1f} us per collection, {a.envs * a.steps / dt:.4g} env-steps/s, "
      f"{a.envs * (a.steps * (176 + (171 if a.final else 0)) + 172) / dt / 1e9:.0f} GB/s")
