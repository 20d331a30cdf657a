#!/usr/bin/env python
"""Small all-kernel workload for compute-sanitizer (racecheck / memcheck), e.g.
   compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402

for n in (1000, 37):
    v = gobblet_v1.vec_env(n, seed=3)
    v.rollout_random(9, ring=3, per_step=True, log_actions=True)
    obs, mask, agent = v.observe()
    acts = mask.to(torch.float32).argmax(1)
    fobs, fmask = torch.zeros_like(obs), torch.zeros_like(mask)
    v.step(acts, final=(fobs, fmask))
    gobblet_v1.greedy_actions(v.obs, v.mask, depth=2)
    sq, ag = v.squares()
    v.set_squares(sq, ag)
    v.step_packed(acts)
    act = torch.zeros(n, dtype=torch.int32, device="cuda")
    from gobblet_rl_b200 import ops
    ops.sample_legal(v.mask, 1, 0, 0, act)
    v.rollout_random(6, ring=1, final=True, per_step=True)          # slots rewritten inside the launch (bulk-store ordering)
    v.rollout_random(6, ring=2, block_hint=32)
torch.cuda.synchronize()
print("sanitizer case done")
