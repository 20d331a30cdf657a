#!/usr/bin/env python
"""Run ON THE B200: produce tests/golden/gpu_rollout_log.npz, a small action log WITH the outputs the CUDA
engine produced, so that CPU-only test runs can replay GPU-made trajectories through the oracle and the
reference itself (tests/test_gpu_log_replay.py).  Regenerate with:  gpurun -- python tools/make_gpu_log_fixture.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from gobblet_rl_b200 import gobblet_v1, trajectory_io  # noqa: E402

out_dir = os.path.join(REPO, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
vec = gobblet_v1.vec_env(96, seed=2026, env_id_base=(1 << 40) + 17)
T = 48
res = vec.rollout_random(T, ring=T, per_step=True, log_actions=True)
trajectory_io.save_action_log(os.path.join(out_dir, "gpu_rollout_log.npz"), vec, res, step_base=0)
print("wrote", os.path.join(out_dir, "gpu_rollout_log.npz"), vec.stats_dict())
