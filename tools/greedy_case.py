#!/usr/bin/env python
"""BASELINE config 4 as a short command (for ncu): 65 536 non-terminal boards at plies 2..12, greedy depth 2.
   python tools/greedy_case.py [--iters K]     prints boards/s (CUDA events)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402


def boards(dev, n=65536):
    src = gobblet_v1.vec_env(1 << 18, device=dev, seed=7, autoreset="off")
    bo, bm = [], []
    for plies in range(2, 13, 2):
        src.rollout_random(2, emit=False)
        obs, mask, _ = src.observe()
        live = (src.state[:, 0] >> 55 & 1) == 0
        bo.append(obs[live][:11000].clone())
        bm.append(mask[live][:11000].clone())
    return torch.cat(bo)[:n].contiguous(), torch.cat(bm)[:n].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    obs, mask = boards(dev)
    for depth in (2, 1):
        gobblet_v1.greedy_actions(obs, mask, None, depth=depth)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            gobblet_v1.greedy_actions(obs, mask, None, depth=depth)
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3 / a.iters
        print(f"greedy depth {depth}: {obs.shape[0]} boards, {dt * 1e6:.1f} us/launch, {obs.shape[0] / dt:.4g} boards/s")


if __name__ == "__main__":
    main()
