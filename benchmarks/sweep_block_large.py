#!/usr/bin/env python
"""Fused rollout with a full trajectory (ring = T = 64, every byte reaches HBM once) for 64 Ki ... 1 Mi envs and every
block size: env-steps/s and the fraction of the measured copy peak.  Complements sweep_small_batch.py (ring = 4:
the L2-resident, latency-bound regime)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    out = {}
    T = 64
    for n in (65536, 131072, 262144, 524288, 1048576):
        vec = gobblet_v1.vec_env(n, device=dev, seed=0)
        res = {}
        for hint in (0, 32, 64, 128, 256):
            for _ in range(2):
                vec.rollout_random(T, ring=T, block_hint=hint)
            torch.cuda.synchronize()
            reps = max(3, (1 << 22) // n)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                vec.rollout_random(T, ring=T, block_hint=hint)
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b) * 1e-3 / reps
            res["auto" if hint == 0 else str(hint)] = {"env_steps_per_s": n * T / dt, "frac_of_copy_peak": n * T * 171 / dt / 6552.6e9}
        out[str(n)] = res
        del vec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
