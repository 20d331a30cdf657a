#!/usr/bin/env python
"""BASELINE config 2 regime (a few thousand lockstep envs: latency-bound, the 0.7 MB of output per step stays in
L2): microseconds per lockstep step of the fused rollout for every block size (32 / 64 / 128 / 256 threads) and,
for the small blocks, with the observation image leaving through the copy engine (TMA bulk store) or through a
lane copy (LDS.128 -> STG.128), and with two warps per 32 envs (one emitting the observations, one the masks).
Prints one JSON object; the best variant per size is what
`rollout_block_for` / the default flags in gobblet_engine.cu wire in."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gobblet_rl_b200 import gobblet_v1  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for n in (1024, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
        T = 4096 if n <= 16384 else 512 if n <= 65536 else 128
        vec = gobblet_v1.vec_env(n, device=dev, seed=0)
        res = {}
        for hint in (0, 32, 64, 128, 256):
            for no_bulk in ((False, True) if hint in (0, 32, 64) else (False,)):
                for _ in range(2):
                    vec.rollout_random(T, ring=4, block_hint=hint, no_bulk=no_bulk)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    vec.rollout_random(T, ring=4, block_hint=hint, no_bulk=no_bulk)
                b.record()
                torch.cuda.synchronize()
                res[f"{'auto' if hint == 0 else hint}{'_lanecopy' if no_bulk else '_bulk'}"] = a.elapsed_time(b) * 1e3 / 3 / T
        for no_bulk in (False, True):                    # two warps per 32 envs: observation warp / mask warp
            if n > 32768:
                break
            for _ in range(2):
                vec.rollout_random(T, ring=4, block_hint=32, no_bulk=no_bulk, split=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                vec.rollout_random(T, ring=4, block_hint=32, no_bulk=no_bulk, split=True)
            b.record()
            torch.cuda.synchronize()
            res[f"32{'_lanecopy' if no_bulk else '_bulk'}_split"] = a.elapsed_time(b) * 1e3 / 3 / T
        out[str(n)] = {"us_per_lockstep_step": res, "hbm_time_us": n * 171 / 6552.6e9 * 1e6,
                       "best": min(res, key=res.get), "env_steps_per_s_best": n / (min(res.values()) * 1e-6)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
