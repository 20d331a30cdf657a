#!/usr/bin/env python
"""Secondary measurements of the hot path's other kernels (not the driver's bench contract):
step_kernel / observe_kernel against the HBM roof, the greedy search kernel (BASELINE config 4), the
GPU collection loop with an MLP policy (config 5), and the fused rollout with emission switched off
(compute ceiling).  Prints one JSON object; run on a B200:  python benchmarks/bench_paths.py
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from gobblet_rl_b200 import adapters, gobblet_v1  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def main():
    dev = torch.device("cuda", 0)
    peak = 6552.6
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    out = {"peak_hbm_gbs": peak}
    n = 1 << 20

    # -- step_kernel, device resident: replay logged legal actions (uint8) -------------------------------
    K = 48
    logger = gobblet_v1.vec_env(n, device=dev, seed=2)
    log = logger.rollout_random(K + 3, emit=False, log_actions=True)["actions"]
    vec = gobblet_v1.vec_env(n, device=dev, seed=2)
    # two alternating output slots > L2 so the stores go to DRAM
    slots = [(torch.zeros((n, 3, 3, 13), dtype=torch.int8, device=dev), torch.zeros((n, 54), dtype=torch.int8, device=dev)) for _ in range(2)]
    k = [0]

    def one_step():
        vec.step(log[k[0]], out=slots[k[0] & 1])
        k[0] += 1

    for _ in range(3):
        one_step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        one_step()
    b.record()
    torch.cuda.synchronize()
    dt = a.elapsed_time(b) * 1e-3 / K
    assert torch.equal(vec.state, logger.state)
    bytes_step = 171 + 32 + 1 + 5
    out["step_kernel"] = {"envs": n, "env_steps_per_s": n / dt, "us_per_launch": dt * 1e6,
                          "algorithmic_bytes_per_env_step": bytes_step, "achieved_gbs": n * bytes_step / dt / 1e9,
                          "frac_of_peak": n * bytes_step / dt / 1e9 / peak}

    # -- observe_kernel ------------------------------------------------------------------------------------
    dt = timed(lambda: vec.observe(), 50)
    out["observe_kernel"] = {"envs": n, "env_obs_per_s": n / dt, "achieved_gbs": n * (171 + 16 + 1) / dt / 1e9,
                             "frac_of_peak": n * 188 / dt / 1e9 / peak}

    # -- fused rollout, emission off (compute ceiling) and with a 4-slot ring (L2-absorbed stores) ---------
    v2 = gobblet_v1.vec_env(n, device=dev, seed=0)
    dt = timed(lambda: v2.rollout_random(64, emit=False), 20)
    out["rollout_no_emit"] = {"env_steps_per_s": n * 64 / dt}
    dt = timed(lambda: v2.rollout_random(64, ring=4), 20)
    out["rollout_ring4_l2_absorbed"] = {"env_steps_per_s": n * 64 / dt, "note": "not an HBM number: L2 absorbs the rewrites"}
    del v2

    # -- greedy search, BASELINE config 4: 65536 non-terminal boards at plies 2..12 ------------------------
    src = gobblet_v1.vec_env(1 << 18, device=dev, seed=7, autoreset="off")
    boards_obs, boards_mask = [], []
    for plies in range(2, 13, 2):
        src.rollout_random(2, emit=False)
        obs, mask, _ = src.observe()
        live = ~src.terminated if plies > 2 else torch.ones_like(src.terminated)
        sq, _ = src.squares()
        st = src.state[:, 0] >> 55 & 1                                   # done bit of the packed state
        live = st == 0
        boards_obs.append(obs[live][:11000].clone()); boards_mask.append(mask[live][:11000].clone())
    gobs, gmask = torch.cat(boards_obs)[:65536].contiguous(), torch.cat(boards_mask)[:65536].contiguous()
    nb = gobs.shape[0]
    for depth in (1, 2):
        dt = timed(lambda: gobblet_v1.greedy_actions(gobs, gmask, None, depth=depth), 20)
        out[f"greedy_depth{depth}"] = {"boards": nb, "boards_per_s": nb / dt, "us_per_launch": dt * 1e6}

    # -- config 5: collection with an MLP policy (example_tianshou_DQN.py:48,161-166: 117 -> 128x4 -> 54) --
    n5, T5 = 131072, 16
    net = torch.nn.Sequential(torch.nn.Linear(117, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                              torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                              torch.nn.Linear(128, 54)).to(dev).half()
    eps_sampler = adapters.RandomLegalPolicy(seed=3)

    @torch.no_grad()
    def policy(obs, mask, agent):
        q = net(obs.reshape(obs.shape[0], 117).half())
        q = q.masked_fill(mask == 0, float("-inf"))
        greedy = q.argmax(1).to(torch.int32)
        rnd = eps_sampler(obs, mask)
        explore = torch.rand(obs.shape[0], device=obs.device) < 0.1
        return torch.where(explore, rnd, greedy)

    v5 = gobblet_v1.vec_env(n5, device=dev, seed=1)
    buf = adapters.TrajectoryBuffer(T5, n5, device=dev)
    col = adapters.VecCollector(v5, policy, buf)

    def collect():
        col.collect()
        col.roll()

    dt = timed(collect, 10)
    out["collect_mlp_policy_c5"] = {"envs": n5, "steps": T5, "env_steps_per_s": n5 * T5 / dt, "buffer_bytes": buf.nbytes(),
                                    "illegal_moves": int(v5.stats[5]), "note": "includes fp16 MLP inference + masked eps-greedy in torch"}
    dt = timed(lambda: adapters.VecCollector(v5, adapters.RandomLegalPolicy(seed=3), buf).collect(), 10)
    out["collect_random_policy"] = {"envs": n5, "steps": T5, "env_steps_per_s": n5 * T5 / dt}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
