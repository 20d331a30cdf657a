#!/usr/bin/env python
"""Bit-exactness at BASELINE's full size (2^20 envs) without shipping gigabytes of observations:

  phase 1 (B200):   python tools/full_size_parity.py gpu  -> gpurun_out/full_size_log.npz
      fused rollout of N envs x T steps with action logging; per step, position-weighted 64-bit checksums of the
      emitted obs / mask tensors (every byte contributes with its own weight) and the rew / terminated / agent sums
  phase 2 (any CPU): python tools/full_size_parity.py cpu gpurun_out/full_size_log.npz
      replays the logged actions of ALL envs through the CPU oracle (oracle/gobblet_oracle.c, one process per
      core) and recomputes the same checksums.  TEST TOOLING (imports oracle/ in phase 2 only).
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
MULT = 2654435761
MOD = (1 << 31) - 1


def weights(lo, hi):
    """weight of flat byte index i in [lo, hi): (i * MULT) mod (2^31 - 1) + 1, exact in int64"""
    i = np.arange(lo, hi, dtype=np.int64)
    return (i % MOD) * (MULT % MOD) % MOD + 1


def gpu_phase(n=1 << 20, T=16, seed=7):
    import torch
    from gobblet_rl_b200 import gobblet_v1
    dev = torch.device("cuda")
    vec = gobblet_v1.vec_env(n, device=dev, seed=seed)
    out = vec.rollout_random(T, ring=T, per_step=True, log_actions=True)
    w_obs = torch.from_numpy(weights(0, n * 117)).to(dev)
    w_mask = torch.from_numpy(weights(0, n * 54)).to(dev)
    cs = np.zeros((T, 5), np.int64)
    for t in range(T):
        cs[t, 0] = int((out["obs"][t].reshape(-1).long() * w_obs).sum())
        cs[t, 1] = int((out["mask"][t].reshape(-1).long() * w_mask).sum())
        cs[t, 2] = int((out["rew"][t].reshape(-1).long() * w_mask[: 2 * n]).sum())
        cs[t, 3] = int((out["terminated"][t].long() * w_mask[:n]).sum())
        cs[t, 4] = int((out["agent_id"][t].long() * w_mask[:n]).sum())
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    path = os.path.join(REPO, "gpurun_out", "full_size_log.npz")
    np.savez_compressed(path, actions=out["actions"].cpu().numpy(), checksums=cs, stats=vec.stats.cpu().numpy(),
                        seed=seed, num_envs=n)
    print("wrote", path, os.path.getsize(path), "bytes;", vec.stats_dict())


def _replay_block(args):
    lo, hi, actions = args
    from oracle import oracle as O
    n, T = hi - lo, actions.shape[0]
    v = O.VecOracle(n, "terminate", "same_step")
    part = np.zeros((T, 5), np.int64)
    w_obs, w_mask = weights(lo * 117, hi * 117), weights(lo * 54, hi * 54)
    w_rew, w_env = weights(lo * 2, hi * 2), weights(lo, hi)
    for t in range(T):
        obs, mask, rew, term, trunc, agent = v.step(actions[t].astype(np.int64))
        part[t] = [(obs.reshape(-1).astype(np.int64) * w_obs).sum(), (mask.reshape(-1).astype(np.int64) * w_mask).sum(),
                   (rew.reshape(-1).astype(np.int64) * w_rew).sum(), (term.astype(np.int64) * w_env).sum(),
                   (agent.astype(np.int64) * w_env).sum()]
    return part, v.stats


def cpu_phase(path):
    import multiprocessing as mp
    d = np.load(path)
    actions, want, n = d["actions"], d["checksums"], int(d["num_envs"])
    T = actions.shape[0]
    cores = len(os.sched_getaffinity(0))
    blocks = 4 * cores
    bounds = [n * i // blocks for i in range(blocks + 1)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_replay_block, [(a, b, actions[:, a:b]) for a, b in zip(bounds[:-1], bounds[1:])])
    got = sum(r[0] for r in res)
    stats = sum(r[1][:7] for r in res)
    ok = np.array_equal(got, want) and stats.tolist() == d["stats"][:7].tolist()
    print(f"{n} envs x {T} steps = {n * T} env-steps replayed through the oracle on {cores} cores in {time.time() - t0:.0f}s")
    print("checksums (obs, mask, rew, terminated, agent_id) per step:", "ALL EQUAL" if np.array_equal(got, want) else "MISMATCH")
    print("episode statistics:", "EQUAL" if stats.tolist() == d["stats"][:7].tolist() else "MISMATCH", d["stats"].tolist())
    return 0 if ok else 1


if __name__ == "__main__":
    if sys.argv[1] == "gpu":
        gpu_phase(*(int(x) for x in sys.argv[2:]))
    else:
        sys.exit(cpu_phase(sys.argv[2]))
