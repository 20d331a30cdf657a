#!/usr/bin/env python
"""Experiment: the end-to-end host path (pinned actions in, reference-shaped host arrays out) with the two
wire formats -- "dense" (176 B/env over PCIe, round 1) vs "packed" (24 B/env + host-side expansion) -- against
their ceilings measured in the same run: pinned D2H bandwidth (plain cudaMemcpyAsync of the same buffers) and
the host thread pool's fill bandwidth (what the expander's stores are bound by).
    python benchmarks/exp_e2e_wire.py [--envs N] [--steps K]
Prints one JSON object."""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from gobblet_rl_b200 import gobblet_v1, ops  # noqa: E402


def d2h_probe(dev, nbytes, reps=5):
    """GB/s of one pinned device->host cudaMemcpyAsync of nbytes."""
    d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    h = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True)
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(dev)
    return reps * nbytes / (time.perf_counter() - t0) / 1e9


def fill_probe(nbytes, threads, mode, reps=5):
    h = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True)
    ops.host_fill(h, threads, mode)
    t0 = time.perf_counter()
    for _ in range(reps):
        ops.host_fill(h, threads, mode)
    return reps * nbytes / (time.perf_counter() - t0) / 1e9


def run(host, h_log, warm, steps, dev):
    host.reset()
    for k in range(warm):
        host.step(h_log[k])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(warm, warm + steps):
        host.step(h_log[k])
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, K, W = a.envs, a.steps, 3
    cores = len(os.sched_getaffinity(0))
    out = {"envs": n, "steps": K, "host_cores": cores, "host_simd": ops.host_simd(), "pool_threads": ops.host_threads(0)}
    out["probe_d2h_gbs"] = {"dense_176B_per_env": d2h_probe(dev, n * 176), "packed_24B_per_env": d2h_probe(dev, n * 24)}
    out["probe_host_fill_gbs"] = {f"threads{t}_{'nt' if m == 0 else 'regular'}": fill_probe(n * 176, t, m)
                                  for t in sorted({1, 4, 8, cores}) if t <= cores for m in (0, 1)}
    logger = gobblet_v1.vec_env(n, device=dev, seed=1)
    log = logger.rollout_random(K + W, emit=False, log_actions=True)["actions"]
    h_log = torch.zeros(log.shape, dtype=torch.uint8, pin_memory=True)
    h_log.copy_(log)
    res = {}
    variants = [("dense_chunks2", dict(wire="dense", chunks=2)), ("packed_geometric_schedule", dict(wire="packed")),
                ("packed_chunks1", dict(wire="packed", chunks=1)), ("packed_chunks4", dict(wire="packed", chunks=4)),
                ("packed_chunks8", dict(wire="packed", chunks=8)), ("packed_chunks16", dict(wire="packed", chunks=16)),
                ("packed_chunks32", dict(wire="packed", chunks=32)),
                ("packed_chunks8_threads4", dict(wire="packed", chunks=8, host_threads=4)),
                ("packed_chunks8_threads8", dict(wire="packed", chunks=8, host_threads=8)),
                ("packed_chunks8_threads12", dict(wire="packed", chunks=8, host_threads=12)),
                ("packed_chunks8_noexpand", dict(wire="packed", chunks=8, expand=False))]
    for rep in range(2):                                  # two rounds: the host side is noisy
        for name, kw in variants:
            if kw.get("host_threads", 0) > cores:
                continue
            host = gobblet_v1.HostVecEnv(n, device=dev, seed=1, **kw)
            dt = run(host, h_log, W, K, dev)
            assert torch.equal(host.env.state, logger.state), name
            r = {"env_steps_per_s": n / dt, "ms_per_step": dt * 1e3, "pcie_gbs": host.d2h_bytes_per_step / dt / 1e9,
                 "host_gbs": (host.host_bytes_per_step if host.expand else host.d2h_bytes_per_step) / dt / 1e9}
            if name not in res or r["env_steps_per_s"] > res[name]["env_steps_per_s"]:
                res[name] = r
            del host
    # store mode of the expander: staged + non-temporal (default) vs direct regular stores
    ops.LIB.gbl_host_set_store_mode(1)
    host = gobblet_v1.HostVecEnv(n, device=dev, seed=1, wire="packed", chunks=8)
    dt = run(host, h_log, W, K, dev)
    res["packed_chunks8_regular_stores"] = {"env_steps_per_s": n / dt, "ms_per_step": dt * 1e3}
    ops.LIB.gbl_host_set_store_mode(0)
    # expander alone (records already in host memory)
    rec = host.h_rec
    outs = (host.h_obs, host.h_mask, host.h_rew, host.h_term.view(torch.uint8), host.h_trunc.view(torch.uint8), host.h_agent)
    for t in sorted({1, 4, 8, cores}):
        if t > cores:
            continue
        ops.host_unpack(rec, *outs, threads=t)
        t0 = time.perf_counter()
        for _ in range(5):
            ops.host_unpack(rec, *outs, threads=t)
        dt = (time.perf_counter() - t0) / 5
        res[f"expander_alone_threads{t}"] = {"env_per_s": n / dt, "ms": dt * 1e3, "host_gbs": n * 176 / dt / 1e9}
    out["paths"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
