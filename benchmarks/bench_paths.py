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

    # -- step_kernel, device resident: replay logged legal actions (uint8).  The 32 launches are captured in a
    #    CUDA graph (state restored at its head) so the figure is the kernel's, not Python's launch rate. ----
    K = 32
    logger = gobblet_v1.vec_env(n, device=dev, seed=2)
    log = logger.rollout_random(K, emit=False, log_actions=True)["actions"]
    vec = gobblet_v1.vec_env(n, device=dev, seed=2)
    state0 = vec.state.clone()
    # four alternating output slots (717 MB > L2) so the stores go to DRAM
    slots = [(torch.zeros((n, 3, 3, 13), dtype=torch.int8, device=dev), torch.zeros((n, 54), dtype=torch.int8, device=dev)) for _ in range(4)]

    def replay_steps():
        vec.state.copy_(state0)
        for k in range(K):
            vec.step(log[k], out=slots[k & 3])

    replay_steps()
    assert torch.equal(vec.state, logger.state)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        replay_steps()
        with torch.cuda.graph(g, stream=side):
            replay_steps()
    torch.cuda.current_stream().wait_stream(side)
    dt_graph = timed(g.replay, 10) / K
    assert torch.equal(vec.state, logger.state)
    dt_eager = timed(replay_steps, 5) / K
    bytes_step = 171 + 32 + 1 + 5
    out["step_kernel"] = {"envs": n, "env_steps_per_s": n / dt_graph, "us_per_launch_graph": dt_graph * 1e6,
                          "us_per_launch_python_loop": dt_eager * 1e6,
                          "algorithmic_bytes_per_env_step": bytes_step, "achieved_gbs": n * bytes_step / dt_graph / 1e9,
                          "frac_of_peak": n * bytes_step / dt_graph / 1e9 / peak}

    # -- BASELINE config 2 (4096 envs): per-step launches vs a CUDA graph of them vs one fused launch -----
    n2, T2 = 4096, 256
    small = gobblet_v1.vec_env(n2, device=dev, seed=0, graph_safe=True)

    def per_step_launches():
        for _ in range(T2):
            small.rollout_random(1, ring=1)

    dt_launch = timed(per_step_launches, 3) / T2
    g2 = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        per_step_launches()
        with torch.cuda.graph(g2, stream=side):
            per_step_launches()
    torch.cuda.current_stream().wait_stream(side)
    dt_g2 = timed(g2.replay, 10) / T2
    dt_fused = timed(lambda: small.rollout_random(T2, ring=4), 10) / T2
    out["c2_4096_envs"] = {"per_step_launch_env_steps_per_s": n2 / dt_launch, "cuda_graph_env_steps_per_s": n2 / dt_g2,
                           "fused_env_steps_per_s": n2 / dt_fused, "us_per_step": {"launch": dt_launch * 1e6, "graph": dt_g2 * 1e6, "fused": dt_fused * 1e6},
                           "illegal_moves": int(small.stats[5])}

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
    eps_sampler = adapters.RandomLegalPolicy(seed=3, graph_safe_device=dev)

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
    graph5 = col.capture()
    dt_graph = timed(graph5.replay, 10)
    out["collect_mlp_policy_c5"] = {"envs": n5, "steps": T5, "env_steps_per_s": n5 * T5 / dt,
                                    "env_steps_per_s_cuda_graph": n5 * T5 / dt_graph, "buffer_bytes": buf.nbytes(),
                                    "illegal_moves": int(v5.stats[5]), "note": "includes fp16 MLP inference + masked eps-greedy in torch"}
    col_f = adapters.VecCollector(v5, adapters.RandomLegalPolicy(seed=3), buf)
    dt = timed(col_f.collect, 10)
    col_u = adapters.VecCollector(v5, adapters.RandomLegalPolicy(seed=3), buf, fused=False)
    dt_u = timed(lambda: (col_u.collect(), col_u.roll()), 10)
    out["collect_random_policy"] = {"envs": n5, "steps": T5, "env_steps_per_s": n5 * T5 / dt, "launches": 1,
                                    "env_steps_per_s_unfused_sample_step_per_step": n5 * T5 / dt_u,
                                    "note": "fused = ONE rollout launch writing into the buffer slots (incl. terminal observations)"}
    # -- masked-uniform sampler alone (coalesced 128-bit mask loads), 2^20 rows ---------------------------------
    from gobblet_rl_b200 import ops
    act = torch.zeros(n, dtype=torch.int32, device=dev)
    dt = timed(lambda: ops.sample_legal(vec.mask, 1, 0, 0, act), 50)
    out["sample_legal_kernel"] = {"rows": n, "rows_per_s": n / dt, "us_per_launch": dt * 1e6,
                                  "achieved_gbs": n * (54 + 4) / dt / 1e9, "frac_of_peak": n * 58 / dt / 1e9 / peak}
    # -- packed wire format step (24-byte records), device resident -------------------------------------------
    rec = torch.zeros((n, 6), dtype=torch.int32, device=dev)
    a8 = torch.zeros(n, dtype=torch.uint8, device=dev)
    dt = timed(lambda: vec.step_packed(a8, rec=rec), 50)
    out["step_packed_kernel"] = {"envs": n, "env_steps_per_s": n / dt, "us_per_launch": dt * 1e6,
                                 "achieved_gbs": n * (32 + 1 + 24) / dt / 1e9, "frac_of_peak": n * 57 / dt / 1e9 / peak}
    # -- the drop-in facade: gobblet_v1.env() driven by the reference's own loop (example_basic.py:50-67) ------
    import time

    import numpy as np
    env = gobblet_v1.env(render_mode=None)
    np.random.seed(0)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 3.0:
        env.reset()
        for agent in env.agent_iter():
            obs, reward, term, trunc, info = env.last()
            if term or trunc:
                env.step(None)
            else:
                m = obs["action_mask"]
                env.step(np.random.choice(np.arange(len(m)), p=m / np.sum(m)))
                steps += 1
    out["facade_env_loop"] = {"env_steps_per_s": steps / (time.perf_counter() - t0),
                              "note": "batch-of-1 AEC facade: one kernel launch + one stream sync per step (zero-copy pinned buffer), rest is Python"}
    # -- the reference's CPU numbers for the same rows (1 core), when its tree travelled in baseline/_ref ------
    try:
        from oracle import reference_loader as RL
        if RL.available():
            Board = RL.load_board().Board
            gp = RL.load_greedy()
            rng = np.random.default_rng(0)
            t0, steps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 5.0:          # Board-only rollout: 54 x is_legal + play_turn + winner
                b, agent = Board(), 0
                while True:
                    legal = [a for a in range(54) if b.is_legal(a, agent)]
                    b.play_turn(agent, int(rng.choice(legal)))
                    steps += 1
                    agent = 1 - agent
                    if b.check_game_over():
                        break
            out["cpu_reference_board_only"] = {"env_steps_per_s": steps / (time.perf_counter() - t0), "cores": 1}
            pol = gp.GreedyGobbletPolicy(depth=2)
            o_np, m_np = gobs[:24].cpu().numpy(), gmask[:24].cpu().numpy()
            t0 = time.perf_counter()
            for i in range(24):
                pol.compute_action(o_np[i], m_np[i])
            out["cpu_reference_greedy_depth2"] = {"boards_per_s": 24 / (time.perf_counter() - t0), "cores": 1}
    except Exception as exc:  # measurement aid only
        out["cpu_reference_error"] = repr(exc)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
