#!/usr/bin/env python
"""GPU-resident rollout collection in the shape Tianshou's DQN example consumes
(gobblet_rl/examples/example_tianshou_DQN.py:161-166, :401-409): an MLP 117 -> 128x4 -> 54 picks masked
epsilon-greedy actions, the step kernel writes obs / mask / reward / flags straight into a TrajectoryBuffer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import torch

from gobblet_rl_b200 import adapters, gobblet_v1

if __name__ == "__main__":
    dev = torch.device("cuda")
    n, horizon, eps = 131072, 16, 0.1
    net = torch.nn.Sequential(torch.nn.Linear(117, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                              torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                              torch.nn.Linear(128, 54)).to(dev)
    explore = adapters.RandomLegalPolicy(seed=0)

    @torch.no_grad()
    def policy(obs, mask, agent_id):
        q = net(obs.reshape(obs.shape[0], 117).float()).masked_fill(mask == 0, float("-inf"))
        greedy = q.argmax(1).to(torch.int32)
        return torch.where(torch.rand(obs.shape[0], device=dev) < eps, explore(obs, mask), greedy)

    vec = gobblet_v1.vec_env(n, seed=0)
    buf = adapters.TrajectoryBuffer(horizon, n)
    collector = adapters.VecCollector(vec, policy, buf)
    for it in range(8):
        collector.collect()          # buf.obs[t], buf.mask[t], buf.act[t], buf.rew[t], buf.terminated[t], buf.final_obs[t]
        collector.roll()
    torch.cuda.synchronize()
    print(f"collected {8 * horizon * n} env-steps; buffer {buf.nbytes() / 1e6:.0f} MB;", vec.stats_dict())
