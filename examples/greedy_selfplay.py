#!/usr/bin/env python
"""Greedy vs greedy, the loop of tutorials/GreedyAgent/tutorial_greedy.py:24-50 (first two plies random),
then the batched search: one warp per board for 65 536 boards."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import numpy as np
import torch

from gobblet_rl_b200 import gobblet_v1

if __name__ == "__main__":
    np.random.seed(0)
    env = gobblet_v1.env(render_mode=None)
    env.reset()
    policy = gobblet_v1.GreedyGobbletPolicy(depth=2)
    iteration = 0
    for agent in env.agent_iter():
        observation, reward, termination, truncation, info = env.last()
        if termination or truncation:
            print(f"Agent: ({agent}), Reward: {reward}, info: {info}")
            env.step(None)
            continue
        if iteration < 2:
            mask = observation["action_mask"]
            action = np.random.choice(np.arange(len(mask)), p=mask / np.sum(mask))
        else:
            action = policy.compute_action(observation["observation"], observation["action_mask"])
        env.step(action)
        iteration += 1
    print("plies:", iteration)

    vec = gobblet_v1.vec_env(1 << 16, seed=1, autoreset="off")
    vec.rollout_random(6, emit=False)
    obs, mask, agent_id = vec.observe()
    act = gobblet_v1.greedy_actions(obs, mask, depth=2)
    torch.cuda.synchronize()
    print("greedy moves for 65536 boards:", act[:16].tolist(), "...")
