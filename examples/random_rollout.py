#!/usr/bin/env python
"""The reference's random-legal rollout (gobblet_rl/examples/example_basic.py:44-67, render_mode=None),
unchanged except for the import -- then the same workload on the vectorised entry point."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import argparse

import numpy as np

from gobblet_rl_b200 import gobblet_v1

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--games", type=int, default=3)
    ap.add_argument("--render_mode", default=None, choices=[None, "text", "text_full"])
    ap.add_argument("--num-envs", type=int, default=1 << 16)
    args = ap.parse_args()
    np.random.seed(args.seed)

    env = gobblet_v1.env(render_mode=args.render_mode, args=args)
    for _ in range(args.games):
        env.reset()
        for agent in env.agent_iter():
            observation, reward, termination, truncation, info = env.last()
            if termination or truncation:
                print(f"Agent: ({agent}), Reward: {reward}, info: {info}")
                env.step(None)
            else:
                action_mask = observation["action_mask"]
                action = np.random.choice(np.arange(len(action_mask)), p=action_mask / np.sum(action_mask))
                env.step(action)

    vec = gobblet_v1.vec_env(args.num_envs, seed=args.seed)
    vec.rollout_random(256, ring=1)
    print(f"{args.num_envs} envs x 256 lockstep steps:", vec.stats_dict())
