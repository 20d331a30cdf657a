"""GreedyGobbletPolicy on the GPU (warp-per-board search kernel, gobblet_greedy.cu).

Same constructor, attributes and entry points as gobblet_rl/game/greedy_policy.py:8-36:
`GreedyGobbletPolicy(depth=2, seed=0)`, `.compute_action(obs, mask)`, `.compute_action_tianshou(obs)`,
`.compute_actions_rllib(obs_batch)`, `.prev_actions`; plus the batched `greedy_actions(...)`.
"""
from typing import Any, Optional

import numpy as np
import torch

from . import ops


def greedy_actions(obs: torch.Tensor, mask: torch.Tensor, prev3: Optional[torch.Tensor] = None, depth: int = 2,
                   seed: int = 0, ctr_base: int = 0, details: bool = False):
    """Depth-1/2 greedy move for a batch of boards given as observations.

    obs int8 [N,3,3,13] (or [N,117]), mask int8/bool [N,54], prev3 int16 [N,3] (-1 = none): the agent's
    last three actions, which trigger the random fallback when the choice repeats one of them
    (greedy_policy.py:211-214).  Returns act int32 [N]; with details=True also (chosen, cand, used_fallback):
    the choice before the fallback (-1 = None), the candidate set as a uint64 bit mask (what
    np.random.choice would draw from, :217) and whether the fallback fired."""
    dev = obs.device
    n = obs.shape[0]
    obs = obs.to(torch.int8).reshape(n, 117).contiguous()
    mask = mask.to(torch.int8).reshape(n, 54).contiguous()
    if prev3 is not None:
        prev3 = prev3.to(device=dev, dtype=torch.int16).reshape(n, 3).contiguous()
    act = torch.empty(n, dtype=torch.int32, device=dev)
    chosen = torch.empty(n, dtype=torch.int32, device=dev) if details else None
    cand = torch.empty(n, dtype=torch.int64, device=dev) if details else None
    fb = torch.empty(n, dtype=torch.uint8, device=dev) if details else None
    ops.greedy(obs, mask, prev3, int(depth), int(seed), int(ctr_base), act, chosen, cand, fb)
    return (act, chosen, cand, fb.bool()) if details else act


class GreedyGobbletPolicy:
    def __init__(self, depth: Optional[int] = 2, seed: Optional[int] = 0, device="cuda", **kwargs: Any) -> None:
        # the reference tests `depth > 1` (greedy_policy.py:103) and `depth == 3` (:160); its depth-3 branch only
        # re-assigns the choice depth 2 has just made (:198 after :157), edits a local list and breaks out of its own
        # loop, so every depth >= 2 returns what depth 2 returns (recorded: tests/golden/greedy_depth3.npz)
        if not isinstance(depth, (int, np.integer)) or isinstance(depth, bool):
            raise TypeError("depth must be an int (the reference compares it with 1 and 3)")
        self.board = None
        self.depth = int(depth)
        self.device = torch.device(device)
        self.rng = np.random.default_rng()              # kept for attribute parity (unused, as in the reference)
        self.prev_actions = {i: [] for i in range(2)}   # greedy_policy.py:19

    def compute_actions_rllib(self, obs_batch):         # greedy_policy.py:21-31
        observations = obs_batch["observation"]
        observations = observations.reshape(observations.shape[0], 3, 3, -1)
        masks = obs_batch["action_mask"]
        return [self.compute_action(observations[i], masks[i]) for i in range(len(observations))]

    def compute_action_tianshou(self, obs):             # greedy_policy.py:33-36
        mask = obs.mask
        obs = obs.obs if hasattr(obs, "obs") else obs
        return self.compute_action(obs, mask)

    def compute_action(self, obs, mask) -> np.ndarray:
        """One board.  The search runs on the GPU; the random fallback draws from numpy's GLOBAL generator
        with the same candidate list as the reference (greedy_policy.py:216-217), so a seeded script sees the
        same action stream."""
        obs_np = np.asarray(obs)
        mask_np = np.asarray(mask).reshape(-1)
        agent_index = int(obs_np[..., 12].max())        # greedy_policy.py:58-60
        prev = (self.prev_actions[agent_index][-3:] + [-1, -1, -1])[:3]
        o = torch.as_tensor(obs_np.astype(np.int8).reshape(1, 117), device=self.device)
        m = torch.as_tensor((mask_np != 0).astype(np.int8).reshape(1, 54), device=self.device)
        p = torch.tensor([prev], dtype=torch.int16, device=self.device)
        act, chosen, cand, fb = greedy_actions(o, m, p, depth=1 if self.depth <= 1 else 2, details=True)
        res = torch.stack([act.long(), chosen.long(), cand, fb.long()]).cpu().numpy()[:, 0]
        if res[3]:
            bits = int(res[2]) & (2**64 - 1)
            actions_depth1 = [a for a in range(54) if (bits >> a) & 1]
            chosen_action = np.random.choice(actions_depth1)
        else:
            chosen_action = int(res[1])
        self.prev_actions[agent_index].append(chosen_action)
        return np.array(chosen_action)
