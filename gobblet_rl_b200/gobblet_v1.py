"""`gobblet_v1` -- the reference's public module surface (gobblet_rl/gobblet_v1.py:1-3) on the GPU engine.

    env(render_mode=None, args=None)       wrapper stack of gobblet.py:110-117 over raw_env
    raw_env(render_mode=None, args=None)   PettingZoo AEC env (gobblet.py:123-290), a batch-of-1 view
    GreedyGobbletPolicy(depth=2, seed=0)   greedy_policy.py:8-221
    vec_env(num_envs, ...)                 NEW: vectorised entry point over torch CUDA tensors
pygame rendering (gobblet.py:430-573) is out of scope; the `text` / `text_full` debug views are provided
(`text_view.py`); with render_mode=None `render()` only warns, as the reference does (gobblet.py:293-297).
"""
import warnings

import numpy as np
import torch

from . import _aec, _spaces, ops
from ._aec import AECEnv, agent_selector
from .greedy_policy import GreedyGobbletPolicy, greedy_actions  # noqa: F401
from .vec_env import HostVecEnv, VecEnv  # noqa: F401


def vec_env(num_envs, device="cuda", seed=0, illegal_mode="terminate", autoreset="same_step", **kw):
    return VecEnv(num_envs, device=device, seed=seed, illegal_mode=illegal_mode, autoreset=autoreset, **kw)


def env(render_mode=None, args=None, device="cuda"):
    e = raw_env(render_mode=render_mode, args=args, device=device)
    e = _aec.TerminateIllegalWrapper(e, illegal_reward=-1)     # gobblet.py:114
    e = _aec.AssertOutOfBoundsWrapper(e)                       # gobblet.py:115
    e = _aec.OrderEnforcingWrapper(e)                          # gobblet.py:116
    return e


class ManualGobbletPolicy:
    """Name kept for import compatibility (gobblet_rl/gobblet_v1.py:3).  The reference's class is a pygame
    mouse / keyboard UI (manual_policy.py); interactive play is out of scope of the batched engine."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("ManualGobbletPolicy is a pygame UI in the reference; not part of the B200 engine")


def parallel_env(**kwargs):
    raise NotImplementedError("gobblet is turn based; the reference skips the parallel API too "
                              "(tests/test_gobblet_env.py:37-42)")


class _BoardView:
    """`env.board` as callers of the reference read it: `.squares` float64[27] (board.py:33)."""

    def __init__(self, owner):
        self._owner = owner

    @property
    def squares(self):
        sq, _ = self._owner._vec.squares()
        return sq[0].cpu().numpy().astype(np.float64)

    def get_flatboard(self):
        """Top-of-stack view (board.py:159-177) of the engine state; formatting only."""
        from .text_view import flatboard_from_squares
        return flatboard_from_squares(self.squares)

    def __str__(self):
        return str(self.squares.reshape(3, 3, 3))


class raw_env(AECEnv):
    metadata = {
        "render_modes": ["human", "rgb_array", "text", "text_full"],
        "name": "gobblet_v1",
        "is_parallelizable": True,
        "render_fps": 60,
        "has_manual_policy": True,
    }

    def __init__(self, render_mode=None, args=None, device="cuda"):
        super().__init__()
        # one env of the batched engine; illegal moves pass the turn like the reference's raw_env
        self._vec = VecEnv(1, device=device, illegal_mode="pass", autoreset="off")
        self._scratch = None
        self._bind_outputs()
        self.board = _BoardView(self)
        self.board_size = 3
        self.agents = ["player_1", "player_2"]
        self.possible_agents = self.agents[:]
        self.action_spaces = {i: _spaces.Discrete(54) for i in self.agents}
        self.observation_spaces = {
            i: _spaces.Dict({
                "observation": _spaces.Box(low=0, high=1, shape=(3, 3, 13), dtype=np.int8),
                "action_mask": _spaces.Box(low=0, high=1, shape=(54,), dtype=np.int8),
            }) for i in self.agents
        }
        self.rewards = {i: 0 for i in self.agents}
        self.terminations = {i: False for i in self.agents}
        self.truncations = {i: False for i in self.agents}
        self.infos = {i: {"legal_moves": list(range(0, 9))} for i in self.agents}
        self._agent_selector = agent_selector(self.agents)
        self.agent_selection = self._agent_selector.reset()
        self.render_mode = render_mode
        self.debug = args.debug if hasattr(args, "debug") else False
        self.screen_width = args.screen_width if hasattr(args, "screen_width") else 640
        self.screen_height = self.screen_width
        self.screen = None
        self._engine_observe()

    # Batch-of-1 fast path: the action and ALL per-step outputs live in ONE pinned host buffer that the kernels
    # address directly (UVA zero-copy), and the facade calls the C ABI itself -- a step is one kernel launch and
    # one stream synchronise, no staging copies and no tensor dispatch.  Layout: obs @0 (117 B) | mask @128 (54 B)
    # | rew @192 | terminated @194 | truncated @195 | agent_id @196 | action (int64) @208.
    def _bind_outputs(self):
        v = self._vec
        self._hbuf = torch.zeros(256, dtype=torch.uint8).pin_memory()
        self._host = self._hbuf.numpy()
        self._host_action = self._host[208:216].view(np.int64)
        base = self._hbuf.data_ptr()
        self._ptr = {"obs": base, "mask": base + 128, "rew": base + 192, "term": base + 194, "trunc": base + 195,
                     "agent": base + 196, "action": base + 208, "state": v.state.data_ptr(), "stats": v.stats.data_ptr()}

    def _finish(self, rc, stream):
        if rc != 0:
            raise ops.GobbletError(f"libgobblet_b200 error {rc}: {ops.LIB.gbl_last_error().decode()}")
        stream.synchronize()
        self._pull()

    def _engine_observe(self):
        p, dev = self._ptr, self._vec.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            self._finish(ops.LIB.gbl_observe(p["state"], p["obs"], p["mask"], p["agent"], 1, stream.cuda_stream), stream)

    def _engine_step(self, action):
        p, v = self._ptr, self._vec
        self._host_action[0] = action
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device)
            self._finish(ops.LIB.gbl_step(p["state"], p["action"], 8, p["obs"], p["mask"], p["rew"], p["term"], p["trunc"],
                                          p["agent"], None, None, p["stats"], 1, v.flags, stream.cuda_stream), stream)

    def _pull(self):
        host = self._host
        self._obs = host[:117].view(np.int8).reshape(3, 3, 13).copy()
        self._mask = host[128:182].view(np.int8).copy()
        self._rew = host[192:194].view(np.int8).astype(int)
        self._term = bool(host[194])
        self._sel = int(host[196])

    def observe(self, agent):                                   # gobblet.py:179-215
        # the reference indexes the LIVE agent list (gobblet.py:182, :199, :225): once a dead step has
        # removed player_1, player_2 is observed (and its mask built) as index 0
        idx = self.agents.index(agent)
        if idx == self._sel:
            planes = self._obs.copy()
        else:   # the other agent's perspective: sign flip only (gobblet.py:182-185) = own/opponent planes swapped
            planes = np.empty_like(self._obs)
            planes[..., 0:6], planes[..., 6:12] = self._obs[..., 6:12], self._obs[..., 0:6]
            planes[..., 12] = 1 - self._obs[..., 12]
        if agent != self.agent_selection:                       # all-zero mask (gobblet.py:209-213)
            return {"observation": planes, "action_mask": np.zeros(54, "int8")}
        if idx == self._sel:
            return {"observation": planes, "action_mask": self._mask.copy()}
        # agent_selection moved on without a board move (dead steps after the game ended): the reference
        # still builds a live mask for it (gobblet.py:209, :223-228).  Evaluate the same board with that
        # agent to move on a scratch env of the engine.
        if self._scratch is None:
            self._scratch = VecEnv(1, device=self._vec.device, illegal_mode="pass", autoreset="off")
        sq, _ = self._vec.squares()
        _, mask, _ = self._scratch.set_squares(sq, torch.tensor([idx], dtype=torch.uint8))
        return {"observation": planes, "action_mask": mask[0].cpu().numpy().copy()}

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    def _legal_moves(self):                                     # gobblet.py:223-228
        return [int(a) for a in np.flatnonzero(self._mask)]

    def step(self, action):                                     # gobblet.py:231-273
        if self.terminations[self.agent_selection] or self.truncations[self.agent_selection]:
            return self._was_dead_step(action)
        self._engine_step(int(action))
        next_agent = self._agent_selector.next()
        if self._term:
            self.rewards[self.agents[0]] += int(self._rew[0])   # += / -=, zeroed only in reset (:255-260)
            self.rewards[self.agents[1]] += int(self._rew[1])
            self.terminations = {i: True for i in self.agents}
        self._cumulative_rewards[self.agent_selection] = 0
        self.agent_selection = next_agent
        self._accumulate_rewards()
        self.turn += 1
        self.action = action
        if self.render_mode in ["human", "text", "text_full", "rgb_array"]:
            self.render()

    def reset(self, seed=None, return_info=False, options=None):   # gobblet.py:275-290
        ops.reset(self._vec.state, None)
        self._engine_observe()
        self.agents = self.possible_agents[:]
        self.rewards = {i: 0 for i in self.agents}
        self._cumulative_rewards = {i: 0 for i in self.agents}
        self.terminations = {i: False for i in self.agents}
        self.truncations = {i: False for i in self.agents}
        self.infos = {i: {} for i in self.agents}
        self._agent_selector.reinit(self.agents)
        self._agent_selector.reset()
        self.agent_selection = self._agent_selector.reset()
        self.turn = 0
        self.action = -1

    def render(self):
        if self.render_mode is None:
            warnings.warn("You are calling render method without specifying any render mode.")
            return
        if self.render_mode in ("text", "text_full"):           # debug views, gobblet.py:316-429
            from .text_view import render_text
            print(render_text(self.board.squares, self.turn, self.agent_selection, self.action,
                              full=self.render_mode == "text_full"), end="")
            return
        raise NotImplementedError("pygame rendering (human / rgb_array) is out of scope of the B200 engine "
                                  "(SURVEY.md section 2 #3)")

    def close(self):
        self.screen = None
