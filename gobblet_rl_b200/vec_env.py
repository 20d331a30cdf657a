"""Vectorised batched entry point: N lockstep Gobblet environments on one GPU.

Mirrors what a Tianshou-style driver does with N copies of the reference env -- `env.step(a);
env.last()` per env (gobblet.py:231-273, :179-215) -- as ONE kernel launch over
torch CUDA tensors, plus the fused random-legal rollout of example_basic.py:50-67.
"""
from typing import Optional

import torch

from . import ops

STAT_NAMES = ("episodes", "player_1_wins", "player_2_wins", "steps", "sum_episode_length", "illegal_moves",
              "both_line_endings", "max_episode_length")


class VecEnv:
    """`num_envs` lockstep environments whose state lives in HBM (16 B / env).

    illegal_mode  "terminate": `env()` semantics, TerminateIllegalWrapper(-1) (gobblet.py:110-117)
                  "pass":      `raw_env` semantics, illegal move is a no-op and the turn passes
                               (board.py:125-126, gobblet.py:244-270)
    autoreset     "same_step": a finished env is reset inside the step that ended it; the step returns the
                               terminal reward / flags and the RESET observation (final_obs optional)
                  "next_step": the call after a terminal step only resets that env
                  "off":       finished envs stay finished until `reset(ids)` (Tianshou's order)
    Outputs are views of persistent buffers, overwritten by the next call (pass `out=` to redirect).
    Global env ids `env_id_base + i` key the Philox sampler, so any sharding gives the same games.
    graph_safe=True keeps the sampler's step counter on the device so rollouts can be CUDA-graph captured.
    """

    def __init__(self, num_envs: int, device="cuda", seed: int = 0, illegal_mode: str = "terminate",
                 autoreset: str = "same_step", env_id_base: int = 0, streaming_stores: bool = True,
                 skip255: bool = False, graph_safe: bool = False):
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops.GobbletError("VecEnv needs a CUDA device: the engine has no CPU path")
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        self.illegal_mode, self.autoreset = illegal_mode, autoreset
        self.flags = ops.make_flags(illegal_mode, autoreset, streaming_stores, skip255)
        n, dev = self.num_envs, self.device
        self.state = torch.zeros((n, 2), dtype=torch.int64, device=dev)
        self.obs = torch.zeros((n, 3, 3, 13), dtype=torch.int8, device=dev)
        self.mask = torch.zeros((n, 54), dtype=torch.int8, device=dev)
        self.rew = torch.zeros((n, 2), dtype=torch.int8, device=dev)
        self.terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.truncated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.agent_id = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.stats = torch.zeros(8, dtype=torch.int64, device=dev)
        self.step_count = 0          # absolute lockstep step index (Philox counter)
        # graph_safe: the Philox step counter lives in device memory and is advanced by a torch op after
        # every rollout, so `rollout_random` can be captured in a CUDA graph and replayed (small-N regime)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev) if graph_safe else None
        self.kernel_launches = 0
        self.reset()

    # -- reference surface, batched ------------------------------------------------------------------
    def reset(self, ids: Optional[torch.Tensor] = None):
        """raw_env.reset() for all envs, or for the envs selected by a bool/uint8 mask or index tensor."""
        which = None
        if ids is not None:
            ids = torch.as_tensor(ids, device=self.device)
            if ids.dtype in (torch.bool, torch.uint8) and ids.numel() == self.num_envs:
                which = ids.to(torch.uint8).contiguous()
            else:
                which = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
                which[ids.long()] = 1
        ops.reset(self.state, which)
        self.kernel_launches += 1
        return self.observe()

    def observe(self):
        ops.observe(self.state, self.obs, self.mask, self.agent_id)
        self.kernel_launches += 1
        return self.obs, self.mask, self.agent_id

    def step(self, actions: torch.Tensor, final: Optional[tuple] = None, out: Optional[tuple] = None,
             aux_out: Optional[tuple] = None):
        """-> (obs[N,3,3,13] i8, mask[N,54] i8, rew[N,2] i8, terminated[N] bool, truncated[N] bool, agent_id[N] u8)

        `final=(final_obs, final_mask)` receives the observation before a same-step reset replaces it.
        `out=(obs, mask)` and `aux_out=(rew, terminated, truncated, agent_id)` redirect the kernel's stores,
        e.g. straight into the slot of a trajectory / replay buffer (no copy).
        With `illegal_mode="terminate"` the observation after an illegal move is the MOVER's (live mask);
        PettingZoo's `last()` would select player_1 there -- `adapters.PettingZooVecEnv` applies that rule."""
        actions = torch.as_tensor(actions, device=self.device)
        if actions.dtype not in (torch.uint8, torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        obs, mask = (self.obs, self.mask) if out is None else out
        rew, term, trunc, agent = (self.rew, self.terminated, self.truncated, self.agent_id) if aux_out is None else aux_out
        fobs, fmask = (None, None) if final is None else final
        ops.step(self.state, actions.contiguous(), obs, mask, rew, term.view(torch.uint8), trunc.view(torch.uint8),
                 agent, fobs, fmask, self.stats, self.flags)
        self.step_count += 1
        self.kernel_launches += 1
        return obs, mask, rew, term, trunc, agent

    def rollout_random(self, T: int, ring: int = 1, emit: bool = True, per_step: bool = False,
                       log_actions: bool = False):
        """T fused lockstep steps with uniform random legal actions (example_basic.py:50-67) in ONE launch.

        emit      write the next observation + mask of every step to ring slot (step % ring)
        per_step  also write rew [ring,N,2], terminated [ring,N], agent_id [ring,N]
        Returns a dict of the buffers that were requested; `self.stats` accumulates episode statistics."""
        n, dev, out = self.num_envs, self.device, {}
        obs_out = mask_out = rew_out = term_out = agent_out = log = None
        if emit:
            obs_out, mask_out = self._ring_buffers(ring)
            out["obs"], out["mask"] = obs_out, mask_out
        if per_step:
            rew_out = torch.zeros((ring, n, 2), dtype=torch.int8, device=dev)
            term_out = torch.zeros((ring, n), dtype=torch.uint8, device=dev)
            agent_out = torch.zeros((ring, n), dtype=torch.uint8, device=dev)
            out["rew"], out["terminated"], out["agent_id"] = rew_out, term_out.view(torch.bool), agent_out
        if log_actions:
            log = torch.zeros((T, n), dtype=torch.uint8, device=dev)
            out["actions"] = log
        ops.rollout_random(self.state, int(T), self.seed, self.env_id_base, self.step_count, obs_out, mask_out,
                           rew_out, term_out, agent_out, log, self.stats, self.flags, self.step_dev)
        if self.step_dev is not None:
            self.step_dev += int(T)
        self.step_count += int(T)
        self.kernel_launches += 1
        return out

    def _ring_buffers(self, ring):
        key = int(ring)
        cache = getattr(self, "_rings", None)
        if cache is None or cache[0] != key:
            n = self.num_envs
            pad = -(-n // 16) * 16                      # slot strides must be multiples of 16 bytes
            obs = torch.zeros((key, pad, 3, 3, 13), dtype=torch.int8, device=self.device)[:, :n]
            mask = torch.zeros((key, pad, 54), dtype=torch.int8, device=self.device)[:, :n]
            self._rings = (key, obs, mask)
        return self._rings[1], self._rings[2]

    # -- views in the reference's own state layout (board.py:33) ----------------------------------------
    def squares(self):
        """int8 [N,27] signed piece numbers = `env.board.squares` of every env, and agent_selection [N]."""
        sq = torch.zeros((self.num_envs, 27), dtype=torch.int8, device=self.device)
        agent = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        ops.export_squares(self.state, sq, agent)
        return sq, agent

    def set_squares(self, squares, agent=None):
        sq = torch.as_tensor(squares, device=self.device).to(torch.int8).reshape(self.num_envs, 27).contiguous()
        ag = None if agent is None else torch.as_tensor(agent, device=self.device).to(torch.uint8).contiguous()
        ops.import_squares(self.state, sq, ag)
        return self.observe()

    def stats_dict(self):
        return dict(zip(STAT_NAMES, self.stats.tolist()))

    # -- checkpoint: the state tensor + counters are the whole env state ------------
    def state_dict(self):
        return {"state": self.state.clone(), "stats": self.stats.clone(), "step_count": self.step_count,
                "seed": self.seed, "env_id_base": self.env_id_base}

    def load_state_dict(self, sd):
        self.state.copy_(sd["state"])
        self.stats.copy_(sd["stats"])
        self.step_count, self.seed, self.env_id_base = sd["step_count"], sd["seed"], sd["env_id_base"]


class HostVecEnv:
    """The same step through HOST buffers: pinned host actions in, pinned host obs/mask/rew/flags out,
    copies chunked over CUDA streams so PCIe transfers overlap the kernels.  This is the end-to-end
    path `bench.py` reports as `e2e` (what a CPU-side PettingZoo/Tianshou driver would see)."""

    def __init__(self, num_envs, device="cuda", chunks=2, **kw):
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        bounds = [self.num_envs * i // chunks for i in range(chunks + 1)]
        self.parts = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        base = kw.pop("env_id_base", 0)
        self.envs = [VecEnv(b - a, device=device, env_id_base=base + a, **kw) for a, b in self.parts]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.parts]
        n = self.num_envs
        pin = dict(pin_memory=True)
        self.h_actions = torch.zeros(n, dtype=torch.uint8, **pin)
        self.h_obs = torch.zeros((n, 3, 3, 13), dtype=torch.int8, **pin)
        self.h_mask = torch.zeros((n, 54), dtype=torch.int8, **pin)
        self.h_rew = torch.zeros((n, 2), dtype=torch.int8, **pin)
        self.h_term = torch.zeros(n, dtype=torch.bool, **pin)
        self.h_trunc = torch.zeros(n, dtype=torch.bool, **pin)
        self.h_agent = torch.zeros(n, dtype=torch.uint8, **pin)
        self.d_actions = [torch.zeros(b - a, dtype=torch.uint8, device=self.device) for a, b in self.parts]
        self.h2d_bytes_per_step = n
        self.d2h_bytes_per_step = n * (117 + 54 + 2 + 1 + 1 + 1)

    def reset(self):
        for (a, b), e in zip(self.parts, self.envs):
            obs, mask, agent = e.reset()
            self.h_obs[a:b].copy_(obs); self.h_mask[a:b].copy_(mask); self.h_agent[a:b].copy_(agent)
        torch.cuda.synchronize(self.device)
        return self.h_obs, self.h_mask, self.h_agent

    def step(self, actions_host: torch.Tensor):
        """actions_host: pinned uint8 [N].  Returns pinned host tensors; synchronises before returning."""
        cur = torch.cuda.current_stream(self.device)
        for (a, b), e, s, da in zip(self.parts, self.envs, self.streams, self.d_actions):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                da.copy_(actions_host[a:b], non_blocking=True)
                obs, mask, rew, term, trunc, agent = e.step(da)
                self.h_obs[a:b].copy_(obs, non_blocking=True)
                self.h_mask[a:b].copy_(mask, non_blocking=True)
                self.h_rew[a:b].copy_(rew, non_blocking=True)
                self.h_term[a:b].copy_(term, non_blocking=True)
                self.h_trunc[a:b].copy_(trunc, non_blocking=True)
                self.h_agent[a:b].copy_(agent, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return self.h_obs, self.h_mask, self.h_rew, self.h_term, self.h_trunc, self.h_agent

    @property
    def kernel_launches(self):
        return sum(e.kernel_launches for e in self.envs)
