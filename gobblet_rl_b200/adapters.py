"""Callers either side of the hot path (SURVEY.md section 8f "next" rows), on the GPU engine.

* `PettingZooVecEnv`  -- what `tianshou.env.DummyVectorEnv([lambda: PettingZooEnv(gobblet_v1.env())] * N)`
  hands a Tianshou Collector (contract: collector_manual_policy.py:80-145, example_tianshou_DQN.py:212-224):
  `reset(id) -> (obs_batch, info)`, `step(action, id) -> (obs_next, rew[N,2], terminated, truncated, info)` with
  `obs = {"agent_id", "obs", "mask"}`; resets are the caller's job (Tianshou resets finished envs itself).
* `TrajectoryBuffer` + `VecCollector` -- GPU-resident collection (BASELINE config 5): the step kernel writes
  the next observation / mask / reward / flags straight into the buffer slot.
* `GreedyPolicy`      -- greedy_policy_tianshou.py:12-98 (`forward(batch) -> Batch(act=...)`).
* `RandomAdmissiblePolicy` -- random_admissible_policy_rllib.py:15-40 (`compute_actions(obs_batch)`).
"""
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops
from .greedy_policy import GreedyGobbletPolicy, greedy_actions
from .vec_env import VecEnv

try:  # pragma: no cover - tianshou is not in the build image
    from tianshou.data import Batch
except ImportError:

    class Batch(SimpleNamespace):
        """Minimal attribute bag standing in for tianshou.data.Batch (`batch.obs.mask`, `Batch(act=...)`)."""

        def __getitem__(self, key):
            return getattr(self, key)

        def __len__(self):
            for v in self.__dict__.values():
                try:
                    return len(v)
                except TypeError:
                    continue
            return 0


AGENTS = ("player_1", "player_2")


class PettingZooVecEnv:
    """N reference envs behind Tianshou's vector-env calling convention, stepped by ONE kernel launch.

    to_numpy=True returns host numpy arrays shaped like Tianshou's (`agent_id` as names, `mask` as bool);
    to_numpy=False keeps everything as CUDA tensors (`agent_id` as uint8 indices) for GPU-side learners."""

    def __init__(self, num_envs: int, device="cuda", to_numpy: bool = True, **kw):
        self.vec = VecEnv(num_envs, device=device, illegal_mode="terminate", autoreset="off", skip255=True, **kw)
        self.env_num = self.num_envs = int(num_envs)
        self.to_numpy = to_numpy
        self.agents = list(AGENTS)
        self._all = torch.arange(self.num_envs, device=self.vec.device)

    def __len__(self):
        return self.num_envs

    def _ids(self, id):
        if id is None:
            return None
        return torch.as_tensor(np.atleast_1d(id) if not torch.is_tensor(id) else id, device=self.vec.device).long()

    def _pack(self, obs, mask, agent, ids):
        if ids is not None:
            obs, mask, agent = obs[ids], mask[ids], agent[ids]
        if not self.to_numpy:
            return {"agent_id": agent, "obs": obs, "mask": mask.bool()}
        a = agent.cpu().numpy()
        return {"agent_id": np.array(AGENTS, dtype=object)[a], "obs": obs.cpu().numpy(), "mask": mask.bool().cpu().numpy()}

    def reset(self, id=None, **kwargs):
        ids = self._ids(id)
        obs, mask, agent = self.vec.reset(ids)
        n = self.num_envs if ids is None else len(ids)
        return self._pack(obs, mask, agent, ids), [{} for _ in range(n)]

    def step(self, action, id=None):
        """PettingZooEnv.step for the envs in `id` (all when None): env.step(a); env.last()."""
        v = self.vec
        ids = self._ids(id)
        if not (torch.is_tensor(action) and action.is_cuda):
            # AssertOutOfBoundsWrapper (gobblet.py:115): an action outside Discrete(54) raises.  Host-side actions
            # are checked here; CUDA tensors are not read back -- the kernel treats out-of-range as an illegal move.
            a_host = np.asarray(action.cpu() if torch.is_tensor(action) else action).reshape(-1)
            assert ((a_host >= 0) & (a_host < 54)).all(), "action is not in action space"
        act = torch.as_tensor(np.asarray(action) if not torch.is_tensor(action) else action, device=v.device).long().reshape(-1)
        if ids is None:
            full = act
        else:
            full = torch.full((self.num_envs,), 255, dtype=torch.int64, device=v.device)   # 255 = not stepped
            full[ids] = act
        obs, mask, rew, term, trunc, agent = v.step(full)
        # PettingZoo's last() after an illegal-move termination selects the FIRST dead agent = player_1
        # (TerminateIllegalWrapper -> _deads_step_first); when player_2 was the mover that is the other
        # player's view with an all-zero mask (gobblet.py:209-213).
        fix = trunc & (agent == 1)
        if bool(fix.any()):
            obs, mask, agent = obs.clone(), mask.clone(), agent.clone()
            o = obs[fix]
            obs[fix] = torch.cat([o[..., 6:12], o[..., 0:6], 1 - o[..., 12:13]], dim=-1)
            mask[fix] = 0
            agent[fix] = 0
        batch = self._pack(obs, mask, agent, ids)
        sel = slice(None) if ids is None else ids
        rew, term, trunc = rew[sel], term[sel], trunc[sel]
        if self.to_numpy:
            rew_np = rew.cpu().numpy().astype(np.float64)      # env.rewards values; -1.0 on illegal moves
            return batch, rew_np, term.cpu().numpy(), trunc.cpu().numpy(), [{} for _ in range(len(rew_np))]
        return batch, rew, term, trunc, [{} for _ in range(rew.shape[0])]

    def close(self):
        pass


class TrajectoryBuffer:
    """[T+1, N] observation / mask / agent slots and [T, N] action / reward / flag slots in HBM.
    Slot t+1 holds what the policy sees after step t (the reset observation when step t ended the game);
    `final_obs[t]` / `final_mask[t]` hold the terminal observation of step t (Tianshou's obs_next)."""

    def __init__(self, horizon: int, num_envs: int, device="cuda", keep_final: bool = True):
        T, n, dev = int(horizon), int(num_envs), torch.device(device)
        self.horizon, self.num_envs = T, n
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        pad = -(-n // 16) * 16                                # keep every [n,...] slot 16-byte aligned
        self.obs = z((T + 1, pad, 3, 3, 13), torch.int8)[:, :n]
        self.mask = z((T + 1, pad, 54), torch.int8)[:, :n]
        self.agent_id = z((T + 1, n), torch.uint8)
        self.act = z((T, n), torch.uint8)
        self.rew = z((T, n, 2), torch.int8)
        self.terminated = z((T, n), torch.bool)
        self.truncated = z((T, n), torch.bool)
        self.final_obs = z((T, pad, 3, 3, 13), torch.int8)[:, :n] if keep_final else None
        self.final_mask = z((T, pad, 54), torch.int8)[:, :n] if keep_final else None

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in vars(self).values() if torch.is_tensor(t))


class VecCollector:
    """Collect `horizon` lockstep steps of N envs with an arbitrary GPU policy
    `policy(obs[N,3,3,13], mask[N,54], agent_id[N]) -> actions[N]` (example_tianshou_DQN.py:401-409 shape).
    Same-step auto-reset + final observation capture = Tianshou's reset-after-done ordering in one launch."""

    def __init__(self, vec: VecEnv, policy, buffer: TrajectoryBuffer, fused: bool = True):
        assert vec.autoreset == "same_step", "VecCollector drives a same-step auto-reset VecEnv"
        self.vec, self.policy, self.buf = vec, policy, buffer
        # the built-in masked-uniform policy is the sampler of the fused rollout kernel: the whole collection
        # (sample -> step -> emit, T times) is then ONE launch writing straight into the buffer slots
        self.fused = bool(fused) and type(policy) is RandomLegalPolicy
        obs, mask, agent = vec.observe()
        buffer.obs[0].copy_(obs); buffer.mask[0].copy_(mask); buffer.agent_id[0].copy_(agent)

    def _collect_fused(self):
        b, v, p = self.buf, self.vec, self.policy
        # slot 0 (what the policy saw before step 0) is emitted by the same launch from the state itself
        ops.rollout_random(v.state, b.horizon, p.seed, p.env_id_base, p.step, b.obs, b.mask, b.rew,
                           b.terminated.view(torch.uint8), b.agent_id, b.act, b.final_obs, b.final_mask, v.stats,
                           v.flags | ops.SLOT_FROM_ZERO | ops.EMIT_INITIAL, p.step_dev)
        v._advance(b.horizon)
        p.advance(b.horizon)
        return b

    def collect(self):
        if self.fused:
            return self._collect_fused()
        b, v = self.buf, self.vec
        for t in range(b.horizon):
            act = self.policy(b.obs[t], b.mask[t], b.agent_id[t])
            b.act[t].copy_(act)
            final = None if b.final_obs is None else (b.final_obs[t], b.final_mask[t])
            v.step(b.act[t], final=final, out=(b.obs[t + 1], b.mask[t + 1]),
                   aux_out=(b.rew[t], b.terminated[t], b.truncated[t], b.agent_id[t + 1]))
        return b

    def roll(self):
        """Make the last slot the first one of the next collection.  (A fused collection re-emits slot 0 itself from
        the state the previous one left, so back-to-back fused collections do not need this copy.)"""
        b = self.buf
        b.obs[0].copy_(b.obs[-1]); b.mask[0].copy_(b.mask[-1]); b.agent_id[0].copy_(b.agent_id[-1])

    def capture(self):
        """Record one collect() + roll() in a CUDA graph: replaying it removes the per-op launch cost that
        dominates policy-in-the-loop collection at moderate N.  The policy must be graph-safe (device-side RNG
        state: torch's generator is; `RandomLegalPolicy(graph_safe_device=...)` keeps its counter on the GPU)."""
        side = torch.cuda.Stream(device=self.vec.device)
        side.wait_stream(torch.cuda.current_stream(self.vec.device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            self.collect(); self.roll()                    # warm-up outside the capture (allocations, lazy init)
            with torch.cuda.graph(graph, stream=side):
                self.collect(); self.roll()
        torch.cuda.current_stream(self.vec.device).wait_stream(side)
        return graph


class RandomLegalPolicy:
    """Uniform over the mask on the GPU (Philox): the vectorised form of example_basic.py:58-61."""

    def __init__(self, seed: int = 0, env_id_base: int = 0, graph_safe_device=None):
        self.seed, self.env_id_base, self.step = int(seed), int(env_id_base), 0
        # graph_safe_device: keep the step counter in device memory (CUDA-graph capturable)
        self.step_dev = None if graph_safe_device is None else torch.zeros(1, dtype=torch.int64, device=graph_safe_device)

    def __call__(self, obs, mask, agent_id=None):
        act = torch.empty(mask.shape[0], dtype=torch.int32, device=mask.device)
        m = mask if mask.dtype in (torch.int8, torch.uint8, torch.bool) and mask.is_contiguous() else mask.to(torch.int8).contiguous()
        ops.sample_legal(m, self.seed, self.env_id_base, self.step, act, self.step_dev)
        self.advance(1)
        return act

    def advance(self, steps: int):
        if self.step_dev is not None:
            self.step_dev += int(steps)
        self.step += int(steps)


class GreedyVecPolicy:
    """Depth-1/2 greedy policy for a whole VecEnv: one warp per board, with the repetition history of
    greedy_policy.py:211-219 kept ON THE GPU as int16 [N, 2, 3] (last three actions per env and per agent).
    The reference shares one history list per agent across everything a policy object is asked;
    a vectorised env has N independent games, so the history is per env here, and `reset_history(done)`
    clears it for finished games."""

    def __init__(self, num_envs: int, depth: int = 2, seed: int = 0, device="cuda"):
        self.depth, self.seed, self.calls = int(depth), int(seed), 0
        self.prev = torch.full((int(num_envs), 2, 3), -1, dtype=torch.int16, device=device)
        self._rows = torch.arange(int(num_envs), device=device)

    def __call__(self, obs, mask, agent_id):
        who = agent_id.long()
        prev3 = self.prev[self._rows, who]                               # [N,3] history of the agent to move
        act = greedy_actions(obs, mask, prev3, depth=self.depth, seed=self.seed, ctr_base=self.calls * obs.shape[0])
        self.prev[self._rows, who] = torch.cat([prev3[:, 1:], act.to(torch.int16)[:, None]], dim=1)
        self.calls += 1
        return act

    def reset_history(self, done):
        self.prev[done] = -1


class RandomAdmissiblePolicy:
    """random_admissible_policy_rllib.py:15-40: `compute_actions(obs_batch) -> (actions, [], {})`."""

    def __init__(self, seed: int = 0, device="cuda"):
        self._sampler, self.device = RandomLegalPolicy(seed), torch.device(device)

    def compute_actions(self, obs_batch, state_batches=None, prev_action_batch=None, prev_reward_batch=None, **kw):
        mask = torch.as_tensor(np.asarray(obs_batch["action_mask"]), device=self.device)
        return [int(a) for a in self._sampler(None, mask).cpu()], [], {}


class GreedyPolicy:
    """greedy_policy_tianshou.py:12-98: `forward(batch) -> Batch(act=np.ndarray)`; depth is honoured here
    (the reference drops it, SURVEY Q12a, and always searches depth 2 -- the default below)."""

    def __init__(self, depth: Optional[int] = 2, device="cuda", **kwargs):
        self.depth = depth
        self.policy = GreedyGobbletPolicy(depth=depth, device=device)

    def forward(self, batch, state=None, input: str = "obs", **kwargs):
        """The whole batch in one kernel launch.  The reference loops the batch through ONE policy object, so its
        3-move repetition history is shared by every element and consulted element by element;
        that order dependence is kept: the search runs on the GPU for all boards, the history rule and the
        numpy fallback are applied in batch order on the host."""
        ob = batch[input]
        obs = torch.as_tensor(np.asarray(ob.obs), device=self.policy.device)
        mask = torch.as_tensor(np.asarray(ob.mask), device=self.policy.device)
        if obs.dim() == 3:
            obs, mask = obs[None], mask[None]
        _, chosen, cand, _ = greedy_actions(obs, mask, None, depth=self.depth, details=True)
        chosen, cand = chosen.cpu().numpy(), cand.cpu().numpy().astype(np.uint64)
        agent = obs[..., 12].reshape(obs.shape[0], -1).amax(1).cpu().numpy()
        acts = []
        for i in range(obs.shape[0]):
            hist = self.policy.prev_actions[int(agent[i])]
            a = int(chosen[i])
            if a < 0 or a in hist[-3:]:                                   # greedy_policy.py:211-217
                a = np.random.choice([k for k in range(54) if (int(cand[i]) >> k) & 1])
            hist.append(a)
            acts.append(a)
        return Batch(act=np.array(acts))

    def learn(self, batch, **kwargs):
        return {}
