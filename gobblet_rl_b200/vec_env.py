"""Vectorised batched entry point: N lockstep Gobblet environments on one GPU.

Mirrors what a Tianshou-style driver does with N copies of the reference env -- `env.step(a);
env.last()` per env (gobblet.py:231-273, :179-215) -- as ONE kernel launch over
torch CUDA tensors, plus the fused random-legal rollout of example_basic.py:50-67.
"""
from typing import Optional

import torch

from . import ops

_ACTION_BYTES = {torch.uint8: 1, torch.int32: 4, torch.int64: 8}
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
        self._plan = None            # data pointers of the persistent buffers (filled on the first plain step)
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
        if (final is None and out is None and aux_out is None and type(actions) is torch.Tensor
                and actions.dtype in _ACTION_BYTES and actions.device == self.state.device and actions.is_contiguous()
                and actions.numel() == self.num_envs and not torch.compiler.is_compiling()):
            # the common call (own buffers, a well-formed CUDA action tensor): pre-marshalled pointers, ONE ctypes call
            # -- at a few thousand envs the Python side of a step costs more than the kernel
            plan = self._plan
            if plan is None:
                plan = self._plan = tuple(t.data_ptr() for t in (self.state, self.obs, self.mask, self.rew, self.terminated,
                                                                 self.truncated, self.agent_id, self.stats))
            st, ob, mk, rw, te, tr, ag, ss = plan
            dev = self.state.device
            with ops._on_device(dev):
                ops._check(ops.LIB.gbl_step(st, actions.data_ptr(), _ACTION_BYTES[actions.dtype], ob, mk, rw, te, tr, ag, None, None,
                                            ss, self.num_envs, self.flags, torch.cuda.current_stream(dev).cuda_stream))
            self._advance(1)
            return self.obs, self.mask, self.rew, self.terminated, self.truncated, self.agent_id
        actions = torch.as_tensor(actions, device=self.device)
        if actions.dtype not in (torch.uint8, torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        obs, mask = (self.obs, self.mask) if out is None else out
        rew, term, trunc, agent = (self.rew, self.terminated, self.truncated, self.agent_id) if aux_out is None else aux_out
        fobs, fmask = (None, None) if final is None else final
        ops.step(self.state, actions.contiguous(), obs, mask, rew, term.view(torch.uint8), trunc.view(torch.uint8),
                 agent, fobs, fmask, self.stats, self.flags)
        self._advance(1)
        return obs, mask, rew, term, trunc, agent

    def _advance(self, steps: int):
        """One more kernel launch that consumed `steps` lockstep steps: host and device counters move together."""
        if self.step_dev is not None:
            self.step_dev += int(steps)
        self.step_count += int(steps)
        self.kernel_launches += 1

    def rollout_random(self, T: int, ring: int = 1, emit: bool = True, per_step: bool = False,
                       log_actions: bool = False, final: bool = False, block_hint: int = 0, no_bulk: bool = False,
                       split: Optional[bool] = None):
        """T fused lockstep steps with uniform random legal actions (example_basic.py:50-67) in ONE launch.

        emit      write the next observation + mask of every step to ring slot (step % ring)
        per_step  also write rew [ring,N,2], terminated [ring,N], agent_id [ring,N]
        final     also write the observation before a same-step reset replaces it (final_obs / final_mask)
        block_hint  threads per block (32/64/128/256, 0 = chosen from N) -- tuning aid
        split     two warps per 32 envs, one emitting the observations, one the masks (None = chosen from N) -- tuning aid
        Returns a dict of the buffers that were requested; `self.stats` accumulates episode statistics."""
        n, dev, out = self.num_envs, self.device, {}
        obs_out = mask_out = rew_out = term_out = agent_out = log = fobs = fmask = None
        if emit:
            obs_out, mask_out = self._ring_buffers(ring)
            out["obs"], out["mask"] = obs_out, mask_out
            if final:
                fobs, fmask = self._ring_buffers(ring, "_final_rings")
                out["final_obs"], out["final_mask"] = fobs, fmask
        if per_step:
            rew_out = torch.zeros((ring, n, 2), dtype=torch.int8, device=dev)
            term_out = torch.zeros((ring, n), dtype=torch.uint8, device=dev)
            agent_out = torch.zeros((ring, n), dtype=torch.uint8, device=dev)
            out["rew"], out["terminated"], out["agent_id"] = rew_out, term_out.view(torch.bool), agent_out
        if log_actions:
            log = torch.zeros((T, n), dtype=torch.uint8, device=dev)
            out["actions"] = log
        hint = ({0: 0, 32: 1, 64: 2, 128: 3, 256: 4}[int(block_hint)] << ops.BLOCK_HINT_SHIFT) | (ops.NO_BULK_STORE_HINT if no_bulk else 0)
        hint |= 0 if split is None else ops.SPLIT_HINT if split else ops.NO_SPLIT_HINT
        self._launch_rollout(T, obs_out, mask_out, rew_out, term_out, agent_out, log, fobs, fmask, self.flags | hint)
        return out

    def rollout_random_into(self, T: int, obs, mask, rew=None, terminated=None, agent_id=None, actions=None,
                            final_obs=None, final_mask=None):
        """The fused rollout writing step t into slot t of CALLER buffers (e.g. the slots of a
        `adapters.TrajectoryBuffer`): obs [T,N,3,3,13] / mask [T,N,54] int8 (slot stride a multiple of 16 bytes),
        rew [T,N,2] int8, terminated [T,N] bool/uint8, agent_id [T,N] uint8, actions [T,N] uint8,
        final_obs / final_mask like obs / mask.  One launch for the whole collection."""
        u8 = lambda t: None if t is None else t.view(torch.uint8)  # noqa: E731
        self._launch_rollout(T, obs, mask, rew, u8(terminated), agent_id, actions, final_obs, final_mask,
                             self.flags | ops.SLOT_FROM_ZERO)

    def _launch_rollout(self, T, obs_out, mask_out, rew_out, term_out, agent_out, log, fobs, fmask, flags):
        ops.rollout_random(self.state, int(T), self.seed, self.env_id_base, self.step_count, obs_out, mask_out,
                           rew_out, term_out, agent_out, log, fobs, fmask, self.stats, flags, self.step_dev)
        self._advance(T)

    def _ring_buffers(self, ring, attr="_rings"):
        key = int(ring)
        cache = getattr(self, attr, None)
        if cache is None or cache[0] != key:
            n = self.num_envs
            pad = -(-n // 16) * 16                      # slot strides must be multiples of 16 bytes
            obs = torch.zeros((key, pad, 3, 3, 13), dtype=torch.int8, device=self.device)[:, :n]
            mask = torch.zeros((key, pad, 54), dtype=torch.int8, device=self.device)[:, :n]
            cache = (key, obs, mask)
            setattr(self, attr, cache)
        return cache[1], cache[2]

    # -- packed wire format (24 B / env, include/gobblet_b200.h): the step for host-side consumers ------------
    def step_packed(self, actions: torch.Tensor, rec: Optional[torch.Tensor] = None,
                    final_rec: Optional[torch.Tensor] = None):
        """`step` emitting int32 [N, 6] records (observation + mask as BITS, rewards and flags in the spare bits)
        instead of the expanded tensors; `ops.host_unpack` / `unpack_records` expand them."""
        actions = torch.as_tensor(actions, device=self.device)
        if actions.dtype not in (torch.uint8, torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        if rec is None:
            if getattr(self, "_rec", None) is None:
                self._rec = torch.zeros((self.num_envs, ops.REC_WORDS), dtype=torch.int32, device=self.device)
            rec = self._rec
        ops.step_packed(self.state, actions.contiguous(), rec, final_rec, self.stats, self.flags)
        self._advance(1)
        return rec

    def observe_packed(self, rec: Optional[torch.Tensor] = None):
        if rec is None:
            rec = torch.zeros((self.num_envs, ops.REC_WORDS), dtype=torch.int32, device=self.device)
        ops.observe_packed(self.state, rec)
        self.kernel_launches += 1
        return rec

    # -- views in the reference's own state layout (board.py:33) ----------------------------------------
    def squares(self):
        """int8 [N,27] signed piece numbers = `env.board.squares` of every env, and agent_selection [N]."""
        sq = torch.zeros((self.num_envs, 27), dtype=torch.int8, device=self.device)
        agent = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        ops.export_squares(self.state, sq, agent)
        return sq, agent

    def set_squares(self, squares, agent=None, validate: bool = True):
        """Load positions in the reference's layout.  Like `Board.is_legal` (board.py:94-95) a piece placed twice is
        refused, and so is a piece on a level that is not its size's: ValueError (validate=False skips the check's
        device->host sync; offending envs are loaded as the empty board either way)."""
        sq = torch.as_tensor(squares, device=self.device).to(torch.int8).reshape(self.num_envs, 27).contiguous()
        ag = None if agent is None else torch.as_tensor(agent, device=self.device).to(torch.uint8).contiguous()
        bad = torch.zeros(1, dtype=torch.int32, device=self.device) if validate else None
        ops.import_squares(self.state, sq, ag, bad)
        if validate and int(bad) > 0:
            raise ValueError(f"set_squares: {int(bad)} position(s) hold a piece twice or on the wrong level "
                             "(board.py:94-95); they were loaded as empty boards")
        return self.observe()

    def stats_dict(self):
        return dict(zip(STAT_NAMES, self.stats.tolist()))

    # -- checkpoint: the state tensor + counters are the whole env state ------------
    def sampler_step(self) -> int:
        """Absolute lockstep step index = the Philox counter of the next rollout step.  With graph_safe=True the
        device counter is the truth (CUDA-graph replays advance it without passing through Python)."""
        if self.step_dev is not None:
            self.step_count = int(self.step_dev)
        return self.step_count

    def state_dict(self):
        return {"state": self.state.clone(), "stats": self.stats.clone(), "step_count": self.sampler_step(),
                "seed": self.seed, "env_id_base": self.env_id_base}

    def load_state_dict(self, sd):
        self.state.copy_(sd["state"])
        self.stats.copy_(sd["stats"])
        self.step_count, self.seed, self.env_id_base = int(sd["step_count"]), sd["seed"], sd["env_id_base"]
        if self.step_dev is not None:
            self.step_dev.fill_(self.step_count)


def unpack_records(rec):
    """Reference-shaped views of packed records WITHOUT the library (numpy only; for consumers that keep the
    wire format and expand lazily, and for tests): rec int32/uint32 [n, 6] host tensor or array ->
    (obs int8 [n,3,3,13], mask int8 [n,54], rew int8 [n,2], terminated bool [n], truncated bool [n], agent_id uint8 [n])."""
    import numpy as np

    r = np.ascontiguousarray(rec.numpy() if torch.is_tensor(rec) else rec).view(np.uint32).reshape(-1, ops.REC_WORDS)
    bits = np.unpackbits(r.view(np.uint8).reshape(-1, 24), axis=1, bitorder="little")
    f = r[:, 3] >> 21
    rew = np.stack([(f & 3).astype(np.int8) - 1, ((f >> 2) & 3).astype(np.int8) - 1], axis=1)
    return (bits[:, :117].astype(np.int8).reshape(-1, 3, 3, 13), bits[:, 128:182].astype(np.int8), rew,
            ((f >> 4) & 1).astype(bool), ((f >> 5) & 1).astype(bool), ((f >> 6) & 1).astype(np.uint8))


class HostVecEnv:
    """The same step through HOST buffers: pinned host actions in, pinned host obs/mask/rew/flags out -- what a
    CPU-side PettingZoo / Tianshou driver of N envs sees (gobblet.py:179-215 per env).  This is the end-to-end path
    `bench.py` reports as `e2e`.

    wire="packed" (default): the step kernel emits 24-byte records (observation and mask as bits), the records
        cross PCIe in `chunks` stream-pipelined pieces, and the library's host thread pool expands each chunk into
        the reference-shaped int8 arrays as soon as its copy has landed (`gbl_host_unpack_chunked`).  7.3x fewer
        PCIe bytes; the bound becomes host-memory write bandwidth.
    wire="dense": the expanded tensors themselves are copied (176 B/env over PCIe) -- the round-1 path, kept for
        the side-by-side measurement.
    expand=False (packed wire only): `step` returns the pinned records int32 [N,6] for consumers that eat bits
        (`unpack_records` documents the layout).
    blocking_events=True makes the calling thread sleep instead of spin while it waits for a chunk (frees a core, but
        the wake-up costs about 150 us per step on the B200 boxes: measured slower, so off by default)."""

    def __init__(self, num_envs, device="cuda", chunks=None, wire="packed", expand=True, host_threads=0,
                 blocking_events=False, **kw):
        assert wire in ("packed", "dense")
        self.device = torch.device(device)
        self.num_envs = n = int(num_envs)
        self.wire, self.expand, self.host_threads = wire, bool(expand), int(host_threads)
        # chunk boundaries on multiples of 1024 envs (the expander's work item: every chunk starts 64-byte aligned)
        if chunks is None and wire == "packed":
            # small first chunks (the host cores start expanding a few microseconds after the kernel), large later
            # ones (few API calls): 1/64, 1/32, 1/16, 1/8, 1/4, 1/2, 3/4, 1 of the envs
            fr = (1 / 64, 1 / 32, 1 / 16, 1 / 8, 1 / 4, 1 / 2, 3 / 4)
            bounds = sorted({0, *(min(n, -(-int(n * f) // 1024) * 1024) for f in fr)} - {n}) + [n]
        else:
            chunks = max(1, min(int(2 if chunks is None else chunks), n // 1024))
            bounds = [min(n, -(-n * i // chunks // 1024) * 1024) for i in range(chunks)] + [n]
        self.parts = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        self.env = VecEnv(n, device=device, **kw)
        self.envs = [self.env]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.parts]
        self.events = [torch.cuda.Event(blocking=bool(blocking_events)) for _ in self.parts]
        pin = dict(pin_memory=True)
        self.h_actions = torch.zeros(n, dtype=torch.uint8, **pin)
        self.h_obs = torch.zeros((n, 3, 3, 13), dtype=torch.int8, **pin)
        self.h_mask = torch.zeros((n, 54), dtype=torch.int8, **pin)
        self.h_rew = torch.zeros((n, 2), dtype=torch.int8, **pin)
        self.h_term = torch.zeros(n, dtype=torch.bool, **pin)
        self.h_trunc = torch.zeros(n, dtype=torch.bool, **pin)
        self.h_agent = torch.zeros(n, dtype=torch.uint8, **pin)
        self.d_actions = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self.h2d_bytes_per_step = n
        self.host_bytes_per_step = n * (117 + 54 + 2 + 1 + 1 + 1)      # what the consumer finally reads
        if wire == "packed":
            self.h_rec = torch.zeros((n, ops.REC_WORDS), dtype=torch.int32, **pin)
            self.d_rec = torch.zeros((n, ops.REC_WORDS), dtype=torch.int32, device=self.device)
            self.d2h_bytes_per_step = n * 4 * ops.REC_WORDS
            outs = (self.h_obs, self.h_mask, self.h_rew, self.h_term.view(torch.uint8), self.h_trunc.view(torch.uint8),
                    self.h_agent) if self.expand else None
            self.plan = ops.HostStepPlan(self.env.state, self.d_actions, self.d_rec, self.h_rec, [b for _, b in self.parts],
                                         self.streams[0], self.events, outs, self.env.stats, self.env.flags, self.host_threads)
        else:
            self.d2h_bytes_per_step = self.host_bytes_per_step
        torch.cuda.synchronize(self.device)

    def _outputs(self):
        return self.h_obs, self.h_mask, self.h_rew, self.h_term, self.h_trunc, self.h_agent

    def reset(self):
        obs, mask, agent = self.env.reset()
        self.h_obs.copy_(obs); self.h_mask.copy_(mask); self.h_agent.copy_(agent)
        self.h_rew.zero_(); self.h_term.zero_(); self.h_trunc.zero_()
        torch.cuda.synchronize(self.device)
        return self.h_obs, self.h_mask, self.h_agent

    def step(self, actions_host: torch.Tensor):
        """actions_host: pinned uint8 [N].  Returns pinned host tensors (obs, mask, rew, terminated, truncated,
        agent_id) -- or the pinned records int32 [N,6] when expand=False; everything has landed on return."""
        e = self.env
        if self.wire == "packed":
            self.plan.run(actions_host)            # ONE C-ABI call: gbl_step_host (copies, kernel, expansion)
            e.step_count += 1
            e.kernel_launches += 1
            return self._outputs() if self.expand else self.h_rec
        cur = torch.cuda.current_stream(self.device)
        u8 = lambda t: t.view(torch.uint8)  # noqa: E731
        for (a, b), s in zip(self.parts, self.streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                da = self.d_actions[a:b]
                da.copy_(actions_host[a:b], non_blocking=True)
                ops.step(e.state[a:b], da, e.obs[a:b], e.mask[a:b], e.rew[a:b], u8(e.terminated)[a:b], u8(e.truncated)[a:b],
                         e.agent_id[a:b], None, None, e.stats, e.flags)
                self.h_obs[a:b].copy_(e.obs[a:b], non_blocking=True)
                self.h_mask[a:b].copy_(e.mask[a:b], non_blocking=True)
                self.h_rew[a:b].copy_(e.rew[a:b], non_blocking=True)
                self.h_term[a:b].copy_(e.terminated[a:b], non_blocking=True)
                self.h_trunc[a:b].copy_(e.truncated[a:b], non_blocking=True)
                self.h_agent[a:b].copy_(e.agent_id[a:b], non_blocking=True)
        for s in self.streams:
            s.synchronize()
        e.step_count += 1
        e.kernel_launches += len(self.parts)
        return self._outputs()

    @property
    def kernel_launches(self):
        return self.env.kernel_launches
