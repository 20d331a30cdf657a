"""CPU-side checks: the C-ABI library loads and exports every symbol include/gobblet_b200.h declares,
the product never touches oracle/, the AEC plumbing behaves, and the N>1 path (sharding + the single
statistics all-reduce) works over gloo with world_size 2.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gobblet_rl_b200 import ops
    header = open(os.path.join(REPO, "include", "gobblet_b200.h")).read()
    declared = sorted(set(re.findall(r"GBL_API\s+(?:const\s+)?\w+\s*\*?\s*(gbl_\w+)\s*\(", header)))
    assert len(declared) == 21 and set(declared) == set(ops.EXPORTED_SYMBOLS)
    lib = C.CDLL(ops.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gbl_abi_version() == 3
    # the shipped SASS is sm_100a only, built from hand-written kernels (no PTX JIT, no other arch)
    out = subprocess.run(["cuobjdump", "-lelf", ops.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_argument_validation_needs_no_gpu():
    from gobblet_rl_b200 import ops
    lib = ops.LIB
    assert lib.gbl_reset(None, 5, None) == -1 and b"gbl_reset" in lib.gbl_last_error()
    assert lib.gbl_reset(None, 0, None) == 0                                   # empty batch is a no-op
    assert lib.gbl_observe(None, None, None, None, -1, None) == -1
    assert lib.gbl_greedy(None, None, None, 3, 0, 0, None, None, None, None, 1, None) == -1
    with pytest.raises(ops.GobbletError):
        ops.observe(torch.zeros((4, 2), dtype=torch.int64), torch.zeros((4, 3, 3, 13), dtype=torch.int8),
                    torch.zeros((4, 54), dtype=torch.int8), None)              # CPU tensors: no fallback


def _random_records(n, rng):
    rec = rng.integers(0, 2**32, (n, 6), dtype=np.uint64).astype(np.uint32)
    rec[:, 3] &= np.uint32(0x0FFFFFFF)                       # spare bits 124-127 are zero on the wire
    f = (rec[:, 3] >> 21) & 0xF
    bad = ((f & 3) == 3) | (((f >> 2) & 3) == 3)             # rewards are -1 / 0 / +1 => fields 0 / 1 / 2
    rec[bad, 3] &= np.uint32(~(0xF << 21) & 0xFFFFFFFF)
    rec[:, 5] &= np.uint32((1 << 22) - 1)                    # mask bits 54-63 are zero
    return rec


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 1024, 70001])
def test_host_unpack_matches_the_wire_format_definition(n):
    """gbl_host_unpack (thread pool; AVX-512 staged / direct stores, whatever the thread count) == the numpy
    statement of the record layout in include/gobblet_b200.h (`vec_env.unpack_records`).  Host only: no GPU."""
    from gobblet_rl_b200 import ops
    from gobblet_rl_b200.vec_env import unpack_records
    rec = _random_records(n, np.random.default_rng(n))
    want = unpack_records(rec)
    t = torch.from_numpy(rec.view(np.int32))
    for mode in (0, 1):
        assert ops.LIB.gbl_host_set_store_mode(mode) == 0
        for threads in (1, 3, 0):
            obs = torch.full((n, 3, 3, 13), 7, dtype=torch.int8); mask = torch.full((n, 54), 7, dtype=torch.int8)
            rew = torch.full((n, 2), 7, dtype=torch.int8)
            term, trunc, agent = (torch.full((n,), 7, dtype=torch.uint8) for _ in range(3))
            ops.host_unpack(t, obs, mask, rew, term, trunc, agent, threads=threads)
            got = (obs, mask, rew, term.bool(), trunc.bool(), agent)
            for g, w in zip(got, want):
                assert np.array_equal(g.numpy(), w), (n, mode, threads)
            # chunked form without events (chunks are ready immediately) and with optional outputs left out
            obs.fill_(7); mask.fill_(7)
            ends = sorted({min(n, max(2, n // 3 // 2 * 2)), n})
            ops.host_unpack(t, obs, mask, threads=threads, chunk_end=ends)
            assert np.array_equal(obs.numpy(), want[0]) and np.array_equal(mask.numpy(), want[1])
    ops.LIB.gbl_host_set_store_mode(0)


def test_host_unpack_table_path_in_a_fresh_process(tmp_path):
    """The portable (no AVX-512) expander is selected at first use: force it with GBL_HOST_SIMD=0 in a child."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np, torch\n"
            "from gobblet_rl_b200 import ops\n"
            "from gobblet_rl_b200.vec_env import unpack_records\n"
            "assert ops.host_simd() == 'table'\n"
            "rec = np.random.default_rng(1).integers(0, 2**32, (5000, 6), dtype=np.uint64).astype(np.uint32)\n"
            "rec[:, 3] &= np.uint32(0x001FFFFF | (5 << 21) | (7 << 25)); rec[:, 5] &= np.uint32((1 << 22) - 1)\n"
            "obs = torch.zeros((5000, 3, 3, 13), dtype=torch.int8); mask = torch.zeros((5000, 54), dtype=torch.int8)\n"
            "rew = torch.zeros((5000, 2), dtype=torch.int8); f = [torch.zeros(5000, dtype=torch.uint8) for _ in range(3)]\n"
            "ops.host_unpack(torch.from_numpy(rec.view(np.int32)), obs, mask, rew, *f, threads=2)\n"
            "w = unpack_records(rec)\n"
            "assert all(np.array_equal(g.numpy().astype(x.dtype), x) for g, x in zip((obs, mask, rew, *f), w))\n"
            "print('OK')\n") % REPO
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GBL_HOST_SIMD="0"))
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_product_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "gobblet_rl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|gobblet_oracle|reference_loader", text, re.M):
                    bad.append(f)
    assert not bad, bad


def test_aec_plumbing_with_a_scripted_env():
    """_aec (used when PettingZoo is absent): last(), reward accumulation, dead steps, the three wrappers."""
    from gobblet_rl_b200 import _aec, _spaces

    class Scripted(_aec.AECEnv):
        metadata = {"name": "scripted"}

        def __init__(self):
            super().__init__()
            self.possible_agents = ["a", "b"]
            self.action_spaces = {x: _spaces.Discrete(3) for x in self.possible_agents}
            self.observation_spaces = {x: _spaces.Dict({"action_mask": _spaces.Box(0, 1, (3,), np.int8)})
                                       for x in self.possible_agents}
            self._sel = _aec.agent_selector(self.possible_agents)

        def reset(self, seed=None, return_info=False, options=None):
            self.agents = self.possible_agents[:]
            self.rewards = {x: 0 for x in self.agents}
            self._cumulative_rewards = {x: 0 for x in self.agents}
            self.terminations = {x: False for x in self.agents}
            self.truncations = {x: False for x in self.agents}
            self.infos = {x: {} for x in self.agents}
            self._sel.reinit(self.agents)
            self.agent_selection = self._sel.reset()
            self.count = 0

        def observe(self, agent):
            return {"action_mask": np.array([1, 1, 0], np.int8)}

        def step(self, action):
            if self.terminations[self.agent_selection]:
                return self._was_dead_step(action)
            self.count += 1
            nxt = self._sel.next()
            if action == 1:
                self.rewards = {"a": 1, "b": -1}
                self.terminations = {x: True for x in self.agents}
            self._cumulative_rewards[self.agent_selection] = 0
            self.agent_selection = nxt
            self._accumulate_rewards()

    env = _aec.OrderEnforcingWrapper(_aec.AssertOutOfBoundsWrapper(_aec.TerminateIllegalWrapper(Scripted(), -1)))
    with pytest.raises(AssertionError):
        env.step(0)
    env.reset()
    assert env.agent_selection == "a" and env.last()[1:4] == (0, False, False)
    env.step(0)
    assert env.agent_selection == "b"
    env.step(1)                                            # b ends the game: a +1, b -1
    assert [env.last()[1]] == [1] and env.agent_selection == "a" and all(env.terminations.values())
    env.step(None)
    assert env.agents == ["b"] and env.last()[1] == -1
    env.step(None)
    assert env.agents == [] and list(env.agent_iter()) == []
    env.reset()
    env.step(2)                                            # masked out -> illegal-move termination
    assert env.rewards == {"a": -1.0, "b": 0} and all(env.truncations.values()) and env.agent_selection == "a"
    with pytest.raises(ValueError):
        env.step(0)
    env.reset()
    with pytest.raises(AssertionError):
        env.step(7)                                        # out of bounds
    s = _spaces.Discrete(54)
    assert s.contains(np.int64(5)) and s.contains(np.array(53)) and not s.contains(54) and not s.contains(1.0)


def test_shard_ranges_cover_all_envs():
    from gobblet_rl_b200.sharding import shard_range
    for total in (0, 1, 7, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {repo!r})
import numpy as np, torch, torch.distributed as dist
from gobblet_rl_b200.sharding import init_from_env, shard_range, all_reduce_stats
from oracle import oracle as O            # the CPU stand-in for the per-rank engine in this host-logic test
rank, local, world = init_from_env(backend="gloo")
total, T, seed = 300, 40, 17
lo, hi = shard_range(total, rank, world)
v = O.VecOracle(hi - lo)
out = v.rollout_random(T, seed=seed, env_id_base=lo, per_step=False)
merged = all_reduce_stats(torch.from_numpy(v.stats.copy()))
if rank == 0:
    whole = O.VecOracle(total)
    w = whole.rollout_random(T, seed=seed, per_step=False)
    assert merged.tolist() == whole.stats.tolist(), (merged.tolist(), whole.stats.tolist())
    assert np.array_equal(out["actions"], w["actions"][:, lo:hi])
    print("OK", merged.tolist())
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_statistics_all_reduce(tmp_path):
    """world_size 2 over gloo: shards keyed by GLOBAL env id replay the same games as one rank, and the
    single SUM/MAX all-reduce reproduces the whole-job statistics."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(repo=REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "OK" in res.stdout


# ---- the wrapper stack, pinned by a literal transition table ----------------------------------
# PettingZoo 1.22.3 is absent from the image, so the product's _aec.py and the test-side stand-ins (a rule-table
# model, oracle/standins) are two independent statements of its published behaviour.  The table
# below is a third one: hand-written expected `last()` tuples and bookkeeping for scripted calls, derived from
# App. D, not from either implementation.  Both implementations must reproduce every row.
# Scripted env: agents a, b; action 1 ends the game (a +1, b -1); action 2 is masked out (mask = [1, 1, 0]).
_AEC_TABLE = [
    # (call, argument)                 -> expected (agent_selection, agents, cumulative reward of the selection, terminated, truncated)
    (("reset", None),                  ("a", ["a", "b"], 0, False, False)),
    (("step", 0),                      ("b", ["a", "b"], 0, False, False)),
    (("step", 1),                      ("a", ["a", "b"], 1, True, False)),     # b ended the game: a is served first, +1
    (("step", None),                   ("b", ["b"], -1, True, False)),         # dead step removes a; b sees its -1
    (("step", None),                   (None, [], None, None, None)),          # cycle over
    (("reset", None),                  ("a", ["a", "b"], 0, False, False)),
    (("step", 2),                      ("a", ["a", "b"], -1.0, True, True)),   # masked action: illegal-move termination, mover -1
    (("step_raises", (0, ValueError)), ("a", ["a", "b"], -1.0, True, True)),   # a dead agent may only step(None)
    (("step", None),                   ("b", ["b"], 0, True, True)),           # other agent: 0, also terminated AND truncated
    (("step", None),                   (None, [], None, None, None)),
    (("reset", None),                  ("a", ["a", "b"], 0, False, False)),
    (("step", 0),                      ("b", ["a", "b"], 0, False, False)),
    (("step", 2),                      ("a", ["a", "b"], 0, True, True)),      # illegal by b: dead agents are served in `agents` order
    (("step", None),                   ("b", ["b"], -1.0, True, True)),
    (("step_raises", (7, AssertionError)), ("b", ["b"], -1.0, True, True)),    # out of bounds (and not None) for a dead agent
    (("step", None),                   (None, [], None, None, None)),
]


def _scripted_env(base, spaces, selector):
    class Scripted(base):
        metadata = {"name": "scripted"}

        def __init__(self):
            super().__init__()
            self.possible_agents = ["a", "b"]
            self.action_spaces = {x: spaces.Discrete(3) for x in self.possible_agents}
            self.observation_spaces = {x: spaces.Dict({"action_mask": spaces.Box(0, 1, (3,), np.int8)})
                                       for x in self.possible_agents}
            self._sel = selector(self.possible_agents)

        def reset(self, seed=None, return_info=False, options=None):
            self.agents = self.possible_agents[:]
            self.rewards = {x: 0 for x in self.agents}
            self._cumulative_rewards = {x: 0 for x in self.agents}
            self.terminations = {x: False for x in self.agents}
            self.truncations = {x: False for x in self.agents}
            self.infos = {x: {} for x in self.agents}
            self._sel.reinit(self.agents)
            self.agent_selection = self._sel.reset()

        def observe(self, agent):
            return {"action_mask": np.array([1, 1, 0], np.int8)}

        def step(self, action):
            if self.terminations[self.agent_selection]:
                return self._was_dead_step(action)
            nxt = self._sel.next()
            if action == 1:
                mover = self.agent_selection
                self.rewards = {x: (-1 if x == mover else 1) for x in self.agents}
                self.terminations = {x: True for x in self.agents}
            self._cumulative_rewards[self.agent_selection] = 0
            self.agent_selection = nxt
            self._accumulate_rewards()

    return Scripted()


@pytest.mark.parametrize("impl", ["product_aec", "standin_model"])
def test_aec_semantics_table(impl):
    if impl == "product_aec":
        from gobblet_rl_b200 import _aec as A, _spaces as S
        base, sel, W = A.AECEnv, A.agent_selector, A
    else:
        import importlib
        from oracle import reference_loader as RL
        if RL.STANDINS not in sys.path:                        # left in place, exactly as reference_loader does
            sys.path.insert(0, RL.STANDINS)
        pz = importlib.import_module("pettingzoo")
        W = importlib.import_module("pettingzoo.utils.wrappers")
        sel = importlib.import_module("pettingzoo.utils").agent_selector
        S = importlib.import_module("gymnasium.spaces")
        assert "standin" in pz.__version__
        base = pz.AECEnv
    env = W.OrderEnforcingWrapper(W.AssertOutOfBoundsWrapper(W.TerminateIllegalWrapper(_scripted_env(base, S, sel), -1)))
    with pytest.raises(AssertionError):
        env.step(0)                                            # OrderEnforcingWrapper: reset first
    for row, ((call, arg), want) in enumerate(_AEC_TABLE):
        if call == "reset":
            env.reset()
        elif call == "step":
            env.step(arg)
        else:
            with pytest.raises(arg[1]):
                env.step(arg[0])
        sel_want, agents_want, rew_want, term_want, trunc_want = want
        assert list(env.agents) == agents_want, (impl, row)
        assert next(iter(env.agent_iter()), None) == (sel_want if agents_want else None), (impl, row)
        if agents_want:
            assert env.agent_selection == sel_want, (impl, row)
            obs, rew, term, trunc, info = env.last()
            assert (rew, term, trunc) == (rew_want, term_want, trunc_want) and type(rew) is type(rew_want), (impl, row, rew, term, trunc)
            assert obs["action_mask"].tolist() == [1, 1, 0] and info == {}
