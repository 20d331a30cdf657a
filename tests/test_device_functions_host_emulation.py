"""CPU check of the DEVICE FUNCTIONS in gobblet_rl_b200/csrc/gobblet_core.cuh.

tests/emul compiles the same header with g++ behind a small shim (one warp = 32 lanes in lockstep)
and replays the rollout / step kernel bodies on the host, so the bitboard rules, the sampler, the
perspective swap and the warp staging / bit->byte expansion are compared with the oracle without a
GPU.  This is a test harness, not a product path (the product has no CPU fallback); the -m gpu
tests repeat the comparisons through the real kernels and the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_core.cpp")
OUT = os.path.join(HERE, "emul", "_build", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-o", OUT, SRC])
    return C.CDLL(OUT)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rollout(lib, n, T, seed, flags, base=0, step_base=0, state=None):
    state = np.zeros((n, 2), np.uint64) if state is None else state
    out = dict(obs=np.full((T, n, 3, 3, 13), 77, np.int8), mask=np.full((T, n, 54), 77, np.int8),
               rew=np.zeros((T, n, 2), np.int8), terminated=np.zeros((T, n), np.uint8),
               agent_id=np.zeros((T, n), np.uint8), actions=np.zeros((T, n), np.uint8))
    stats = np.zeros(8, np.int64)
    lib.emul_rollout(_p(state), C.c_int64(n), C.c_int32(T), C.c_uint64(seed), C.c_uint64(base),
                     C.c_uint64(step_base), C.c_uint32(flags), _p(out["obs"]), _p(out["mask"]),
                     _p(out["rew"]), _p(out["terminated"]), _p(out["agent_id"]), _p(out["actions"]), _p(stats))
    out["terminated"] = out["terminated"].astype(bool)
    return out, stats, state


@pytest.mark.parametrize("mode", ["same_step", "off", "next_step"])
@pytest.mark.parametrize("n,T", [(70, 60), (32, 33), (1, 50), (33, 9)])
def test_rollout_matches_oracle(emul, mode, n, T):
    fl = O.flags("terminate", mode)
    got, stats, _ = _rollout(emul, n, T, 5, fl, base=1000, step_base=3)
    v = O.VecOracle(n, "terminate", mode)
    want = v.rollout_random(T, seed=5, env_id_base=1000, step_base=3)
    for k in ("actions", "obs", "mask", "rew", "terminated", "agent_id"):
        assert np.array_equal(got[k], want[k]), k
    assert stats.tolist() == v.stats.tolist()


def test_rollout_resumes_from_packed_state(emul):
    fl = O.flags("terminate", "same_step")
    a, sa, st = _rollout(emul, 40, 21, 9, fl)
    b, sb, _ = _rollout(emul, 40, 30, 9, fl, step_base=21, state=st)
    v = O.VecOracle(40)
    want = v.rollout_random(51, seed=9)
    assert np.array_equal(np.concatenate([a["obs"], b["obs"]]), want["obs"])
    assert np.array_equal(np.concatenate([a["actions"], b["actions"]]), want["actions"])
    assert (sa + sb)[:7].tolist() == v.stats[:7].tolist()


@pytest.mark.parametrize("illegal_mode", ["terminate", "pass"])
@pytest.mark.parametrize("autoreset", ["same_step", "off", "next_step"])
def test_step_with_arbitrary_actions_matches_oracle(emul, illegal_mode, autoreset):
    n, T = 67, 80
    rng = np.random.default_rng(3)
    fl = O.flags(illegal_mode, autoreset)
    v = O.VecOracle(n, illegal_mode, autoreset)
    state = np.zeros((n, 2), np.uint64)
    stats = np.zeros(8, np.int64)
    _, mask, _ = v.reset()
    for t in range(T):
        acts = np.array([rng.choice(np.flatnonzero(m)) for m in mask], np.int64)
        bad = rng.random(n) < 0.1                        # illegal, out-of-range and negative actions
        acts[bad] = rng.integers(-3, 60, bad.sum())
        obs = np.full((n, 3, 3, 13), 77, np.int8); msk = np.full((n, 54), 77, np.int8)
        fobs = np.full((n, 3, 3, 13), 77, np.int8); fmsk = np.full((n, 54), 77, np.int8)
        rew = np.zeros((n, 2), np.int8); term = np.zeros(n, np.uint8); trunc = np.zeros(n, np.uint8)
        agent = np.zeros(n, np.uint8)
        emul.emul_step(_p(state), C.c_int64(n), _p(acts), C.c_uint32(fl), _p(obs), _p(msk), _p(rew), _p(term),
                       _p(trunc), _p(agent), _p(fobs), _p(fmsk), _p(stats))
        w = v.step(acts, want_final=True)
        assert np.array_equal(obs, w[0]) and np.array_equal(msk, w[1]), t
        assert np.array_equal(rew, w[2]) and np.array_equal(term.astype(bool), w[3])
        assert np.array_equal(trunc.astype(bool), w[4]) and np.array_equal(agent, w[5])
        if autoreset == "same_step":
            assert np.array_equal(fobs, w[6]) and np.array_equal(fmsk, w[7])
        mask = w[1]
    assert stats.tolist() == v.stats.tolist()
    assert stats[5] > 0


def test_hypothesis_action_sequences(emul):
    """Property test (the reference lists hypothesis as a dev dependency but never uses it): ANY sequence of
    integers fed as actions -- legal, illegal, out of range -- leaves engine and oracle in lockstep."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(st.lists(st.integers(min_value=-2, max_value=58), min_size=1, max_size=60),
           st.sampled_from(["terminate", "pass"]), st.sampled_from(["same_step", "off", "next_step"]))
    def run(actions, illegal_mode, autoreset):
        fl = O.flags(illegal_mode, autoreset)
        v = O.VecOracle(1, illegal_mode, autoreset)
        state = np.zeros((1, 2), np.uint64)
        stats = np.zeros(8, np.int64)
        for a in actions:
            acts = np.array([a], np.int64)
            obs = np.zeros((1, 3, 3, 13), np.int8); msk = np.zeros((1, 54), np.int8)
            rew = np.zeros((1, 2), np.int8); term = np.zeros(1, np.uint8); trunc = np.zeros(1, np.uint8)
            agent = np.zeros(1, np.uint8)
            emul.emul_step(_p(state), C.c_int64(1), _p(acts), C.c_uint32(fl), _p(obs), _p(msk), _p(rew), _p(term),
                           _p(trunc), _p(agent), None, None, _p(stats))
            w = v.step(acts)
            assert np.array_equal(obs, w[0]) and np.array_equal(msk, w[1]) and np.array_equal(rew, w[2])
            assert bool(term[0]) == bool(w[3][0]) and bool(trunc[0]) == bool(w[4][0]) and agent[0] == w[5][0]
        assert stats.tolist() == v.stats.tolist()

    run()
