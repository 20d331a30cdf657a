"""ctypes view of oracle/gobblet_oracle.c.  TEST INFRASTRUCTURE ONLY -- never imported by the product."""
import ctypes as C

import numpy as np

from . import build as _build

ILLEGAL_PASS = 0x1
AUTORESET = {"off": 0 << 1, "same_step": 1 << 1, "next_step": 2 << 1}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.gbo_draw.restype = C.c_uint32
        _lib.gbo_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        _lib.gbo_pick.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        _lib.gbo_env_sizeof.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


ACTION_SKIP_255 = 0x10


def flags(illegal_mode="terminate", autoreset="same_step", skip255=False):
    return (ILLEGAL_PASS if illegal_mode == "pass" else 0) | AUTORESET[autoreset] | (ACTION_SKIP_255 if skip255 else 0)


# ---- Board-level -------------------------------------------------------------------------------
def is_legal(squares, action, agent_index):
    sq = np.ascontiguousarray(squares, dtype=np.int8)
    return int(lib().gbo_is_legal(_p(sq), C.c_int(int(action)), C.c_int(int(agent_index))))


def play_turn(squares, agent_index, action):
    sq = np.array(squares, dtype=np.int8)
    lib().gbo_play_turn(_p(sq), C.c_int(int(agent_index)), C.c_int(int(action)))
    return sq


def check_for_winner(squares):
    sq = np.ascontiguousarray(squares, dtype=np.int8)
    return int(lib().gbo_check_for_winner(_p(sq)))


def observe(squares, agent, selected):
    sq = np.ascontiguousarray(squares, dtype=np.int8)
    obs = np.zeros(117, np.int8)
    mask = np.zeros(54, np.int8)
    lib().gbo_observe(_p(sq), C.c_int(int(agent)), C.c_int(int(selected)), _p(obs), _p(mask))
    return obs.reshape(3, 3, 13), mask


# ---- vectorised env mirroring the engine's C ABI ---------------------------------------------
class VecOracle:
    def __init__(self, n, illegal_mode="terminate", autoreset="same_step", skip255=False):
        self.n = int(n)
        self.flags = flags(illegal_mode, autoreset, skip255)
        self._buf = np.zeros(self.n * lib().gbo_env_sizeof(), np.uint8)
        self.stats = np.zeros(8, np.int64)
        lib().gbo_vec_reset(_p(self._buf), C.c_int64(self.n))

    def reset(self):
        lib().gbo_vec_reset(_p(self._buf), C.c_int64(self.n))
        return self.observe()

    def observe(self):
        obs = np.zeros((self.n, 3, 3, 13), np.int8)
        mask = np.zeros((self.n, 54), np.int8)
        agent = np.zeros(self.n, np.uint8)
        lib().gbo_vec_observe(_p(self._buf), C.c_int64(self.n), _p(obs), _p(mask), _p(agent))
        return obs, mask, agent

    def set(self, squares, sel):
        sq = np.ascontiguousarray(squares, dtype=np.int8).reshape(self.n, 27)
        s = np.ascontiguousarray(sel, dtype=np.uint8).reshape(self.n)
        lib().gbo_vec_set(_p(self._buf), C.c_int64(self.n), _p(sq), _p(s))

    def squares(self):
        sq = np.zeros((self.n, 27), np.int8)
        lib().gbo_vec_get_squares(_p(self._buf), C.c_int64(self.n), _p(sq))
        return sq

    def step(self, actions, want_final=False):
        a = np.ascontiguousarray(actions, dtype=np.int64).reshape(self.n)
        obs = np.zeros((self.n, 3, 3, 13), np.int8)
        mask = np.zeros((self.n, 54), np.int8)
        rew = np.zeros((self.n, 2), np.int8)
        term = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        agent = np.zeros(self.n, np.uint8)
        fobs = np.zeros((self.n, 3, 3, 13), np.int8) if want_final else None
        fmask = np.zeros((self.n, 54), np.int8) if want_final else None
        lib().gbo_vec_step(_p(self._buf), C.c_int64(self.n), _p(a), C.c_uint32(self.flags), _p(obs),
                           _p(mask), _p(rew), _p(term), _p(trunc), _p(agent), _p(fobs), _p(fmask),
                           _p(self.stats))
        out = (obs, mask, rew, term.astype(bool), trunc.astype(bool), agent)
        return out + (fobs, fmask) if want_final else out

    def rollout_random(self, T, seed=0, env_id_base=0, step_base=0, per_step=True):
        n, T = self.n, int(T)
        obs = np.zeros((T, n, 3, 3, 13), np.int8) if per_step else None
        mask = np.zeros((T, n, 54), np.int8) if per_step else None
        rew = np.zeros((T, n, 2), np.int8)
        term = np.zeros((T, n), np.uint8)
        agent = np.zeros((T, n), np.uint8)
        log = np.zeros((T, n), np.uint8)
        lib().gbo_rollout_random(_p(self._buf), C.c_int64(n), C.c_int32(T), C.c_uint64(seed),
                                 C.c_uint64(env_id_base), C.c_uint64(step_base),
                                 C.c_uint32(self.flags), _p(obs), _p(mask), _p(rew), _p(term),
                                 _p(agent), _p(log), _p(self.stats))
        return dict(obs=obs, mask=mask, rew=rew, terminated=term.astype(bool), agent_id=agent,
                    actions=log)


# ---- sampler -------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().gbo_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def draw(seed, env_id, step, tag=0):
    return int(lib().gbo_draw(int(seed), int(env_id), int(step), int(tag)))


def pick(mask, d):
    m = np.ascontiguousarray(mask, np.int8)
    return int(lib().gbo_pick(_p(m), C.c_int(m.size), C.c_uint32(int(d))))


# ---- greedy ----------------------------------------------------------------------------------------
def greedy(obs, mask, prev3=(-1, -1, -1), depth=2):
    """-> (chosen_before_fallback or -1, candidate list, used_fallback)"""
    o = np.ascontiguousarray(obs, np.int8).reshape(117)
    m = np.ascontiguousarray(mask, np.int8).reshape(54)
    p = np.asarray(prev3, np.int16)
    chosen, ncand, fb = C.c_int(), C.c_int(), C.c_int()
    cand = np.zeros(54, np.int32)
    rc = lib().gbo_greedy(_p(o), _p(m), _p(p), C.c_int(int(depth)), C.byref(chosen), _p(cand),
                          C.byref(ncand), C.byref(fb))
    if rc != 0:
        raise ValueError("empty action mask")
    return chosen.value, cand[: ncand.value].tolist(), bool(fb.value)
