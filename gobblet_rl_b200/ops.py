"""ctypes binding of libgobblet_b200.so (C ABI: include/gobblet_b200.h) + torch custom ops.

PyTorch is plumbing here: it owns device memory and streams; every computation happens in the
sm_100a kernels behind the C ABI.  There is NO fallback: if the library cannot be loaded (or built
with nvcc) importing this module raises.
"""
import ctypes as C
import functools
from typing import Optional

import torch

from .csrc import build as _build

OBS_BYTES, MASK_BYTES, STATE_BYTES, NUM_ACTIONS = 117, 54, 16, 54
ILLEGAL_TERMINATE, ILLEGAL_PASS = 0x0, 0x1
AUTORESET_OFF, AUTORESET_SAME_STEP, AUTORESET_NEXT_STEP = 0 << 1, 1 << 1, 2 << 1
STORE_DEFAULT_POLICY = 0x8
ACTION_SKIP_255 = 0x10
SLOT_FROM_ZERO = 0x20
EMIT_INITIAL = 0x40
BLOCK_HINT_SHIFT = 12
NO_BULK_STORE_HINT = 0x8000
SPLIT_HINT, NO_SPLIT_HINT = 0x10000, 0x20000
REC_WORDS = 6                      # packed wire format: 6 x u32 = 24 bytes per env (include/gobblet_b200.h)
ABI_VERSION = 3

_ILLEGAL = {"terminate": ILLEGAL_TERMINATE, "pass": ILLEGAL_PASS}
_AUTORESET = {"off": AUTORESET_OFF, "same_step": AUTORESET_SAME_STEP, "next_step": AUTORESET_NEXT_STEP}


def make_flags(illegal_mode="terminate", autoreset="same_step", streaming_stores=True, skip255=False):
    return (_ILLEGAL[illegal_mode] | _AUTORESET[autoreset] | (0 if streaming_stores else STORE_DEFAULT_POLICY)
            | (ACTION_SKIP_255 if skip255 else 0))


def _load():
    try:
        path = _build.build()
    except Exception as exc:  # no nvcc and no prebuilt library: fail loudly, never fall back
        raise RuntimeError(
            "gobblet_rl_b200: libgobblet_b200.so is missing and could not be built with nvcc "
            f"({exc}). The engine has no CPU fallback.") from exc
    lib = C.CDLL(path)
    vp, i64, i32, u64, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32
    sigs = {
        "gbl_abi_version": (C.c_int, []),
        "gbl_last_error": (C.c_char_p, []),
        "gbl_reset": (C.c_int, [vp, i64, vp]),
        "gbl_reset_masked": (C.c_int, [vp, vp, i64, vp]),
        "gbl_observe": (C.c_int, [vp, vp, vp, vp, i64, vp]),
        "gbl_step": (C.c_int, [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, u32, vp]),
        "gbl_rollout_random": (C.c_int, [vp, i64, i32, u64, u64, u64, vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, u32, vp]),
        "gbl_step_packed": (C.c_int, [vp, vp, i32, vp, vp, vp, i64, u32, vp]),
        "gbl_observe_packed": (C.c_int, [vp, vp, i64, vp]),
        "gbl_host_unpack": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, i32]),
        "gbl_host_unpack_chunked": (C.c_int, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "gbl_step_host": (C.c_int, [vp, vp, i64, u32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "gbl_host_fill": (C.c_int, [vp, i64, i32, i32]),
        "gbl_host_threads": (C.c_int, [i32]),
        "gbl_host_last_timing": (C.c_int, [vp, i32]),
        "gbl_host_simd": (C.c_int, []),
        "gbl_host_set_store_mode": (C.c_int, [i32]),
        "gbl_sample_legal": (C.c_int, [vp, u64, u64, u64, vp, vp, i64, vp]),
        "gbl_greedy": (C.c_int, [vp, vp, vp, i32, u64, u64, vp, vp, vp, vp, i64, vp]),
        "gbl_export_squares": (C.c_int, [vp, vp, vp, i64, vp]),
        "gbl_import_squares": (C.c_int, [vp, vp, vp, i64, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.gbl_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libgobblet_b200.so ABI {lib.gbl_abi_version()} != expected {ABI_VERSION}")
    return lib, path


LIB, LIB_PATH = _load()
EXPORTED_SYMBOLS = ("gbl_abi_version", "gbl_last_error", "gbl_reset", "gbl_reset_masked", "gbl_observe",
                    "gbl_step", "gbl_rollout_random", "gbl_sample_legal", "gbl_greedy", "gbl_export_squares",
                    "gbl_import_squares", "gbl_step_packed", "gbl_observe_packed", "gbl_host_unpack",
                    "gbl_host_unpack_chunked", "gbl_step_host", "gbl_host_last_timing", "gbl_host_fill", "gbl_host_threads", "gbl_host_simd",
                    "gbl_host_set_store_mode")


class GobbletError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise GobbletError(f"libgobblet_b200 error {rc}: {LIB.gbl_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class _on_device:
    """Make `dev` the current CUDA device around a launch; costs nothing when it already is (the usual case --
    torch.cuda.device() pays two cudaSetDevice round trips per call)."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.prev)
        return False


def _need_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise GobbletError("gobblet_rl_b200 ops need CUDA tensors (there is no CPU path)")
        if not t.is_contiguous():
            raise GobbletError("gobblet_rl_b200 ops need contiguous tensors")
        if dev is not None and t.device != dev:
            raise GobbletError("all tensors of one call must live on the same device")
        dev = t.device
    return dev


def _need_bytes(name, t: Optional[torch.Tensor], nbytes: int, itemsize: int = 1):
    """Buffers cross the C ABI as raw pointers: refuse tensors that are too small or of the wrong width."""
    if t is None:
        return
    if t.element_size() != itemsize:
        raise GobbletError(f"{name}: expected {itemsize}-byte elements, got {t.dtype}")
    if t.numel() * t.element_size() < nbytes:
        raise GobbletError(f"{name}: needs {nbytes} bytes, tensor holds {t.numel() * t.element_size()}")


def _need_state(state: torch.Tensor):
    if state.dtype != torch.int64 or state.dim() != 2 or state.shape[1] != 2:
        raise GobbletError("state must be an int64 [n, 2] tensor (16 bytes per env)")
    return state.shape[0]


# ---- torch custom ops (schema + fake impls so the engine composes with torch.compile / graphs) ------
def _engine_op(name: str, mutates):
    """Register `fn` as the torch custom op gobblet_b200::<name> (schema with mutated arguments + a fake impl, so a
    caller's step function composes with torch.compile) and return a thin dispatcher: under tracing the custom op
    is recorded, in eager mode the function is called DIRECTLY -- the custom-op dispatcher costs tens of
    microseconds per call, as long as the kernels themselves."""
    def deco(fn):
        op = torch.library.custom_op(f"gobblet_b200::{name}", mutates_args=tuple(mutates))(fn)
        op.register_fake(lambda *a, **k: None)

        @functools.wraps(fn)
        def call(*a, **k):
            if torch.compiler.is_compiling():
                return op(*a, **k)
            return fn(*a, **k)

        call.op = op
        return call
    return deco

@_engine_op("reset", ("state",))
def reset(state: torch.Tensor, which: Optional[torch.Tensor] = None) -> None:
    dev = _need_cuda(state, which)
    _need_bytes("which", which, _need_state(state))
    with _on_device(dev):
        _check(LIB.gbl_reset_masked(_ptr(state), _ptr(which), state.shape[0], _stream(state)))


@_engine_op("observe", ("obs", "mask", "agent_id"))
def observe(state: torch.Tensor, obs: torch.Tensor, mask: torch.Tensor, agent_id: Optional[torch.Tensor]) -> None:
    dev = _need_cuda(state, obs, mask, agent_id)
    n = _need_state(state)
    _need_bytes("obs", obs, n * OBS_BYTES); _need_bytes("mask", mask, n * MASK_BYTES); _need_bytes("agent_id", agent_id, n)
    with _on_device(dev):
        _check(LIB.gbl_observe(_ptr(state), _ptr(obs), _ptr(mask), _ptr(agent_id), state.shape[0], _stream(state)))


@_engine_op("step", ("state", "obs", "mask", "rew", "terminated", "truncated",
                                                             "agent_id", "final_obs", "final_mask", "stats"))
def step(state: torch.Tensor, actions: torch.Tensor, obs: torch.Tensor, mask: torch.Tensor,
         rew: Optional[torch.Tensor], terminated: Optional[torch.Tensor], truncated: Optional[torch.Tensor],
         agent_id: Optional[torch.Tensor], final_obs: Optional[torch.Tensor], final_mask: Optional[torch.Tensor],
         stats: Optional[torch.Tensor], flags: int) -> None:
    dev = _need_cuda(state, actions, obs, mask, rew, terminated, truncated, agent_id, final_obs, final_mask, stats)
    n = _need_state(state)
    if actions.dtype not in (torch.uint8, torch.int32, torch.int64) or actions.numel() != n:
        raise GobbletError("actions must be uint8 / int32 / int64 with one entry per env")
    for name, t, per_env in (("obs", obs, OBS_BYTES), ("mask", mask, MASK_BYTES), ("rew", rew, 2), ("terminated", terminated, 1),
                             ("truncated", truncated, 1), ("agent_id", agent_id, 1), ("final_obs", final_obs, OBS_BYTES),
                             ("final_mask", final_mask, MASK_BYTES)):
        _need_bytes(name, t, n * per_env)
    _need_bytes("stats", stats, 64, 8)
    with _on_device(dev):
        _check(LIB.gbl_step(_ptr(state), _ptr(actions), actions.element_size(), _ptr(obs), _ptr(mask), _ptr(rew),
                            _ptr(terminated), _ptr(truncated), _ptr(agent_id), _ptr(final_obs), _ptr(final_mask),
                            _ptr(stats), state.shape[0], flags, _stream(state)))


@_engine_op("rollout_random", ("state", "obs_out", "mask_out", "rew_out", "term_out", "agent_out",
                                       "action_log", "stats", "final_obs_out", "final_mask_out"))
def rollout_random(state: torch.Tensor, T: int, seed: int, env_id_base: int, step_base: int,
                   obs_out: Optional[torch.Tensor], mask_out: Optional[torch.Tensor],
                   rew_out: Optional[torch.Tensor], term_out: Optional[torch.Tensor],
                   agent_out: Optional[torch.Tensor], action_log: Optional[torch.Tensor],
                   final_obs_out: Optional[torch.Tensor], final_mask_out: Optional[torch.Tensor],
                   stats: Optional[torch.Tensor], flags: int, step_dev: Optional[torch.Tensor] = None) -> None:
    """obs_out [ring, n, 3, 3, 13] / mask_out [ring, n, 54] (possibly views of padded slots).
    step_dev: optional int64[1] CUDA tensor holding the absolute step (replaces step_base; graph-capturable).
    final_obs_out / final_mask_out: same shape and slot strides as obs_out / mask_out (terminal observations)."""
    dev = _need_cuda(state, rew_out, term_out, agent_out, action_log, stats, step_dev)
    n = state.shape[0]
    ring, so, sm = 1, 0, 0
    if obs_out is not None:
        if not (obs_out.is_cuda and mask_out is not None and mask_out.is_cuda):
            raise GobbletError("obs_out / mask_out must both be CUDA tensors")
        ring = obs_out.shape[0]
        so, sm = obs_out.stride(0), mask_out.stride(0)          # int8 => element stride == bytes
        if obs_out[0].numel() != n * OBS_BYTES or not obs_out[0].is_contiguous() or not mask_out[0].is_contiguous():
            raise GobbletError("each ring slot must be a contiguous [n,3,3,13] / [n,54] int8 block")
    _need_state(state)
    _need_bytes("stats", stats, 64, 8)
    _need_bytes("step_dev", step_dev, 8, 8)
    if obs_out is not None and (obs_out.element_size() != 1 or mask_out.element_size() != 1
                                or mask_out.shape[0] != ring or mask_out[0].numel() != n * MASK_BYTES):
        raise GobbletError("obs_out / mask_out must be int8 [ring, n, 3, 3, 13] / [ring, n, 54]")
    aux = [t for t in (rew_out, term_out, agent_out) if t is not None]
    if obs_out is None and aux:
        ring = aux[0].shape[0]
    initial = 1 if flags & EMIT_INITIAL else 0      # [T+1] observation / agent slots next to [T] reward / flag slots
    for t in aux:
        if t.shape[0] != ring - (0 if t is agent_out else initial):
            raise GobbletError("per-step outputs must share one ring length (that of obs_out when it is given)")
    _need_bytes("rew_out", rew_out, (ring - initial) * n * 2); _need_bytes("term_out", term_out, (ring - initial) * n)
    _need_bytes("agent_out", agent_out, ring * n); _need_bytes("action_log", action_log, T * n)
    if (final_obs_out is None) != (final_mask_out is None):
        raise GobbletError("final_obs_out / final_mask_out go together")
    if final_obs_out is not None:
        if obs_out is None or not (final_obs_out.is_cuda and final_mask_out.is_cuda):
            raise GobbletError("final_obs_out / final_mask_out need obs_out / mask_out and must be CUDA tensors")
        if (final_obs_out.shape[1:] != obs_out.shape[1:] or final_mask_out.shape[1:] != mask_out.shape[1:]
                or final_obs_out.shape[0] != ring - initial or final_mask_out.shape[0] != ring - initial or final_obs_out.stride(0) != so
                or final_mask_out.stride(0) != sm or final_obs_out.element_size() != 1 or final_mask_out.element_size() != 1
                or not final_obs_out[0].is_contiguous() or not final_mask_out[0].is_contiguous()):
            raise GobbletError("final_obs_out / final_mask_out must match obs_out / mask_out in shape and slot stride")
    with _on_device(dev):
        _check(LIB.gbl_rollout_random(_ptr(state), n, T, seed & (2**64 - 1), env_id_base, step_base, _ptr(step_dev), _ptr(obs_out),
                                      _ptr(mask_out), so, sm, ring, _ptr(rew_out), _ptr(term_out), _ptr(agent_out),
                                      _ptr(action_log), _ptr(final_obs_out), _ptr(final_mask_out), _ptr(stats), flags,
                                      _stream(state)))


@_engine_op("step_packed", ("state", "rec", "final_rec", "stats"))
def step_packed(state: torch.Tensor, actions: torch.Tensor, rec: torch.Tensor, final_rec: Optional[torch.Tensor],
                stats: Optional[torch.Tensor], flags: int) -> None:
    """gbl_step in the packed wire format: rec int32 [n, 6] (24 bytes per env, include/gobblet_b200.h)."""
    dev = _need_cuda(state, actions, rec, final_rec, stats)
    n = _need_state(state)
    if actions.dtype not in (torch.uint8, torch.int32, torch.int64) or actions.numel() != n:
        raise GobbletError("actions must be uint8 / int32 / int64 with one entry per env")
    _need_bytes("rec", rec, 24 * n, 4); _need_bytes("final_rec", final_rec, 24 * n, 4); _need_bytes("stats", stats, 64, 8)
    with _on_device(dev):
        _check(LIB.gbl_step_packed(_ptr(state), _ptr(actions), actions.element_size(), _ptr(rec), _ptr(final_rec),
                                   _ptr(stats), n, flags, _stream(state)))


@_engine_op("observe_packed", ("rec",))
def observe_packed(state: torch.Tensor, rec: torch.Tensor) -> None:
    dev = _need_cuda(state, rec)
    n = _need_state(state)
    _need_bytes("rec", rec, 24 * n, 4)
    with _on_device(dev):
        _check(LIB.gbl_observe_packed(_ptr(state), _ptr(rec), n, _stream(state)))


@_engine_op("sample_legal", ("act",))
def sample_legal(mask: torch.Tensor, seed: int, env_id_base: int, step: int, act: torch.Tensor,
                 step_dev: Optional[torch.Tensor] = None) -> None:
    dev = _need_cuda(mask, act, step_dev)
    _need_bytes("mask", mask, act.numel() * MASK_BYTES); _need_bytes("act", act, 4 * act.numel(), 4)
    _need_bytes("step_dev", step_dev, 8, 8)
    with _on_device(dev):
        _check(LIB.gbl_sample_legal(_ptr(mask), seed & (2**64 - 1), env_id_base, step, _ptr(step_dev), _ptr(act),
                                    act.numel(), _stream(mask)))


@_engine_op("greedy", ("act", "chosen", "cand", "used_fallback"))
def greedy(obs: torch.Tensor, mask: torch.Tensor, prev3: Optional[torch.Tensor], depth: int, seed: int,
           ctr_base: int, act: torch.Tensor, chosen: Optional[torch.Tensor], cand: Optional[torch.Tensor],
           used_fallback: Optional[torch.Tensor]) -> None:
    dev = _need_cuda(obs, mask, prev3, act, chosen, cand, used_fallback)
    n = act.numel()
    _need_bytes("obs", obs, n * OBS_BYTES); _need_bytes("mask", mask, n * MASK_BYTES); _need_bytes("prev3", prev3, 6 * n, 2)
    _need_bytes("act", act, 4 * n, 4); _need_bytes("chosen", chosen, 4 * n, 4); _need_bytes("cand", cand, 8 * n, 8)
    _need_bytes("used_fallback", used_fallback, n)
    with _on_device(dev):
        _check(LIB.gbl_greedy(_ptr(obs), _ptr(mask), _ptr(prev3), depth, seed & (2**64 - 1), ctr_base, _ptr(act),
                              _ptr(chosen), _ptr(cand), _ptr(used_fallback), act.numel(), _stream(obs)))


@_engine_op("export_squares", ("squares", "agent"))
def export_squares(state: torch.Tensor, squares: torch.Tensor, agent: Optional[torch.Tensor]) -> None:
    dev = _need_cuda(state, squares, agent)
    n = _need_state(state)
    _need_bytes("squares", squares, 27 * n); _need_bytes("agent", agent, n)
    with _on_device(dev):
        _check(LIB.gbl_export_squares(_ptr(state), _ptr(squares), _ptr(agent), state.shape[0], _stream(state)))


@_engine_op("import_squares", ("state", "invalid_count"))
def import_squares(state: torch.Tensor, squares: torch.Tensor, agent: Optional[torch.Tensor],
                   invalid_count: Optional[torch.Tensor]) -> None:
    """invalid_count: optional int32[1] CUDA tensor, incremented per env whose squares the reference would reject
    (a piece placed twice, board.py:94-95) or could never hold; such envs are loaded as the empty board."""
    dev = _need_cuda(state, squares, agent, invalid_count)
    n = _need_state(state)
    _need_bytes("squares", squares, 27 * n); _need_bytes("agent", agent, n); _need_bytes("invalid_count", invalid_count, 4, 4)
    with _on_device(dev):
        _check(LIB.gbl_import_squares(_ptr(state), _ptr(squares), _ptr(agent), state.shape[0], _ptr(invalid_count),
                                      _stream(state)))




# ---- host side of the packed wire format (plain host memory; no torch custom op: nothing here is traced) ----
def host_unpack(rec, obs, mask, rew=None, terminated=None, truncated=None, agent_id=None, threads: int = 0,
                chunk_end=None, events=None):
    """Expand packed records (host tensor int32 [n, 6]) into reference-shaped HOST tensors obs int8 [n,3,3,13],
    mask int8 [n,54], rew int8 [n,2], terminated / truncated / agent_id (1 byte each) on the library's thread
    pool.  chunk_end (list of env counts) + events (torch.cuda.Event per chunk): envs of chunk c are expanded
    once events[c] has completed, overlapping the expansion with the remaining D2H copies."""
    n = rec.shape[0]
    for name, t, per_env in (("rec", rec, 24), ("obs", obs, OBS_BYTES), ("mask", mask, MASK_BYTES), ("rew", rew, 2),
                             ("terminated", terminated, 1), ("truncated", truncated, 1), ("agent_id", agent_id, 1)):
        if t is None:
            continue
        if t.is_cuda or not t.is_contiguous():
            raise GobbletError(f"host_unpack: {name} must be a contiguous HOST tensor")
        if t.numel() * t.element_size() < n * per_env:
            raise GobbletError(f"host_unpack: {name} too small")
    if chunk_end is None:
        rc = LIB.gbl_host_unpack(_ptr(rec), n, _ptr(obs), _ptr(mask), _ptr(rew), _ptr(terminated), _ptr(truncated),
                                 _ptr(agent_id), threads)
    else:
        k = len(chunk_end)
        ends = (C.c_int64 * k)(*[int(x) for x in chunk_end])
        evs = (C.c_void_p * k)(*[None if e is None else e.cuda_event for e in (events or [None] * k)])
        rc = LIB.gbl_host_unpack_chunked(_ptr(rec), n, k, ends, evs, _ptr(obs), _ptr(mask), _ptr(rew), _ptr(terminated),
                                         _ptr(truncated), _ptr(agent_id), threads)
    _check(rc)


class HostStepPlan:
    """Pre-marshalled arguments of gbl_step_host (chunk table, stream / event handles): the per-step cost on the
    Python side is ONE ctypes call."""

    def __init__(self, state, d_actions, d_rec, h_rec, chunk_end, stream, events, outputs, stats, flags, threads=0):
        k = len(chunk_end)
        self.keep = (state, d_actions, d_rec, h_rec, stream, events, outputs, stats)     # keep the buffers alive
        self.n, self.k, self.flags, self.threads = state.shape[0], k, int(flags), int(threads)
        self.ends = (C.c_int64 * k)(*[int(x) for x in chunk_end])
        self.stream = stream.cuda_stream
        for e in events:
            e.record(stream)                     # torch creates the cudaEvent lazily: make the handles exist
        self.events = (C.c_void_p * k)(*[e.cuda_event for e in events])
        self.ptrs = [_ptr(t) for t in (state, d_actions, d_rec, h_rec)]
        self.out = [_ptr(t) for t in outputs] if outputs is not None else [None] * 6
        self.stats = _ptr(stats)
        self.device = state.device

    def run(self, actions_host: torch.Tensor):
        if actions_host.is_cuda or actions_host.dtype != torch.uint8 or actions_host.numel() != self.n:
            raise GobbletError("actions must be a host uint8 tensor with one entry per env")
        st, da, dr, hr = self.ptrs
        with _on_device(self.device):
            _check(LIB.gbl_step_host(st, _ptr(actions_host), self.n, self.flags, da, dr, hr, self.k, self.ends, self.stream,
                                     self.events, *self.out, self.stats, self.threads))


def host_fill(t: torch.Tensor, threads: int = 0, mode: int = 0):
    """Measurement aid: fill a host tensor with the expander's thread pool (mode 0 = non-temporal stores)."""
    if t.is_cuda or not t.is_contiguous():
        raise GobbletError("host_fill needs a contiguous host tensor")
    _check(LIB.gbl_host_fill(_ptr(t), t.numel() * t.element_size(), threads, mode))


def host_last_timing():
    """Marks of the calling thread's last gbl_step_host (seconds since entry): enqueue done, chunks published, finished."""
    buf = (C.c_double * 66)()
    k = LIB.gbl_host_last_timing(buf, 66)
    return [buf[i] for i in range(min(k, 66))]


def host_threads(threads: int = 0) -> int:
    return int(LIB.gbl_host_threads(threads))


def host_simd() -> str:
    return "avx512bw" if LIB.gbl_host_simd() else "table"
