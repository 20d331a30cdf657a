"""Parity of the CUDA engine (through the C ABI / torch custom ops) with the CPU oracle. Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gv1():
    from gobblet_rl_b200 import gobblet_v1
    assert torch.cuda.is_available()
    return gobblet_v1


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("autoreset", ["same_step", "off", "next_step"])
@pytest.mark.parametrize("n,T", [(1000, 64), (32, 20), (1, 40), (4097, 12)])
def test_fused_rollout_bit_exact(gv1, autoreset, n, T):
    """obs, masks, rewards, terminations, agent ids, sampled actions and statistics of the fused
    rollout kernel == oracle replay (same seed, same global env ids)."""
    v = gv1.vec_env(n, seed=11, autoreset=autoreset, env_id_base=12345)
    v.step_count = 6
    got = v.rollout_random(T, ring=T, per_step=True, log_actions=True)
    o = O.VecOracle(n, "terminate", autoreset)
    want = o.rollout_random(T, seed=11, env_id_base=12345, step_base=6)
    # ring slot of absolute step s is s % ring
    order = [(6 + t) % T for t in range(T)]
    assert np.array_equal(_np(got["actions"]), want["actions"])
    assert np.array_equal(_np(got["obs"])[order], want["obs"])
    assert np.array_equal(_np(got["mask"])[order], want["mask"])
    assert np.array_equal(_np(got["rew"])[order], want["rew"])
    assert np.array_equal(_np(got["terminated"])[order], want["terminated"])
    assert np.array_equal(_np(got["agent_id"])[order], want["agent_id"])
    assert v.stats.tolist() == o.stats.tolist()
    sq, agent = v.squares()
    assert np.array_equal(_np(sq), o.squares())


def test_rollout_split_equals_whole(gv1):
    """State round-trips through HBM: 3 launches of 7+1+25 steps == one launch of 33."""
    a = gv1.vec_env(777, seed=3)
    b = gv1.vec_env(777, seed=3)
    whole = a.rollout_random(33, ring=33, log_actions=True)
    parts = [b.rollout_random(t, ring=1, log_actions=True)["actions"] for t in (7, 1, 25)]
    assert torch.equal(whole["actions"], torch.cat(parts))
    assert torch.equal(a.state, b.state) and torch.equal(a.stats, b.stats)


@pytest.mark.parametrize("illegal_mode", ["terminate", "pass"])
@pytest.mark.parametrize("autoreset", ["same_step", "off", "next_step"])
@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8])
def test_step_arbitrary_actions_bit_exact(gv1, illegal_mode, autoreset, dtype):
    n, T = 333, 50
    rng = np.random.default_rng(7)
    v = gv1.vec_env(n, illegal_mode=illegal_mode, autoreset=autoreset)
    o = O.VecOracle(n, illegal_mode, autoreset)
    _, mask, _ = o.reset()
    fobs = torch.zeros((n, 3, 3, 13), dtype=torch.int8, device="cuda")
    fmask = torch.zeros((n, 54), dtype=torch.int8, device="cuda")
    for t in range(T):
        acts = np.array([rng.choice(np.flatnonzero(m)) for m in mask], np.int64)
        bad = rng.random(n) < 0.1
        lo = -3 if dtype == torch.int64 else 0
        acts[bad] = rng.integers(lo, 60, bad.sum())
        g = v.step(torch.as_tensor(acts).to(dtype).cuda(), final=(fobs, fmask))
        w = o.step(acts, want_final=True)
        for i, name in enumerate(("obs", "mask", "rew", "terminated", "truncated", "agent_id")):
            assert np.array_equal(_np(g[i]), w[i]), (t, name)
        if autoreset == "same_step":
            assert np.array_equal(_np(fobs), w[6]) and np.array_equal(_np(fmask), w[7])
        mask = w[1]
    assert v.stats.tolist() == o.stats.tolist() and v.stats[5] > 0


def test_int32_actions_and_reset_ids(gv1):
    v = gv1.vec_env(64, autoreset="off")
    o = O.VecOracle(64, "terminate", "off")
    rng = np.random.default_rng(1)
    _, mask, _ = o.reset()
    for t in range(9):
        acts = np.array([rng.choice(np.flatnonzero(m)) for m in mask], np.int64)
        g = v.step(torch.as_tensor(acts, dtype=torch.int32).cuda())
        w = o.step(acts)
        assert np.array_equal(_np(g[0]), w[0]) and np.array_equal(_np(g[3]), w[3])
        mask = w[1]
    done = g[3].clone()
    assert done.any() and not done.all()
    obs, mask_t, agent = v.reset(done)                      # Tianshou order: reset finished envs only
    assert (obs[done] == 0).all() and (mask_t[done] == 1).all() and (agent[done] == 0).all()
    assert np.array_equal(_np(obs[~done]), w[0][~_np(done)])
    idx = torch.nonzero(~done).flatten()[:3]
    v.reset(idx)
    assert (v.obs[idx] == 0).all()


def test_reference_known_answers_and_golden_traces(gv1, golden):
    k = golden("reference_kat")
    v = gv1.vec_env(1, illegal_mode="pass", autoreset="off")
    _, mask, _ = v.reset()
    assert np.array_equal(_np(mask).astype(bool), k["output0"].astype(bool))
    for i, a in enumerate(k["actions"]):
        _, mask, *_ = v.step(torch.tensor([a]))
        assert np.array_equal(_np(mask).astype(bool), k[f"output{i + 1}"].astype(bool))
    assert np.flatnonzero(_np(mask)[0]).tolist() == k["output6"].tolist()
    v.step(torch.tensor([int(k["illegal_action"])]))
    assert np.array_equal(_np(v.squares()[0])[0].reshape(3, 3, 3), k["output8"].astype(np.int8))
    for name in ("env_traces", "env_traces_illegal"):
        g = golden(name)
        starts = g["game_start"]
        for gi in range(len(starts) - 1):
            v.reset()
            for i in range(starts[gi], starts[gi + 1]):
                obs, mask, rew, term, trunc, agent = v.step(torch.tensor([g["actions"][i]]))
                assert np.array_equal(_np(obs)[0], g["obs"][i]) and np.array_equal(_np(mask)[0], g["mask"][i])
                assert _np(rew)[0].tolist() == g["rew"][i].tolist() and bool(term[0]) == bool(g["term"][i])
                assert int(agent[0]) == int(g["agent_id"][i])
                assert np.array_equal(_np(v.squares()[0])[0], g["squares"][i])


def test_import_export_and_observe_random_positions(gv1):
    """Arbitrary reachable positions (incl. finished games) loaded through import_squares."""
    o = O.VecOracle(3000, "terminate", "off")
    o.rollout_random(14, seed=2, per_step=False)
    sq = o.squares()
    _, _, agent = o.observe()
    v = gv1.vec_env(3000, autoreset="off")
    obs, mask, ag = v.set_squares(sq, agent)
    wobs, wmask, wagent = o.observe()
    assert np.array_equal(_np(obs), wobs) and np.array_equal(_np(mask), wmask) and np.array_equal(_np(ag), wagent)
    back, ag2 = v.squares()
    assert np.array_equal(_np(back), sq) and np.array_equal(_np(ag2), agent)


def test_sample_legal_matches_oracle(gv1):
    from gobblet_rl_b200 import ops
    rng = np.random.default_rng(0)
    mask = (rng.random((500, 54)) < 0.4).astype(np.int8)
    mask[0] = 0
    mask[1] = 1
    act = torch.zeros(500, dtype=torch.int32, device="cuda")
    ops.sample_legal(torch.as_tensor(mask).cuda(), 99, 1 << 33, 6, act)
    want = [O.pick(m, O.draw(99, (1 << 33) + i, 6)) if m.any() else -1 for i, m in enumerate(mask)]
    assert _np(act).tolist() == want


@pytest.mark.parametrize("depth", [1, 2])
def test_greedy_matches_oracle_on_rollout_positions(gv1, depth):
    """Warp-per-board search == literal restatement of greedy_policy.py on 4000 positions,
    including positions of finished games and histories that trip the repetition rule."""
    n = 4000
    o = O.VecOracle(n, "terminate", "off")
    o.rollout_random(9, seed=4, per_step=False)
    obs, mask, _ = o.observe()
    rng = np.random.default_rng(5)
    prev3 = rng.integers(-1, 54, (n, 3)).astype(np.int16)
    act, chosen, cand, fb = gv1.greedy_actions(torch.as_tensor(obs).cuda(), torch.as_tensor(mask).cuda(),
                                               torch.as_tensor(prev3), depth=depth, seed=8, ctr_base=100, details=True)
    act, chosen, cand, fb = _np(act), _np(chosen), _np(cand).astype(np.uint64), _np(fb)
    n_fb = 0
    for i in range(n):
        wc, wcand, wfb = O.greedy(obs[i], mask[i], prev3[i], depth)
        bits = sum(1 << a for a in wcand)
        assert (int(chosen[i]), int(cand[i]), bool(fb[i])) == (wc, bits, wfb), i
        if wfb:
            n_fb += 1
            j = (O.philox4x32_10([100 + i, 0, 0, 1], [8, 0])[0].item() * len(wcand)) >> 32
            assert act[i] == wcand[j]
        else:
            assert act[i] == wc
    assert 0 < n_fb < n


def test_greedy_golden_from_reference(gv1, golden):
    g = golden("greedy")
    for depth in (1, 2):
        sel = g["depth"] == depth
        act, chosen, cand, fb = gv1.greedy_actions(torch.as_tensor(g["obs"][sel]).cuda(),
                                                   torch.as_tensor(g["mask"][sel]).cuda(),
                                                   torch.as_tensor(g["prev3"][sel]), depth=depth, details=True)
        assert _np(chosen).tolist() == g["chosen"][sel].tolist()
        want_bits = [sum(1 << int(a) for a in np.flatnonzero(c)) for c in g["cand"][sel]]
        assert [int(x) & (2**64 - 1) for x in _np(cand)] == want_bits
        assert _np(fb).tolist() == g["fallback"][sel].tolist()


def test_greedy_depth3_golden_from_reference(gv1, golden):
    """depth=3 through the kernel, the policy object and the adapter == what the reference returned at depth=3."""
    g = golden("greedy_depth3")
    obs, mask, prev3 = (torch.as_tensor(g[k]) for k in ("obs", "mask", "prev3"))
    act, chosen, cand, fb = gv1.greedy_actions(obs.cuda(), mask.cuda(), prev3, depth=3, details=True)
    assert _np(chosen).tolist() == g["chosen"].tolist() and _np(fb).tolist() == g["fallback"].tolist()
    assert [int(x) & (2**64 - 1) for x in _np(cand)] == [sum(1 << int(a) for a in np.flatnonzero(c)) for c in g["cand"]]
    import numpy.random as npr
    real = npr.choice
    npr.choice = lambda a, *k, **kw: a[0]                       # the fixture was recorded with this stand-in for the fallback draw
    try:
        for k in range(len(g["row"])):
            pol = gv1.GreedyGobbletPolicy(depth=3)
            agent = int(g["obs"][k].reshape(3, 3, 13)[0, 0, 12])
            pol.prev_actions[agent] = [int(x) for x in g["prev3"][k] if x >= 0]
            assert int(pol.compute_action(g["obs"][k].reshape(3, 3, 13), g["mask"][k])) == int(g["returned"][k]), k
    finally:
        npr.choice = real
    with pytest.raises(TypeError):
        gv1.GreedyGobbletPolicy(depth=None)                     # the reference fails on `None > 1` (greedy_policy.py:103)


def test_abi_rejects_bad_arguments(gv1):
    from gobblet_rl_b200 import ops
    v = gv1.vec_env(40)
    buf = torch.zeros(40 * 117 + 16, dtype=torch.int8, device="cuda")
    with pytest.raises(ops.GobbletError):
        ops.observe(v.state, buf[1:1 + 40 * 117], v.mask, None)         # misaligned obs
    with pytest.raises(ops.GobbletError):
        ops.step(v.state, torch.zeros(40, dtype=torch.int64, device="cuda"), v.obs, v.mask, None, None, None, None,
                 None, None, None, 3 << 1)                                # bad autoreset mode
    with pytest.raises(ops.GobbletError):
        ops.observe(v.state.cpu(), v.obs, v.mask, None)                  # no CPU path
    with pytest.raises(ops.GobbletError):
        ops.observe(v.state, v.obs[:39], v.mask, None)                   # obs too small for 40 envs
    with pytest.raises(ops.GobbletError):
        ops.observe(v.state, v.obs.to(torch.int32), v.mask, None)        # wrong element width
    with pytest.raises(ops.GobbletError):
        ops.step(v.state, torch.zeros(40, dtype=torch.int64, device="cuda"), v.obs, v.mask, v.rew[:10], None, None,
                 None, None, None, None, v.flags)                        # rew too small
    with pytest.raises(ops.GobbletError):
        ops.greedy(v.obs, v.mask[:20], None, 2, 0, 0, torch.zeros(40, dtype=torch.int32, device="cuda"), None, None, None)
    with pytest.raises(ops.GobbletError):
        ops.rollout_random(v.state.view(-1), 1, 0, 0, 0, None, None, None, None, None, None, None, None, None, v.flags)   # state shape


def test_million_env_properties(gv1):
    """BASELINE config 3 size: invariants that need no oracle replay."""
    n, T = 1 << 20, 48
    a = gv1.vec_env(n, seed=0)
    a.rollout_random(T, ring=2)
    ep, p1, p2, steps, sumlen, illegal, both, maxlen = a.stats.tolist()
    assert steps == n * T and illegal == 0 and ep == p1 + p2
    assert 11.0 < sumlen / ep < 12.2 and 0.535 < p1 / ep < 0.555 and 0.005 < both / ep < 0.02
    # determinism + sharding independence: two half-size shards with global ids == the whole
    lo = gv1.vec_env(n // 2, seed=0, env_id_base=0)
    hi = gv1.vec_env(n // 2, seed=0, env_id_base=n // 2)
    lo.rollout_random(T, ring=2)
    hi.rollout_random(T, ring=2)
    assert torch.equal(torch.cat([lo.state, hi.state]), a.state)
    assert (lo.stats[:7] + hi.stats[:7]).tolist() == a.stats[:7].tolist()
    assert max(lo.stats[7].item(), hi.stats[7].item()) == maxlen
    # the last emitted observation is the observation of the final state
    obs_ring, mask_ring = a._rings[1], a._rings[2]
    slot = (T - 1) % 2
    obs, mask, _ = a.observe()
    assert torch.equal(obs_ring[slot], obs) and torch.equal(mask_ring[slot], mask)
    # masks are 0/1 and every env has a legal move
    assert int(mask.max()) == 1 and int(mask.sum(1).min()) >= 1


def test_cuda_graph_replay_of_rollouts_matches_oracle(gv1):
    """graph_safe=True keeps the Philox step counter on the device: a captured launch replayed k times
    walks the same random stream as k eager launches (BASELINE config 2, CUDA-graph variant)."""
    n, T, reps = 300, 5, 4
    v = gv1.vec_env(n, seed=13, graph_safe=True)
    log = torch.zeros((reps + 1, T, n), dtype=torch.uint8, device="cuda")

    def one(i):
        out = v.rollout_random(T, ring=1, log_actions=True)
        log[i].copy_(out["actions"])

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        one(0)                                           # eager warm-up launch: steps 0..T-1
        idx = torch.zeros((), dtype=torch.int64, device="cuda")
        with torch.cuda.graph(g, stream=side):
            out = v.rollout_random(T, ring=1, log_actions=True)
            captured = out["actions"]
    torch.cuda.current_stream().wait_stream(side)
    for i in range(1, reps + 1):
        g.replay()
        log[i].copy_(captured)
    torch.cuda.synchronize()
    o = O.VecOracle(n)
    want = o.rollout_random(T * (reps + 1), seed=13)
    assert np.array_equal(log.cpu().numpy().reshape(-1, n), want["actions"])
    assert v.stats.tolist()[:7] == o.stats.tolist()[:7]
    assert int(v.step_dev) == T * (reps + 1)


def test_checkpoint_resume(gv1):
    """The state tensor + counters are the whole env state: save, keep running, restore,
    rerun -> identical trajectories."""
    v = gv1.vec_env(500, seed=31)
    v.rollout_random(17, emit=False)
    ckpt = v.state_dict()
    a = v.rollout_random(23, ring=1, log_actions=True)["actions"].clone()
    end_state, end_stats = v.state.clone(), v.stats.clone()
    w = gv1.vec_env(500, seed=0)
    w.load_state_dict(ckpt)
    b = w.rollout_random(23, ring=1, log_actions=True)["actions"]
    assert torch.equal(a, b) and torch.equal(w.state, end_state) and torch.equal(w.stats, end_stats)


def test_empty_and_degenerate_sizes(gv1):
    """n = 0 and T = 0 are no-ops at the C ABI; ring shorter than T keeps the last `ring` steps."""
    from gobblet_rl_b200 import ops
    lib = ops.LIB
    assert lib.gbl_reset(None, 0, None) == 0 and lib.gbl_observe(None, None, None, None, 0, None) == 0
    assert lib.gbl_step(None, None, 8, None, None, None, None, None, None, None, None, None, 0, 0, None) == 0
    assert lib.gbl_greedy(None, None, None, 2, 0, 0, None, None, None, None, 0, None) == 0
    v = gv1.vec_env(77, seed=2)
    before = v.state.clone()
    ops.rollout_random(v.state, 0, 2, 0, 0, None, None, None, None, None, None, None, None, v.stats, v.flags)
    assert torch.equal(v.state, before)
    out = v.rollout_random(10, ring=3)
    o = O.VecOracle(77)
    want = o.rollout_random(10, seed=2)
    for s in (7, 8, 9):                                   # ring slot of absolute step s is s % 3
        assert np.array_equal(_np(out["obs"][s % 3]), want["obs"][s])
        assert np.array_equal(_np(out["mask"][s % 3]), want["mask"][s])


def test_empty_mask_and_full_board_greedy_inputs(gv1):
    """gbl_greedy on degenerate inputs: an empty mask yields act = -1 with the fallback flag set (the reference
    raises inside np.random.choice); a mask inconsistent with the board only ever returns masked actions."""
    obs = torch.zeros((3, 3, 3, 13), dtype=torch.int8, device="cuda")
    mask = torch.zeros((3, 54), dtype=torch.int8, device="cuda")
    mask[1, 5] = 1
    mask[2] = 1
    act, chosen, cand, fb = gv1.greedy_actions(obs, mask, None, depth=2, details=True)
    assert act[0].item() == -1 and fb[0].item() and cand[0].item() == 0
    assert act[1].item() == 5 and act[2].item() in range(54)
    wc, wcand, wfb = O.greedy(_np(obs)[2], _np(mask)[2], (-1, -1, -1), 2)
    assert (chosen[2].item(), bool(fb[2])) == (wc, wfb)


def test_custom_ops_trace_under_torch_compile(gv1):
    """The ops carry schemas with mutated arguments and fake impls, so a caller's step function can be traced
    (aot_eager: functionalisation + fake tensors, no code generation) with the engine call inside the graph."""
    from gobblet_rl_b200 import ops
    v = gv1.vec_env(256, seed=3, autoreset="off")
    ref = gv1.vec_env(256, seed=3, autoreset="off")

    def fn(state, actions, obs, mask, rew, term, trunc, agent, stats):
        ops.step(state, actions, obs, mask, rew, term, trunc, agent, None, None, stats, v.flags)
        return obs.sum(dim=(1, 2, 3)) + mask.sum(dim=1)

    compiled = torch.compile(fn, backend="aot_eager", fullgraph=True)
    acts = torch.full((256,), 4, dtype=torch.int64, device="cuda")
    got = compiled(v.state, acts, v.obs, v.mask, v.rew, v.terminated.view(torch.uint8), v.truncated.view(torch.uint8),
                   v.agent_id, v.stats)
    want_obs, want_mask, *_ = ref.step(acts)
    assert torch.equal(v.obs, want_obs) and torch.equal(v.mask, want_mask) and torch.equal(v.state, ref.state)
    assert torch.equal(got, want_obs.sum(dim=(1, 2, 3)) + want_mask.sum(dim=1))


def _random_boards(rng, n):
    """Arbitrary (not necessarily reachable) placements: every piece in hand or on a random square, at most one
    piece per (level, square) cell -- stresses covered pieces, double lines and full columns."""
    sq = np.zeros((n, 27), np.int8)
    for i in range(n):
        for sign in (1, -1):
            for k in range(1, 7):
                if rng.random() < 0.7:
                    lvl = (k - 1) // 2
                    free = np.flatnonzero(sq[i, 9 * lvl: 9 * lvl + 9] == 0)
                    if len(free):
                        sq[i, 9 * lvl + rng.choice(free)] = sign * k
    return sq


def test_synthetic_positions_observe_step_and_greedy(gv1):
    rng = np.random.default_rng(11)
    n = 3000
    sq = _random_boards(rng, n)
    agent = rng.integers(0, 2, n).astype(np.uint8)
    v = gv1.vec_env(n, illegal_mode="pass", autoreset="off")
    o = O.VecOracle(n, "pass", "off")
    obs, mask, ag = v.set_squares(sq, agent)
    o.set(sq, agent)
    wobs, wmask, wag = o.observe()
    assert np.array_equal(_np(obs), wobs) and np.array_equal(_np(mask), wmask) and np.array_equal(_np(ag), wag)
    # greedy on these positions, with masks that are sometimes INCONSISTENT with the board (random extra / missing bits)
    gmask = wmask.copy()
    flip = rng.random(gmask.shape) < 0.05
    gmask[flip] ^= 1
    gmask[gmask.sum(1) == 0, 0] = 1
    prev3 = rng.integers(-1, 54, (n, 3)).astype(np.int16)
    for depth in (1, 2):
        act, chosen, cand, fb = gv1.greedy_actions(obs, torch.as_tensor(gmask).cuda(), torch.as_tensor(prev3), depth=depth, details=True)
        chosen, cand, fb = _np(chosen), _np(cand).astype(np.uint64), _np(fb)
        for i in range(n):
            wc, wcand, wfb = O.greedy(wobs[i], gmask[i], prev3[i], depth)
            assert (int(chosen[i]), int(cand[i]), bool(fb[i])) == (wc, sum(1 << a for a in wcand), wfb), (depth, i)
    # one step with arbitrary actions from these positions (winner by line order, uncovering, illegal = pass)
    acts = rng.integers(0, 54, n)
    g = v.step(torch.as_tensor(acts).cuda())
    w = o.step(acts)
    for i in range(6):
        assert np.array_equal(_np(g[i]), w[i]), i
    assert np.array_equal(_np(v.squares()[0]), o.squares())


# ---- round 2: states the fast rollout path did not produce itself, the timed template instance, wire format ----
def test_fast_rollout_from_a_checkpoint_with_finished_envs(gv1):
    """An autoreset="off" run leaves finished envs in the state tensor; loaded into a same-step env the fused
    (kFast) rollout must treat them as the general path and the oracle do: the first step only resets them."""
    n, T = 2000, 30
    src = gv1.vec_env(n, seed=4, autoreset="off")
    src.rollout_random(12, emit=False)
    assert 0 < int((src.state[:, 0] >> 55 & 1).sum()) < n                 # done bit of the packed state
    o = O.VecOracle(n, "terminate", "off")
    o.rollout_random(12, seed=4, per_step=False)
    v = gv1.vec_env(n, seed=77)
    v.load_state_dict({**src.state_dict(), "seed": 77})
    assert v.step_count == 12
    got = v.rollout_random(T, ring=T, per_step=True, log_actions=True)
    o.flags = O.flags("terminate", "same_step")
    stats_before = o.stats.copy()
    want = o.rollout_random(T, seed=77, step_base=12)
    order = [(12 + t) % T for t in range(T)]
    assert (want["actions"][0] == 255).any()                     # reset-only first steps exist
    assert np.array_equal(_np(got["actions"]), want["actions"])
    for k in ("obs", "mask", "rew", "terminated", "agent_id"):
        assert np.array_equal(_np(got[k])[order], want[k]), k
    assert v.stats.tolist() == o.stats.tolist() and (o.stats != stats_before).any()
    assert np.array_equal(_np(v.squares()[0]), o.squares())
    # the per-step kernel agrees on the same imported state (an env that arrives finished is reset, not stuck)
    s = gv1.vec_env(n, seed=77)
    s.load_state_dict({**src.state_dict(), "seed": 77})
    o2 = O.VecOracle(n, "terminate", "off")
    o2.rollout_random(12, seed=4, per_step=False)
    o2.flags = O.flags("terminate", "same_step")
    acts = np.where(want["actions"][0] == 255, 0, want["actions"][0]).astype(np.int64)
    g = s.step(torch.as_tensor(acts).cuda())
    w = o2.step(acts)
    for i in range(6):
        assert np.array_equal(_np(g[i]), w[i]), i


def test_rollout_from_finished_positions_set_by_squares(gv1):
    """set_squares of positions that already hold a complete line, then the fused rollout (kFast)."""
    n, T = 1500, 16
    o = O.VecOracle(n, "terminate", "off")
    o.rollout_random(14, seed=8, per_step=False)
    sq, agent = o.squares(), o.observe()[2]
    assert any(O.check_for_winner(s) != 0 for s in sq[:200])
    v = gv1.vec_env(n, seed=3)
    v.set_squares(sq, agent)
    w = O.VecOracle(n, "terminate", "same_step")
    w.set(sq, agent)
    got = v.rollout_random(T, ring=T, per_step=True, log_actions=True)
    want = w.rollout_random(T, seed=3)
    for k in ("actions", "obs", "mask", "rew", "terminated", "agent_id"):
        assert np.array_equal(_np(got[k]), want[k]), k
    assert v.stats.tolist() == w.stats.tolist()


def test_timed_template_instance_equals_the_checked_one_at_full_size(gv1):
    """bench.py times rollout_kernel<fast, streaming, no-aux>; the oracle comparisons run the aux instance.
    Device-side equality of both instances at BASELINE config 3 size: 2^20 envs x 64 steps, every observation and
    mask byte of the trajectory, the state and the statistics."""
    n, T = 1 << 20, 64
    a = gv1.vec_env(n, seed=0)
    b = gv1.vec_env(n, seed=0)
    plain = a.rollout_random(T, ring=T)                                        # <true, true, false>
    aux = b.rollout_random(T, ring=T, per_step=True, log_actions=True)         # <true, true, true>
    for t in range(T):
        assert torch.equal(plain["obs"][t], aux["obs"][t]) and torch.equal(plain["mask"][t], aux["mask"][t]), t
    assert torch.equal(a.state, b.state) and torch.equal(a.stats, b.stats)
    # and the aux instance is the oracle's on a slice small enough to replay on the CPU
    k = 4096
    o = O.VecOracle(k)
    want = o.rollout_random(T, seed=0)
    assert np.array_equal(_np(aux["actions"][:, :k]), want["actions"])
    assert np.array_equal(_np(plain["obs"][:, :k]), want["obs"]) and np.array_equal(_np(plain["mask"][:, :k]), want["mask"])


@pytest.mark.parametrize("n", [100, 4096, 40000])
def test_rollout_block_sizes_and_short_rings_agree(gv1, n):
    """Every block size of the fused rollout (32/64/128/256 threads, chosen from N by default) and every ring
    length (slots rewritten within the launch: bulk stores ordered by wait_group) gives the same bytes."""
    T = 24
    ref = gv1.vec_env(n, seed=6)
    base = ref.rollout_random(T, ring=T, per_step=True, log_actions=True, final=True)
    for hint, ring, no_bulk, split in ((32, T, False, False), (64, T, False, None), (128, 2, False, None), (256, 1, False, None),
                                       (0, 3, False, None), (32, 1, False, False), (32, T, True, False), (64, 2, True, None),
                                       (0, 1, True, None), (0, T, False, False), (32, T, False, True), (32, 2, True, True),
                                       (32, 1, False, True), (0, 3, True, True)):
        # split: two warps per 32 envs (observation warp / mask warp); None = the library's choice for this N
        v = gv1.vec_env(n, seed=6)
        out = v.rollout_random(T, ring=ring, per_step=True, log_actions=True, final=True, block_hint=hint, no_bulk=no_bulk,
                               split=split)
        assert torch.equal(out["actions"], base["actions"]) and torch.equal(v.state, ref.state) and torch.equal(v.stats, ref.stats)
        for s in range(max(0, T - ring), T):
            for key in ("obs", "mask", "final_obs", "final_mask", "rew", "terminated", "agent_id"):
                assert torch.equal(out[key][s % ring], base[key][s]), (hint, ring, no_bulk, split, s, key)


def test_skip255_only_skips_exactly_255(gv1):
    """GBL_ACTION_SKIP_255 reserves exactly 255; -1, 254, 256, 300 are illegal moves (AssertOutOfBoundsWrapper /
    TerminateIllegalWrapper territory, gobblet.py:110-117), never silent no-ops."""
    n = 64
    v = gv1.vec_env(n, autoreset="off", skip255=True)
    o = O.VecOracle(n, "terminate", "off", skip255=True)
    acts = np.zeros(n, np.int64)
    acts[0:8] = [-1, 300, 254, 256, 255, 54, -(2**40), 2**40]
    g = v.step(torch.as_tensor(acts).cuda())
    w = o.step(acts)
    for i in range(6):
        assert np.array_equal(_np(g[i]), w[i]), i
    term = _np(g[3])
    assert term[[0, 1, 2, 3, 5, 6, 7]].all() and not term[4]           # 255 alone was skipped
    assert _np(g[2])[0].tolist() == [-1, 0] and int(v.stats[5]) == 7 and v.stats.tolist() == o.stats.tolist()
    # uint8 actions: 255 skips, 254 is illegal
    v2 = gv1.vec_env(4, autoreset="off", skip255=True)
    g2 = v2.step(torch.tensor([255, 254, 0, 53], dtype=torch.uint8).cuda())
    assert _np(g2[3]).tolist() == [False, True, False, False]


def test_graph_safe_checkpoint_and_resume(gv1):
    """graph_safe=True keeps the Philox step counter on the device; a checkpoint taken after CUDA-graph replays
    (which advance only the device counter) must resume on the same random stream."""
    n, T = 400, 4
    v = gv1.vec_env(n, seed=9, graph_safe=True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        v.rollout_random(T, emit=False)
        with torch.cuda.graph(g, stream=side):
            v.rollout_random(T, emit=False)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ckpt = v.state_dict()
    assert ckpt["step_count"] == 4 * T == v.sampler_step()      # 1 eager launch + 3 replays (a capture runs nothing)
    cont = v.rollout_random(9, ring=1, log_actions=True)["actions"].clone()
    for graph_safe in (True, False):
        w = gv1.vec_env(n, seed=1, graph_safe=graph_safe)
        w.load_state_dict(ckpt)
        again = w.rollout_random(9, ring=1, log_actions=True)["actions"]
        assert torch.equal(cont, again) and torch.equal(w.state, v.state)
    o = O.VecOracle(n)
    want = o.rollout_random(4 * T + 9, seed=9)
    assert np.array_equal(_np(cont), want["actions"][4 * T:])
    # externally driven steps keep both counters in lockstep
    v.step(torch.zeros(n, dtype=torch.int64, device="cuda"))
    assert int(v.step_dev) == v.step_count == 4 * T + 10


@pytest.mark.parametrize("illegal_mode", ["terminate", "pass"])
@pytest.mark.parametrize("autoreset", ["same_step", "off", "next_step"])
def test_packed_wire_format_step_bit_exact(gv1, illegal_mode, autoreset):
    """gbl_step_packed / gbl_observe_packed: the 24-byte records expand (numpy statement of the layout AND the
    library's host expander) to exactly what the oracle's step returns, terminal observation included."""
    from gobblet_rl_b200 import ops
    from gobblet_rl_b200.vec_env import unpack_records
    n, T = 777, 40
    rng = np.random.default_rng(5)
    v = gv1.vec_env(n, illegal_mode=illegal_mode, autoreset=autoreset)
    o = O.VecOracle(n, illegal_mode, autoreset)
    wobs, mask, wagent = o.reset()
    r0 = unpack_records(v.observe_packed().cpu())
    assert np.array_equal(r0[0], wobs) and np.array_equal(r0[1], mask) and np.array_equal(r0[5], wagent) and not r0[2].any()
    frec = torch.zeros((n, 6), dtype=torch.int32, device="cuda")
    host = [torch.zeros((n, 3, 3, 13), dtype=torch.int8), torch.zeros((n, 54), dtype=torch.int8), torch.zeros((n, 2), dtype=torch.int8),
            torch.zeros(n, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8)]
    for t in range(T):
        acts = np.array([rng.choice(np.flatnonzero(m)) for m in mask], np.int64)
        bad = rng.random(n) < 0.1
        acts[bad] = rng.integers(-3, 60, bad.sum())
        rec = v.step_packed(torch.as_tensor(acts).cuda(), final_rec=frec).cpu()
        w = o.step(acts, want_final=True)
        got = unpack_records(rec)
        for i in range(6):
            assert np.array_equal(got[i], w[i]), (t, i)
        ops.host_unpack(rec, *host, threads=3)
        for i in range(6):
            assert np.array_equal(host[i].numpy().astype(w[i].dtype), w[i]), (t, i)
        if autoreset == "same_step":
            f = unpack_records(frec.cpu())
            assert np.array_equal(f[0], w[6]) and np.array_equal(f[1], w[7])
        assert (rec[:, 3].numpy().view(np.uint32) >> 28 == 0).all() and (rec[:, 5].numpy().view(np.uint32) >> 22 == 0).all()
        mask = w[1]
    assert v.stats.tolist() == o.stats.tolist() and v.stats[5] > 0


def test_import_squares_rejects_what_the_reference_rejects(gv1):
    """Board.is_legal raises when a piece sits on two squares (board.py:94-95); a piece on a level that is not
    its size's cannot exist at all.  Both are refused; valid positions import unchanged."""
    v = gv1.vec_env(4, autoreset="off")
    sq = np.zeros((4, 27), np.int8)
    sq[0, 0] = 1; sq[0, 4] = 1                 # piece 1 twice
    sq[1, 3] = 5                               # a large piece on the small level
    sq[2, 20] = -6; sq[2, 2] = 2               # fine
    with pytest.raises(ValueError, match="2 position"):
        v.set_squares(sq, np.zeros(4, np.uint8))
    back = _np(v.squares()[0])
    assert not back[0].any() and not back[1].any() and np.array_equal(back[2], sq[2])
    v.set_squares(sq[[2, 2, 3, 3]], np.zeros(4, np.uint8))


def test_sample_legal_coalesced_and_fallback_paths(gv1):
    """Full warps take the 128-bit load + shared bit-stream path, the ragged last warp and a misaligned base the
    byte path; int8, uint8 and bool masks; all equal the oracle's pick."""
    from gobblet_rl_b200 import ops
    rng = np.random.default_rng(3)
    n = 32 * 40 + 13
    mask = (rng.random((n, 54)) < 0.5).astype(np.int8)
    mask[5] = 0
    mask[70] = 1
    mask[100] *= 77                                             # any non-zero byte counts as legal
    want = [O.pick(m, O.draw(7, 1000 + i, 11)) if m.any() else -1 for i, m in enumerate(mask)]
    dev = torch.as_tensor(mask).cuda()
    for m in (dev, dev.view(torch.uint8), dev != 0):
        act = torch.zeros(n, dtype=torch.int32, device="cuda")
        ops.sample_legal(m.contiguous(), 7, 1000, 11, act)
        assert _np(act).tolist() == want
    pad = torch.zeros(n * 54 + 64, dtype=torch.int8, device="cuda")
    off = pad[2:2 + n * 54].view(n, 54)
    off.copy_(dev)
    act = torch.zeros(n, dtype=torch.int32, device="cuda")
    ops.sample_legal(off, 7, 1000, 11, act)
    assert _np(act).tolist() == want
