"""Tianshou-shaped vector env, GPU trajectory collection and the
policy adaptors, checked against the reference traces (tests/golden) and the oracle. Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ad():
    from gobblet_rl_b200 import adapters
    assert torch.cuda.is_available()
    return adapters


def test_pettingzoo_vec_env_matches_reference_trace(ad, golden):
    """PettingZooEnv.step == env.step(a); env.last(): replay the recorded env() games (illegal-move
    terminations by either player included) and compare with the NEXT recorded last() tuple."""
    g = golden("env_wrapped")
    starts = g["game_start"]
    venv = ad.PettingZooVecEnv(1)
    seen_illegal_p2 = False
    for gi in range(len(starts) - 1):
        batch, info = venv.reset()
        i = starts[gi]
        assert batch["agent_id"][0] == "player_1" and batch["mask"].dtype == bool and batch["mask"].all()
        while g["action"][i] >= 0:
            batch, rew, term, trunc, info = venv.step(np.array([g["action"][i]]))
            i += 1
            who = int(g["agent"][i])
            assert batch["agent_id"][0] == ("player_1", "player_2")[who]
            assert np.array_equal(batch["obs"][0], g["obs"][i]) and np.array_equal(batch["mask"][0], g["mask"][i].astype(bool))
            assert bool(term[0]) == bool(g["term"][i]) and bool(trunc[0]) == bool(g["trunc"][i])
            assert rew.shape == (1, 2) and rew[0, who] == g["reward"][i]
            seen_illegal_p2 |= bool(trunc[0]) and not g["mask"][i].any()
    assert seen_illegal_p2


def test_subset_step_and_reset_ids(ad):
    """Tianshou steps / resets only the ready env ids; the others must not move."""
    n = 50
    venv = ad.PettingZooVecEnv(n, to_numpy=False)
    o = O.VecOracle(n, "terminate", "off", skip255=True)
    venv.reset()
    _, mask, _ = o.reset()
    rng = np.random.default_rng(3)
    for t in range(25):
        ids = np.sort(rng.choice(n, size=int(rng.integers(1, n)), replace=False))
        acts = np.array([rng.choice(np.flatnonzero(mask[i])) for i in ids])
        full = np.full(n, 255, np.int64)
        full[ids] = acts
        batch, rew, term, trunc, _ = venv.step(torch.as_tensor(acts), id=ids)
        w = o.step(full)
        assert np.array_equal(batch["obs"].cpu().numpy(), w[0][ids]) and np.array_equal(batch["mask"].cpu().numpy(), w[1][ids].astype(bool))
        assert np.array_equal(rew.cpu().numpy(), w[2][ids]) and np.array_equal(term.cpu().numpy(), w[3][ids])
        mask = w[1]
        done = np.flatnonzero(w[3])
        if len(done):
            venv.reset(done)
            sq = o.squares(); sel = o.observe()[2]
            sq[done] = 0; sel[done] = 0
            o.set(sq, sel)
            mask = o.observe()[1]
    assert np.array_equal(venv.vec.squares()[0].cpu().numpy(), o.squares())


def test_collector_trajectory_matches_oracle_replay(ad):
    """BASELINE config 5 shape: obs/mask/rew/flags batches written by the step kernel straight into the
    trajectory buffer == oracle replay of the collected actions (terminal observations included)."""
    from gobblet_rl_b200 import gobblet_v1
    n, T = 1500, 20
    vec = gobblet_v1.vec_env(n, seed=5)
    buf = ad.TrajectoryBuffer(T, n)
    col = ad.VecCollector(vec, ad.RandomLegalPolicy(seed=21), buf)
    col.collect()
    o = O.VecOracle(n, "terminate", "same_step")
    obs0, mask0, agent0 = o.reset()
    assert np.array_equal(buf.obs[0].cpu().numpy(), obs0) and np.array_equal(buf.mask[0].cpu().numpy(), mask0)
    acts = buf.act.cpu().numpy()
    mask = mask0
    for t in range(T):
        want_act = [O.pick(mask[i], O.draw(21, i, t)) for i in range(0, n, 97)]
        assert acts[t][::97].tolist() == want_act                       # Philox masked-uniform sampler
        w = o.step(acts[t].astype(np.int64), want_final=True)
        assert np.array_equal(buf.obs[t + 1].cpu().numpy(), w[0]) and np.array_equal(buf.mask[t + 1].cpu().numpy(), w[1])
        assert np.array_equal(buf.rew[t].cpu().numpy(), w[2]) and np.array_equal(buf.terminated[t].cpu().numpy(), w[3])
        assert not buf.truncated[t].any() and np.array_equal(buf.agent_id[t + 1].cpu().numpy(), w[5])
        assert np.array_equal(buf.final_obs[t].cpu().numpy(), w[6]) and np.array_equal(buf.final_mask[t].cpu().numpy(), w[7])
        mask = w[1]
    assert buf.terminated.any()
    col.roll()
    assert torch.equal(buf.obs[0], buf.obs[-1])


def test_greedy_policy_forward_shares_history_like_the_reference(ad):
    """greedy_policy_tianshou.py:63-84 loops the batch through ONE GreedyGobbletPolicy, so the repetition
    history is shared and order dependent; expected values come from the oracle restatement
    driven the same way with the same numpy seed."""
    o = O.VecOracle(64, "terminate", "off")
    o.rollout_random(7, seed=9, per_step=False)
    obs, mask, agent = o.observe()
    live = np.flatnonzero(~np.array([O.check_for_winner(s) != 0 for s in o.squares()]))[:40]
    obs, mask = obs[live], mask[live]
    pol = ad.GreedyPolicy(depth=2)
    batch = ad.Batch(obs=ad.Batch(obs=obs, mask=mask.astype(bool), agent_id=np.array(["player_1"] * len(live))))
    np.random.seed(123)
    got = [pol.forward(batch).act for _ in range(2)]                      # second call trips the history rule
    np.random.seed(123)
    hist = {0: [], 1: []}
    for rnd in range(2):
        want = []
        for i in range(len(live)):
            me = int(obs[i][..., 12].max())
            chosen, cand, fb = O.greedy(obs[i], mask[i], (hist[me][-3:] + [-1, -1, -1])[:3], 2)
            a = int(np.random.choice(cand)) if fb else chosen
            hist[me].append(a)
            want.append(a)
        assert got[rnd].tolist() == want
    assert got[0].shape == (len(live),) and pol.learn(batch) == {}


def test_random_admissible_policy(ad):
    rng = np.random.default_rng(0)
    masks = (rng.random((200, 54)) < 0.5).astype(np.int8)
    masks[:, 0] = 1
    acts, state, info = ad.RandomAdmissiblePolicy(seed=4).compute_actions({"action_mask": masks})
    assert state == [] and info == {} and all(masks[i, a] == 1 for i, a in enumerate(acts))
    assert len(set(acts)) > 20


def test_greedy_vec_policy_self_play_matches_oracle(ad):
    """Greedy-vs-greedy self play of 400 envs on the GPU (tutorial_greedy.py:30-44 shape, two random plies
    first): every move must be the one the oracle restatement of greedy_policy.py allows given the per-env
    history, and the env transitions must match the oracle env."""
    from gobblet_rl_b200 import gobblet_v1
    n, T = 400, 14
    vec = gobblet_v1.vec_env(n, seed=8, autoreset="off")
    vec.rollout_random(2, emit=False)
    o = O.VecOracle(n, "terminate", "off")
    o.rollout_random(2, seed=8, per_step=False)
    pol = ad.GreedyVecPolicy(n, depth=2, seed=4)
    hist = [[[], []] for _ in range(n)]
    obs, mask, agent = vec.observe()
    n_fb = 0
    for t in range(T):
        act = pol(obs, mask, agent)
        a_np, ag_np, obs_np, mask_np = act.cpu().numpy(), agent.cpu().numpy(), obs.cpu().numpy(), mask.cpu().numpy()
        done = vec.terminated.cpu().numpy() if t else np.zeros(n, bool)
        for i in range(0, n, 3):
            if done[i]:
                continue
            h = hist[i][ag_np[i]]
            chosen, cand, fb = O.greedy(obs_np[i], mask_np[i], (h[-3:] + [-1, -1, -1])[:3] if len(h) < 3 else h[-3:], 2)
            assert (a_np[i] in cand) if fb else (a_np[i] == chosen), (t, i)
            n_fb += fb
        for i in range(n):
            hist[i][ag_np[i]].append(int(a_np[i]))
        obs, mask, rew, term, trunc, agent = vec.step(act)
        w = o.step(a_np.astype(np.int64))
        assert np.array_equal(obs.cpu().numpy(), w[0]) and np.array_equal(term.cpu().numpy(), w[3])
    assert vec.stats[5] == 0 and vec.stats[0] > 0 and n_fb > 0


def test_captured_collection_graph_matches_eager(ad):
    """VecCollector.capture(): replaying the CUDA graph of collect()+roll() continues the same trajectories
    as eager collection (device-side Philox counter), checked against an eager twin."""
    from gobblet_rl_b200 import gobblet_v1
    n, T = 600, 6

    def make():
        vec = gobblet_v1.vec_env(n, seed=12)
        buf = ad.TrajectoryBuffer(T, n, keep_final=False)
        return vec, buf, ad.VecCollector(vec, ad.RandomLegalPolicy(seed=5, graph_safe_device="cuda"), buf)

    va, ba, ca = make()
    vb, bb, cb = make()
    g = ca.capture()                      # runs one eager collect+roll (warm-up), then captures
    cb.collect(); cb.roll()
    for _ in range(3):
        g.replay()
        cb.collect(); cb.roll()
        torch.cuda.synchronize()
        assert torch.equal(ba.act, bb.act) and torch.equal(ba.obs, bb.obs) and torch.equal(ba.rew, bb.rew)
    assert torch.equal(va.state, vb.state) and va.stats.tolist() == vb.stats.tolist() and va.stats[5] == 0


@pytest.mark.parametrize("wire,n,chunks", [("packed", 1003, 5), ("packed", 9000, 4), ("dense", 1003, 5), ("dense", 9000, 3)])
def test_host_vec_env_matches_oracle(ad, wire, n, chunks):
    """The host-buffer path bench.py reports as `e2e`: pinned actions in, pinned obs / mask / rew / flags out,
    chunked over CUDA streams -- same results as the oracle, whatever the chunking and whatever crosses PCIe
    (24-byte packed records expanded by the host thread pool, or the expanded tensors themselves)."""
    from gobblet_rl_b200 import gobblet_v1
    T = 25
    host = gobblet_v1.HostVecEnv(n, chunks=chunks, wire=wire, seed=0, autoreset="same_step")
    assert len(host.parts) == (1 if n < 2048 else chunks) and host.parts[-1][1] == n
    o = O.VecOracle(n, "terminate", "same_step")
    obs, mask, agent = host.reset()
    wobs, wmask, wagent = o.reset()
    assert np.array_equal(obs.numpy(), wobs) and np.array_equal(mask.numpy(), wmask)
    rng = np.random.default_rng(2)
    acts = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for t in range(T):
        a = np.array([rng.choice(np.flatnonzero(m)) for m in wmask], np.int64)
        a[rng.random(n) < 0.05] = 60                                  # some illegal moves
        acts.copy_(torch.as_tensor(a, dtype=torch.uint8))
        got = host.step(acts)
        w = o.step(a)
        for g, ww in zip(got, w):
            assert np.array_equal(g.numpy(), ww), t
        wmask = w[1]
    assert host.h2d_bytes_per_step == n and host.host_bytes_per_step == n * 176
    assert host.d2h_bytes_per_step == n * (24 if wire == "packed" else 176)
    assert host.env.stats.tolist() == o.stats.tolist()


def test_host_vec_env_packed_consumer(ad):
    """expand=False: the consumer keeps the wire format; the records expand (numpy) to the oracle's arrays."""
    from gobblet_rl_b200 import gobblet_v1
    from gobblet_rl_b200.vec_env import unpack_records
    n = 4096
    host = gobblet_v1.HostVecEnv(n, chunks=2, expand=False, seed=0)
    o = O.VecOracle(n)
    _, wmask, _ = o.reset()
    host.reset()
    acts = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for t in range(6):
        a = np.array([np.flatnonzero(m)[t % m.sum()] for m in wmask], np.int64)
        acts.copy_(torch.as_tensor(a, dtype=torch.uint8))
        rec = host.step(acts)
        assert rec.shape == (n, 6) and rec.is_pinned()
        w = o.step(a)
        for g, ww in zip(unpack_records(rec), w):
            assert np.array_equal(g, ww), t
        wmask = w[1]


def test_fused_collection_equals_step_by_step_collection(ad):
    """The built-in masked-uniform policy collects in ONE launch (fused rollout writing into the buffer slots,
    slot 0 and the terminal observations included); it must fill the buffer exactly like sample_legal + step
    per step."""
    from gobblet_rl_b200 import gobblet_v1
    n, T = 3000, 12
    bufs = []
    for fused in (True, False):
        vec = gobblet_v1.vec_env(n, seed=5, env_id_base=40)
        buf = ad.TrajectoryBuffer(T, n)
        col = ad.VecCollector(vec, ad.RandomLegalPolicy(seed=21, env_id_base=40), buf, fused=fused)
        assert col.fused == fused
        launches = vec.kernel_launches
        for _ in range(2):
            col.collect(); col.roll()
        assert vec.kernel_launches - launches == (2 if fused else 2 * T)
        bufs.append((vec, buf, col))
    (va, a, ca), (vb, b, cb) = bufs
    fields = ("obs", "mask", "agent_id", "act", "rew", "terminated", "truncated", "final_obs", "final_mask")
    for name in fields:
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(va.state, vb.state) and va.stats.tolist() == vb.stats.tolist() and a.terminated.any()
    # a fused collection re-emits slot 0 from the state itself: no roll() needed between collections
    last = (a.obs[-1].clone(), a.mask[-1].clone(), a.agent_id[-1].clone())
    ca.collect(); a.obs[0].fill_(7); a.mask[0].fill_(7); a.agent_id[0].fill_(7)          # scribble over slot 0, no roll()
    cb.collect(); cb.roll()
    va2 = va.state.clone()
    ca.collect()
    cb.collect()
    for name in fields:
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert not torch.equal(va2, va.state) and torch.equal(va.state, vb.state)
    assert va.step_count == vb.step_count == 4 * T and not torch.equal(last[0], a.obs[0])
