"""Pins oracle/gobblet_oracle.c against fixtures produced by running the reference itself
(oracle/make_golden.py) and against the literal arrays of the reference's own test file."""
import numpy as np

from oracle import oracle as O


def test_reference_test_file_known_answers(golden):
    """tests/test_manual_policy_collector.py:49-509: masks output0..5, _legal_moves output6, board output8."""
    k = golden("reference_kat")
    v = O.VecOracle(1, illegal_mode="pass", autoreset="off")
    _, mask, _ = v.reset()
    assert np.array_equal(mask.astype(bool), k["output0"].astype(bool))
    for i, a in enumerate(k["actions"]):
        _, mask, *_ = v.step([a])
        assert np.array_equal(mask.astype(bool), k[f"output{i + 1}"].astype(bool)), i
    assert np.array_equal(mask.astype(bool), k["output5"].astype(bool))
    assert np.flatnonzero(mask[0]).tolist() == k["output6"].tolist()
    # illegal action 29: board unchanged (test :481-509)
    sq_before = v.squares().copy()
    v.step([int(k["illegal_action"])])
    assert np.array_equal(v.squares(), sq_before)
    assert np.array_equal(v.squares()[0].reshape(3, 3, 3), k["output8"].astype(np.int8))


def test_board_rules_match_reference_games(golden):
    g = golden("board_games")
    starts = g["game_start"]
    for gi in range(len(starts) - 1):
        sq = np.zeros(27, np.int8)
        for i in range(starts[gi], starts[gi + 1]):
            m0 = np.array([O.is_legal(sq, a, 0) for a in range(54)], np.int8)
            m1 = np.array([O.is_legal(sq, a, 1) for a in range(54)], np.int8)
            assert np.array_equal(m0, g["mask_p1"][i]) and np.array_equal(m1, g["mask_p2"][i])
            sq = O.play_turn(sq, int(g["agents"][i]), int(g["actions"][i]))
            assert np.array_equal(sq, g["squares_after"][i])
            assert O.check_for_winner(sq) == g["winner_after"][i]
    # illegal play_turn is a no-op (board.py:125-126)
    for sq, a, ag in zip(g["illegal_squares"], g["illegal_action"], g["illegal_agent"]):
        assert O.is_legal(sq, int(a), int(ag)) == 0
        assert np.array_equal(O.play_turn(sq, int(ag), int(a)), sq)


def _replay_raw(golden, name):
    g = golden(name)
    starts = g["game_start"]
    for gi in range(len(starts) - 1):
        v = O.VecOracle(1, illegal_mode="pass", autoreset="off")
        for i in range(starts[gi], starts[gi + 1]):
            obs, mask, rew, term, trunc, agent = v.step([g["actions"][i]])
            assert np.array_equal(obs[0], g["obs"][i]), (gi, i)
            assert np.array_equal(mask[0], g["mask"][i])
            assert rew[0].tolist() == g["rew"][i].tolist()
            assert bool(term[0]) == bool(g["term"][i]) and bool(trunc[0]) == bool(g["trunc"][i])
            assert int(agent[0]) == int(g["agent_id"][i])
            assert np.array_equal(v.squares()[0], g["squares"][i])
            # the non-selected agent: its own perspective, all-zero mask (gobblet.py:209-213)
            oo, mo = O.observe(v.squares()[0], 1 - int(agent[0]), int(agent[0]))
            assert np.array_equal(oo, g["obs_other"][i]) and not mo.any() and not g["mask_other"][i].any()


def test_raw_env_traces(golden):
    _replay_raw(golden, "env_traces")


def test_raw_env_traces_with_illegal_moves(golden):
    _replay_raw(golden, "env_traces_illegal")


def test_win_kat():
    """0, 21, 10, 31, 20 -> player_1 wins on line (0,1,2); Q3 live terminal mask."""
    v = O.VecOracle(1, illegal_mode="terminate", autoreset="off")
    for a in (0, 21, 10, 31):
        _, _, rew, term, _, _ = v.step([a])
        assert not term[0] and rew[0].tolist() == [0, 0]
    _, mask, rew, term, trunc, agent = v.step([20])
    assert term[0] and not trunc[0] and rew[0].tolist() == [1, -1] and agent[0] == 1
    zeros = set(range(0, 5)) | set(range(9, 14)) | {20, 21, 22} | {29, 30, 31}
    assert set(np.flatnonzero(mask[0] == 0).tolist()) == zeros


def test_greedy_matches_reference(golden):
    g = golden("greedy")
    n_fb = 0
    for i in range(len(g["depth"])):
        chosen, cand, fb = O.greedy(g["obs"][i], g["mask"][i], g["prev3"][i], int(g["depth"][i]))
        assert chosen == int(g["chosen"][i]), i
        assert sorted(cand) == np.flatnonzero(g["cand"][i]).tolist(), i
        assert cand == sorted(cand)
        assert fb == bool(g["fallback"][i]), i
        n_fb += fb
    assert 0 < n_fb < len(g["depth"])


def test_greedy_depth3_matches_reference_and_equals_depth2(golden):
    """The reference's depth-3 branch (greedy_policy.py:160-208), restated literally in the oracle, against what the
    reference itself returned at depth=3 -- and, row by row, against its own depth-2 answers: the branch re-assigns
    the choice depth 2 has made, edits a local list and breaks out of its own loop, nothing else."""
    g3, g = golden("greedy_depth3"), golden("greedy")
    assert len(g3["row"]) >= 16
    for k, i in enumerate(g3["row"]):
        assert int(g["depth"][i]) == 2 and np.array_equal(g3["obs"][k], g["obs"][i]) and np.array_equal(g3["prev3"][k], g["prev3"][i])
        chosen, cand, fb = O.greedy(g3["obs"][k], g3["mask"][k], g3["prev3"][k], 3)
        assert chosen == int(g3["chosen"][k]) and fb == bool(g3["fallback"][k]), k
        assert cand == np.flatnonzero(g3["cand"][k]).tolist(), k
        assert (int(g3["chosen"][k]), bool(g3["fallback"][k])) == (int(g["chosen"][i]), bool(g["fallback"][i]))
        assert np.array_equal(g3["cand"][k], g["cand"][i])
    v = O.VecOracle(300, "terminate", "off")                   # and on fresh positions: literal depth 3 == depth 2
    v.rollout_random(7, seed=31, per_step=False)
    obs, mask, _ = v.observe()
    for i in range(300):
        assert O.greedy(obs[i], mask[i], (-1, -1, -1), 3) == O.greedy(obs[i], mask[i], (-1, -1, -1), 2), i


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert [hex(x) for x in O.philox4x32_10([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_rollout_statistics():
    """SURVEY.md section 0 fingerprint of uniform-random self-play (statistical sanity, not parity)."""
    v = O.VecOracle(2000, autoreset="same_step")
    v.rollout_random(60, seed=1, per_step=False)
    ep, p1, p2, steps, sumlen, illegal, both, maxlen = v.stats.tolist()
    assert illegal == 0 and ep == p1 + p2 and steps == 2000 * 60
    assert 11.3 < sumlen / ep < 12.5
    assert 0.52 < p1 / ep < 0.57
    assert 5 <= maxlen
