"""gobblet_v1.env() / raw_env / GreedyGobbletPolicy drop-in surface on the GPU engine. Needs a B200."""
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


def test_env_loop_matches_reference_last_tuples(gv1, golden):
    """The example_basic.py:50-67 loop on env(): every last() tuple equals the trace recorded from the
    reference's gobblet.py (dead steps and illegal-move terminations included)."""
    g = golden("env_wrapped")
    starts = g["game_start"]
    env = gv1.env(render_mode=None)
    for gi in range(len(starts) - 1):
        env.reset()
        i = starts[gi]
        for agent in env.agent_iter():
            obs, reward, term, trunc, info = env.last()
            assert env.possible_agents.index(agent) == g["agent"][i]
            assert np.array_equal(obs["observation"], g["obs"][i]) and obs["observation"].dtype == np.int8
            assert np.array_equal(obs["action_mask"], g["mask"][i]) and obs["action_mask"].dtype == np.int8
            assert float(reward) == g["reward"][i] and term == bool(g["term"][i]) and trunc == bool(g["trunc"][i])
            assert len(env.agents) == g["n_agents"][i]
            env.step(None if g["action"][i] < 0 else int(g["action"][i]))
            i += 1
        assert i == starts[gi + 1] and env.agents == []


def test_raw_env_reset_and_board_view(gv1):
    """tests/test_gobblet_env.py:18-28 of the reference."""
    env = gv1.raw_env(render_mode=None)
    env.reset()
    assert (env.board.squares == np.zeros(27)).all() and env.board.squares.dtype == np.float64
    env.step(18)
    env.step(36)
    assert env.board.squares[9] == 3 and env.board.squares[18] == -5
    assert env.agent_selection == "player_1" and env.turn == 2 and env.action == 36
    other = env.observe("player_2")
    assert not other["action_mask"].any() and other["observation"][..., 12].all()
    assert env._legal_moves() == np.flatnonzero(env.observe("player_1")["action_mask"]).tolist()
    env.render()
    env.close()


def test_conformance_checklist(gv1):
    """what pettingzoo.test.api_test exercises (tests/test_gobblet_env.py:32-34)."""
    env = gv1.env()
    with pytest.raises((AttributeError, AssertionError)):
        env.step(0)                                                   # step before reset
    for _ in range(2):
        env.reset(seed=42)
        assert env.agents == env.possible_agents == ["player_1", "player_2"]
        assert env.agent_selection == "player_1"
        for d in (env.rewards, env.terminations, env.truncations, env.infos):
            assert list(d) == env.agents
    assert env.observation_space("player_1") is env.observation_space("player_1")
    assert env.action_space("player_2") is env.action_space("player_2")
    assert env.action_space("player_1").n == 54
    rng = np.random.default_rng(0)
    for agent in env.agent_iter():
        obs, rew, term, trunc, info = env.last()
        assert env.observation_space(agent).contains(obs)
        assert isinstance(term, bool) and isinstance(trunc, bool) and isinstance(info, dict)
        if term or trunc:
            with pytest.raises(ValueError):
                env.step(3)                                           # dead agents only accept None
            env.step(None)
        else:
            a = rng.choice(np.flatnonzero(obs["action_mask"]))
            env.step([int(a), np.int64(a), np.int32(a), np.array(a)][int(rng.integers(4))])
    assert env.agents == []
    env.reset()
    with pytest.raises(AssertionError):
        env.step(54)                                                  # AssertOutOfBoundsWrapper
    assert env.metadata["name"] == "gobblet_v1" and env.unwrapped.__class__.__name__ == "raw_env"


def test_illegal_move_kats(gv1):
    """after 18, 36, 28, 46 player_1 plays 29 (its medium piece is covered)."""
    env = gv1.env()
    env.reset()
    for a in (18, 36, 28, 46):
        env.step(a)
    env.step(29)
    assert env.rewards == {"player_1": -1.0, "player_2": 0}
    assert all(env.terminations.values()) and all(env.truncations.values())
    assert env.agent_selection == "player_1"
    raw = gv1.raw_env()
    raw.reset()
    for a in (18, 36, 28, 46):
        raw.step(a)
    before = raw.board.squares.copy()
    raw.step(29)
    assert (raw.board.squares == before).all() and raw.agent_selection == "player_2" and raw.turn == 5
    assert raw.rewards == {"player_1": 0, "player_2": 0} and not any(raw.terminations.values())
    zeros = {0, 1, 9, 10, 18, 19, 27, 28, 36, 37, 45, 46}
    assert set(np.flatnonzero(raw.observe("player_2")["action_mask"] == 0)) == zeros


def test_win_kat(gv1):
    env = gv1.env()
    env.reset()
    for a in (0, 21, 10, 31, 20):
        env.step(a)
    assert env.rewards == {"player_1": 1, "player_2": -1} and all(env.terminations.values())
    assert not any(env.truncations.values())
    obs, rew, term, trunc, _ = env.last()
    assert env.agent_selection == "player_2" and rew == -1 and term
    env.step(None)
    assert env.last()[1] == 1
    env.step(None)
    assert env.agents == []


def test_greedy_policy_object_matches_oracle_stream(gv1):
    """GreedyGobbletPolicy.compute_action driven like tutorials/GreedyAgent/tutorial_greedy.py:30-44:
    each returned action must be what the reference algorithm (oracle restatement) allows -- the chosen
    move, or a member of the candidate list when the random fallback fires -- and prev_actions must be
    maintained per agent (greedy_policy.py:219)."""
    np.random.seed(0)
    env = gv1.env()
    for game in range(3):
        env.reset()
        pol = gv1.GreedyGobbletPolicy(depth=2)
        it = 0
        for agent in env.agent_iter():
            obs, rew, term, trunc, _ = env.last()
            if term or trunc:
                env.step(None)
                continue
            idx = env.possible_agents.index(agent)
            if it < 2:
                a = int(np.random.choice(np.flatnonzero(obs["action_mask"])))
            else:
                prev = (pol.prev_actions[idx][-3:] + [-1, -1, -1])[:3]
                chosen, cand, fb = O.greedy(obs["observation"], obs["action_mask"], prev, 2)
                a = pol.compute_action(obs["observation"], obs["action_mask"])
                assert a.shape == () and pol.prev_actions[idx][-1] == int(a)
                assert (int(a) in cand) if fb else (int(a) == chosen)
                a = int(a)
            env.step(a)
            it += 1
