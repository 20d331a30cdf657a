"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # needs /root/reference (build container only)

The reference's board.py / greedy_policy.py run as they are (numpy only); gobblet.py runs unchanged
behind oracle/standins (pettingzoo / gymnasium / pygame are not installed).  The fixtures are what
pins oracle/gobblet_oracle.c (tests/test_oracle_golden.py) and, through it, the CUDA engine.

Fixtures
  reference_kat.npz   masks output0..5, legal list output6 and board output8 parsed from the literal
                      arrays of tests/test_manual_policy_collector.py:49-509 (the only golden values
                      the reference's own tests hold), plus the action sequence that test plays.
  board_games.npz     random Board-level games: per ply the action, the is_legal mask of BOTH agents
                      before the move, squares after, check_for_winner after; plus illegal attempts.
  env_traces.npz      raw_env (no wrappers) driven like Tianshou's PettingZooEnv: per step obs/mask
                      seen by the next agent, obs of the other agent, rewards, terminations, and the
                      same with illegal actions injected ("pass" semantics, board.py:125-126).
  env_wrapped.npz     env() (wrapper stack, gobblet.py:110-117, stand-in wrappers): full agent_iter
                      loops incl. dead steps and illegal-move termination: the last() 5-tuples.
  render_text.npz     stdout of render_mode="text" / "text_full" along one game (debug views, gobblet.py:299-429).
  greedy.npz          GreedyGobbletPolicy(depth 1 and 2).compute_action on sampled positions with
                      np.random.choice patched to expose (chosen-before-fallback, candidates).
  greedy_depth3.npz   GreedyGobbletPolicy(depth=3) on 16 of those positions (equal to the depth-2 rows, as recorded)
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import reference_loader as RL  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def kat_from_reference_tests(root):
    path = os.path.join(root, "tests", "test_manual_policy_collector.py")
    tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if not (name.startswith("output") and name[6:].isdigit()) or name == "output7":
                continue
            val = node.value
            if isinstance(val, ast.Call):          # np.array([...])
                val = val.args[0]
            found[name] = np.array(ast.literal_eval(val))
    out = {k: v.astype(np.int8) if v.dtype != np.float64 else v for k, v in found.items()}
    out["actions"] = np.array([18, 36, 27 + 1, 45 + 1], np.int64)   # test :112, :179, :247, :314
    out["illegal_action"] = np.array(27 + 2, np.int64)              # test :382
    return out


def board_games(Board, n_games, rng):
    acts, agents, mask0, mask1, sq_after, winner, start = [], [], [], [], [], [], [0]
    ill_sq, ill_act, ill_agent = [], [], []
    for g in range(n_games):
        b = Board()
        agent = 0
        for ply in range(200):
            m = [np.array([b.is_legal(a, ag) for a in range(54)], np.int8) for ag in (0, 1)]
            legal = np.flatnonzero(m[agent])
            illegal = np.flatnonzero(m[agent] == 0)
            if len(illegal) and rng.random() < 0.15:               # play_turn must be a no-op
                a = int(rng.choice(illegal))
                before = b.squares.copy()
                b.play_turn(agent, a)
                assert (b.squares == before).all()
                ill_sq.append(before.astype(np.int8)); ill_act.append(a); ill_agent.append(agent)
            a = int(rng.choice(legal))
            b.play_turn(agent, a)
            acts.append(a); agents.append(agent); mask0.append(m[0]); mask1.append(m[1])
            sq_after.append(b.squares.astype(np.int8)); winner.append(b.check_for_winner())
            agent = 1 - agent
            if b.check_game_over():
                break
        start.append(len(acts))
    return dict(actions=np.array(acts, np.int64), agents=np.array(agents, np.int8),
                mask_p1=np.array(mask0), mask_p2=np.array(mask1), squares_after=np.array(sq_after),
                winner_after=np.array(winner, np.int8), game_start=np.array(start, np.int64),
                illegal_squares=np.array(ill_sq), illegal_action=np.array(ill_act, np.int64),
                illegal_agent=np.array(ill_agent, np.int8))


def raw_env_traces(gob, n_games, rng, p_illegal):
    """Tianshou-style driving of raw_env: env.step(a); obs of new agent_selection; env.rewards."""
    rec = {k: [] for k in ("actions", "obs", "mask", "obs_other", "mask_other", "rew", "term", "trunc",
                           "agent_id", "squares", "turn", "cum_next")}
    start = [0]
    for g in range(n_games):
        env = gob.raw_env(render_mode=None)
        env.reset()
        for t in range(300):
            sel = env.agent_selection
            mask = env.observe(sel)["action_mask"]
            if rng.random() < p_illegal and (mask == 0).any():
                a = int(rng.choice(np.flatnonzero(mask == 0)))
            else:
                a = int(rng.choice(np.flatnonzero(mask)))
            env.step(a)
            nxt = env.agent_selection
            other = [x for x in env.possible_agents if x != nxt][0]
            o, oo = env.observe(nxt), env.observe(other)
            rec["actions"].append(a)
            rec["obs"].append(o["observation"]); rec["mask"].append(o["action_mask"])
            rec["obs_other"].append(oo["observation"]); rec["mask_other"].append(oo["action_mask"])
            rec["rew"].append([env.rewards["player_1"], env.rewards["player_2"]])
            rec["term"].append(env.terminations[nxt]); rec["trunc"].append(env.truncations[nxt])
            rec["agent_id"].append(env.possible_agents.index(nxt))
            rec["squares"].append(env.board.squares.astype(np.int8)); rec["turn"].append(env.turn)
            rec["cum_next"].append(env._cumulative_rewards[nxt])
            if env.terminations[nxt]:
                break
        start.append(len(rec["actions"]))
    out = {k: np.array(v) for k, v in rec.items()}
    out["obs"] = out["obs"].astype(np.int8); out["obs_other"] = out["obs_other"].astype(np.int8)
    out["rew"] = out["rew"].astype(np.int8)
    out["game_start"] = np.array(start, np.int64)
    return out


def wrapped_env_traces(gob, n_games, rng, p_illegal):
    """The example_basic.py:50-67 loop on env(): every last() tuple, dead steps included."""
    rec = {k: [] for k in ("agent", "obs", "mask", "reward", "term", "trunc", "action", "n_agents")}
    start = [0]
    for g in range(n_games):
        env = gob.env(render_mode=None)
        env.reset()
        for agent in env.agent_iter():
            obs, reward, term, trunc, info = env.last()
            rec["agent"].append(env.possible_agents.index(agent))
            rec["obs"].append(obs["observation"]); rec["mask"].append(obs["action_mask"])
            rec["reward"].append(float(reward)); rec["term"].append(term); rec["trunc"].append(trunc)
            rec["n_agents"].append(len(env.agents))
            if term or trunc:
                rec["action"].append(-1)
                env.step(None)
            else:
                mask = obs["action_mask"]
                if rng.random() < p_illegal and (mask == 0).any():
                    a = int(rng.choice(np.flatnonzero(mask == 0)))
                else:
                    a = int(rng.choice(np.flatnonzero(mask)))
                rec["action"].append(a)
                env.step(a)
        start.append(len(rec["action"]))
    out = {k: np.array(v) for k, v in rec.items()}
    out["obs"] = out["obs"].astype(np.int8)
    out["game_start"] = np.array(start, np.int64)
    return out


def greedy_cases(Board, gp_mod, n_pos, rng):
    """Positions from random play (any ply, finished games included), random prev-action history."""
    import numpy.random as npr

    captured = {}
    real_choice = npr.choice

    def fake_choice(a, *args, **kw):
        captured["cand"] = [int(x) for x in a]
        return a[0]

    def observe(b, agent):
        board = b.squares.reshape(3, 3, 3) * (-1 if agent == 1 else 1)
        layers = [board[(i - 1) // 2] == i for i in range(1, 7)] + [board[(i - 1) // 2] == -i for i in range(1, 7)]
        layers.append(np.ones((3, 3)) if agent == 1 else np.zeros((3, 3)))
        obs = np.stack(layers, axis=2).astype(np.int8)
        mask = np.array([b.is_legal(a, agent) for a in range(54)], np.int8)
        return obs, mask

    rec = {k: [] for k in ("obs", "mask", "prev3", "depth", "chosen", "cand", "fallback", "returned")}
    while len(rec["depth"]) < n_pos:
        b = Board()
        agent = 0
        hist = {0: [], 1: []}
        nply = int(rng.integers(0, 16))
        for ply in range(nply):
            legal = [a for a in range(54) if b.is_legal(a, agent)]
            a = int(rng.choice(legal))
            b.play_turn(agent, a)
            hist[agent].append(a)
            agent = 1 - agent
            if b.check_game_over() and rng.random() < 0.9:
                break
        obs, mask = observe(b, agent)
        def run(depth, prev):
            pol = gp_mod.GreedyGobbletPolicy(depth=depth)
            pol.prev_actions[agent] = list(prev)
            captured.clear()
            npr.choice = fake_choice
            gp_mod.np.random.choice = fake_choice
            try:
                act = int(pol.compute_action(obs, mask))
            finally:
                npr.choice = real_choice
                gp_mod.np.random.choice = real_choice
            assert pol.prev_actions[agent][-1] == act                      # greedy_policy.py:219
            return act, captured.get("cand")

        if not mask.any():
            continue
        for depth in (1, 2):
            act0, cand0 = run(depth, [])                  # empty history: fallback iff chosen is None
            chosen = -1 if cand0 is not None else act0
            if cand0 is None:                             # force the fallback to expose the candidates
                _, cand0 = run(depth, [chosen] * 3)
            for variant in range(2):
                if variant == 0:
                    prev = list(hist[agent][-3:])
                else:                                     # history that trips the repetition rule
                    prev = [int(x) for x in rng.choice(np.flatnonzero(mask), size=3)]
                    if chosen >= 0 and rng.random() < 0.5:
                        prev[int(rng.integers(0, 3))] = chosen
                fb = chosen < 0 or chosen in prev[-3:]
                cand = np.zeros(54, np.int8)
                cand[cand0] = 1
                rec["obs"].append(obs.reshape(117)); rec["mask"].append(mask)
                rec["prev3"].append((prev + [-1, -1, -1])[:3])
                rec["depth"].append(depth)
                rec["chosen"].append(chosen)
                rec["cand"].append(cand); rec["fallback"].append(fb); rec["returned"].append(act0)
    out = {k: np.array(v) for k, v in rec.items()}
    out["prev3"] = out["prev3"].astype(np.int16)
    return out


def greedy_depth3_cases(gp_mod, cases, n_pos):
    """GreedyGobbletPolicy(depth=3) on the first `n_pos` depth-2 rows of `cases` (same observation, mask and history).
    The depth-3 branch (greedy_policy.py:160-208) only re-assigns `chosen_action = action` (already assigned at :157),
    edits a local list and breaks out of its own inner loop, so its results must equal the depth-2 rows -- the fixture
    records what the reference actually returns so that this is a recorded fact, not an argument."""
    import numpy.random as npr

    captured = {}
    real_choice = npr.choice

    def fake_choice(a, *args, **kw):
        captured["cand"] = [int(x) for x in a]
        return a[0]

    rows = [i for i in range(len(cases["depth"])) if cases["depth"][i] == 2]
    rows = rows[::max(1, len(rows) // n_pos)][:n_pos]     # spread over the set; a depth-3 call costs up to a minute
    rec = {k: [] for k in ("row", "obs", "mask", "prev3", "chosen", "cand", "fallback", "returned")}
    for i in rows:
        obs, mask = cases["obs"][i].reshape(3, 3, 13), cases["mask"][i]
        agent = int(obs[0, 0, 12])
        prev = [int(x) for x in cases["prev3"][i] if x >= 0]

        def run(history):
            pol = gp_mod.GreedyGobbletPolicy(depth=3)
            pol.prev_actions[agent] = list(history)
            captured.clear()
            npr.choice = fake_choice
            gp_mod.np.random.choice = fake_choice
            try:
                act = int(pol.compute_action(obs, mask))
            finally:
                npr.choice = real_choice
                gp_mod.np.random.choice = real_choice
            return act, captured.get("cand")

        act, cand0 = run(prev)
        if cand0 is None:                                 # no fallback: the return value IS the choice;
            chosen = act                                  # force the fallback once to expose the candidates
            _, cand0 = run([chosen] * 3)
        else:                                             # fallback fired: empty history tells whether the choice was None
            act_free, cand_free = run([])
            chosen = -1 if cand_free is not None else act_free
        cand = np.zeros(54, np.int8)
        cand[cand0] = 1
        rec["row"].append(i); rec["obs"].append(cases["obs"][i]); rec["mask"].append(mask)
        rec["prev3"].append(cases["prev3"][i]); rec["chosen"].append(chosen); rec["cand"].append(cand)
        rec["fallback"].append(chosen < 0 or chosen in prev[-3:]); rec["returned"].append(act)
    out = {k: np.array(v) for k, v in rec.items()}
    out["prev3"] = out["prev3"].astype(np.int16)
    return out


def render_text_cases(gob):
    """stdout of the reference's text / text_full renderers (gobblet.py:299-429) along one game."""
    import contextlib
    import io
    out = {}
    actions = [18, 36, 28, 46, 2, 40, 13, 47, 26]
    out["actions"] = np.array(actions, np.int64)
    for mode in ("text", "text_full"):
        env = gob.raw_env(render_mode=mode)
        env.reset()
        chunks = []
        for a in actions:
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                env.step(a)
            chunks.append(buf.getvalue())
        out[mode] = np.array(chunks)
    return out


def main():
    root = RL.find_reference_root()
    assert root, "reference not found"
    os.makedirs(OUT, exist_ok=True)
    Board = RL.load_board().Board
    gob = RL.load_gobblet()
    gp = RL.load_greedy()
    rng = np.random.default_rng(20261018)
    np.savez_compressed(os.path.join(OUT, "reference_kat.npz"), **kat_from_reference_tests(root))
    np.savez_compressed(os.path.join(OUT, "board_games.npz"), **board_games(Board, 150, rng))
    np.savez_compressed(os.path.join(OUT, "env_traces.npz"), **raw_env_traces(gob, 60, rng, 0.0))
    np.savez_compressed(os.path.join(OUT, "env_traces_illegal.npz"), **raw_env_traces(gob, 40, rng, 0.2))
    np.savez_compressed(os.path.join(OUT, "env_wrapped.npz"), **wrapped_env_traces(gob, 40, rng, 0.04))
    cases = greedy_cases(Board, gp, 480, rng)
    np.savez_compressed(os.path.join(OUT, "greedy.npz"), **cases)
    np.savez_compressed(os.path.join(OUT, "greedy_depth3.npz"), **greedy_depth3_cases(gp, cases, 16))
    np.savez_compressed(os.path.join(OUT, "render_text.npz"), **render_text_cases(gob))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
