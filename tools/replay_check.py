#!/usr/bin/env python
"""Replay a GPU action log through the CPU oracle and, when the reference tree is available, through the
UNMODIFIED reference (`gobblet_rl/game/gobblet.py` raw_env behind oracle/standins), and report the first
divergence of observations / masks / rewards / terminations.  TEST TOOLING (imports oracle/).

    python tools/replay_check.py tests/golden/gpu_rollout_log.npz [--reference-envs 8]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def load(path):
    d = dict(np.load(path, allow_pickle=False))
    assert str(d["format"]) == "gobblet_b200_action_log_v1", "not an action log"
    return d


def first_divergence(name, got, want):
    if np.array_equal(got, want):
        return None
    bad = np.argwhere(got != want)[0]
    return f"{name}: first divergence at step {bad[0]}, env {bad[1]} (index {tuple(int(x) for x in bad)})"


def check_with_oracle(log):
    from oracle import oracle as O
    n, T = int(log["num_envs"]), log["actions"].shape[0]
    v = O.VecOracle(n, str(log["illegal_mode"]), str(log["autoreset"]), skip255=True)
    problems = []
    for t in range(T):
        w = v.step(log["actions"][t].astype(np.int64))
        for name, got in (("obs", 0), ("mask", 1), ("rew", 2), ("terminated", 3), ("agent_id", 5)):
            if name in log and not np.array_equal(log[name][t], w[got]):
                e = np.argwhere(log[name][t] != w[got])[0][0]
                problems.append(f"{name}: first divergence at step {t}, env {e}")
        if problems:
            break
    # the sampler itself: logged action == pick(mask, Philox draw) for the recorded seed / env ids
    return problems


def check_sampler(log, envs=16):
    from oracle import oracle as O
    if str(log["autoreset"]) != "same_step" or "mask" not in log:
        return []
    n, T = int(log["num_envs"]), log["actions"].shape[0]
    seed, base, sb = int(log["seed"]), int(log["env_id_base"]), int(log["step_base"])
    problems = []
    for e in range(0, n, max(1, n // envs)):
        for t in range(1, T):
            want = O.pick(log["mask"][t - 1][e], O.draw(seed, base + e, sb + t))
            if want != int(log["actions"][t][e]):
                problems.append(f"sampler: env {e} step {t}: logged {int(log['actions'][t][e])}, Philox pick {want}")
                return problems
    return problems


def check_with_reference(log, envs):
    """Each env is one reference raw_env driven by its logged actions, reset where the log's terminations say so."""
    from oracle import reference_loader as RL
    if not RL.available():
        return None
    if str(log["autoreset"]) != "same_step" or str(log["illegal_mode"]) != "terminate":
        return []
    gob = RL.load_gobblet()
    n, T = int(log["num_envs"]), log["actions"].shape[0]
    for e in np.linspace(0, n - 1, min(envs, n)).astype(int):
        env = gob.raw_env(render_mode=None)
        env.reset()
        for t in range(T):
            a = int(log["actions"][t][e])
            env.step(a)
            sel = env.agent_selection
            rew = [env.rewards["player_1"], env.rewards["player_2"]]
            term = env.terminations[sel]
            if "rew" in log and (rew != log["rew"][t][e].tolist() or bool(term) != bool(log["terminated"][t][e])):
                return [f"reference: env {e} step {t}: rewards/termination {rew}/{term} vs log {log['rew'][t][e].tolist()}/{bool(log['terminated'][t][e])}"]
            if term:
                env.reset()                          # same-step auto-reset: the log holds the reset observation
                sel = env.agent_selection
            o = env.observe(sel)
            if "obs" in log and not (np.array_equal(o["observation"], log["obs"][t][e]) and np.array_equal(o["action_mask"], log["mask"][t][e])):
                return [f"reference: env {e} step {t}: observation / mask differ"]
    return []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("log")
    ap.add_argument("--reference-envs", type=int, default=8)
    a = ap.parse_args()
    log = load(a.log)
    problems = check_with_oracle(log) + check_sampler(log)
    ref = check_with_reference(log, a.reference_envs)
    print(f"{a.log}: {log['actions'].shape[0]} steps x {int(log['num_envs'])} envs")
    print("oracle replay:", "bit-exact" if not problems else problems)
    print("reference replay:", "reference tree not available" if ref is None else ("bit-exact" if not ref else ref))
    return 1 if problems or ref else 0


if __name__ == "__main__":
    sys.exit(main())
