"""tests/golden/gpu_rollout_log.npz was written ON A B200 by tools/make_gpu_log_fixture.py: the actions the
fused rollout kernel sampled and every observation / mask / reward / termination it emitted.  Replaying
those actions through the oracle -- and through the UNMODIFIED reference when its tree is present -- must
reproduce the GPU's outputs bit for bit (north_star: "bit-exact ... when replaying identical logged
action sequences").  Runs without a GPU."""
import importlib.util
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(REPO, "tests", "golden", "gpu_rollout_log.npz")

spec = importlib.util.spec_from_file_location("replay_check", os.path.join(REPO, "tools", "replay_check.py"))
replay_check = importlib.util.module_from_spec(spec)
spec.loader.exec_module(replay_check)


def test_gpu_log_replays_through_the_oracle():
    log = replay_check.load(LOG)
    assert log["actions"].shape == (48, 96) and int(log["stats"][0]) > 300 and int(log["stats"][6]) >= 1
    assert replay_check.check_with_oracle(log) == []
    assert replay_check.check_sampler(log, envs=96) == []


def test_gpu_log_replays_through_the_reference_itself():
    log = replay_check.load(LOG)
    res = replay_check.check_with_reference(log, envs=20)
    if res is None:
        pytest.skip("reference tree not present on this box")
    assert res == []


def test_tampered_log_is_caught():
    log = replay_check.load(LOG)
    log["obs"] = log["obs"].copy()
    log["obs"][17, 5, 0, 0, 12] ^= 1
    assert replay_check.check_with_oracle(log) == ["obs: first divergence at step 17, env 5"]


@pytest.mark.gpu
def test_engine_replay_of_the_log_matches_its_outputs():
    from gobblet_rl_b200 import trajectory_io
    log = trajectory_io.load_action_log(LOG)
    outs, vec = trajectory_io.replay_on_engine(log)
    for k in ("obs", "mask", "rew", "terminated", "agent_id"):
        assert np.array_equal(outs[k], log[k]), k
    assert vec.stats.tolist()[:7] == log["stats"].tolist()[:7]
