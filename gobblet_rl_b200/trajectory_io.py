"""Action-log / trajectory files: what makes the bit-exactness claim
auditable.  A log holds the actions the engine took ([T,N] uint8, 255 = none) plus the header needed to
replay them (seed, global env ids, step base, flags) and, optionally, the outputs it produced;
`tools/replay_check.py` feeds the actions through the reference / oracle and reports the first divergence."""
import numpy as np
import torch

FORMAT = "gobblet_b200_action_log_v1"


def save_action_log(path, vec, rollout_out, step_base, include_outputs=True):
    """`rollout_out` = dict returned by `VecEnv.rollout_random(T, ring=T, per_step=True, log_actions=True)`
    started at absolute step `step_base`."""
    T = rollout_out["actions"].shape[0]
    order = [(step_base + t) % T for t in range(T)]            # ring slot of absolute step s is s % ring
    data = {"format": np.array(FORMAT), "actions": rollout_out["actions"].cpu().numpy(),
            "seed": np.array(vec.seed, np.uint64), "env_id_base": np.array(vec.env_id_base, np.uint64),
            "step_base": np.array(step_base, np.uint64), "num_envs": np.array(vec.num_envs, np.int64),
            "illegal_mode": np.array(vec.illegal_mode), "autoreset": np.array(vec.autoreset)}
    if include_outputs:
        for k in ("obs", "mask", "rew", "terminated", "agent_id"):
            if k in rollout_out:
                data[k] = rollout_out[k].cpu().numpy()[order]
        data["stats"] = vec.stats.cpu().numpy()
    np.savez_compressed(path, **data)


def load_action_log(path):
    d = dict(np.load(path, allow_pickle=False))
    if str(d["format"]) != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} file")
    return d


def replay_on_engine(log, device="cuda"):
    """Feed a log's actions back through `step_kernel` (fresh envs): returns the per-step outputs, so that two
    engines / builds / GPU counts can be diffed without any CPU code."""
    from .vec_env import VecEnv
    n = int(log["num_envs"])
    vec = VecEnv(n, device=device, seed=int(log["seed"]), env_id_base=int(log["env_id_base"]),
                 illegal_mode=str(log["illegal_mode"]), autoreset=str(log["autoreset"]), skip255=True)
    outs = {k: [] for k in ("obs", "mask", "rew", "terminated", "agent_id")}
    for a in torch.as_tensor(log["actions"], device=device):
        o, m, r, t, _, ag = vec.step(a)
        for k, v in zip(outs, (o, m, r, t, ag)):
            outs[k].append(v.cpu().numpy().copy())
    return {k: np.stack(v) for k, v in outs.items()}, vec
