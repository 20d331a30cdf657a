"""gobblet_rl_b200 -- B200-native batched Gobblet Gobblers engine.

Drop-in for the per-step hot path of elliottower/gobblet-rl (`Board.play_turn / is_legal /
check_for_winner`, `raw_env.step / observe`, the 54-way action mask, the random-legal sampler and the
depth-2 `GreedyGobbletPolicy`), computed by hand-written sm_100a CUDA kernels behind a C ABI
(include/gobblet_b200.h).  There is no CPU fallback: importing `ops` without the CUDA library fails.

    from gobblet_rl_b200 import gobblet_v1
    env = gobblet_v1.env()                      # PettingZoo-style AEC env (reference surface)
    vec = gobblet_v1.vec_env(1 << 20)           # vectorised entry point: torch CUDA tensors
"""
__version__ = "0.1.0"

from . import gobblet_v1  # noqa: E402,F401
