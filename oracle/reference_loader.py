"""Import the UNMODIFIED reference (board.py / gobblet.py / greedy_policy.py).  TEST INFRASTRUCTURE.

Search order for the reference tree (a directory containing gobblet_rl/):
  1. $GOBBLET_REFERENCE_ROOT
  2. /root/reference                     (the build container)
  3. <repo>/baseline/_ref                (git-ignored copy that travels to the GPU box)
`gobblet.py` needs pettingzoo / gymnasium / pygame (gobblet.py:98-104); when they are not installed
the stand-ins under oracle/standins are put on sys.path first.
"""
import importlib
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
STANDINS = os.path.join(HERE, "standins")


def find_reference_root():
    for root in (os.environ.get("GOBBLET_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(REPO, "baseline", "_ref")):
        if root and os.path.isfile(os.path.join(root, "gobblet_rl", "game", "board.py")):
            return root
    return None


def available():
    return find_reference_root() is not None


def _ensure_paths():
    root = find_reference_root()
    if root is None:
        raise ImportError("reference tree not found (set GOBBLET_REFERENCE_ROOT)")
    if importlib.util.find_spec("pettingzoo") is None and STANDINS not in sys.path:
        sys.path.insert(0, STANDINS)
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def load_board():
    _ensure_paths()
    return importlib.import_module("gobblet_rl.game.board")


def load_greedy():
    _ensure_paths()
    return importlib.import_module("gobblet_rl.game.greedy_policy")


def load_gobblet():
    """gobblet_rl.game.gobblet: env(), raw_env (gobblet.py:110-132)."""
    _ensure_paths()
    return importlib.import_module("gobblet_rl.game.gobblet")
