"""The boundary is a C ABI: a plain C program (tests/c_abi/c_abi_client.c) that includes include/gobblet_b200.h
and links libgobblet_b200.so -- no Python, no torch in the process -- must produce what the oracle produces."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_matches_oracle(tmp_path):
    from gobblet_rl_b200 import ops
    libdir = os.path.dirname(ops.LIB_PATH)
    exe = str(tmp_path / "c_abi_client")
    subprocess.check_call(["nvcc", "-x", "cu", "-o", exe, os.path.join(REPO, "tests", "c_abi", "c_abi_client.c"),
                           f"-L{libdir}", "-lgobblet_b200", f"-Xlinker=-rpath={libdir}"])
    n, T, seed = 1237, 20, 5
    out = subprocess.run([exe, str(n), str(T), str(seed)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    v = O.VecOracle(n)
    v.rollout_random(T, seed=seed, per_step=False)
    obs, mask, rew, term, trunc, agent = v.step(np.zeros(n, np.int64))
    w = lambda a: int((a.reshape(-1).astype(np.int64) * (np.arange(a.size) % 251 + 1)).sum())  # noqa: E731
    assert [int(x) for x in lines["stats"].split()] == v.stats.tolist()
    assert [int(x) for x in lines["sums"].split()] == [w(obs), w(mask), w(rew), int(term.sum())]
    # gbl_step_host: the same step through host buffers (packed wire format + host expander), actions i % 54
    obs, mask, rew, term, trunc, agent = v.step(np.arange(n, dtype=np.int64) % 54)
    assert [int(x) for x in lines["host"].split()] == [w(obs), w(mask), w(rew), int(term.sum()), int(trunc.sum()), int(agent.sum())]
