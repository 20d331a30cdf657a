"""Real NCCL path: two ranks on two GPUs own halves of the global env-id range and reduce their statistics
once at the end; the result must equal one GPU running all envs (skipped with fewer than 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import sys
sys.path.insert(0, {repo!r})
import torch, torch.distributed as dist
from gobblet_rl_b200 import gobblet_v1, sharding
rank, local, world = sharding.init_from_env()
dev = torch.device("cuda", local)
total, T = 200_000, 40
vec = sharding.sharded_vec_env(total, rank, world, device=dev, seed=5)
out = vec.rollout_random(T, ring=1)
stats = sharding.all_reduce_stats(vec.stats)                 # the run's single collective (NCCL)
states = [torch.empty_like(vec.state) for _ in range(world)]
dist.all_gather(states, vec.state)
if rank == 0:
    whole = gobblet_v1.vec_env(total, device=dev, seed=5)
    whole.rollout_random(T, ring=1)
    assert stats.tolist() == whole.stats.tolist(), (stats.tolist(), whole.stats.tolist())
    assert torch.equal(torch.cat(states), whole.state)
    print("OK", stats.tolist())
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_equal_one_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(repo=REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "OK" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices_launches_on_the_tensor_device():
    """Every op makes the device of its tensors current around the launch (`ops._on_device`) and puts it back: an env on
    cuda:1 driven from a process whose current device is cuda:0 plays the same games as one on cuda:0."""
    from gobblet_rl_b200 import gobblet_v1
    assert torch.cuda.current_device() == 0
    a = gobblet_v1.vec_env(5000, device="cuda:0", seed=21)
    b = gobblet_v1.vec_env(5000, device="cuda:1", seed=21)
    oa = a.rollout_random(20, ring=20, per_step=True, log_actions=True)
    ob = b.rollout_random(20, ring=20, per_step=True, log_actions=True)
    assert torch.cuda.current_device() == 0
    for k in oa:
        assert torch.equal(oa[k].cpu(), ob[k].cpu()), k
    acts = oa["mask"][-1].to(torch.float32).argmax(1)
    ra, rb = a.step(acts), b.step(acts.to("cuda:1"))
    assert all(torch.equal(x.cpu(), y.cpu()) for x, y in zip(ra, rb))
    ga = gobblet_v1.greedy_actions(ra[0], ra[1], None, depth=2)
    gb = gobblet_v1.greedy_actions(rb[0], rb[1], None, depth=2)
    assert torch.equal(ga.cpu(), gb.cpu()) and torch.cuda.current_device() == 0
    assert a.stats.tolist() == b.stats.tolist()
