"""Multi-GPU sharding: environments are independent, so ranks own contiguous blocks of global env ids
and exchange NOTHING per step.  The only collective is one end-of-run reduction of the int64[8]
episode statistics (slots 0-6 summed, slot 7 = max episode length maxed): a single 64-byte all-gather over
NCCL (gloo in CPU tests).
Because the Philox counters are keyed by GLOBAL env id, any world size plays the same games."""
import os

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of global env ids owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend=None):
    """One process per GPU, launched by torchrun: reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin this process to the CPU cores NVML reports as local to the GPU, so that pinned host buffers allocated
    afterwards (first touch) live on the GPU's NUMA node -- matters for the host-buffer path (`HostVecEnv`) when
    several ranks stream 50 GB/s each over PCIe.  Returns False when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.replace(",", "").isdigit() else device_index
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:
        return False


def partition_host_cores(rank: int, world_size: int):
    """Give every rank its own slice of the host cores.  After `bind_to_gpu_numa_node` a rank's affinity mask is its
    GPU's NUMA-local core set, which the ranks of that node SHARE; the host-side expander of the packed wire format
    (`HostVecEnv`) runs one thread per core of the mask, so ranks with the same mask split it into contiguous blocks
    (8 GPUs on a 32-core box: 4 cores per rank).  Returns the cores this process is now confined to."""
    mine = sorted(os.sched_getaffinity(0))
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        return mine
    masks = [None] * world_size
    dist.all_gather_object(masks, mine)
    same = [r for r in range(world_size) if masks[r] == mine]
    k, m = same.index(rank), len(same)
    if len(mine) < m:
        return mine                                  # fewer cores than ranks: nothing sensible to split
    per = len(mine) // m
    cores = mine[k * per:(k + 1) * per] if k < m - 1 else mine[k * per:]
    os.sched_setaffinity(0, cores)
    return cores


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Whole-job statistics from per-rank int64[8] vectors -- the run's single collective: ONE all-gather of
    64 bytes per rank (slots 0-6 are summed, slot 7 is a maximum, so a plain SUM all-reduce would not do)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats.clone()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(parts, stats.contiguous(), group=group)       # works on NCCL and gloo alike
    gathered = torch.stack(parts)
    out = gathered.sum(0)
    out[7] = gathered[:, 7].max()
    return out


def sharded_vec_env(total_envs: int, rank: int, world_size: int, device=None, **kw):
    from .vec_env import VecEnv

    lo, hi = shard_range(total_envs, rank, world_size)
    return VecEnv(hi - lo, device=device if device is not None else "cuda", env_id_base=kw.pop("env_id_base", 0) + lo, **kw)
