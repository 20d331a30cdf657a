#!/usr/bin/env python
"""bench.py -- env-steps/s of the per-step Gobblet hot path (step + observation + 54-way mask).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N --steps K --warmup W]  # the reference's CPU path

Workload (BASELINE.json config 3, at every N so the 1->8 run is weak scaling): 2^20 lockstep envs per
GPU, uniform random legal actions (Philox4x32-10 keyed by global env id), same-step auto-reset; one
bench "step" = ONE launch of the fused rollout kernel = 64 lockstep env-steps of all envs, the next
observation (117 B) and action mask (54 B) of every env written every env-step into a [64, N, ...]
trajectory buffer (11.5 GB per launch, each byte written once per launch and far larger than the
126 MB L2, so every emitted byte travels to HBM; a short ring would let L2 absorb the rewrites).  One process per GPU;
envs shard by global id with no per-step communication; the only collective is the end-of-run
all-reduce of the episode statistics (outside the timed region).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ENVS_PER_GPU = 1 << 20
FUSED_STEPS = 64
RING = FUSED_STEPS   # one trajectory slot per fused step: every emitted byte is written exactly once per launch
BYTES_PER_ENV_STEP = 117 + 54
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = ("c3: 2^20 lockstep envs per GPU, uniform random legal actions, same-step auto-reset, "
            "obs[3,3,13]+mask[54] int8 emitted every env-step")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--fused-steps", type=int, default=FUSED_STEPS)
    ap.add_argument("--ring", type=int, default=0, help="trajectory slots (0 = one per fused step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--plain-stores", action="store_true", help="st.global instead of st.global.cs")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA-local cores")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=20):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms), "-i", str(gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, windows):
        """windows: [(t0, t1), ...] wall-clock intervals during which the GPU was under the benchmark's load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 8]
        rows = [r for t, r in ok if any(a <= t <= b + 0.03 for a, b in windows)] or [r for _, r in ok]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": reasons, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------
# CPU arms
def _reference_worker(args):
    """The reference's own loop (gobblet_rl/examples/example_basic.py:50-67, render_mode=None) on the
    UNMODIFIED reference package, behind oracle/standins for the missing pettingzoo / gymnasium / pygame."""
    wid, budget_s = args
    import numpy as np
    from oracle import reference_loader as RL
    gob = RL.load_gobblet()
    np.random.seed(wid)
    env = gob.env(render_mode=None)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        env.reset()
        for agent in env.agent_iter():
            obs, reward, term, trunc, info = env.last()
            if term or trunc:
                env.step(None)
            else:
                m = obs["action_mask"]
                env.step(np.random.choice(np.arange(len(m)), p=m / np.sum(m)))
                steps += 1
            if time.perf_counter() - t0 >= budget_s:
                break
    return steps, time.perf_counter() - t0


def _port_worker(args):
    """The C restatement (oracle/gobblet_oracle.c): same per-step work, one core."""
    wid, budget_s = args
    from oracle import oracle as O
    v = O.VecOracle(2048)
    steps, t0, k = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        v.rollout_random(8, seed=wid, step_base=8 * k)
        steps += 2048 * 8
        k += 1
    return steps, time.perf_counter() - t0


def cpu_rate(kind, cores, budget_s, pool=None):
    import multiprocessing as mp
    fn = _reference_worker if kind == "reference" else _port_worker
    jobs = [(i, budget_s) for i in range(cores)]
    if cores == 1:
        res = [fn(jobs[0])]
    elif pool is not None:
        res = pool.map(fn, jobs, chunksize=1)
    else:
        with mp.get_context("spawn").Pool(cores) as p:
            res = p.map(fn, jobs, chunksize=1)
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return steps / wall, steps, wall


def reference_kind():
    from oracle import reference_loader as RL
    return "reference" if RL.available() else "port"


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    kind = reference_kind()
    budget = max(0.5, min(3.0, 150.0 / max(1, a.steps + a.warmup)))
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:       # one worker per host core, reused by every step
        for _ in range(a.warmup):
            cpu_rate(kind, cores, min(budget, 1.0), pool)
        tot_steps, tot_wall = 0, 0.0
        for _ in range(a.steps):
            _, s, w = cpu_rate(kind, cores, budget, pool)
            tot_steps += s
            tot_wall += w
    value = tot_steps / tot_wall
    sample = (f"{a.steps} x {budget:.2f}s of the example_basic.py:50-67 loop per core "
              f"({'unmodified reference gobblet.py+board.py behind stand-in PettingZoo' if kind == 'reference' else 'C oracle port'})")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot_wall / max(1, a.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if kind == "reference" else "i8",
            "data": "synthetic", "config": {"workload": WORKLOAD, "envs": "one env per host core"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


# ---------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit_line(obj):
    """The contract is ONE JSON line on stdout: everything else a library prints (e.g. NCCL's version
    banner) is diverted to stderr by main(); the line goes to the original stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                     # stray prints (NCCL banner, warnings) -> stderr
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from gobblet_rl_b200 import gobblet_v1, sharding

    rank, local, world = sharding.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa_bound = False if a.no_numa_bind else sharding.bind_to_gpu_numa_node(local)
    n, T = a.envs_per_gpu, a.fused_steps
    ring = a.ring if a.ring > 0 else T
    vec = gobblet_v1.vec_env(n, device=dev, seed=0, env_id_base=rank * n, streaming_stores=not a.plain_stores)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(3, a.warmup)):                 # >= 3 untimed warm-up steps
        vec.rollout_random(T, ring=ring)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = vec.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    t_wall0 = time.time()
    ev[0].record()
    for k in range(a.steps):
        vec.rollout_random(T, ring=ring)
        ev[k + 1].record()
    torch.cuda.synchronize(dev)
    t_wall1 = time.time()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    per_launch = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps))
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks
    elapsed_ms = t.item()
    gpu_launches = vec.kernel_launches - launches0
    load_windows = [(t_wall0, t_wall1)]
    stats = sharding.all_reduce_stats(vec.stats)      # the run's only collective (NCCL), untimed

    total_env_steps = world * n * T * a.steps
    value = total_env_steps / (elapsed_ms * 1e-3)

    # ---- end-to-end through host buffers: logged legal actions H2D, obs/mask/rew/flags D2H -------------
    e2e = None
    if not a.no_e2e:
        ne, ke, we = n, 16, 3
        logger = gobblet_v1.vec_env(ne, device=dev, seed=1, env_id_base=rank * ne)
        log = logger.rollout_random(ke + we, emit=False, log_actions=True)["actions"]
        h_log = torch.zeros(log.shape, dtype=torch.uint8, pin_memory=True)
        h_log.copy_(log)
        host = gobblet_v1.HostVecEnv(ne, device=dev, chunks=2, seed=1, env_id_base=rank * ne)
        host.reset()
        for k in range(we):
            host.step(h_log[k])
        sync_all()
        t0, w0 = time.perf_counter(), time.time()
        for k in range(we, we + ke):
            host.step(h_log[k])                       # synchronises: results are in host memory
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        load_windows.append((w0, time.time()))
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        final = torch.cat([e.state for e in host.envs])
        assert torch.equal(final, logger.state), "host-buffer replay diverged from the fused rollout"
        e2e = {"value": world * ne * ke / dt.item(), "unit": UNIT, "h2d_bytes_per_step": world * host.h2d_bytes_per_step,
               "d2h_bytes_per_step": world * host.d2h_bytes_per_step, "lockstep_steps": ke, "envs_per_step": world * ne, "numa_bound": numa_bound,
               "api": "HostVecEnv.step(pinned uint8 actions) -> pinned obs/mask/rew/terminated/truncated/agent_id"}

    # clocks are sampled over the timed launches (and the e2e steps, which keep the GPU busy too)
    clocks = sampler.stop(load_windows) if sampler else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- same-run write-only reference: a plain fill of the same trajectory buffer ---------
    ring_obs, ring_mask = vec._rings[1], vec._rings[2]
    flat = ring_obs.reshape(-1) if ring_obs.is_contiguous() else ring_obs[0].reshape(-1)
    flat = flat[: flat.numel() // 16 * 16].view(torch.int64)          # 8-byte elements: the fill kernel's widest stores
    for _ in range(2):
        flat.fill_(0)
    torch.cuda.synchronize(dev)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        flat.fill_(0)
    f1.record()
    torch.cuda.synchronize(dev)
    fill_gbs = 5 * flat.numel() * 8 / (f0.elapsed_time(f1) * 1e-3) / 1e9

    # ---- roofline of the fused rollout kernel -----------------------------------------------------------
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    launch_ms = elapsed_ms / a.steps
    achieved = BYTES_PER_ENV_STEP * n * T / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("envs") == n:
            traffic = tj["dram_bytes_per_env_step"] * n * T   # ncu --set full capture scaled to this launch
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "gbl::rollout_kernel<true,true,false>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n * T,
                "launch_ms_avg": launch_ms, "launch_ms_min": per_launch[0], "launch_ms_median": per_launch[len(per_launch) // 2],
                "same_run_fill_gbs": fill_gbs, "frac_of_same_run_fill": achieved / fill_gbs,
                "fill_note": f"torch fill_ (int64 view) of the {flat.numel() * 8 / 1e9:.1f} GB observation trajectory buffer, write-only"}

    # ---- config 2 (4096 envs): latency-bound, reported beside the headline --------------------------------
    small = gobblet_v1.vec_env(4096, device=dev, seed=0)
    small.rollout_random(512, ring=4)
    torch.cuda.synchronize(dev)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    small.rollout_random(4096, ring=4)
    s1.record()
    torch.cuda.synchronize(dev)
    small_batch = {"workload": "c2: 4096 lockstep envs, 4096 fused steps, 1 launch", "value": 4096 * 4096 / (s0.elapsed_time(s1) * 1e-3),
                   "unit": UNIT, "note": "0.7 MB/step stays in L2: launch/latency-bound, no HBM roofline applies"}

    # ---- CPU side by side (rank 0, N=1 only): bounded samples on the host cores ---------------------------
    cpu = None
    if world == 1 and not a.no_cpu:
        # parity gate: the throughput above is only reported for a build whose logged
        # actions replay bit-exactly through the CPU checker (and through the reference itself when present)
        gate = gobblet_v1.vec_env(512, device=dev, seed=99)
        got = gate.rollout_random(24, ring=24, per_step=True, log_actions=True)
        from oracle import oracle as O
        from oracle import reference_loader as RL
        ora = O.VecOracle(512)
        want = ora.rollout_random(24, seed=99)
        for key in ("actions", "obs", "mask", "rew", "terminated", "agent_id"):
            assert (got[key].cpu().numpy() == want[key]).all(), f"parity gate failed on {key}"
        assert gate.stats.tolist() == ora.stats.tolist(), "parity gate failed on statistics"
        parity = "512 envs x 24 fused steps bit-exact vs oracle (obs, mask, rew, terminated, agent_id, actions, stats)"
        if RL.available():
            gob = RL.load_gobblet()
            acts = got["actions"].cpu().numpy()
            for e_i in range(0, 512, 128):
                env = gob.raw_env(render_mode=None)
                env.reset()
                for t in range(24):
                    env.step(int(acts[t, e_i]))
                    if env.terminations[env.agent_selection]:
                        env.reset()
                    o = env.observe(env.agent_selection)
                    assert (o["observation"] == want["obs"][t, e_i]).all() and (o["action_mask"] == want["mask"][t, e_i]).all(), \
                        "parity gate failed against the reference"
            parity += "; 4 envs replayed through the unmodified reference"
        kind = reference_kind()
        rate1, s1_, w1_ = cpu_rate(kind, 1, 10.0)
        cpu = {"value": rate1, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{s1_} live steps in {w1_:.1f}s of the example_basic.py:50-67 loop, 1 process",
               "parity_gate": parity}
        if kind == "reference":
            pr, ps, pw = cpu_rate("port", 1, 4.0)
            cpu["port"] = {"value": pr, "unit": UNIT, "cores": 1, "kind": "port",
                           "sample": f"{ps} steps in {pw:.1f}s of oracle/gobblet_oracle.c gbo_rollout_random"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "fused_env_steps_per_launch": T, "ring_slots": ring,
                       "l2": f"each launch writes a {ring * n * BYTES_PER_ENV_STEP / 1e9:.1f} GB trajectory once (>> 126 MB L2, no flush needed)",
                       "parallelism": f"{world} independent shards by global env id, no per-step communication",
                       "stores": "st.global" if a.plain_stores else "st.global.cs"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "small_batch": small_batch,
            "episode_stats": dict(zip(("episodes", "p1_wins", "p2_wins", "steps", "sum_len", "illegal", "both_lines", "max_len"), stats.tolist()))}
    emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
