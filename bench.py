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

Beside the headline the line carries
  e2e      the same step through HOST buffers (`HostVecEnv.step` = ONE C-ABI call, gbl_step_host): pinned actions
           H2D, packed 24-byte records D2H, expanded into the reference-shaped int8 arrays by the library's host
           thread pool; with its own roofline (same-run pinned-D2H and host-fill probes) and the round-1 dense
           path (176 B/env over PCIe) measured beside it;
  configs  BASELINE.json configs 2 (4096 envs), 4 (65 536 boards, greedy depth 2) and 5 (collection of
           131 072 envs x 16 steps per GPU into a trajectory buffer), each behind a small oracle gate;
  multi_gpu_equality (N > 1)  rank 0 replays the first 4096 global env ids of every other rank and compares states.
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
HOST_BYTES_PER_ENV_STEP = 117 + 54 + 2 + 1 + 1 + 1
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
WORKLOAD = ("c3: 2^20 lockstep envs per GPU, uniform random legal actions, same-step auto-reset, "
            "obs[3,3,13]+mask[54] int8 emitted every env-step")
SM_COUNT, SCHEDULERS_PER_SM = 148, 4


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
    ap.add_argument("--no-configs", action="store_true", help="skip the c2 / c4 / c5 side measurements")
    ap.add_argument("--plain-stores", action="store_true", help="st.global instead of st.global.cs")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA-local cores")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="D2H chunks of the packed e2e path (0 = HostVecEnv's schedule)")
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


def device_time(fn, iters, warmup=2):
    """seconds per call, CUDA events on the current stream, synchronised on both sides"""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


# ---------------------------------------------------------------------------------------------------
def measure_e2e(a, ctx):
    """End to end through host buffers.  Every rank runs its own shard; value = all envs / max-over-ranks time."""
    import torch
    import torch.distributed as dist
    from gobblet_rl_b200 import gobblet_v1, ops
    dev, rank, world, n = ctx["dev"], ctx["rank"], ctx["world"], ctx["n"]
    sync_all, windows = ctx["sync_all"], ctx["load_windows"]
    ne, ke, we = n, 16, 3
    threads = ctx["host_threads"]

    def gmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [x]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [p.item() for p in parts]

    logger = gobblet_v1.vec_env(ne, device=dev, seed=1, env_id_base=rank * ne)
    log = logger.rollout_random(ke + we, emit=False, log_actions=True)["actions"]
    h_log = torch.zeros(log.shape, dtype=torch.uint8, pin_memory=True)
    h_log.copy_(log)
    torch.cuda.synchronize(dev)

    def run(host, reps=3):
        """median over `reps` passes of: reset, `we` warm-up steps, `ke` timed lockstep steps replaying the logged actions
        (the host side shares its memory system with whatever else runs on the box: one 20 ms pass is noisy)"""
        dts = []
        for _ in range(reps):
            host.reset()
            host.env.stats.zero_()
            for k in range(we):
                host.step(h_log[k])
            sync_all()
            t0, w0 = time.perf_counter(), time.time()
            for k in range(we, we + ke):
                host.step(h_log[k])                   # returns with the results in host memory
            torch.cuda.synchronize(dev)
            dts.append(gmax(time.perf_counter() - t0))
            windows.append((w0, time.time()))
            assert torch.equal(host.env.state, logger.state), "host-buffer replay diverged from the fused rollout"
        return sorted(dts)[len(dts) // 2], dts

    res = {}
    chunks = a.e2e_chunks if a.e2e_chunks > 0 else None
    variants = [("packed", dict(wire="packed", chunks=chunks, host_threads=threads)),
                ("dense", dict(wire="dense", chunks=2)),
                ("packed_consumer", dict(wire="packed", chunks=chunks, expand=False))]
    for name, kw in variants:
        host = gobblet_v1.HostVecEnv(ne, device=dev, seed=1, env_id_base=rank * ne, **kw)
        l0 = host.kernel_launches
        dt, dts = run(host)
        res[name] = {"dt": dt, "dts": dts, "h2d": host.h2d_bytes_per_step, "d2h": host.d2h_bytes_per_step, "chunks": len(host.parts),
                     "launches": (host.kernel_launches - l0) // (3 * (we + ke))}
        if name == "packed":                          # spot-check the expanded arrays against the device path
            obs_d, mask_d, _ = host.env.observe()
            assert torch.equal(host.h_obs[:4096], obs_d[:4096].cpu()) and torch.equal(host.h_mask[-4096:], mask_d[-4096:].cpu())
            # ceiling probe: the expander alone on records already resident in host memory (no PCIe), all ranks at once
            outs = (host.h_obs, host.h_mask, host.h_rew, host.h_term.view(torch.uint8), host.h_trunc.view(torch.uint8), host.h_agent)
            ops.host_unpack(host.h_rec, *outs, threads=threads)
            sync_all()
            t0 = time.perf_counter()
            for _ in range(4):
                ops.host_unpack(host.h_rec, *outs, threads=threads)
            expander_alone = ne * HOST_BYTES_PER_ENV_STEP * 4 / (time.perf_counter() - t0) / 1e9
        del host

    # ---- ceilings, same run: pinned D2H (one plain cudaMemcpyAsync of the same size) and the pool's fill bandwidth
    def d2h_gbs(nbytes, reps=4):
        d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        h = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True)
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize(dev)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize(dev)
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    def fill_gbs(nbytes, reps=4):
        h = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True)
        ops.host_fill(h, threads, 0)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            ops.host_fill(h, threads, 0)
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    dense_bytes, packed_bytes = ne * HOST_BYTES_PER_ENV_STEP, ne * 24
    d2h_conc = gather(d2h_gbs(dense_bytes))
    d2h_packed_conc = gather(d2h_gbs(packed_bytes))
    fill_conc = gather(fill_gbs(dense_bytes))
    d2h_alone = fill_alone = None
    if world > 1:                                     # rank 0 alone: what one link / one pool does without neighbours
        dist.barrier()
        if rank == 0:
            d = torch.zeros(dense_bytes, dtype=torch.uint8, device=dev)
            h = torch.zeros(dense_bytes, dtype=torch.uint8, pin_memory=True)
            h.copy_(d, non_blocking=True)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(4):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize(dev)
            d2h_alone = 4 * dense_bytes / (time.perf_counter() - t0) / 1e9
            ops.host_fill(h, threads, 0)
            t0 = time.perf_counter()
            for _ in range(4):
                ops.host_fill(h, threads, 0)
            fill_alone = 4 * dense_bytes / (time.perf_counter() - t0) / 1e9
        dist.barrier()

    exp_conc = gather(expander_alone)
    total = world * ne
    p, dn, pc = res["packed"], res["dense"], res["packed_consumer"]
    value = total / (p["dt"] / ke)
    host_gbs = value * HOST_BYTES_PER_ENV_STEP / 1e9
    peak_gbs, peak_src = max((sum(exp_conc), "the expander alone on host-resident records (no PCIe), same thread pools, all ranks at once"),
                             (sum(fill_conc), "non-temporal fill of a pinned buffer by the same thread pools, all ranks at once"))
    e2e = {"value": value, "unit": UNIT, "h2d_bytes_per_step": world * p["h2d"], "d2h_bytes_per_step": world * p["d2h"],
           "host_bytes_delivered_per_step": world * dense_bytes, "lockstep_steps": ke, "envs_per_step": total,
           "ms_per_lockstep_step": 1e3 * p["dt"] / ke, "numa_bound": ctx["numa_bound"], "host_threads_per_rank": ops.host_threads(threads),
           "host_simd": ops.host_simd(), "chunks": p["chunks"],
           "timing": f"median of 3 passes of {ke} lockstep steps (each after {we} warm-up steps), max over ranks per pass",
           "passes_env_steps_per_s": [total / (x / ke) for x in p["dts"]],
           "api": ("HostVecEnv.step(pinned uint8 actions) -> pinned obs[N,3,3,13]/mask[N,54]/rew/terminated/truncated/agent_id "
                   "= one C-ABI call gbl_step_host: H2D actions, step kernel -> 24-B packed records, D2H, host thread pool expands"),
           "roofline": {"bound": "host DRAM write bandwidth: 176 B/env-step of int8 arrays written by the host cores (+ 24 B/env-step of "
                                 "records written by the DMA engine and read back by the expander); PCIe carries 24 B/env-step",
                        "achieved_gbs": host_gbs, "peak_gbs": peak_gbs, "frac": host_gbs / peak_gbs, "peak_source": "same run: " + peak_src,
                        "host_dram_traffic_gbs": value * (HOST_BYTES_PER_ENV_STEP + 48) / 1e9,
                        "probe_expander_alone_gbs_sum": sum(exp_conc), "probe_nt_fill_gbs_sum": sum(fill_conc),
                        "probe_nt_fill_gbs_per_rank": fill_conc, "probe_nt_fill_gbs_rank0_alone": fill_alone,
                        "pcie_achieved_gbs": value * 24 / 1e9, "pcie_d2h_probe_gbs_sum": sum(d2h_packed_conc),
                        "note": "the box's host memory is shared by all ranks: the ceiling does not grow with the GPU count"},
           "dense_wire": {"value": total / (dn["dt"] / ke), "unit": UNIT, "d2h_bytes_per_step": world * dn["d2h"],
                          "ms_per_lockstep_step": 1e3 * dn["dt"] / ke,
                          "roofline": {"bound": "pcie d2h", "achieved_gbs": total / (dn["dt"] / ke) * HOST_BYTES_PER_ENV_STEP / 1e9,
                                       "peak_gbs": sum(d2h_conc),
                                       "frac": total / (dn["dt"] / ke) * HOST_BYTES_PER_ENV_STEP / 1e9 / sum(d2h_conc),
                                       "peak_source": "same-run pinned cudaMemcpyAsync D2H of the same bytes, all ranks at once",
                                       "per_rank_d2h_gbs": d2h_conc, "d2h_gbs_rank0_alone": d2h_alone},
                          "note": "round-1 path: expanded tensors cross PCIe (176 B/env)"},
           "packed_consumer": {"value": total / (pc["dt"] / ke), "unit": UNIT, "d2h_bytes_per_step": world * pc["d2h"],
                               "ms_per_lockstep_step": 1e3 * pc["dt"] / ke,
                               "note": "HostVecEnv(expand=False): the consumer reads the 24-byte records (bits) itself"}}
    e2e["gpu_launches_per_lockstep_step"] = p["launches"]
    return e2e, world * p["launches"] * ke


# ---------------------------------------------------------------------------------------------------
def measure_configs(a, ctx):
    """BASELINE configs 2, 4, 5 beside the headline (c3).  c5 runs on every rank (weak scaling: 131 072 envs x 16
    steps per GPU, 16.8 M env-steps at 8 GPUs); c2 and c4 are single-GPU configs and run on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from gobblet_rl_b200 import adapters, gobblet_v1
    from oracle import oracle as O
    dev, rank, world, peak = ctx["dev"], ctx["rank"], ctx["world"], ctx["peak"]
    out = {}

    # ---- c5: collection into a trajectory buffer ------------------------------------------------------------
    n5, T5 = 131072, 16

    def collector(n, T, keep_final, seed=3, base=0):
        vec = gobblet_v1.vec_env(n, device=dev, seed=1, env_id_base=base)
        buf = adapters.TrajectoryBuffer(T, n, device=dev, keep_final=keep_final)
        return vec, buf, adapters.VecCollector(vec, adapters.RandomLegalPolicy(seed=seed, env_id_base=base), buf)

    if rank == 0:     # gate: a small collection replays bit-exactly through the oracle (terminal observations included)
        vec, buf, col = collector(512, 8, True)
        col.collect()
        o = O.VecOracle(512, "terminate", "same_step")
        o.reset()
        acts = buf.act.cpu().numpy()
        for t in range(8):
            w = o.step(acts[t].astype(np.int64), want_final=True)
            ok = (np.array_equal(buf.obs[t + 1].cpu().numpy(), w[0]) and np.array_equal(buf.mask[t + 1].cpu().numpy(), w[1])
                  and np.array_equal(buf.rew[t].cpu().numpy(), w[2]) and np.array_equal(buf.terminated[t].cpu().numpy(), w[3])
                  and np.array_equal(buf.agent_id[t + 1].cpu().numpy(), w[5]) and np.array_equal(buf.final_obs[t].cpu().numpy(), w[6])
                  and np.array_equal(buf.final_mask[t].cpu().numpy(), w[7]))
            assert ok, f"c5 gate: collected trajectory diverges from the oracle replay at step {t}"
        del vec, buf, col
    c5 = {"workload": f"c5: {n5} envs x {T5} lockstep steps per GPU collected into a TrajectoryBuffer (obs, mask, rew, terminated, "
                      f"agent_id, action per step) by VecCollector, masked-uniform policy, {world} GPU(s)",
          "env_steps_total": world * n5 * T5, "gate": "512 envs x 8 steps == oracle replay (terminal observations included)"}
    for key, keep_final, per_step_bytes in (("no_final", False, 117 + 54 + 2 + 1 + 1 + 1), ("with_final_obs", True, 2 * 171 + 5)):
        vec, buf, col = collector(n5, T5, keep_final, base=rank * n5)

        def once():
            col.collect()                             # fused: one launch, slot 0 re-emitted from the state (no roll() copy needed)
            if not col.fused:
                col.roll()
        ctx["sync_all"]()
        dt = device_time(once, 10)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
        rate = world * n5 * T5 / dt
        gbs = n5 * T5 * per_step_bytes / dt / 1e9
        c5[key] = {"value": rate, "unit": UNIT, "ms_per_collection": dt * 1e3, "launches_per_collection": 1 if col.fused else 3 * T5,
                   "algorithmic_bytes_per_env_step": per_step_bytes,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "per": "GPU"}}
        del vec, buf, col
    out["c5"] = c5
    if rank != 0:
        return out

    # ---- c2: 4096 lockstep envs (latency-bound: 0.7 MB per step stays in L2) -----------------------------------
    n2, T2 = 4096, 4096
    small = gobblet_v1.vec_env(n2, device=dev, seed=0)
    per_block = {}
    for hint in (0, 32, 64, 128, 256):
        dt = device_time(lambda: small.rollout_random(T2, ring=4, block_hint=hint), 3, warmup=1)
        per_block["auto" if hint == 0 else str(hint)] = dt / T2 * 1e6
    best = min(per_block.values())
    # gate: the timed instance (chosen from N: two warps per 32 envs, no aux outputs) == the 256-thread instance on the
    # device, and its logged twin == the CPU oracle, 4096 envs x 48 steps
    ga, gb, gc = (gobblet_v1.vec_env(n2, device=dev, seed=5) for _ in range(3))
    oa = ga.rollout_random(48, ring=48)
    ob = gb.rollout_random(48, ring=48, block_hint=256)
    oc = gc.rollout_random(48, ring=48, per_step=True, log_actions=True)
    assert torch.equal(oa["obs"], ob["obs"]) and torch.equal(oa["mask"], ob["mask"]) and torch.equal(ga.state, gb.state), "c2 gate"
    assert torch.equal(oa["obs"], oc["obs"]) and torch.equal(oa["mask"], oc["mask"]) and torch.equal(ga.stats, gc.stats), "c2 gate"
    ora = O.VecOracle(n2)
    want = ora.rollout_random(48, seed=5)
    for key in ("actions", "obs", "mask", "rew", "terminated", "agent_id"):
        assert np.array_equal(oc[key].cpu().numpy(), want[key]), f"c2 gate: {key} differs from the oracle"
    assert gc.stats.tolist() == ora.stats.tolist(), "c2 gate: statistics differ from the oracle"
    del ga, gb, gc, oa, ob, oc
    gsmall = gobblet_v1.vec_env(n2, device=dev, seed=0, graph_safe=True)

    def per_step_launches():
        for _ in range(256):
            gsmall.rollout_random(1, ring=1)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        per_step_launches()
        with torch.cuda.graph(g2, stream=side):
            per_step_launches()
    torch.cuda.current_stream(dev).wait_stream(side)
    dt_graph = device_time(g2.replay, 5) / 256
    out["c2"] = {"workload": "c2: 4096 lockstep envs, random legal actions, same-step auto-reset, obs+mask emitted every step",
                 "value": n2 / (per_block["auto"] * 1e-6), "unit": UNIT, "us_per_lockstep_step": per_block["auto"],
                 "us_per_lockstep_step_by_block_threads": per_block, "best_us": best,
                 "us_per_lockstep_step_cuda_graph_of_1step_launches": dt_graph * 1e6,
                 "hbm_time_us": n2 * BYTES_PER_ENV_STEP / (peak * 1e9) * 1e6,
                 "gate": "4096 envs x 48 steps: timed instance == 256-thread instance == logged instance (device), logged instance == oracle",
                 "note": "one fused launch of 4096 steps; latency-bound (a dependent chain per warp; two warps per 32 envs share "
                         "the emission: observation warp / mask warp), the HBM time is shown for scale"}

    # ---- c4: greedy depth 2 on 65 536 boards (warp per board; issue-bound) ------------------------------------------
    src = gobblet_v1.vec_env(1 << 18, device=dev, seed=7, autoreset="off")
    bo, bm = [], []
    for plies in range(2, 13, 2):
        src.rollout_random(2, emit=False)
        obs, mask, _ = src.observe()
        live = (src.state[:, 0] >> 55 & 1) == 0
        bo.append(obs[live][:11000].clone())
        bm.append(mask[live][:11000].clone())
    gobs, gmask = torch.cat(bo)[:65536].contiguous(), torch.cat(bm)[:65536].contiguous()
    nb = gobs.shape[0]
    act, chosen, cand, fb = gobblet_v1.greedy_actions(gobs, gmask, None, depth=2, details=True)
    idx = torch.arange(0, nb, nb // 256, device=dev)[:256]
    o_np, m_np = gobs[idx].cpu().numpy(), gmask[idx].cpu().numpy()
    for j, i in enumerate(idx.tolist()):
        wc, wcand, wfb = O.greedy(o_np[j], m_np[j], (-1, -1, -1), 2)
        assert (int(chosen[i]), bool(fb[i])) == (wc, wfb) and [k for k in range(54) if (int(cand[i]) >> k) & 1] == wcand, \
            "c4 gate: greedy move choice differs from the oracle"
    dt = device_time(lambda: gobblet_v1.greedy_actions(gobs, gmask, None, depth=2), 20)
    issue_peak = SM_COUNT * SCHEDULERS_PER_SM * ctx["sm_max_mhz"] * 1e6
    c4 = {"workload": "c4: GreedyGobbletPolicy depth 2 for 65 536 boards at plies 2..12 (one warp per board)", "boards": nb,
          "value": nb / dt, "unit": "boards/s", "us_per_launch": dt * 1e6, "gate": "256 boards: chosen / candidates / fallback == oracle"}
    gpath = os.path.join(REPO, "profiles", "greedy_issue.json")
    if os.path.exists(gpath):
        gj = json.load(open(gpath))
        wi = gj["warp_instructions_per_board"]
        c4["roofline"] = {"bound": "issue", "warp_instr_per_board": wi, "achieved": wi * nb / dt, "peak": issue_peak,
                          "unit": "warp-instr/s", "frac": wi * nb / dt / issue_peak,
                          "source": "profiles/greedy_issue.json (static ncu capture of this kernel: smsp__inst_executed.sum / boards); "
                                    "peak = 148 SMs x 4 schedulers x SM max clock"}
    out["c4"] = c4
    return out


# ---------------------------------------------------------------------------------------------------
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
    my_cores = sharding.partition_host_cores(rank, world)          # ranks sharing a NUMA node split its cores
    n, T = a.envs_per_gpu, a.fused_steps
    ring = a.ring if a.ring > 0 else T
    vec = gobblet_v1.vec_env(n, device=dev, seed=0, env_id_base=rank * n, streaming_stores=not a.plain_stores)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    W = max(3, a.warmup)
    for _ in range(W):                                # >= 3 untimed warm-up steps
        vec.rollout_random(T, ring=ring)

    # ---- multi-GPU equality (N > 1): rank 0 replays the first 4096 global env ids of EVERY other rank through the
    #      warm-up steps and compares with their states -- "any world size plays the same games", on NCCL ----------
    multi_eq = None
    if world > 1:
        k = min(4096, n)
        mine = vec.state[:k].contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            for r in range(1, world):
                twin = gobblet_v1.vec_env(k, device=dev, seed=0, env_id_base=r * n)
                for _ in range(W):
                    twin.rollout_random(T, emit=False)
                assert torch.equal(twin.state, parts[r]), f"rank {r}: shard state differs from a single-GPU replay of the same global env ids"
            multi_eq = {"ranks_checked": list(range(1, world)), "envs_per_rank": k, "lockstep_steps": W * T, "equal": True,
                        "how": "NCCL all_gather of state[:4096] after the warm-up launches == rank 0's replay of the same global env ids"}
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
    gpu_launches = world * (vec.kernel_launches - launches0)      # whole job: every rank launches the same kernels
    load_windows = [(t_wall0, t_wall1)]
    total_env_steps = world * n * T * a.steps
    value = total_env_steps / (elapsed_ms * 1e-3)

    # ---- sustained: the same launch back to back for >= 1.2 s (the driver's K may be a few tens of ms) ----------
    k_sus = max(a.steps, int(1.2 / max(1e-6, elapsed_ms * 1e-3 / a.steps)) + 1)
    sync_all()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    s0.record()
    for _ in range(k_sus):
        vec.rollout_random(T, ring=ring)
    s1.record()
    torch.cuda.synchronize(dev)
    load_windows.append((w0, time.time()))
    ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sustained = {"value": world * n * T * k_sus / (ts.item() * 1e-3), "unit": UNIT, "launches": k_sus, "seconds": ts.item() * 1e-3}
    stats = sharding.all_reduce_stats(vec.stats)      # the run's only data collective (NCCL), untimed

    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    sm_max_mhz = 1965.0
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_src, sm_max_mhz = pk["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (copy, read+write)", pk.get("sm_max_mhz", 1965.0)
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ctx = {"dev": dev, "rank": rank, "world": world, "n": n, "sync_all": sync_all, "load_windows": load_windows,
           "numa_bound": numa_bound, "host_threads": len(my_cores) if my_cores else 0, "peak": peak, "sm_max_mhz": sm_max_mhz}

    # ---- end-to-end through host buffers ---------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        ring_bufs = vec._rings
        e2e, e2e_launches = measure_e2e(a, ctx)
        e2e["host_cores_per_rank"] = len(my_cores) if my_cores else len(os.sched_getaffinity(0))

    # ---- BASELINE configs 2 / 4 / 5 (c5 on every rank) ---------------------------------------------------------
    configs = None
    if not a.no_configs:
        ring_obs_mask = (vec._rings[1], vec._rings[2])
        configs = measure_configs(a, ctx)

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
    fill_gbs = flat.numel() * 8 / device_time(lambda: flat.fill_(0), 5) / 1e9

    # ---- roofline of the fused rollout kernel -----------------------------------------------------------
    launch_ms = elapsed_ms / a.steps
    achieved = BYTES_PER_ENV_STEP * n * T / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("envs") == n:
            traffic = tj["dram_bytes_per_env_step"] * n * T   # ncu --set full capture scaled to this launch
            traffic_src = "profiles/traffic.json (static ncu --set full capture of this kernel, not measured in this run)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "gbl::rollout_kernel<fast=true, streaming=true, aux=false, block=256, bulk=true>",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n * T,
                "launch_ms_avg": launch_ms, "launch_ms_min": per_launch[0], "launch_ms_median": per_launch[len(per_launch) // 2],
                "same_run_fill_gbs": fill_gbs, "frac_of_same_run_fill": achieved / fill_gbs,
                "fill_note": f"torch fill_ (int64 view) of the {flat.numel() * 8 / 1e9:.1f} GB observation trajectory buffer, write-only"}

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
        # the TIMED template instance (no per-step aux outputs) against the checked one, at full size, on the device
        twin = gobblet_v1.vec_env(n, device=dev, seed=0, env_id_base=rank * n, streaming_stores=not a.plain_stores)
        chk = twin.rollout_random(T, ring=ring, per_step=True, log_actions=True)
        plain = gobblet_v1.vec_env(n, device=dev, seed=0, env_id_base=rank * n, streaming_stores=not a.plain_stores)
        plain._rings = vec._rings                      # reuse the timed run's trajectory buffer
        out = plain.rollout_random(T, ring=ring)
        same = all(torch.equal(out["obs"][s], chk["obs"][s]) and torch.equal(out["mask"][s], chk["mask"][s]) for s in range(ring))
        assert same and torch.equal(plain.state, twin.state) and torch.equal(plain.stats, twin.stats), \
            "the timed kernel instance differs from the oracle-checked one"
        parity += f"; timed instance <fast,streaming,no-aux> == checked instance on all {n} envs x {T} steps (device-side compare)"
        del twin, chk, plain, out
        if RL.available():
            gob = RL.load_gobblet()
            acts = got["actions"].cpu().numpy()
            for e_i in range(0, 512, 128):
                env = gob.raw_env(render_mode=None)
                env.reset()
                for t_ in range(24):
                    env.step(int(acts[t_, e_i]))
                    if env.terminations[env.agent_selection]:
                        env.reset()
                    o = env.observe(env.agent_selection)
                    assert (o["observation"] == want["obs"][t_, e_i]).all() and (o["action_mask"] == want["mask"][t_, e_i]).all(), \
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

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "fused_env_steps_per_launch": T, "ring_slots": ring,
                       "l2": f"each launch writes a {ring * n * BYTES_PER_ENV_STEP / 1e9:.1f} GB trajectory once (>> 126 MB L2, no flush needed)",
                       "parallelism": f"{world} independent shards by global env id, no per-step communication",
                       "stores": "st.global" if a.plain_stores else "st.global.cs"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "gpu_launches_per_rank": gpu_launches // world, "clocks": clocks,
            "sustained": sustained, "configs": configs, "multi_gpu_equality": multi_eq,
            "small_batch": None if not configs or "c2" not in configs else
            {"workload": configs["c2"]["workload"], "value": configs["c2"]["value"], "unit": UNIT},
            "episode_stats": dict(zip(("episodes", "p1_wins", "p2_wins", "steps", "sum_len", "illegal", "both_lines", "max_len"), stats.tolist()))}
    emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
