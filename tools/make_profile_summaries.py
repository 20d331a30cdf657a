#!/usr/bin/env python
"""Turn the .ncu-rep captures of tools/final_evidence.sh (gpurun_out/<tag>_*.ncu-rep) into the tracked summaries
under profiles/: raw-page CSVs, traffic.json (timed rollout kernel), greedy_issue.json, small_batch_issue.json,
collect_issue.json and the SASS excerpt of the timed instance.   python tools/make_profile_summaries.py r2f"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(REPO, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return out, dict(zip(rows[0], rows[2]))


def f(m, k):
    return float(m[k].replace(",", ""))


def common(m):
    return {"kernel": m["Kernel Name"], "grid": int(f(m, "launch__grid_size")), "block": int(f(m, "launch__block_size")),
            "registers_per_thread": int(f(m, "launch__registers_per_thread")),
            "duration_us_under_ncu": f(m, "gpu__time_duration.sum") * (1e3 if "ms" in m.get("gpu__time_duration.sum__unit", "") else 1),
            "smsp__inst_executed.sum": f(m, "smsp__inst_executed.sum"),
            "issue_active_pct": f(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": f(m, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": f(m, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct_of_pin_peak": f(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}


def units(rep):
    """ncu prints durations / bytes in scaled units: read the unit row too"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[1]))


SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
    g = os.path.join(REPO, "gpurun_out")
    cmd = "ncu --set full --clock-control none --import-source on"
    # ---- timed rollout kernel
    rep = f"{g}/{tag}_rollout.ncu-rep"
    text, m = raw(rep)
    u = units(rep)
    open(f"{P}/r2_rollout_kernel_ncu_full.csv", "w").write(text)
    dur_us = f(m, "gpu__time_duration.sum") * SCALE[u["gpu__time_duration.sum"]]
    rd = f(m, "dram__bytes_read.sum") * SCALE[u["dram__bytes_read.sum"]]
    wr = f(m, "dram__bytes_write.sum") * SCALE[u["dram__bytes_write.sum"]]
    n, T = 1 << 20, 64
    c = common(m)
    c["duration_us_under_ncu"] = dur_us
    json.dump({**c, "envs": n, "fused_steps": T, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_env_step": (rd + wr) / (n * T), "algorithmic_bytes_per_env_step": 171,
               "warp_instructions_per_warp_step": c["smsp__inst_executed.sum"] / (n / 32 * T), "duration_ms": dur_us / 1e3,
               "dynamic_smem_bytes_per_block": f(m, "launch__shared_mem_per_block_dynamic") * SCALE.get(u["launch__shared_mem_per_block_dynamic"].split("/")[0], 1.0),
               "smem_store_bank_conflicts": f(m, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
               "smem_store_wavefronts": f(m, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
               "source": f"profiles/r2_rollout_kernel_ncu_full.csv ({cmd} --kernel-name-base demangled -k 'regex:rollout_kernel<.bool.1, .bool.1, "
                         ".bool.0, .int.256' -s 4 -c 1, bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-configs; round-2 final binary)"},
              open(f"{P}/traffic.json", "w"), indent=1)
    # ---- greedy, small batch, collection
    for name, csvname, jname, per, nunits, how in (
            ("greedy", "r2_greedy_depth2_ncu_full.csv", "greedy_issue.json", "warp_instructions_per_board", 65536,
             "-k regex:greedy_kernel -s 1 -c 1 python tools/greedy_case.py --iters 1"),
            ("small", "r2_small_batch_kernel_ncu_full.csv", "small_batch_issue.json", "warp_instructions_per_32env_step_both_warps", 128 * 512,
             "-k regex:rollout_kernel -s 4 -c 1 python tools/small_batch_case.py  (4096 envs x 512 fused steps; two warps per 32 envs: "
             "observation warp + mask warp)"),
            ("collect", "r2_collect_kernel_ncu_full.csv", "collect_issue.json", "warp_instructions_per_warp_slot", 4096 * 17,
             "-k regex:rollout_kernel -s 3 -c 1 python tools/collect_case.py --iters 1  (131 072 envs x (1 + 16) emitted slots)")):
        rep = f"{g}/{tag}_{name}.ncu-rep"
        text, m = raw(rep)
        u = units(rep)
        open(f"{P}/{csvname}", "w").write(text)
        c = common(m)
        c["duration_us_under_ncu"] = f(m, "gpu__time_duration.sum") * SCALE[u["gpu__time_duration.sum"]]
        c[per] = c["smsp__inst_executed.sum"] / nunits
        if name == "greedy":
            c["boards"] = 65536
            c["round1_warp_instructions_per_board"] = 2950
        if name == "small":
            c["sm__cycles_elapsed.max"] = f(m, "sm__cycles_elapsed.max")
            c["cycles_per_lockstep_step"] = c["sm__cycles_elapsed.max"] / 512
        c["source"] = f"profiles/{csvname} ({cmd} {how})"
        json.dump(c, open(f"{P}/{jname}", "w"), indent=1)
    # ---- SASS excerpt of the timed instance
    so = os.path.join(REPO, "gobblet_rl_b200", "csrc", "libgobblet_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn = "_ZN3gbl14rollout_kernelILb1ELb1ELb0ELi256ELb1ELb0EEEvNS_13RolloutParamsE"
    body = sass.split("Function : " + fn)[1].split("Function : ")[0].splitlines()
    ins = [ln for ln in body if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    keep = re.compile(r"STS\.U8|STS\.128|FENCE|UBLKCP|UTMACMDFLUSH|DEPBAR|STG\.E|LDG|WARPSYNC|HMMA|UTCMMA|SHFL")
    with open(f"{P}/r2_rollout_sass.txt", "w") as fo:
        fo.write("# SASS excerpt of the TIMED kernel instance gbl::rollout_kernel<fast, streaming, no-aux, 256 threads, bulk> (sm_100a, round-2 final build)\n"
                 f"# cuobjdump -sass gobblet_rl_b200/csrc/libgobblet_b200.so, function {fn}\n"
                 "# what to look for: predicated STS.U8 (observation byte scatter), FENCE.VIEW.ASYNC.S (generic -> async proxy),\n"
                 "# UBLKCP.G.S + UTMACMDFLUSH (cp.async.bulk.global.shared::cta = TMA bulk store of the 3744-byte observation image),\n"
                 "# DEPBAR.LE SB0 (cp.async.bulk.wait_group.read, now BEHIND the next step's game logic), STS.128 (re-zeroing the image),\n"
                 "# STG.E.EF.128 (st.global.cs.v4: the expanded action mask).\n"
                 f"# total instructions in the function: {len(ins)} (the step body appears three times: peeled first step + loop unrolled by two);\n"
                 f"# tensor-core ops (HMMA / UTCMMA): {sum(bool(re.search('HMMA|UTCMMA', x)) for x in ins)} -- nothing is a contraction\n\n")
        for i, ln in enumerate(ins):
            if keep.search(ln):
                fo.write(f"{i}:{ln.rstrip()[:110]}\n")
    print("profiles/ updated from", tag)


if __name__ == "__main__":
    main()
