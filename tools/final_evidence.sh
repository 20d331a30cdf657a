#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, the bench line + reference arm, ncu launch list and full
# captures of the timed kernels.  Everything lands in gpurun_out/<tag>_*; summaries are copied to profiles/ by hand.
tag=${1:-r2f}
o=gpurun_out
python -m pytest tests -m gpu -q > $o/${tag}_tests.log 2>&1; tail -2 $o/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench_1gpu.json 2> $o/${tag}_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err; echo "ref rc=$?"
python benchmarks/bench_paths.py > $o/${tag}_bench_paths.json 2> $o/${tag}_bench_paths.err; echo "paths rc=$?"
B="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-configs"
$B > $o/${tag}_plain_steps6.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_bench_steps6.csv $B > $o/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rollout_kernel<.bool.1, .bool.1, .bool.0, .int.256' -s 4 -c 1 -o $o/${tag}_rollout -f $B > $o/${tag}_ncu_rollout.log 2>&1; tail -1 $o/${tag}_ncu_rollout.log
python tools/small_batch_case.py > $o/${tag}_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -o $o/${tag}_small -f python tools/small_batch_case.py > $o/${tag}_ncu_small.log 2>&1; tail -1 $o/${tag}_ncu_small.log
python tools/collect_case.py > $o/${tag}_collect_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $o/${tag}_collect -f python tools/collect_case.py --iters 1 > $o/${tag}_ncu_collect.log 2>&1; tail -1 $o/${tag}_ncu_collect.log
python tools/greedy_case.py > $o/${tag}_greedy_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:greedy_kernel -s 1 -c 1 -o $o/${tag}_greedy -f python tools/greedy_case.py --iters 1 > $o/${tag}_ncu_greedy.log 2>&1; tail -1 $o/${tag}_ncu_greedy.log
python tools/full_size_parity.py gpu > $o/${tag}_full_size_gpu.log 2>&1; tail -1 $o/${tag}_full_size_gpu.log
