#!/bin/bash
# one short GPU session: parity tests, then the device-timed headline + side configs (no CPU leg, no e2e)
tag=${1:-q}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
c = d["configs"]
print("value %.4g  ms/step %.4f  frac %.4f  sustained %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["sustained"]["value"]))
print("c2 us/step", c["c2"]["us_per_lockstep_step_by_block_threads"], "graph", c["c2"]["us_per_lockstep_step_cuda_graph_of_1step_launches"])
print("c4 boards/s %.4g  us %.1f" % (c["c4"]["value"], c["c4"]["us_per_launch"]))
print("c5 no_final %.4g (%.1f us)  with_final %.4g (%.1f us)" % (c["c5"]["no_final"]["value"], 1e3 * c["c5"]["no_final"]["ms_per_collection"], c["c5"]["with_final_obs"]["value"], 1e3 * c["c5"]["with_final_obs"]["ms_per_collection"]))
PY
