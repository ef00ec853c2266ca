#!/bin/bash
# batch L (1 GPU): the default bench line after the clock sampler was gated behind the launch thread's enqueue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python bench.py > $O/r2l_bench_default.json 2> $O/r2l_bench_default.err
python - <<P
import json
d=json.loads(open("$O/r2l_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["ms_per_step"], "e2e", round(d["e2e"]["value"],1), d["step_ms"], "sv", d["single_view"]["value"], d["clocks"], d["gpu_launches"], d["cpu_baseline"])
P
tail -c 400 $O/r2l_bench_default.err
