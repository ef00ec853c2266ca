#!/bin/bash
# batch D (1 GPU): host-stall probe, then the bench with the new defaults (tile_order, pre_tma on)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python tools/stall_probe.py 300 > $O/r2d_stall_probe.json 2> $O/r2d_stall_probe.err
python bench.py --steps 20 --warmup 3 > $O/r2d_bench_default.json 2> $O/r2d_bench_default.err
python - <<P
import json
d=json.loads(open("$O/r2d_stall_probe.json").read().strip().splitlines()[-1])
for k,v in d.items(): print(k, v)
d=json.loads(open("$O/r2d_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["step_ms"], d["single_view"]["value"], d["cpu_baseline"])
P
tail -c 300 $O/r2d_stall_probe.err
