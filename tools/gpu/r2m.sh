#!/bin/bash
# batch M (2 GPUs): the bench as the driver launches it at N=2, after the sampler change
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-single-view > $O/r2m_bench_n2.json 2> $O/r2m_bench_n2.err
python - <<P
import json
d=json.loads(open("$O/r2m_bench_n2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["ms_per_step"], "e2e", round(d["e2e"]["value"],1), d["step_ms"], d["clocks"])
P
tail -c 300 $O/r2m_bench_n2.err
