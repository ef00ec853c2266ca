#!/bin/bash
# batch K (N GPUs of one box, N = $1): final multi-GPU check with the round's last kernels: the two-GPU hardware tests (N = 2 only),
# the bench exactly as the driver launches it, and the fused reduce-scatter variant
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
if [ "$N" = "2" ]; then python -m pytest tests/test_round2_gpu.py -q -k two_gpu 2>&1 | tail -5 > $O/r2k_pytest_n$N.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-single-view > $O/r2k_bench_n${N}_default.json 2> $O/r2k_bench_n${N}_default.err
$TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --reduce peer --no-single-view > $O/r2k_bench_n${N}_peer.json 2> $O/r2k_bench_n${N}_peer.err
[ -f $O/r2k_pytest_n$N.log ] && tail -3 $O/r2k_pytest_n$N.log
for f in default peer; do [ -f $O/r2k_bench_n${N}_$f.json ] && python - <<P
import json
try:
    d=json.loads(open("$O/r2k_bench_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["step_ms"])
except Exception as e:
    print("$f", "FAILED", e); print(open("$O/r2k_bench_n${N}_$f.err").read()[-1500:])
P
done
