#!/bin/bash
# batch E: N GPUs: hardware test (chunked + fused reduce-scatter), bench with --reduce peer against the all-reduce default
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_round2_gpu.py -q -s -k two_gpu 2>&1 | tail -25 > $O/r2e_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --reduce peer --no-single-view > $O/r2e_bench_n${N}_peer.json 2> $O/r2e_bench_n${N}_peer.err
$TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-single-view > $O/r2e_bench_n${N}_allreduce.json 2> $O/r2e_bench_n${N}_allreduce.err
$TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --reduce peer --no-single-view > $O/r2e_bench_n${N}_peer2.json 2> $O/r2e_bench_n${N}_peer2.err
grep -v "^\[W\|^W0" $O/r2e_pytest_n$N.log | tail -12
for f in peer allreduce peer2; do python - <<P
import json
try:
    d=json.loads(open("$O/r2e_bench_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["step_ms"]["resident"], d["config"].get("reduction"), d["kernel_ms_per_view"]["preprocess_bwd"])
except Exception as e:
    print("$f", "FAILED", e); print(open("$O/r2e_bench_n${N}_$f.err").read()[-2500:])
P
done
