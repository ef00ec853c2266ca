#!/bin/bash
# round-2 GPU batch C: N GPUs of one box (N = $1): hardware test of the chunked reduction, then the bench at N in three
# configurations -- sync-free + 4 overlapped chunks (default), sync-free + one all-reduce, round-1 behaviour (synchronous)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2c_topo_n$N.txt 2>&1
python -m pytest tests/test_round2_gpu.py -q -k two_gpu 2>&1 | tail -15 > $O/r2c_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2c_bench_n${N}_default.json 2> $O/r2c_bench_n${N}_default.err
$TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --grad-chunks 1 --no-single-view > $O/r2c_bench_n${N}_chunks1.json 2> $O/r2c_bench_n${N}_chunks1.err
$TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --sync --grad-chunks 1 --no-single-view --no-pin > $O/r2c_bench_n${N}_sync.json 2> $O/r2c_bench_n${N}_sync.err
$TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --no-single-view > $O/r2c_bench_n${N}_default2.json 2> $O/r2c_bench_n${N}_default2.err
tail -5 $O/r2c_pytest_n$N.log
for f in default chunks1 sync default2; do python - <<P
import json
try:
    d=json.loads(open("$O/r2c_bench_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f N=$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["step_ms"], d["config"].get("pinned_cores"), d["attempts"]["resident"])
except Exception as e:
    print("$f", "FAILED", e); print(open("$O/r2c_bench_n${N}_$f.err").read()[-1500:])
P
done
