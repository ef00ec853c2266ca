#!/bin/bash
# round-2 GPU batch A: GPU tests, default bench, A/B of the prepared options, parity report, config 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -3 > $O/r2a_gpu.txt
python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/r2a_pytest.log
python bench.py --steps 20 --warmup 3 > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-single-view"
$B --sync > $O/r2a_bench_sync.json 2> $O/r2a_bench_sync.err
GS_OPTS=tile_order=1 $B > $O/r2a_bench_tile_order.json 2> $O/r2a_bench_tile_order.err
GS_OPTS=fused_ranges=1 $B > $O/r2a_bench_fused_ranges.json 2> $O/r2a_bench_fused_ranges.err
GS_OPTS=sort_big_ipt=8 $B > $O/r2a_bench_ipt8.json 2> $O/r2a_bench_ipt8.err
GS_OPTS=fused_ranges=1,sort_big_ipt=8 $B > $O/r2a_bench_fused_ipt8.json 2> $O/r2a_bench_fused_ipt8.err
python tools/parity_report.py $O/r2_parity.json > $O/r2a_parity.log 2>&1
python bench.py --workload train6m --iterations 1000 > $O/r2a_train6m.json 2> $O/r2a_train6m.err
tail -c 600 $O/r2a_pytest.log
for f in default sync tile_order fused_ranges ipt8 fused_ipt8; do python - <<P
import json
try:
    d=json.loads(open("$O/r2a_bench_$f.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_view"]
    print("$f", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["step_ms"]["resident"], "sv", (d.get("single_view") or {}).get("value"),
          {n: k[n] for n in ("render_fwd","render_bwd","sort_scatter","sort_hist","emit","tile_ranges","ranges_from_counts","tile_order","preprocess_fwd","preprocess_bwd")})
except Exception as e:
    print("$f", "FAILED", e)
P
done
tail -c 400 $O/r2a_train6m.json; tail -c 300 $O/r2a_train6m.err
