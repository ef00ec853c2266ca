#!/bin/bash
# round-2 GPU batch B: GPU tests (TMA row path, finite differences), default / pre_tma / tile_order bench lines with the
# single-view leg, memcheck of the sync-free overflow path, ncu launch list + full captures of the dominant kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2b_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > $O/r2b_bench_default.json 2> $O/r2b_bench_default.err
GS_OPTS=pre_tma=1 $B > $O/r2b_bench_pre_tma.json 2> $O/r2b_bench_pre_tma.err
GS_OPTS=tile_order=1 $B > $O/r2b_bench_tile_order.json 2> $O/r2b_bench_tile_order.err
GS_OPTS=pre_tma=1,tile_order=1 $B > $O/r2b_bench_both.json 2> $O/r2b_bench_both.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_round2_gpu.py -q -k "sync_free or chunked" > $O/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2b_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -k "tile_order_off or sort_4" > $O/r2b_memcheck_tma.log 2>&1; echo "memcheck rc=$?" >> $O/r2b_memcheck_tma.log
# launch list of two batched steps (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file $O/r2b_launches.csv python tools/profile_one.py 3 > $O/r2b_ncu_a.log 2>&1
# full captures: blend kernels + both preprocess kernels of the batched step
ncu --set full --clock-control none --import-source on -k regex:"render_bwd|render_fwd|preprocess_bwd|preprocess_fwd" -s 8 -c 8 -o $O/r2b_prof python tools/profile_one.py 3 > $O/r2b_ncu_b.log 2>&1
GS_OPTS=pre_tma=1 ncu --set full --clock-control none --import-source on -k regex:"preprocess_bwd|preprocess_fwd" -s 4 -c 4 -o $O/r2b_prof_tma python tools/profile_one.py 3 > $O/r2b_ncu_c.log 2>&1
# single-view kernels (GS_V=1 -> V=1 batches use the batch kernels; the drop-in path is profiled through bench --api render)
GS_V=1 ncu --set full --clock-control none -k regex:"preprocess_bwd|preprocess_fwd|render_bwd|render_fwd" -s 8 -c 4 -o $O/r2b_prof_v1 python tools/profile_one.py 4 > $O/r2b_ncu_d.log 2>&1
tail -c 500 $O/r2b_pytest.log
tail -3 $O/r2b_memcheck.log; tail -3 $O/r2b_memcheck_tma.log
for f in default pre_tma tile_order both; do python - <<P
import json
try:
    d=json.loads(open("$O/r2b_bench_$f.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_view"]
    print("$f", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["step_ms"]["resident"], "sv", (d.get("single_view") or {}).get("value"), (d.get("single_view") or {}).get("library_kernel_ms_per_view"),
          {n: k[n] for n in ("render_fwd","render_bwd","sort_scatter","tile_order","preprocess_fwd","preprocess_bwd")})
except Exception as e:
    print("$f", "FAILED", e)
P
done
ls -la $O | grep r2b
