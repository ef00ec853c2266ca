#!/bin/bash
# batch J (1 GPU): final validation at the round's defaults: GPU suite, smoke, bench lines (default / whole iteration / full loss /
# drop-in API / unbatched), the CPU reference arm, final ncu launch list and full captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2j_pytest.log
python __graft_entry__.py smoke > $O/r2j_smoke.log 2>&1
python bench.py > $O/r2j_bench_default.json 2> $O/r2j_bench_default.err
python bench.py --optimizer --no-cpu-baseline --steps 20 > $O/r2j_bench_optimizer.json 2> $O/r2j_bench_optimizer.err
python bench.py --loss l1_ssim --no-cpu-baseline --no-single-view --steps 20 > $O/r2j_bench_l1_ssim.json 2> $O/r2j_bench_l1_ssim.err
python bench.py --api render --no-cpu-baseline --no-single-view --steps 10 > $O/r2j_bench_api_render.json 2> $O/r2j_bench_api_render.err
python bench.py --no-batch --no-cpu-baseline --no-single-view --steps 10 > $O/r2j_bench_no_batch.json 2> $O/r2j_bench_no_batch.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2j_bench_reference.json 2> $O/r2j_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/r2j_launches.csv python tools/profile_one.py 3 > $O/r2j_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_bwd|render_fwd" -s 2 -c 4 -o $O/r2j_prof python tools/profile_one.py 3 > $O/r2j_ncu_b.log 2>&1
tail -3 $O/r2j_pytest.log; tail -2 $O/r2j_smoke.log
for f in default optimizer l1_ssim api_render no_batch reference; do python - <<P
import json
try:
    d=json.loads(open("$O/r2j_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), (d.get("step_ms") or {}).get("resident"), "sv", ((d.get("single_view") or {}).get("value")), "launches", d.get("gpu_launches"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e:
    print("$f", "FAILED", e); print(open("$O/r2j_bench_$f.err").read()[-1500:])
P
done
