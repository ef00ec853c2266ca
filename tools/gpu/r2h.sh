#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2h_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2h_bench.json 2> $O/r2h_bench.err
ncu --set full --clock-control none --import-source on -k regex:"render_bwd|render_fwd" -s 2 -c 2 -o $O/r2h_prof python tools/profile_one.py 3 > $O/r2h_ncu.log 2>&1
tail -3 $O/r2h_pytest.log
python - <<P
import json
d=json.loads(open("$O/r2h_bench.json").read().strip().splitlines()[-1])
k=d["kernel_ms_per_view"]; sv=d["single_view"]
print(round(d["value"],1), d["step_ms"]["resident"], "e2e", round(d["e2e"]["value"],1), "sv", round(sv["value"],1), {n:k[n] for n in ("render_fwd","render_bwd")}, {n: sv["library_kernel_ms_per_view"][n] for n in ("render_fwd","render_bwd","sum")})
P
tail -c 300 $O/r2h_bench.err
