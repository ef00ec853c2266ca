#!/bin/bash
# batch F (1 GPU): A/B of the forward half-tile CTAs and the backward register budgets; full GPU suite at the current defaults
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r2f_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
for v in base fwd_half=1 bwd_minb=14 bwd_minb=12 fwd_half=1,bwd_minb=14; do
  n=$(echo $v | tr '=,' '__')
  if [ "$v" = "base" ]; then $B > $O/r2f_bench_$n.json 2> $O/r2f_bench_$n.err; else GS_OPTS=$v $B > $O/r2f_bench_$n.json 2> $O/r2f_bench_$n.err; fi
  python - <<P
import json
try:
    d=json.loads(open("$O/r2f_bench_$n.json").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_view"]; sv=d.get("single_view") or {}
    print("$v", round(d["value"],1), d["step_ms"]["resident"]["median"], "sv", round(sv.get("value",0),1), {n: k[n] for n in ("render_fwd","render_bwd")}, {n: sv["library_kernel_ms_per_view"][n] for n in ("render_fwd","render_bwd","sum")})
except Exception as e:
    print("$v", "FAILED", e)
P
done
tail -4 $O/r2f_pytest.log
