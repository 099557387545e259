#!/bin/bash
# tools/r2_call4.sh -- 1 GPU: the second-generation resident-row kernels (csrc/rows2.cuh) on hardware for the first time.
# 1) whole GPU suite, 2) A/B row_kernels = 1 | 2 (ms per sweep and per phase), 3) launch lists of both for per-bin times.
set -u
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/c4_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
for opt in ${OPTS:-"row_kernels=1" "row_kernels=2" "row_kernels=2,l2_budget_mb=64"}; do
  i=$((i+1))
  timeout 300 $B --options "$opt" > $O/c4_$i.json 2> $O/c4_$i.err
  python - "$opt" $O/c4_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-32s %.3f ms/sweep  user %.3f item %.3f  top %.1f us  rmse %.6f" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"], d["rmse_after_timed"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
B2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
for v in 1 2; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_launches_v$v.csv $B2 --options row_kernels=$v > $O/c4_ncu_v$v.log 2>&1; echo "ncu list v$v rc=$?"
  python tools/launch_summary.py $O/c4_launches_v$v.csv > $O/c4_launches_v${v}_summary.txt 2>&1; head -30 $O/c4_launches_v${v}_summary.txt
done
