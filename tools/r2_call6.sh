#!/bin/bash
# tools/r2_call6.sh -- 1 GPU: (1) A/B of the resident-row kernel generations (row_kernels = 1 | 2 | 3), (2) the general FM Gibbs path
# measured for the first time: both FM workloads of bench.py, ours and the unmodified libFM on the host cores, + a launch list
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
for opt in ${OPTS:-"row_kernels=1" "row_kernels=2" "row_kernels=3"}; do
  i=$((i+1))
  timeout 300 $B --options "$opt" > $O/c6_$i.json 2> $O/c6_$i.err
  python - "$opt" $O/c6_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-32s %.3f ms/sweep  user %.3f item %.3f  top %.1f us  rmse %.6f" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"], d["rmse_after_timed"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
for w in fm_mf_ml10m fm_wide_ml1m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $O/c6_$w.json 2> $O/c6_$w.err; echo "$w ours rc=$?"; cut -c1-600 $O/c6_$w.json
  timeout 900 python bench.py --workload $w --impl reference > $O/c6_${w}_ref.json 2> $O/c6_${w}_ref.err; echo "$w reference rc=$?"; cut -c1-400 $O/c6_${w}_ref.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/c6_fm_launches.csv python tools/fm_devbench.py --shape ml1m -K 8 --wide 4 --iters 1 --warmup 1 > $O/c6_fm_ncu.log 2>&1; echo "fm ncu rc=$?"
