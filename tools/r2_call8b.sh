#!/bin/bash
# tools/r2_call8b.sh -- 1 GPU A/B: option sets + the popularity-relabel experiment (bench.py --dev-relabel)
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
run() {
  i=$((i+1))
  timeout 400 $B "$@" > $O/c8b_$i.json 2> $O/c8b_$i.err
  python - "$*" $O/c8b_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-52s %.3f ms/sweep  user %.3f item %.3f  top %.1f us  rmse %.6f" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"], d["rmse_after_timed"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
}
run --options row_kernels=1
run --options row_kernels=3,alt_bins=1
run --options row_kernels=1,alt_bins=1
run --options resident_max_item=1024
run --dev-relabel
run --dev-relabel --options row_kernels=3,alt_bins=1
