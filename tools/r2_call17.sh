#!/bin/bash
# tools/r2_call17.sh -- 1 GPU: option sweep on the final (relabelled) build
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
for opt in "" "l2_budget_mb=96" "l2_budget_mb=384" "slice_len=2048" "slice_len=8192" "resident_max_user=1024" "row_kernels=3" "resident_max_item=1536"; do
  i=$((i+1))
  timeout 300 $B --options "$opt" > $O/c17_$i.json 2> $O/c17_$i.err
  python - "$opt" $O/c17_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-28s %.3f ms/sweep  user %.3f item %.3f  top %.1f us" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
