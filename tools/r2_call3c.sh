#!/bin/bash
# tools/r2_call3c.sh -- 1-GPU A/B of the streaming-pipeline variants (bench line summary per option set)
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
for opt in ${OPTS:-"" "fuse_solve=0"}; do
  i=$((i+1))
  timeout 300 $B --options "$opt" > $O/c3c_$i.json 2> $O/c3c_$i.err
  python - "$opt" $O/c3c_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-32s %.3f ms/sweep  user %.3f item %.3f  top %.1f us  frac_same %.3f" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"], d["roofline"]["l2_gather"]["frac_same_form"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
