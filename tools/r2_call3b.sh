#!/bin/bash
# tools/r2_call3b.sh -- 1-GPU A/B batch: streaming-pipeline variants and the resident / streamed threshold per side
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
for opt in "" "fuse_solve=0" "pair_gather=1" "pair_gather=1,fuse_solve=0" "resident_max_item=1024" "resident_max_item=512" "resident_max_item=256" "resident_max_user=1024" "resident_max_user=512" "slice_len=2048" "slice_len=8192"; do
  i=$((i+1))
  timeout 300 $B --options "$opt" > $O/c3b_$i.json 2> $O/c3b_$i.err
  python - "$opt" $O/c3b_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-32s %.3f ms/sweep  user %.3f item %.3f  top %.1f us" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
