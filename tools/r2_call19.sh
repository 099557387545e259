#!/bin/bash
# tools/r2_call19.sh -- 1 GPU: two-chain streaming pipeline (option heavy_chains): heavy-row parity cases, then A/B
set -u
O=gpurun_out; mkdir -p $O
( time python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "zero_noise_heavy_rows or options_keep or baseline or determinism or graph" ) > $O/c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c19_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
for opt in "heavy_chains=1" "" "heavy_chains=1" ""; do
  timeout 300 $B --options "$opt" > $O/c19_tmp.json 2> $O/c19_tmp.err
  python - "$opt" $O/c19_tmp.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-20s %.3f ms/sweep  user %.3f item %.3f  rmse %.6f" % (sys.argv[1] or "default (2 chains)", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["rmse_after_timed"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
done
