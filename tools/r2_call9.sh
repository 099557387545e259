#!/bin/bash
# tools/r2_call9.sh -- 1 GPU: row relabelling (option relabel, default on) on hardware: whole GPU suite, then A/B and set_train cost
set -u
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/c9_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-point --no-cli"
i=0
run() {
  i=$((i+1))
  timeout 400 $B "$@" > $O/c9_$i.json 2> $O/c9_$i.err
  python - "$*" $O/c9_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-40s %.3f ms/sweep  user %.3f item %.3f  top %.1f us  rmse %.6f  e2e %.1f G/s set_train %s" % (sys.argv[1] or "default", d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["roofline"]["us_per_launch"], d["rmse_after_timed"], (e.get("value") or 0) / 1e9, (e.get("breakdown_rank0") or {}).get("set_train_s")))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
}
run --options relabel=0
run
run --options trace=1
grep "sbmf trace" $O/c9_3.err | head -12
