#!/bin/bash
# tools/r2_call14.sh -- 1 GPU: the FM column tiers of a run on concurrent streams: FM parity cases (gating suite) + both FM workloads
set -u
O=gpurun_out; mkdir -p $O
( time python -m pytest tests/test_z_fm_parity_gpu.py tests/test_y_reference_patch.py -m gpu -x -q ) > $O/c14_pytest_fm.log 2>&1; echo "pytest fm rc=$?"; tail -4 $O/c14_pytest_fm.log
for w in fm_mf_ml10m fm_wide_ml1m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $O/c14_$w.json 2> $O/c14_$w.err; echo "$w ours rc=$?"; cut -c1-180 $O/c14_$w.json
done
