#!/bin/bash
# tools/r2_fm.sh -- first GPU call for the general FM Gibbs path: gpurun --timeout 1500 -- 'bash tools/r2_fm.sh'
# 1) every parity case of tests/fm_gpu_cases.py in a process of its own (the path has not run on hardware yet)
# 2) timing on an MF-shaped and on a wide design matrix, 3) launch list of one short run for profiles/
set -u
O=gpurun_out
mkdir -p $O
for c in columns zero_mf zero_general zero_general_k20 zero_als zero_variants long_columns errors cli live; do
  timeout 600 python tests/fm_gpu_cases.py $c > $O/r2_fm_$c.log 2>&1; echo "fm case $c rc=$?"; tail -3 $O/r2_fm_$c.log
done
timeout 900 python tools/fm_devbench.py --shape ml10m -K 8 > $O/r2_fm_bench_ml10m.json 2> $O/r2_fm_bench_ml10m.err; echo "devbench ml10m rc=$?"; cat $O/r2_fm_bench_ml10m.json
timeout 900 python tools/fm_devbench.py --shape ml1m -K 8 --wide 4 > $O/r2_fm_bench_wide.json 2> $O/r2_fm_bench_wide.err; echo "devbench wide rc=$?"; cat $O/r2_fm_bench_wide.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_fm_launches.csv python tools/fm_devbench.py --shape ml1m -K 4 --iters 1 --warmup 1 > $O/r2_fm_ncu.log 2>&1; echo "ncu rc=$?"
