#!/bin/bash
# tools/r2_call16.sh -- 1 GPU: A/B of the FM column passes (previous build in lib_prev vs 4 entries in flight), alternating runs
set -u
O=gpurun_out; mkdir -p $O
P=$PWD/scalable-bayesian-matrix-factorization_b200
for rep in 1 2; do
 for v in prev new; do
  L=$P/lib/libsbmf_cuda.so; [ $v = prev ] && L=$P/lib_prev/libsbmf_cuda.so
  for sh in "ml10m" "ml1m --wide 4" "netflix"; do
    SBMF_LIB_PATH=$L timeout 300 python tools/fm_devbench.py --shape $sh -K 8 --iters 10 --warmup 3 > $O/c16_tmp.json 2> $O/c16_tmp.err
    echo "$rep $v $sh: $(python -c "import json;d=json.load(open('$O/c16_tmp.json'));print(d['ms_per_iteration'],'ms',d['iterations_per_s'],'it/s',d['algorithmic_GBs'],'GB/s')" 2>/dev/null || tail -2 $O/c16_tmp.err)"
  done
 done
done
