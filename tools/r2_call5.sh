#!/bin/bash
# tools/r2_call5.sh -- 1 GPU: ncu --set full of the two generations of the resident-row kernels (same bins), to see WHY rows2 is slower
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
for v in 1 2; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'row_resident2?_kernel<\(int\)6, \(int\)1, \(bool\)1|row_resident2?_kernel<\(int\)3, \(int\)4, \(bool\)1|row_resident2?_kernel<\(int\)6, \(int\)8, \(bool\)0' \
    --launch-skip 6 --launch-count 3 -o $O/c5_rows_v$v -f $B --options row_kernels=$v > $O/c5_ncu_v$v.log 2>&1; echo "ncu v$v rc=$?"; tail -2 $O/c5_ncu_v$v.log
done
ls -la $O/*.ncu-rep
