#!/bin/bash
# tools/r2_ncu.sh -- launch list + one ncu --set full capture (top streaming kernel, biggest resident bins) of the default bench command
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
$B > $O/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2.csv $B > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'heavy_accumulate_kernel<\(int\)2, \(int\)2, \(int\)2, \(int\)64, \(bool\)0, \(bool\)0>|row_resident_kernel<\(int\)6, \(int\)1, \(bool\)1>|row_group_kernel<\(int\)6, \(int\)16, \(bool\)1>|row_resident_kernel<\(int\)3, \(int\)4, \(bool\)1>' \
    --launch-skip 45 --launch-count 15 -o $O/prof_r2_top -f $B > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/ncu_full.log
ls -la $O/*.ncu-rep; du -sh $O
