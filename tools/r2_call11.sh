#!/bin/bash
# tools/r2_call11.sh -- 1 GPU, for profiles/r2: launch list of the default bench command and ncu --set full of the dominant streaming
# kernel + the three biggest resident-row kernels of the user phase (final round-2 build: relabel, alt_bins, row_kernels = 3)
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
$B > $O/c11_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c11_launches.csv $B > $O/c11_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py $O/c11_launches.csv > $O/c11_launches_summary.txt 2>&1; head -12 $O/c11_launches_summary.txt
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'heavy_accumulate_kernel<\(int\)2, \(int\)2, \(int\)2, \(int\)64, \(bool\)0|row_resident2_kernel<\(int\)6, \(int\)1, \(bool\)1|row_group2_kernel<\(int\)6, \(int\)16, \(bool\)1|row_resident2_kernel<\(int\)6, \(int\)2, \(bool\)1' \
    --launch-skip 45 --launch-count 15 -o $O/c11_top -f $B > $O/c11_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/c11_ncu_full.log
ls -la $O/c11_top.ncu-rep
