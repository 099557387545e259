#!/bin/bash
# tools/r2_call15.sh -- 1 GPU: ncu --set full of the FM column kernels on the ML-10M-shaped MF workload (profiles/r2/ncu_fm_r2.txt)
# + launch list of one iteration
set -u
O=gpurun_out; mkdir -p $O
B="python tools/fm_devbench.py --shape ml10m -K 8 --iters 1 --warmup 2"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'col_warp_kernel<\(int\)1>|col_block_kernel<\(int\)1>|slice_apply_kernel<\(int\)1>|slice_reduce_kernel<\(int\)1>|q_rebuild_kernel|predict_kernel' \
    --launch-skip 40 --launch-count 12 -o $O/c15_fm -f $B > $O/c15_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $O/c15_ncu_full.log
ls -la $O/c15_fm.ncu-rep
