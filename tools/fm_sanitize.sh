#!/bin/bash
# tools/fm_sanitize.sh -- the FM Gibbs kernels under AddressSanitizer / ThreadSanitizer (compute-sanitizer is closed on this GPU
# pool, so this is the memory- and race-check of csrc/fm.cu): the host build of the kernels (tools/emu_include, see
# tests/test_fm_simt_emulation.py) instrumented by g++, driven by the parity cases of tests/fm_gpu_cases.py.  CPU only, ~5 minutes.
#   ASan: out-of-bounds / use-after-free in kernels and host orchestration; with detect_leaks=1 every "device" allocation
#         (cudaMalloc = malloc here) must be freed by sbmf_fm_destroy and by the error paths
#   TSan: races between the threads of a CTA (a missing __syncthreads / shuffle barrier); CTAs run one after the other here, so
#         cross-CTA conflicts are NOT covered -- those are excluded by construction (conflict-free runs, test_fm_emulation.py)
# Last run (round 1): all clean.
set -u
cd "$(dirname "$0")/.."
B=tools/build; mkdir -p $B
SRC=scalable-bayesian-matrix-factorization_b200/csrc/fm.cu
COMMON="-O1 -g -std=c++17 -pthread -fPIC -shared -x c++ -DSBMF_SIMT_EMU -I tools/emu_include -I include -I scalable-bayesian-matrix-factorization_b200/csrc"
SHRUNK="-DFM_BLOCK_T=64 -DFM_WARP_COL_MAX=8 -DFM_BLOCK_COL_MAX=40 -DFM_SLICE_LEN=16 -DFM_HYPER_CHUNK=16"
g++ $COMMON $SHRUNK -fsanitize=address -fno-omit-frame-pointer -o $B/libsbmf_fm_emu_asan.so $SRC || exit 1
g++ $COMMON -fsanitize=address -fno-omit-frame-pointer -o $B/libsbmf_fm_emu_prod_asan.so $SRC || exit 1
g++ $COMMON $SHRUNK -fsanitize=thread -o $B/libsbmf_fm_emu_tsan.so $SRC || exit 1
ASAN=$(gcc -print-file-name=libasan.so); TSAN=$(gcc -print-file-name=libtsan.so)
rc=0
for c in zero_general zero_variants zero_als errors; do
  out=$(ASAN_OPTIONS=detect_leaks=1 LD_PRELOAD=$ASAN SBMF_FM_LIB_PATH=$PWD/$B/libsbmf_fm_emu_asan.so python tests/fm_gpu_cases.py $c 2>&1)
  echo "$out" | grep -q "^ok $c" || { echo "asan $c: case failed"; rc=1; }
  echo "$out" | grep -E "ERROR: AddressSanitizer|libsbmf_fm_emu" && { echo "asan $c: finding inside the library"; rc=1; }
  echo "asan $c done"
done
out=$(ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$ASAN SBMF_FM_LIB_PATH=$PWD/$B/libsbmf_fm_emu_prod_asan.so python tests/fm_gpu_cases.py long_columns 2>&1)
echo "$out" | grep -q "^ok long_columns" || { echo "asan long_columns (production geometry): failed"; rc=1; }
echo "$out" | grep "ERROR: AddressSanitizer" && rc=1
echo "asan long_columns done"
for c in zero_mf zero_general; do
  out=$(TSAN_OPTIONS="report_signal_unsafe=0 history_size=2" LD_PRELOAD=$TSAN SBMF_FM_LIB_PATH=$PWD/$B/libsbmf_fm_emu_tsan.so python tests/fm_gpu_cases.py $c 2>&1)
  echo "$out" | grep -q "^ok $c" || { echo "tsan $c: case failed"; rc=1; }
  echo "$out" | grep "WARNING: ThreadSanitizer" && rc=1
  echo "tsan $c done"
done
# UBSan: misaligned accesses, signed overflow, shifts (first finding aborts the case)
g++ $COMMON $SHRUNK -fsanitize=undefined -fno-sanitize-recover=undefined -o $B/libsbmf_fm_emu_ubsan.so $SRC || exit 1
for c in zero_general zero_variants zero_als errors; do
  out=$(LD_PRELOAD=$(gcc -print-file-name=libubsan.so) SBMF_FM_LIB_PATH=$PWD/$B/libsbmf_fm_emu_ubsan.so python tests/fm_gpu_cases.py $c 2>&1)
  echo "$out" | grep -q "^ok $c" || { echo "ubsan $c: case failed"; rc=1; }
  echo "$out" | grep "runtime error" && rc=1
  echo "ubsan $c done"
done
echo "fm_sanitize rc=$rc"; exit $rc
