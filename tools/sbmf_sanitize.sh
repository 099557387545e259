#!/bin/bash
# tools/sbmf_sanitize.sh -- the SBMF sweep kernels (csrc/kernels.cu, storage.cu behind the C ABI of api.cu) under AddressSanitizer /
# ThreadSanitizer / UBSan.  compute-sanitizer is closed on this GPU pool, so this is the memory- and race-check of the hot path: the host
# build of the library's own sources (tools/build_emu.sh: CTAs on host threads, barriers and shuffles are real synchronisation),
# instrumented by g++ and driven by members of tests/test_parity_gpu.py (SBMF_EMULATED=1 + SBMF_LIB_PATH).  CPU only, ~7 minutes.
#   ASan: out-of-bounds / use-after-free in kernels and host orchestration ("device" memory is malloc'ed and filled with 0xCD, so
#         a kernel that relies on cudaMalloc returning zeros fails the parity check as well)
#   TSan: races between the threads of a CTA -- a read that is only ordered before another lane's write by warp lock-step, a
#         missing __syncthreads / __syncwarp.  CTAs run one after the other here, so cross-CTA conflicts are NOT covered.
# Round 1: TSan found the old bias of a row read after the reduction while the row's first lane already stores the new one
# (row_resident_kernel with several warps per row: a real race on the GPU; row_group_kernel: ordered only by lock-step) -- fixed
# by reading before the reduction; clean since.
set -u
cd "$(dirname "$0")/.."
bash tools/build_emu.sh libsbmf_cuda_emu_asan.so -fsanitize=address -fno-omit-frame-pointer -g > /dev/null || exit 1
bash tools/build_emu.sh libsbmf_cuda_emu_tsan.so -fsanitize=thread -g > /dev/null 2>&1 || exit 1
ASAN=$(gcc -print-file-name=libasan.so); TSAN=$(gcc -print-file-name=libtsan.so)
SEL="layout_bit_exact or ragged_random or edge_ or error_behaviour or (zero_noise_heavy_rows and (8-False-1 or 8-True-0 or 20-False-0))"
rc=0
rm -f tools/build/asan_rep.* tools/build/tsan_rep.*
SBMF_EMULATED=1 ASAN_OPTIONS="detect_leaks=0 log_path=$PWD/tools/build/asan_rep" LD_PRELOAD=$ASAN SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu_asan.so \
  python -m pytest tests/test_parity_gpu.py -q -x -k "$SEL" 2>&1 | tail -2 | tee tools/build/asan_pytest.log
grep -q " passed" tools/build/asan_pytest.log && ! grep -q "failed" tools/build/asan_pytest.log || { echo "asan: cases failed"; rc=1; }
if ls tools/build/asan_rep.* > /dev/null 2>&1; then grep -h "^SUMMARY" tools/build/asan_rep.* | sort | uniq -c; echo "asan: findings"; rc=1; fi
SBMF_EMULATED=1 TSAN_OPTIONS="report_signal_unsafe=0 history_size=2 log_path=$PWD/tools/build/tsan_rep" LD_PRELOAD=$TSAN SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu_tsan.so \
  python -m pytest tests/test_parity_gpu.py -q -x -k "ragged_random or edge_latent or edge_single or (zero_noise_heavy_rows and (8-False-1 or 8-True-0))" 2>&1 | tail -2 | tee tools/build/tsan_pytest.log
grep -q " passed" tools/build/tsan_pytest.log && ! grep -q "failed" tools/build/tsan_pytest.log || { echo "tsan: cases failed"; rc=1; }
# the smoke job (ML-100K, K = 20: every resident-row bin of a real data set, zero-noise and live Philox sampling) under TSan
SBMF_EMULATED=1 TSAN_OPTIONS="report_signal_unsafe=0 history_size=2 log_path=$PWD/tools/build/tsan_rep" LD_PRELOAD=$TSAN SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu_tsan.so \
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee tools/build/tsan_smoke.log
grep -q "smoke ok" tools/build/tsan_smoke.log || { echo "tsan: smoke job failed"; rc=1; }
if ls tools/build/tsan_rep.* > /dev/null 2>&1; then grep -h "^SUMMARY" tools/build/tsan_rep.* | sort | uniq -c; echo "tsan: findings"; rc=1; fi
# UBSan (-fno-sanitize-recover: the first finding aborts): misaligned 256-bit accesses (the host cudaMalloc returns 256-byte aligned
# blocks like the real one, so sub-allocation offsets are checked), signed overflow, out-of-range shifts in index arithmetic
bash tools/build_emu.sh libsbmf_cuda_emu_ubsan.so -fsanitize=undefined -fno-sanitize-recover=undefined -g > /dev/null 2>&1 || exit 1
SBMF_EMULATED=1 LD_PRELOAD=$(gcc -print-file-name=libubsan.so) SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu_ubsan.so \
  python -m pytest tests/test_parity_gpu.py -q -x -s -k "$SEL or live_same_philox" > tools/build/ubsan_pytest.log 2>&1
tail -1 tools/build/ubsan_pytest.log
grep -q " passed" tools/build/ubsan_pytest.log && ! grep -q "failed\|runtime error" tools/build/ubsan_pytest.log || { echo "ubsan: findings or failed cases"; grep -m3 "runtime error" tools/build/ubsan_pytest.log; rc=1; }
# WIDE=1: more of the GPU suite on the plain (uninstrumented) host build, ~4 minutes -- tiny-fixture parity in every mode, incremental
# residual, determinism, re-initialisation, burn-in / sqrt mode, live chains on shared Philox streams, checkpoint / resume
if [ "${WIDE:-0}" = 1 ]; then
  bash tools/build_emu.sh > /dev/null || exit 1
  SBMF_EMULATED=1 SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu.so python -m pytest tests/test_parity_gpu.py tests/test_state_resume.py -q \
    -k "((zero_noise_10_sweeps or ng_mode_zero_noise) and tiny) or incremental_residual or determinism or reset_by_init or burn_in_and_sqrt or live_same_philox or (resume and not cli)" \
    2>&1 | tail -2 | tee tools/build/wide_pytest.log
  grep -q " passed" tools/build/wide_pytest.log && ! grep -q "failed" tools/build/wide_pytest.log || { echo "wide: cases failed"; rc=1; }
  # the host CLI (bin/sbmf) with the host build preloaded over libsbmf_cuda.so: CLI == binding, -dump_xt, -save_state / -load_state
  SBMF_EMULATED=1 LD_PRELOAD=$PWD/tools/build/libsbmf_cuda_emu.so SBMF_LIB_PATH=$PWD/tools/build/libsbmf_cuda_emu.so python -m pytest \
    tests/test_cli.py tests/test_xt_format.py tests/test_state_resume.py -q -m gpu -k cli 2>&1 | tail -2 | tee tools/build/wide_cli.log
  grep -q " passed" tools/build/wide_cli.log && ! grep -q "failed" tools/build/wide_cli.log || { echo "wide: CLI cases failed"; rc=1; }
fi
echo "sbmf_sanitize rc=$rc"; exit $rc
