#!/bin/bash
# tools/build_emu.sh [name] [extra g++ flags...] -- TEST INFRASTRUCTURE: the host build of libsbmf_cuda (all kernels of csrc/kernels.cu,
# storage.cu, fm.cu + the C ABI of api.cu) against tools/emu_include, where CTAs execute on host threads.  Output: tools/build/<name>
# (default libsbmf_cuda_emu.so).  Used by tests/test_sbmf_simt_emulation.py and tools/sbmf_sanitize.sh; never part of the product.
set -eu
cd "$(dirname "$0")/.."
NAME=${1:-libsbmf_cuda_emu.so}; shift || true
B=tools/build/obj_${NAME%.so}; mkdir -p $B
C=scalable-bayesian-matrix-factorization_b200/csrc
FLAGS="-O1 -std=c++17 -pthread -fPIC -DSBMF_SIMT_EMU -DSBMF_FFMA2=0 -I tools/emu_include -I include -I $C $*"
pids=""
for f in api kernels storage fm; do g++ $FLAGS -x c++ -c $C/$f.cu -o $B/$f.o & pids="$pids $!"; done
for f in plan checkpoint xt_writer synth_host; do g++ $FLAGS -c $C/$f.cpp -o $B/$f.o & pids="$pids $!"; done
g++ $FLAGS -c tools/emu_stubs.cpp -o $B/emu_stubs.o & pids="$pids $!"
for p in $pids; do wait $p; done
g++ -shared -pthread $* -o tools/build/$NAME $B/*.o
echo tools/build/$NAME
