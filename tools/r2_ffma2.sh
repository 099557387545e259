#!/bin/bash
# tools/r2_ffma2.sh -- A/B of the packed-FP32 builds (csrc/gram.cuh: FFMA2=1 packed FMAs, FFMA2=2 also packed adds) on one B200.
#   here (no GPU):   bash tools/r2_ffma2.sh build      -> scalable-.../lib_f{0,1,2}/libsbmf_cuda.so (git-ignored, travel with gpurun)
#   on the GPU box:  gpurun -- 'bash tools/r2_ffma2.sh' -> gpurun_out/ffma2_*
# The builds must produce IDENTICAL chains (every sum receives the same fused multiply-adds / additions in the same order): the
# script compares RMSE history and final factors bit for bit before it compares speed.
set -u
P=scalable-bayesian-matrix-factorization_b200
if [ "${1:-}" = "build" ]; then
  for lvl in 0 1 2; do
    make -C $P -j8 FFMA2=$lvl OBJ=build_f$lvl LIB=lib_f$lvl/libsbmf_cuda.so lib_f$lvl/libsbmf_cuda.so 2>&1 | grep -i "error" ; ls -la $P/lib_f$lvl/libsbmf_cuda.so
  done
  exit 0
fi
O=gpurun_out; mkdir -p $O
for lvl in 0 1 2; do
  export SBMF_LIB_PATH=$PWD/$P/lib_f$lvl/libsbmf_cuda.so
  python - "$lvl" <<'PY'
import sys, numpy as np
sys.path.insert(0, "scalable-bayesian-matrix-factorization_b200"); import sbmf
lvl = sys.argv[1]
d = sbmf.synth_generate(71567, 10681, 11111111)
m = sbmf.SbmfModel(K=100, seed=5)
m.set_train(d["train_user"], d["train_item"], d["train_rating"], 71567, 10681); m.set_test(d["test_user"], d["test_item"], d["test_rating"])
m.init_factors(); m.sweep(4)
st = m.get_state(with_E=False); r, _ = m.rmse_history(0, 4)
import hashlib
np.savez(f"gpurun_out/ffma2_chain_{lvl}.npz", r=r, **{k: np.frombuffer(hashlib.sha256(np.ascontiguousarray(st[k]).tobytes()).digest(), np.uint8) for k in ("U", "V", "b_i", "b_j")})
print("FFMA2 =", lvl, "rmse", r)
PY
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli > $O/ffma2_bench_$lvl.json 2> $O/ffma2_bench_$lvl.err
done
python - <<'PY'
import json, numpy as np
a = np.load("gpurun_out/ffma2_chain_0.npz")
for lvl in (1, 2):
    b = np.load(f"gpurun_out/ffma2_chain_{lvl}.npz")
    print(f"FFMA2={lvl}: chain bit-identical to the scalar build:", all(np.array_equal(a[k], b[k]) for k in a.files))
for lvl in (0, 1, 2):
    d = json.loads(open(f"gpurun_out/ffma2_bench_{lvl}.json").read().strip().splitlines()[-1])
    print(f"FFMA2={lvl}", "%.3f ms/sweep" % d["ms_per_step"], d["phases_ms"])
PY
