#!/bin/bash
# tools/r2_ffma2.sh -- A/B of the packed (FFMA2) Gram accumulation on one B200 (gpurun -- 'bash tools/r2_ffma2.sh'); outputs in gpurun_out/.
# The two builds must produce IDENTICAL chains (the packed form performs the same fused multiply-adds per sum, csrc/gram.cuh):
# the script compares the RMSE history and the final factors of both builds bit for bit before it compares their speed.
set -u
O=gpurun_out; mkdir -p $O
P=scalable-bayesian-matrix-factorization_b200
run_one() {   # $1 = tag
  python - "$1" <<'PY'
import sys, numpy as np
sys.path.insert(0, "scalable-bayesian-matrix-factorization_b200"); import sbmf
tag = sys.argv[1]
d = sbmf.synth_generate(71567, 10681, 11111111)
m = sbmf.SbmfModel(K=100, seed=5)
m.set_train(d["train_user"], d["train_item"], d["train_rating"], 71567, 10681); m.set_test(d["test_user"], d["test_item"], d["test_rating"])
m.init_factors(); m.sweep(4)
st = m.get_state(with_E=False); r, _ = m.rmse_history(0, 4)
np.savez(f"gpurun_out/ffma2_{tag}.npz", U=st["U"], V=st["V"], b_i=st["b_i"], b_j=st["b_j"], r=r)
print(tag, "rmse", r)
PY
}
make -C $P clean >/dev/null; make -C $P -j8 FFMA2=0 >/dev/null 2>&1; run_one scalar
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ffma2_bench_scalar.json 2> $O/ffma2_bench_scalar.err
make -C $P clean >/dev/null; make -C $P -j8 FFMA2=1 >/dev/null 2>&1; run_one packed
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ffma2_bench_packed.json 2> $O/ffma2_bench_packed.err
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/ffma2_parity_packed.log 2>&1; echo "parity (packed build) rc=$?"; tail -2 $O/ffma2_parity_packed.log
python - <<'PY'
import json, numpy as np
a, b = np.load("gpurun_out/ffma2_scalar.npz"), np.load("gpurun_out/ffma2_packed.npz")
print("bit-identical chains:", all(np.array_equal(a[k], b[k]) for k in a.files))
for t in ("scalar", "packed"):
    d = json.loads(open(f"gpurun_out/ffma2_bench_{t}.json").read().strip().splitlines()[-1])
    print(t, "%.3f ms/sweep" % d["ms_per_step"], d["phases_ms"])
PY
make -C $P clean >/dev/null; make -C $P -j8 >/dev/null 2>&1   # back to the default build
