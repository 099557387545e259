#!/bin/bash
# tools/r2_call12.sh -- 1 GPU: A/B (evict-first stream hints build, fold_item), FM workloads after the final-reduction fix, then
# the complete default bench line (cpu_baseline, e2e, e2e_cli, full-config point) and the reference arm for profiles/r2
set -u
O=gpurun_out; mkdir -p $O
P=scalable-bayesian-matrix-factorization_b200
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
i=0
run() {
  i=$((i+1))
  timeout 400 $B "$@" > $O/c12_$i.json 2> $O/c12_$i.err
  python - "${SBMF_LIB_PATH:-default lib} $*" $O/c12_$i.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-60s %.3f ms/sweep  user %.3f item %.3f exch %.3f  top %.1f us  rmse %.6f" % (sys.argv[1][-60:], d["ms_per_step"], d["phases_ms"]["ms_user_phase"], d["phases_ms"]["ms_item_phase"], d["phases_ms"]["ms_exchange"], d["roofline"]["us_per_launch"], d["rmse_after_timed"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
E
}
run
SBMF_LIB_PATH=$PWD/$P/lib_sh/libsbmf_cuda.so run
run --options fold_item=1
run --workload ml10m_k100
run --workload ml10m_k100 --options slice_len=2048
for w in fm_mf_ml10m fm_wide_ml1m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $O/c12_$w.json 2> $O/c12_$w.err; echo "$w ours rc=$?"; cut -c1-200 $O/c12_$w.json
done
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $O/c12_bench_full.json 2> $O/c12_bench_full.err; echo "full bench rc=$?"; cut -c1-300 $O/c12_bench_full.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/c12_bench_ref.json 2> $O/c12_bench_ref.err; echo "reference arm rc=$?"; cut -c1-600 $O/c12_bench_ref.json
