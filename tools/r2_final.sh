#!/bin/bash
# tools/r2_final.sh -- 1 GPU: the round-end sequence on the final tree: GPU suite, smoke(), default bench (short)
set -u
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/final_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-point --no-cli > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
python - <<'E'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("%.3f ms/sweep %.1f G/s e2e %.1f G/s" % (d["ms_per_step"], d["value"] / 1e9, d["e2e"]["value"] / 1e9), d["phases_ms"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
E
