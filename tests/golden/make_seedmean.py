#!/usr/bin/env python3
"""tests/golden/make_seedmean.py -- the reference's per-sweep test-RMSE DISTRIBUTION over rand() seeds on ML-100K (K=20).

The reference never calls srand, so its one printed trajectory (ref_ml100k_K20_T100_rmse.txt) is a single
draw; at sweep 0 it sits +2.9 sigma from the mean of its own seed distribution.  The live-sampling parity bar (0.003 on the
per-sweep test RMSE) is therefore checked against the seed-MEAN trajectory.  That mean is produced by the oracle in
NOISE_RAND mode, which tests/test_oracle.py proves bit-identical to the unmodified reference (same rand() stream, same
samplers, same 100 printed values at seed 1), run from 128 well-separated srand() seeds.
Writes ref_ml100k_K20_T40_seedmean.json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as orc  # noqa: E402

d = np.load(os.path.join(HERE, "ml100k.npz"))
tu, ti, tr, su, si, sr = [d[k] for k in ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating")]
I, J, K, T, NSEED = 943, 1682, 20, 40, 128
runs = []
for s in range(NSEED):
    o = orc.Oracle(tu, ti, tr, su, si, sr, I, J, K, noise=orc.NOISE_RAND)
    o.srand(1000003 * (s + 1) + 17)
    o.init_factors()
    runs.append(o.sweep(T)[0])
    o.close()
a = np.array(runs)
out = {"case": "ml100k", "K": K, "T": T, "seeds": NSEED, "srand": "1000003*(s+1)+17, s=0..127",
       "mean": [float(x) for x in a.mean(0)], "std": [float(x) for x in a.std(0)]}
json.dump(out, open(os.path.join(HERE, "ref_ml100k_K20_T40_seedmean.json"), "w"), indent=1)
print("mean[:5]", a.mean(0)[:5], "std[:5]", a.std(0)[:5])
