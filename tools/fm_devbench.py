#!/usr/bin/env python3
"""tools/fm_devbench.py -- developer timing of the general FM Gibbs path (csrc/fm.cu); not the bench contract (bench.py).
Casts a synthetic MovieLens/Netflix-shaped rating matrix as a factorization machine (one-hot user + one-hot item [+ W dense
real-valued context attributes]) and times iterations of sbmf_fm_learn on one GPU (wall clock around a synchronising call).
Prints iterations/s and the algorithmic traffic of DESIGN.md 11: per iteration and design-matrix entry 16 B for the w draw,
24 B per factor for the v draws, 4 B per factor for the q rebuild, (8 + 4 K) B for the re-prediction."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
import sbmf  # noqa: E402

SHAPES = {"ml1m": (6040, 3706, 1000209), "ml10m": (71567, 10681, 10000000), "ml20m": (138493, 26744, 20000000),
          "netflix": (480189, 17770, 100480507)}
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="ml10m")
ap.add_argument("-K", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--wide", type=int, default=0, help="dense real-valued context attributes per case")
a = ap.parse_args()
I, J, N = SHAPES[a.shape]
d = sbmf.synth_generate(I, J, N, test_frac=0.1, seed=20151001)


def fm(u, i, r, seed):
    n, W = u.size, a.wide
    cols = [u.astype(np.uint32), (I + i).astype(np.uint32)] + [np.full(n, I + J + k, dtype=np.uint32) for k in range(W)]
    rs = np.random.RandomState(seed)
    vals = [np.ones(n, np.float32), np.ones(n, np.float32)] + [(np.round(rs.standard_normal(n) * 8) / 8).astype(np.float32) for _ in range(W)]
    return {"row_ptr": ((2 + W) * np.arange(n + 1)).astype(np.int64), "attr": np.stack(cols, axis=1).reshape(-1), "x": np.stack(vals, axis=1).reshape(-1),
            "y": r.astype(np.float32)}


tr, te = fm(d["train_user"], d["train_item"], d["train_rating"], 1), fm(d["test_user"], d["test_item"], d["test_rating"], 2)
p = I + J + a.wide + 1
group = np.concatenate([np.zeros(I), np.ones(J), np.full(a.wide + 1, 2)]).astype(np.uint32)
m = sbmf.FmModel(p, a.K, attr_group=group, seed=1)
t0 = time.perf_counter(); m.set_train(tr); m.set_test(te); m.init(); t_setup = time.perf_counter() - t0
m.learn(a.warmup); m.rmse_history(0, a.warmup)
t0 = time.perf_counter(); m.learn(a.iters); r = m.rmse_history(a.warmup, a.iters); dt = (time.perf_counter() - t0) / a.iters
nnz = int(tr["row_ptr"][-1])
alg = nnz * (16 + a.K * 24 + a.K * 4 + 8 + 4 * a.K)
print(json.dumps({"shape": a.shape, "K": a.K, "wide": a.wide, "cases": int(tr["y"].size), "nnz": nnz, "attributes": p, "runs": int(m.get_runs().size - 1),
                  "setup_s": round(t_setup, 3), "ms_per_iteration": round(dt * 1e3, 3), "iterations_per_s": round(1 / dt, 3),
                  "algorithmic_GBs": round(alg / dt / 1e9, 1), "rmse_test": [round(float(x), 5) for x in r[1]]}))
