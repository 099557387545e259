#!/usr/bin/env python3
"""tools/devbench.py -- developer timing loop (not the bench contract; see bench.py): per-phase CUDA-event times of the sweep."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
import sbmf  # noqa: E402

SHAPES = {"ml1m": (6040, 3706, 1000209), "ml10m": (71567, 10681, 10000000), "ml20m": (138493, 26744, 20000000),
          "netflix": (480189, 17770, 100480507), "1b": (10000000, 1000000, 1000000000)}

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="ml10m")
ap.add_argument("-K", type=int, default=100)
ap.add_argument("--sweeps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--s_user", type=float, default=0.8)
ap.add_argument("--s_item", type=float, default=1.0)
a = ap.parse_args()
I, J, N = SHAPES[a.shape]
t0 = time.time()
if a.shape == "1b":   # 1e13 pairs: sparse host sampler (csrc/synth_host.cpp); N counts train + test here
    s = sbmf.synth_generate_host(I, J, N, s_user=a.s_user, s_item=a.s_item)
else:
    s = sbmf.synth_generate(I, J, N, s_user=a.s_user, s_item=a.s_item)
t1 = time.time()
du = np.bincount(s["train_user"], minlength=I); di = np.bincount(s["train_item"], minlength=J)
print(f"synth {a.shape}: train {s['train_user'].size} test {s['test_user'].size} in {t1 - t0:.1f}s; user deg max {du.max()} med {int(np.median(du))}"
      f" item deg max {di.max()} med {int(np.median(di))}", flush=True)
m = sbmf.SbmfModel(K=a.K, sample_mode=a.mode)
t0 = time.time()
m.set_train(s["train_user"], s["train_item"], s["train_rating"], I, J)
m.set_test(s["test_user"], s["test_item"], s["test_rating"])
m.init_factors()
print(f"set_train+init {time.time() - t0:.2f}s", flush=True)
m.sweep(a.warmup)
m.reset_timing()
m.sweep(a.sweeps)
t = m.timing()
n = t["sweeps"]
ms = t["ms_total"] / n
fu = s["train_user"].size * a.K / (ms * 1e-3)
print({k: (round(v / n, 3) if k.startswith("ms_") else v) for k, v in t.items()})
print(f"{ms:.3f} ms/sweep  {1e3 / ms:.2f} sweeps/s  {fu / 1e9:.2f} G factor-updates/s  roofline(24B/fu @6538.6GB/s) frac {fu * 24 / 6538.6e9:.3f}")
print("rmse", m.rmse_history(0, a.warmup + a.sweeps)[0])
