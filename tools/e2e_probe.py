#!/usr/bin/env python3
"""tools/e2e_probe.py -- developer probe: wall-clock pieces of the end-to-end job (set_train, get_pred) outside bench.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
import sbmf  # noqa: E402

I, J, N = 480189, 17770, 100480507
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s = sbmf.synth_generate(I, J, int(round(N / 0.9)))
hb = {}
for k in ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating"):
    hb[k] = sbmf.pinned_empty(s[k].size, s[k].dtype)
    hb[k][:] = s[k]
pred = sbmf.pinned_empty(s["test_user"].size, np.float32)
for rep in range(2):
    m = sbmf.SbmfModel(K=K, sample_mode=0)
    m.set_timing_enabled(0)
    m.synchronize()
    t0 = time.perf_counter()
    m.set_train(hb["train_user"], hb["train_item"], hb["train_rating"], I, J)
    t1 = time.perf_counter()
    m.set_test(hb["test_user"], hb["test_item"], hb["test_rating"])
    t2 = time.perf_counter()
    m.init_factors()
    t3 = time.perf_counter()
    for _ in range(4):
        m.sweep(1)
        m.eval()
    t4 = time.perf_counter()
    tp = []
    for _ in range(3):
        a = time.perf_counter()
        m._ck(m.lib.sbmf_cuda_get_pred(m.h, pred.ctypes.data))
        tp.append(time.perf_counter() - a)
    a = time.perf_counter()
    m.synchronize()
    ts = time.perf_counter() - a
    print(f"rep {rep}: set_train {t1 - t0:.4f}  set_test {t2 - t1:.4f}  init {t3 - t2:.4f}  4 sweeps {t4 - t3:.4f}  get_pred {[round(x, 4) for x in tp]}  sync {ts:.4f}",
          flush=True)
    m.close()
