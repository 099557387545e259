#!/usr/bin/env python3
"""tests/golden/make_split_fixture.py -- the reference's SECOND real fixture as a golden case: data/m1m/m100k/{train,test}_sbpmf
(the ML-100K 80k/20k split whose item ids were never re-based, so items 0..942 are EMPTY rows; 79,999 train
and 19,999 test ratings; the test set only touches 459 users).  Needs /root/reference and `make -C oracle ref`.  Writes

  ml100k_split.npz                         the rating data (no reference source), so the GPU box can run parity on it
  ref_ml100k_split_K20_T100_rmse.txt       100 "rmse is" values printed by the UNMODIFIED gibbs_sbpmf2.cpp on it
  ref_ml100k_split_K20_T10_{live,zero}.json  rmse text + sha256 / sample of the sampler-argument log (shim build, unmodified source)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
from run_ref import ref_binary, run_ref  # noqa: E402
from make_golden import shim_case  # noqa: E402

REF = os.environ.get("SBMF_REF", "/root/reference")
TR, TE = os.path.join(REF, "data/m1m/m100k/train_sbpmf"), os.path.join(REF, "data/m1m/m100k/test_sbpmf")


def load(p):
    a = np.loadtxt(p, dtype=np.float64, ndmin=2)
    assert np.all(a[:, 2] == np.round(a[:, 2])) and a[:, :2].max() < 65536
    return a[:, 0].astype(np.uint16), a[:, 1].astype(np.uint16), a[:, 2].astype(np.uint8)


tu, ti, tr = load(TR)
su, si, sr = load(TE)
np.savez_compressed(os.path.join(HERE, "ml100k_split.npz"), train_user=tu, train_item=ti, train_rating=tr, test_user=su, test_item=si, test_rating=sr)
I, J = int(max(tu.max(), su.max())) + 1, int(max(ti.max(), si.max())) + 1
print("ml100k_split.npz:", tu.size, "train,", su.size, "test,", I, "x", J, "empty item rows:", J - np.unique(ti).size)
r = run_ref(ref_binary(20, 100), TR, TE, threads=1)
assert (r["num_rows"], r["num_users"], r["num_items"]) == (tu.size, I, J) and len(r["rmse_text"]) == 100
with open(os.path.join(HERE, "ref_ml100k_split_K20_T100_rmse.txt"), "w") as f:
    f.write("# unmodified reference gibbs_sbpmf2.cpp, data/m1m/m100k, D=20, T=100, OMP_NUM_THREADS=1, glibc rand seed 1\n")
    f.write("\n".join(r["rmse_text"]) + "\n")
shim_case("ml100k_split", TR, TE, I * 20 + 20 * J)
