#!/usr/bin/env python3
"""tests/golden/make_ml100k_npz.py -- pack the reference's ML-100K triple fixtures (data/m100k/{train,test}_sbpmf,
SURVEY.md App. C.1) into one small npz so the GPU box (which has no /root/reference) can run the parity tests.
Rating DATA only, no reference source.  Run in the build container: python tests/golden/make_ml100k_npz.py"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SBMF_REF", "/root/reference")


def load(p):
    a = np.loadtxt(p, dtype=np.float64, ndmin=2)
    assert np.all(a[:, 2] == np.round(a[:, 2])) and a[:, :2].max() < 65536
    return a[:, 0].astype(np.uint16), a[:, 1].astype(np.uint16), a[:, 2].astype(np.uint8)


tu, ti, tr = load(os.path.join(REF, "data/m100k/train_sbpmf"))
su, si, sr = load(os.path.join(REF, "data/m100k/test_sbpmf"))
assert tu.size == 90570 and su.size == 9430
np.savez_compressed(os.path.join(HERE, "ml100k.npz"), train_user=tu, train_item=ti, train_rating=tr, test_user=su, test_item=si,
                    test_rating=sr)
print("wrote ml100k.npz", os.path.getsize(os.path.join(HERE, "ml100k.npz")), "bytes")
