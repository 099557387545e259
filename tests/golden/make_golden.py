#!/usr/bin/env python3
"""tests/golden/make_golden.py -- (re)generate the golden vectors from the REFERENCE ITSELF.

Needs /root/reference and `make -C oracle ref` (oracle/_ref/*).  Runs the reference's own gibbs_sbpmf2.cpp
(unmodified; `_shim` builds are the same unmodified source compiled against oracle/shim_random.h, which logs
every sampler call's arguments and offers the zero-noise mode) and writes small fixtures next to this file:

  ref_ml100k_K20_T100_rmse.txt          100 "rmse is" values, unmodified binary, ML-100K, 1 thread, glibc seed 1
  ref_<case>_K20_T10_{live,zero}.json   rmse text + sha256 and a 1-in-997 sample of the sampler-argument log
  tiny_unsorted.{train,test}            a 50x40 synthetic fixture: unsorted file order, empty rows, ids only in test

The reference has no tests or golden vectors of its own; these are the pins.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from run_ref import ref_binary, run_ref  # noqa: E402

ML = os.path.join(ROOT, "oracle", "_ref", "data", "m100k")


def write_tiny():
    rs = np.random.RandomState(20151001)
    I, J, n = 50, 40, 640
    pairs = set()
    while len(pairs) < n:
        u = int(rs.zipf(1.6)) % (I - 3)        # users 47..49 never appear in train
        j = int(rs.randint(0, J - 2))          # items 38, 39 never appear in train
        if j == 7:
            continue                            # item 7 empty
        pairs.add((u, j))
    pairs = list(pairs)
    rs.shuffle(pairs)                           # file order is NOT sorted by user
    r = rs.randint(1, 6, size=len(pairs))
    test_idx = set(rs.choice(len(pairs), 80, replace=False).tolist())
    with open(os.path.join(HERE, "tiny_unsorted.train"), "w") as ftr, open(os.path.join(HERE, "tiny_unsorted.test"), "w") as fte:
        for t, ((u, j), rr) in enumerate(zip(pairs, r)):
            (fte if t in test_idx else ftr).write(f"{u}\t{j}\t{rr}\n")
        fte.write("49\t39\t3\n")               # max ids come from the test file only ([T]:112-119)
        fte.write("48\t7\t4\n")


def log_summary(path):
    a = np.fromfile(path, dtype=np.float64).reshape(-1, 3)
    h = hashlib.sha256(a.tobytes()).hexdigest()
    idx = np.arange(0, a.shape[0], 997)
    return {"records": int(a.shape[0]), "sha256": h, "sample_stride": 997,
            "sample": [[float(x).hex() for x in a[i]] for i in idx]}


def shim_case(name, train, test, live_init, K=20, T=10, variant="T"):
    """variant "T" = top-level gibbs_sbpmf2.cpp, "S" = src/libfm/gibbs_sbpmf2.cpp (Normal-Gamma hyper-prior, no biases)"""
    for mode in ("live", "zero"):
        log = f"/tmp/sbmf_golden_{name}_{mode}.log"
        env = {"SBMF_SHIM_LOG": log}
        if mode == "zero":
            env.update({"SBMF_SHIM_MODE": "zero", "SBMF_SHIM_LIVE_INIT": str(live_init)})
        r = run_ref(ref_binary(K, T, shim=True, variant=variant), train, test, threads=1, env_extra=env)
        out = {"case": name, "K": K, "T": T, "mode": mode, "num_rows": r["num_rows"], "num_users": r["num_users"],
               "num_items": r["num_items"], "rmse": r["rmse_text"], "log": log_summary(log)}
        tag = "ref" if variant == "T" else "refS"
        with open(os.path.join(HERE, f"{tag}_{name}_K{K}_T{T}_{mode}.json"), "w") as f:
            json.dump(out, f, indent=1)
        os.remove(log)
        print(name, mode, r["rmse_text"][:3], "...")


def main():
    r = run_ref(ref_binary(20, 100), os.path.join(ML, "train_sbpmf"), os.path.join(ML, "test_sbpmf"), threads=1)
    assert (r["num_rows"], r["num_users"], r["num_items"]) == (90570, 943, 1682)
    with open(os.path.join(HERE, "ref_ml100k_K20_T100_rmse.txt"), "w") as f:
        f.write("# unmodified reference gibbs_sbpmf2.cpp, data/m100k, D=20, T=100, OMP_NUM_THREADS=1, glibc rand seed 1\n")
        f.write("\n".join(r["rmse_text"]) + "\n")
    shim_case("ml100k", os.path.join(ML, "train_sbpmf"), os.path.join(ML, "test_sbpmf"), 943 * 20 + 20 * 1682)
    write_tiny()
    shim_case("tiny_unsorted", os.path.join(HERE, "tiny_unsorted.train"), os.path.join(HERE, "tiny_unsorted.test"),
              50 * 20 + 20 * 40)
    # [S]: the sibling program src/libfm/gibbs_sbpmf2.cpp (unmodified; its input names are staged by run_ref)
    r = run_ref(ref_binary(20, 100, variant="S"), os.path.join(ML, "train_sbpmf"), os.path.join(ML, "test_sbpmf"), threads=1)
    with open(os.path.join(HERE, "refS_ml100k_K20_T100_rmse.txt"), "w") as f:
        f.write("# unmodified reference src/libfm/gibbs_sbpmf2.cpp, data/m100k, D=20, 100 sweeps, OMP_NUM_THREADS=1, glibc rand seed 1\n")
        f.write("\n".join(r["rmse_text"]) + "\n")
    shim_case("ml100k", os.path.join(ML, "train_sbpmf"), os.path.join(ML, "test_sbpmf"), 943 * 20 + 20 * 1682, variant="S")


if __name__ == "__main__":
    main()
