#!/usr/bin/env python3
"""tests/golden/make_fm_golden.py -- golden vectors for the general FM Gibbs path, produced by the UNMODIFIED
reference libFM (src/libfm/libfm.cpp built against oracle/shim_random_libfm.h -> oracle/_ref/libFM_shim, `make -C oracle ref`).
Run in the build container (needs /root/reference at build time); the outputs are committed.

  fm_general.{train,test,meta}     a design matrix that is NOT matrix factorisation: one-hot users and items, a multi-hot genre
                                   block with fractional values, two dense real-valued attributes, one attribute that only the
                                   test file mentions; four attribute groups.  All values are dyadic, so text -> float is exact.
  libfm_<fixture>_<mode>.json      per run: the "#Iter= i Train= Test=" values as printed, the number of sampler calls and the
                                   sha256 of the complete sampler-argument stream (shim_random.h log format).
Modes: live (glibc rand() re-seeded by SBMF_SHIM_SEED), zero (SURVEY 8c zero-noise, the factor init stays live), als (libFM's
own -method als: do_sampling = do_multilevel = 0 with -regular).
"""
import hashlib
import json
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIBFM = os.path.join(ROOT, "oracle", "_ref", "libFM_shim")
SEED = 7
ITERS = 10

NU, NI, NG, NC = 30, 40, 6, 2          # users, items, genres, dense context attributes (+ 1 test-only attribute)
OFF_I, OFF_G, OFF_C = NU, NU + NI, NU + NI + NG


def write_general():
    rs = np.random.RandomState(20151001)
    item_genres = [sorted(rs.choice(NG, size=rs.choice([1, 2, 4]), replace=False)) for _ in range(NI)]
    bu, bi, bg = 0.5 * rs.standard_normal(NU), 0.5 * rs.standard_normal(NI), 0.3 * rs.standard_normal(NG)
    fu, fi = 0.6 * rs.standard_normal((NU, 2)), 0.6 * rs.standard_normal((NI, 2))
    wc = np.array([0.8, -0.5])

    def cases(n, test):
        lines = []
        for c in range(n):
            u, i = rs.randint(NU), rs.randint(NI)
            ctx = np.round(rs.standard_normal(NC) * 16) / 16            # dyadic
            g = item_genres[i]
            y = 3.5 + bu[u] + bi[i] + sum(bg[k] for k in g) / len(g) + fu[u] @ fi[i] + ctx @ wc + 0.4 * rs.standard_normal()
            y = min(5.0, max(0.5, round(y * 2) / 2))
            feats = [(u, 1.0), (OFF_I + i, 1.0)] + [(OFF_G + k, 1.0 / len(g)) for k in g] + [(OFF_C + k, ctx[k]) for k in range(NC) if ctx[k] != 0]
            if test and c % 7 == 0:
                feats.append((OFF_C + NC, 0.5))                        # an attribute the train file never mentions
            if c % 3 == 0:
                feats = feats[::-1]                                    # libFM does not need ascending ids within a line
            lines.append(f"{y:g} " + " ".join(f"{a}:{v:g}" for a, v in feats))
        return lines

    with open(os.path.join(HERE, "fm_general.train"), "w") as f:
        f.write("\n".join(cases(600, False)) + "\n")
    with open(os.path.join(HERE, "fm_general.test"), "w") as f:
        f.write("\n".join(cases(150, True)) + "\n")
    with open(os.path.join(HERE, "fm_general.meta"), "w") as f:
        f.write("\n".join(["0"] * NU + ["1"] * NI + ["2"] * NG + ["3"] * (NC + 2)) + "\n")   # + the phantom attribute of [L]:326


def run_libfm(train, test, K, p, mode, meta=None):
    args = [LIBFM, "-task", "r", "-train", train, "-test", test, "-dim", f"1,1,{K}", "-iter", str(ITERS), "-init_stdev", "0.1"]
    if meta:
        args += ["-meta", meta]
    env = dict(os.environ, SBMF_SHIM_SEED=str(SEED), OMP_NUM_THREADS="1")
    if mode == "als":
        args += ["-method", "als", "-regular", "0.25,1,4"]
    else:
        args += ["-method", "mcmc"]
    if mode == "zero":
        env.update(SBMF_SHIM_MODE="zero", SBMF_SHIM_LIVE_INIT=str(K * p + p))
    with tempfile.TemporaryDirectory(prefix="libfm_") as tmp:     # libFM writes v_file.txt and test_rmse_* into its CWD
        log = os.path.join(tmp, "args.bin")
        env["SBMF_SHIM_LOG"] = log
        out = subprocess.run(args, cwd=tmp, env=env, capture_output=True, text=True, check=True).stdout
        raw = open(log, "rb").read()
    rows = re.findall(r"^#Iter=\s*(\d+)\tTrain=(\S+)\tTest=(\S+)$", out, flags=re.M)
    assert len(rows) == ITERS, out[-2000:]
    return {"fixture": os.path.basename(train), "mode": mode, "K": K, "num_attr": p, "iters": ITERS, "seed": SEED,
            "train": [r[1] for r in rows], "test": [r[2] for r in rows], "n_sampler_calls": len(raw) // 24,
            "sha256_args": hashlib.sha256(raw).hexdigest(), "argv": " ".join(os.path.basename(a) if os.sep in a else a for a in args[1:])}


def main():
    write_general()
    jobs = [("tiny_libfm", 4, None), ("fm_general", 3, os.path.join(HERE, "fm_general.meta"))]
    for name, K, meta in jobs:
        tr, te = os.path.join(HERE, name + ".train"), os.path.join(HERE, name + ".test")
        p = 0
        for path in (tr, te):
            for line in open(path):
                for t in line.split()[1:]:
                    p = max(p, int(t.split(":")[0]) + 1)
        p += 1     # [L]:326: num_all_attribute = max(train.num_feature, test.num_feature) + 1 -- one attribute beyond the largest id
        for mode in ("live", "zero", "als"):
            g = run_libfm(tr, te, K, p, mode, meta)
            with open(os.path.join(HERE, f"libfm_{name}_{mode}.json"), "w") as f:
                json.dump(g, f, indent=1)
            print(name, mode, g["n_sampler_calls"], g["test"][-1])


if __name__ == "__main__":
    main()
