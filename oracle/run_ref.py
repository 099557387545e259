#!/usr/bin/env python3
"""oracle/run_ref.py -- run a reference binary from oracle/_ref on given triple files.  TEST INFRASTRUCTURE.

The reference program (gibbs_sbpmf2.cpp, "[T]") takes no arguments: it opens ../../data/ra.train_sbpmf and
../../data/ra.test_sbpmf relative to its CWD ([T]:32, 98) and prints one "rmse is <v>" line per sweep
([T]:635).  This helper stages a scratch directory <tmp>/data/ra.{train,test}_sbpmf (symlinks), runs the
binary from <tmp>/a/b and returns the header values and the RMSE trajectory.
"""
import os
import re
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def ref_binary(K=20, T=100, shim=False, variant="T"):
    name = ("gibbs_ref" if variant == "T" else "gibbs_refS") + ("_shim" if shim else "")
    if not (K == 20 and T == 100):
        name += f"_D{K}_T{T}"
    path = os.path.join(REF_DIR, name)
    return path if os.path.exists(path) else None


def run_ref(binary, train, test, threads=1, env_extra=None, timeout=None):
    """Returns dict(rmse=[...], num_rows=, num_users=, num_items=, wall_s=, stdout=)."""
    with tempfile.TemporaryDirectory(prefix="sbmf_ref_") as tmp:
        os.makedirs(os.path.join(tmp, "data"))
        os.makedirs(os.path.join(tmp, "a", "b"))
        os.symlink(os.path.abspath(train), os.path.join(tmp, "data", "ra.train_sbpmf"))
        os.symlink(os.path.abspath(test), os.path.join(tmp, "data", "ra.test_sbpmf"))
        # [S] = src/libfm/gibbs_sbpmf2.cpp opens ../../data/m100k/train (passes 1-2), .../train_sbpmf (pass 3) and .../test ([S]:33-192)
        os.makedirs(os.path.join(tmp, "data", "m100k"))
        for name in ("train", "train_sbpmf"):
            os.symlink(os.path.abspath(train), os.path.join(tmp, "data", "m100k", name))
        os.symlink(os.path.abspath(test), os.path.join(tmp, "data", "m100k", "test"))
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(threads)
        if env_extra:
            env.update(env_extra)
        t0 = time.perf_counter()
        proc = subprocess.run([binary], cwd=os.path.join(tmp, "a", "b"), env=env, capture_output=True,
                              text=True, timeout=timeout)
        out = proc.stdout
        wall = time.perf_counter() - t0
    # The reference may SIGSEGV in its own cleanup AFTER the last sweep ([T]:648-666 delete[]s rows it never
    # allocated when some id in [0,max] has no train rating).  Every "rmse is" line is flushed (std::endl)
    # before that, so the trajectory is complete; the return code is reported, not trusted.
    res = {"stdout": out, "wall_s": wall, "returncode": proc.returncode}
    res["rmse"] = [float(x) for x in re.findall(r"^rmse is (\S+)$", out, flags=re.M)]
    res["rmse_text"] = re.findall(r"^rmse is (\S+)$", out, flags=re.M)
    for key, pat in (("num_rows", r"number rows =(\d+)"), ("num_users", r"number of user =(\d+)"),
                     ("num_items", r"number of items =(\d+)")):
        m = re.search(pat, out)
        res[key] = int(m.group(1)) if m else None
    return res


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("train")
    ap.add_argument("test")
    ap.add_argument("-K", type=int, default=20)
    ap.add_argument("-T", type=int, default=100)
    ap.add_argument("--shim", action="store_true")
    ap.add_argument("--threads", type=int, default=1)
    a = ap.parse_args()
    b = ref_binary(a.K, a.T, a.shim)
    if b is None:
        sys.exit(f"no reference binary for K={a.K} T={a.T} shim={a.shim} under {REF_DIR} (make -C oracle ref)")
    r = run_ref(b, a.train, a.test, threads=a.threads)
    print(f"number rows ={r['num_rows']}\nnumber of user ={r['num_users']}\nnumber of items ={r['num_items']}")
    for t in r["rmse_text"]:
        print("rmse is", t)
    print(f"# wall {r['wall_s']:.3f}s", file=sys.stderr)
