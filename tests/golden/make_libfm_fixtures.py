#!/usr/bin/env python3
"""tests/golden/make_libfm_fixtures.py -- libFM text + binary renderings of the tiny_unsorted fixture.

The binary files are written by the REFERENCE's own converter (src/libfm/tools/convert.cpp, built unmodified into
oracle/_ref/convert by `make -C oracle ref`), so the host program's binary reader is tested against the reference's bytes.
Item features are numbered after the users (item feature = 50 + item id), like scripts/triple_format_to_libfm.pl does."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CONVERT = os.path.join(ROOT, "oracle", "_ref", "convert")
TRANSPOSE = os.path.join(ROOT, "oracle", "_ref", "transpose")   # src/libfm/tools/transpose.cpp, unmodified
NUM_USERS = 50

for split in ("train", "test"):
    src = os.path.join(HERE, f"tiny_unsorted.{split}")
    txt = os.path.join(HERE, f"tiny_libfm.{split}")
    with open(src) as f, open(txt, "w") as o:
        for line in f:
            u, i, r = line.split()
            o.write(f"{r} {u}:1 {int(i) + NUM_USERS}:1\n")
    subprocess.run([CONVERT, "-ifile", txt, "-ofilex", txt + "_bin.x", "-ofiley", txt + "_bin.y"], check=True, capture_output=True)
    print(split, os.path.getsize(txt + "_bin.x"), os.path.getsize(txt + "_bin.y"))
# the transposed design matrix of the training data, written by the reference's own transpose tool
xt = os.path.join(HERE, "tiny_libfm.train_bin.xt")
subprocess.run([TRANSPOSE, "-ifile", os.path.join(HERE, "tiny_libfm.train_bin.x"), "-ofile", xt], check=True, capture_output=True)
print("xt", os.path.getsize(xt))
