"""tests/fm_oracle_py.py -- ctypes binding of oracle/libfm_oracle.so (the CPU restatement of libFM's MCMC learner)
and a reader for libFM's text format.  TEST INFRASTRUCTURE: only tests/ and __graft_entry__.smoke() import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libfm_oracle.so")

NOISE_RAND, NOISE_ZERO = 0, 1

_lib = None


class _Config(C.Structure):
    _fields_ = [("num_attr", C.c_uint32), ("num_groups", C.c_uint32), ("K", C.c_uint32), ("k0", C.c_int32), ("k1", C.c_int32),
                ("do_sample", C.c_int32), ("do_multilevel", C.c_int32), ("noise", C.c_int32), ("init_stdev", C.c_double),
                ("reg0", C.c_double), ("regw", C.c_double), ("regv", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        L = C.CDLL(LIB)
        L.fm_oracle_create.restype = C.c_void_p
        L.fm_oracle_create.argtypes = [C.POINTER(_Config), C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 5
        L.fm_oracle_destroy.argtypes = [C.c_void_p]
        L.fm_oracle_srand.argtypes = [C.c_uint]
        L.fm_oracle_set_log.argtypes = [C.c_void_p, C.c_char_p]
        L.fm_oracle_set_log.restype = C.c_int
        L.fm_oracle_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fm_oracle_learn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.fm_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.fm_oracle_get_columns.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def read_libfm(path):
    """libFM text (Data.h:184-278): `target id:value id:value ...` per line -> dict(row_ptr int64, attr uint32, x float32, y float32).
    Values go through float32 exactly like sscanf("%f")."""
    row_ptr, attr, x, y = [0], [], [], []
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            tok = line.split()
            y.append(np.float32(tok[0]))
            for t in tok[1:]:
                a, v = t.split(":")
                attr.append(int(a))
                x.append(np.float32(v))
            row_ptr.append(len(attr))
    return {"row_ptr": np.asarray(row_ptr, dtype=np.int64), "attr": np.asarray(attr, dtype=np.uint32), "x": np.asarray(x, dtype=np.float32),
            "y": np.asarray(y, dtype=np.float32)}


def num_attributes(*mats):
    """[L]:326: num_all_attribute = max(train.num_feature, test.num_feature) + 1, and a file's num_feature is its largest id + 1
    (Data.h:205-221): this fork of libFM carries one attribute beyond the largest id (no data; drawn from its prior each sweep)"""
    return max(int(m["attr"].max()) + 1 if m["attr"].size else 0 for m in mats) + 1


class FmOracle:
    def __init__(self, train, test, K, num_attr=None, attr_group=None, k0=1, k1=1, do_sample=1, do_multilevel=1, noise=NOISE_RAND,
                 init_stdev=0.1, reg=(0.0, 0.0, 0.0)):
        self.L = lib()
        self.p = int(num_attr if num_attr is not None else num_attributes(train, test))
        self.K = int(K)
        self.group = None if attr_group is None else np.ascontiguousarray(attr_group, dtype=np.uint32)
        self.G = 1 if self.group is None else int(self.group.max()) + 1
        self.n, self.nt = int(train["y"].size), int(test["y"].size)
        self.nnz = int(train["row_ptr"][-1])
        cfg = _Config(self.p, self.G, self.K, k0, k1, do_sample, do_multilevel, noise, init_stdev, reg[0], reg[1], reg[2])
        self._keep = [np.ascontiguousarray(train[k]) for k in ("row_ptr", "attr", "x", "y")] + \
                     [np.ascontiguousarray(test[k]) for k in ("row_ptr", "attr", "x", "y")]
        a = self._keep
        self.h = self.L.fm_oracle_create(C.byref(cfg), self.n, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), self.nt, _p(a[4]), _p(a[5]),
                                         _p(a[6]), _p(a[7]), _p(self.group))
        if not self.h:
            raise RuntimeError("fm_oracle_create failed (attribute id or group out of range?)")

    def srand(self, seed):
        self.L.fm_oracle_srand(seed)

    def set_log(self, path):
        if self.L.fm_oracle_set_log(self.h, path.encode() if path else None) != 0:
            raise OSError(path)

    def init(self, w=None, v=None):
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        assert w is None or w.shape == (self.p,)
        assert v is None or v.shape == (self.K, self.p)
        self.L.fm_oracle_init(self.h, _p(w), _p(v))

    def learn(self, iters):
        tr, te = np.zeros(iters), np.zeros(iters)
        self.L.fm_oracle_learn(self.h, iters, _p(tr), _p(te))
        return tr, te

    def state(self):
        s = {"w": np.zeros(self.p), "v": np.zeros((self.K, self.p)), "w_mu": np.zeros(self.G), "w_lambda": np.zeros(self.G),
             "v_mu": np.zeros((self.G, self.K)), "v_lambda": np.zeros((self.G, self.K)), "e": np.zeros(self.n),
             "pred_sum": np.zeros(max(self.nt, 1)), "scal": np.zeros(2)}
        self.L.fm_oracle_get_state(self.h, *[_p(s[k]) for k in ("w", "v", "w_mu", "w_lambda", "v_mu", "v_lambda", "e", "pred_sum", "scal")])
        s["w0"], s["alpha"] = float(s["scal"][0]), float(s["scal"][1])
        s["pred_sum"] = s["pred_sum"][:self.nt]
        return s

    def columns(self):
        cp, ci, x = np.zeros(self.p + 1, dtype=np.int64), np.zeros(max(self.nnz, 1), dtype=np.uint32), np.zeros(max(self.nnz, 1), dtype=np.float32)
        self.L.fm_oracle_get_columns(self.h, _p(cp), _p(ci), _p(x))
        return {"col_ptr": cp, "case": ci[:self.nnz], "x": x[:self.nnz]}

    def close(self):
        if self.h:
            self.L.fm_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass
