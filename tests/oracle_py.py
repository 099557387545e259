"""tests/oracle_py.py -- ctypes binding of oracle/liboracle.so (the CPU restatement of gibbs_sbpmf2.cpp).
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

NOISE_RAND, NOISE_ZERO, NOISE_PHILOX = 0, 1, 2
STDEV_REF, STDEV_SQRT = 0, 1

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.sbmf_oracle_create.restype = C.c_void_p
        L.sbmf_oracle_create.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint64]
        L.sbmf_oracle_destroy.argtypes = [C.c_void_p]
        L.sbmf_oracle_srand.argtypes = [C.c_uint]
        L.sbmf_oracle_set_variant.argtypes = [C.c_void_p, C.c_int]
        L.sbmf_oracle_init_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.sbmf_oracle_sweep.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.sbmf_oracle_set_log.argtypes = [C.c_void_p, C.c_char_p]
        L.sbmf_oracle_set_log.restype = C.c_int
        for name, n in (("get_U", 1), ("get_V", 1), ("get_bias", 2), ("get_bias_hypers", 4), ("get_dim_hypers", 4), ("get_scalars", 1),
                        ("get_E", 1), ("get_pred_mean", 1), ("get_layout", 6)):
            getattr(L, "sbmf_oracle_" + name).argtypes = [C.c_void_p] + [C.c_void_p] * n
        L.sbmf_oracle_read_triples.restype = C.c_int64
        L.sbmf_oracle_read_triples.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sbmf_oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sbmf_oracle_philox_normal_f32.restype = C.c_float
        L.sbmf_oracle_philox_normal_f32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.sbmf_oracle_philox_normal_f64.restype = C.c_double
        L.sbmf_oracle_philox_normal_f64.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, tu, ti, tr, su, si, sr, num_users, num_items, K, noise=NOISE_RAND, stdev_mode=STDEV_REF, seed=1, variant=0):
        """variant 0 = top-level gibbs_sbpmf2.cpp ([T]); 1 = src/libfm/gibbs_sbpmf2.cpp ([S], Normal-Gamma hypers, no biases) as
        committed; 2 = [S] with its line-412 slip corrected"""
        self.L = lib()
        self.tu, self.ti = np.ascontiguousarray(tu, np.uint32), np.ascontiguousarray(ti, np.uint32)
        self.tr = np.ascontiguousarray(tr, np.float64)
        self.su, self.si = np.ascontiguousarray(su, np.uint32), np.ascontiguousarray(si, np.uint32)
        self.sr = np.ascontiguousarray(sr, np.float64)
        self.I, self.J, self.K, self.N, self.Nt = num_users, num_items, K, self.tu.size, self.su.size
        self.h = self.L.sbmf_oracle_create(self.N, _p(self.tu), _p(self.ti), _p(self.tr), self.Nt, _p(self.su), _p(self.si), _p(self.sr),
                                           num_users, num_items, K, noise, stdev_mode, seed)
        if variant:
            self.L.sbmf_oracle_set_variant(self.h, variant)

    def close(self):
        if self.h:
            self.L.sbmf_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def srand(self, seed):
        self.L.sbmf_oracle_srand(seed)

    def set_log(self, path):
        return self.L.sbmf_oracle_set_log(self.h, path.encode() if path else None)

    def init_factors(self, U0=None, V0=None, init_stdev=0.1):
        u = np.ascontiguousarray(U0, np.float64) if U0 is not None else None
        v = np.ascontiguousarray(V0, np.float64) if V0 is not None else None
        self.L.sbmf_oracle_init_factors(self.h, _p(u), _p(v), init_stdev)

    def sweep(self, n=1):
        a, b = np.empty(n), np.empty(n)
        self.L.sbmf_oracle_sweep(self.h, n, _p(a), _p(b))
        return a, b

    def state(self):
        I, J, K, N = self.I, self.J, self.K, self.N
        s = {"U": np.empty((I, K)), "V": np.empty((K, J)), "b_i": np.empty(I), "b_j": np.empty(J), "mu_b_i": np.empty(I),
             "sigma_b_i": np.empty(I), "mu_b_j": np.empty(J), "sigma_b_j": np.empty(J), "sigma_u": np.empty(K), "mu_u": np.empty(K),
             "sigma_v": np.empty(K), "mu_v": np.empty(K), "E": np.empty(N)}
        L, h = self.L, self.h
        L.sbmf_oracle_get_U(h, _p(s["U"])); L.sbmf_oracle_get_V(h, _p(s["V"]))
        L.sbmf_oracle_get_bias(h, _p(s["b_i"]), _p(s["b_j"]))
        L.sbmf_oracle_get_bias_hypers(h, _p(s["mu_b_i"]), _p(s["sigma_b_i"]), _p(s["mu_b_j"]), _p(s["sigma_b_j"]))
        L.sbmf_oracle_get_dim_hypers(h, _p(s["sigma_u"]), _p(s["mu_u"]), _p(s["sigma_v"]), _p(s["mu_v"]))
        sc = np.empty(4)
        L.sbmf_oracle_get_scalars(h, _p(sc))
        s["b_0"], s["alpha"], s["mu_b_0"], s["sigma_b_0"] = sc
        L.sbmf_oracle_get_E(h, _p(s["E"]))
        return s

    def pred_mean(self):
        p = np.empty(self.Nt)
        self.L.sbmf_oracle_get_pred_mean(self.h, _p(p))
        return p

    def layout(self):
        N, I, J = self.N, self.I, self.J
        o = {"row_ptr": np.empty(I + 1, np.int64), "col": np.empty(N, np.uint32), "csr_id": np.empty(N, np.uint64),
             "col_ptr": np.empty(J + 1, np.int64), "row": np.empty(N, np.uint32), "csc_id": np.empty(N, np.uint64)}
        self.L.sbmf_oracle_get_layout(self.h, *[_p(o[k]) for k in ("row_ptr", "col", "csr_id", "col_ptr", "row", "csc_id")])
        # perm[csc slot] = csr slot: the composition that replaces [T]'s `.id` back-pointers
        inv = np.empty(N, np.uint64)
        inv[o["csr_id"]] = np.arange(N, dtype=np.uint64)
        o["perm"] = inv[o["csc_id"]]
        return o


def read_triples(path):
    L = lib()
    um, im = C.c_uint32(), C.c_uint32()
    n = L.sbmf_oracle_read_triples(path.encode(), None, None, None, C.byref(um), C.byref(im))
    if n < 0:
        raise FileNotFoundError(path)
    u, i, r = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.float64)
    L.sbmf_oracle_read_triples(path.encode(), _p(u), _p(i), _p(r), None, None)
    return u, i, r, um.value, im.value


def philox(ctr, key):
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.empty(4, np.uint32)
    lib().sbmf_oracle_philox4x32_10(_p(c), _p(k), _p(o))
    return o
