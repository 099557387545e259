"""sbmf.py -- ctypes binding of libsbmf_cuda.so (include/sbmf_cuda.h) for tests and bench.py.

This is a thin mirror of the C ABI, not a second implementation: every method is one sbmf_cuda_* call on
host (numpy) buffers.  There is no CPU path here -- if the shared library is missing or no B200 is visible,
loading / SbmfModel() raises.  The product host program is the C++ CLI (csrc/main.cpp -> bin/sbmf), which
links the same library.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SBMF_LIB_PATH: developer knob for A/B runs of differently built libraries (tools/r2_ffma2.sh); the product is lib/libsbmf_cuda.so
LIB_PATH = os.environ.get("SBMF_LIB_PATH") or os.path.join(HERE, "lib", "libsbmf_cuda.so")

SAMPLE_REF, SAMPLE_SQRT, SAMPLE_ZERO = 0, 1, 2
HYPER_REF_T, HYPER_NG_S, HYPER_NG = 0, 1, 2


class Priors(C.Structure):
    _fields_ = [("alpha", C.c_double * 6), ("beta", C.c_double * 6), ("mu", C.c_double * 6), ("sigma", C.c_double * 6),
                ("alpha_dash", C.c_double), ("beta_dash", C.c_double), ("ng_a_0", C.c_double), ("ng_b_0", C.c_double),
                ("ng_alpha_0", C.c_double), ("ng_beta_0", C.c_double), ("ng_mu_0", C.c_double), ("ng_nu_0", C.c_double)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("K", C.c_uint32), ("device", C.c_int32), ("sample_mode", C.c_int32),
                ("hyper_mode", C.c_int32), ("rebuild_every", C.c_uint32), ("burn_in", C.c_uint32), ("residual_mode", C.c_uint32),
                ("seed", C.c_uint64), ("init_stdev", C.c_double), ("clamp_lo", C.c_double), ("clamp_hi", C.c_double),
                ("priors", Priors), ("rank", C.c_int32), ("world_size", C.c_int32), ("nccl_id", C.c_uint8 * 128)]


class State(C.Structure):
    _fields_ = [("U", C.c_void_p), ("V", C.c_void_p), ("b_i", C.c_void_p), ("b_j", C.c_void_p),
                ("mu_b_i", C.c_void_p), ("sigma_b_i", C.c_void_p), ("mu_b_j", C.c_void_p), ("sigma_b_j", C.c_void_p),
                ("sigma_u", C.c_void_p), ("mu_u", C.c_void_p), ("sigma_v", C.c_void_p), ("mu_v", C.c_void_p),
                ("E", C.c_void_p),
                ("b_0", C.c_double), ("alpha", C.c_double), ("mu_b_0", C.c_double), ("sigma_b_0", C.c_double),
                ("sum_e", C.c_double), ("sum_e2", C.c_double), ("sweeps_done", C.c_uint32), ("reserved0", C.c_uint32)]


class CheckpointDims(C.Structure):
    _fields_ = [("num_users", C.c_uint32), ("num_items", C.c_uint32), ("K", C.c_uint32), ("hyper_mode", C.c_int32),
                ("n_train", C.c_uint64), ("n_test", C.c_uint64), ("sweeps_done", C.c_uint32), ("present", C.c_uint32),
                ("seed", C.c_uint64), ("sample_mode", C.c_int32), ("burn_in", C.c_uint32), ("residual_mode", C.c_uint32),
                ("rebuild_every", C.c_uint32)]


STATE_ARRAYS = ("U", "V", "b_i", "b_j", "mu_b_i", "sigma_b_i", "mu_b_j", "sigma_b_j", "sigma_u", "mu_u", "sigma_v", "mu_v", "E")
STATE_SCALARS = ("b_0", "alpha", "mu_b_0", "sigma_b_0", "sum_e", "sum_e2", "sweeps_done")


class Timing(C.Structure):
    _fields_ = [("ms_rebuild", C.c_double), ("ms_hypers", C.c_double), ("ms_user_phase", C.c_double), ("ms_exchange", C.c_double),
                ("ms_item_phase", C.c_double), ("ms_eval", C.c_double), ("ms_total", C.c_double), ("sweeps", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("nnz_light_user", C.c_uint64), ("nnz_heavy_user", C.c_uint64),
                ("nnz_light_item", C.c_uint64), ("nnz_heavy_item", C.c_uint64), ("ms_top_kernel", C.c_double),
                ("top_kernel_launches", C.c_uint64), ("top_kernel_ratings", C.c_uint64), ("ms_allgather", C.c_double)]


class SynthSpec(C.Structure):
    _fields_ = [("num_users", C.c_uint32), ("num_items", C.c_uint32), ("n_ratings", C.c_uint64), ("s_user", C.c_double),
                ("s_item", C.c_double), ("test_frac", C.c_double), ("seed", C.c_uint64), ("device", C.c_int32),
                ("reserved0", C.c_int32)]


class ProbeResult(C.Structure):
    _fields_ = [("hbm_copy_gbs", C.c_double), ("hbm_read_gbs", C.c_double), ("gather_sectors_per_s", C.c_double * 8),
                ("sm_clock_mhz_max", C.c_double), ("table_bytes", C.c_uint64), ("n_gathers", C.c_uint64), ("sm_count", C.c_int32),
                ("reserved0", C.c_int32)]


PROBE_FORMS = ("g32_lane", "g64_lane", "g64_coop2", "g128_lane", "g128_coop4")

FM_SAMPLE_LIVE, FM_SAMPLE_ZERO = 0, 1


class FmConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_attr", C.c_uint32), ("num_groups", C.c_uint32), ("K", C.c_uint32), ("k0", C.c_int32),
                ("k1", C.c_int32), ("do_sample", C.c_int32), ("do_multilevel", C.c_int32), ("sample_mode", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("init_stdev", C.c_double), ("reg0", C.c_double), ("regw", C.c_double), ("regv", C.c_double)]


class FmState(C.Structure):
    _fields_ = [("w", C.c_void_p), ("v", C.c_void_p), ("w_mu", C.c_void_p), ("w_lambda", C.c_void_p), ("v_mu", C.c_void_p),
                ("v_lambda", C.c_void_p), ("e", C.c_void_p), ("pred_sum", C.c_void_p), ("w0", C.c_double), ("alpha", C.c_double),
                ("iterations", C.c_uint32)]


class SbmfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sbmf_cuda error {code}: {msg}")
        self.code = code


_lib = None


def _bind_fm(lib):
    """argument types of the general FM Gibbs entry points (include/sbmf_fm_cuda.h)"""
    P = C.POINTER
    lib.sbmf_fm_config_default.argtypes = [P(FmConfig)]
    lib.sbmf_fm_create.argtypes = [P(FmConfig), P(C.c_void_p)]
    lib.sbmf_fm_destroy.argtypes = [C.c_void_p]
    lib.sbmf_fm_last_error.argtypes = [C.c_void_p]
    lib.sbmf_fm_last_error.restype = C.c_char_p
    lib.sbmf_fm_set_groups.argtypes = [C.c_void_p, C.c_void_p]
    lib.sbmf_fm_set_train.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    lib.sbmf_fm_set_test.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
    lib.sbmf_fm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sbmf_fm_learn.argtypes = [C.c_void_p, C.c_uint32]
    lib.sbmf_fm_rmse_history.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.sbmf_fm_predict.argtypes = [C.c_void_p, C.c_void_p]
    lib.sbmf_fm_get_state.argtypes = [C.c_void_p, P(FmState)]
    lib.sbmf_fm_get_columns.argtypes = [C.c_void_p] + [C.c_void_p] * 3
    lib.sbmf_fm_get_runs.argtypes = [C.c_void_p, P(C.c_uint32), C.c_void_p]
    lib.sbmf_fm_plan_runs.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.sbmf_fm_plan_runs.restype = C.c_uint32


_fm_lib = None


def load_fm_library():
    """The library that serves sbmf_fm_*: libsbmf_cuda.so.  SBMF_FM_LIB_PATH is a test knob: tests/test_fm_simt_emulation.py points it
    at a host build of csrc/fm.cu whose kernels execute on CPU threads (tools/emu_include) -- never set in production."""
    global _fm_lib
    if _fm_lib is None:
        alt = os.environ.get("SBMF_FM_LIB_PATH")
        if alt:
            _fm_lib = C.CDLL(alt)
            _bind_fm(_fm_lib)
        else:
            _fm_lib = load_library()
    return _fm_lib


def load_library(path=None):
    """Load libsbmf_cuda.so; raises OSError if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib = C.CDLL(path or LIB_PATH)
    # tools/build_emu.sh (test infrastructure: the kernels on host threads, for the race / memory checks) marks its output with this
    # symbol; such a library is only accepted when the test harness asks for it by name, never by accident
    if hasattr(lib, "sbmf_simt_host_emulation") and os.environ.get("SBMF_EMULATED") != "1":
        raise OSError(f"{path or LIB_PATH} is the host-emulation test build of the kernels, not the CUDA library (there is no CPU path)")
    P = C.POINTER
    lib.sbmf_cuda_abi_version.restype = C.c_int
    lib.sbmf_cuda_config_default.argtypes = [P(Config)]
    lib.sbmf_cuda_create.argtypes = [P(Config), P(C.c_void_p)]
    lib.sbmf_cuda_destroy.argtypes = [C.c_void_p]
    lib.sbmf_cuda_last_error.argtypes = [C.c_void_p]
    lib.sbmf_cuda_last_error.restype = C.c_char_p
    lib.sbmf_cuda_nccl_unique_id.argtypes = [C.c_void_p]
    lib.sbmf_cuda_set_train.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.sbmf_cuda_set_test.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_get_layout.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.sbmf_cuda_get_storage_layout.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.sbmf_cuda_get_row_positions.argtypes = [C.c_void_p] * 3
    lib.sbmf_cuda_init_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_get_state.argtypes = [C.c_void_p, P(State)]
    lib.sbmf_cuda_set_state.argtypes = [C.c_void_p, P(State)]
    lib.sbmf_cuda_get_pred_sum.argtypes = [C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_set_pred_sum.argtypes = [C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_checkpoint_write.argtypes = [C.c_char_p, P(CheckpointDims), P(State), C.c_void_p]
    lib.sbmf_cuda_checkpoint_read_dims.argtypes = [C.c_char_p, P(CheckpointDims)]
    lib.sbmf_cuda_checkpoint_read.argtypes = [C.c_char_p, P(CheckpointDims), P(State), C.c_void_p, P(C.c_int)]
    lib.sbmf_cuda_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.sbmf_cuda_get_option.argtypes = [C.c_void_p, C.c_char_p, P(C.c_int64)]
    lib.sbmf_cuda_checkpoint_last_error.restype = C.c_char_p
    lib.sbmf_cuda_sweep.argtypes = [C.c_void_p, C.c_uint32]
    lib.sbmf_cuda_eval.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double)]
    lib.sbmf_cuda_get_rmse_history.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_get_pred.argtypes = [C.c_void_p, C.c_void_p]
    lib.sbmf_cuda_get_timing.argtypes = [C.c_void_p, P(Timing)]
    lib.sbmf_cuda_reset_timing.argtypes = [C.c_void_p]
    lib.sbmf_cuda_synchronize.argtypes = [C.c_void_p]
    lib.sbmf_cuda_set_timing_enabled.argtypes = [C.c_void_p, C.c_int]
    lib.sbmf_cuda_last_sweep_call_ms.argtypes = [C.c_void_p, P(C.c_double)]
    lib.sbmf_cuda_host_alloc.argtypes = [P(C.c_void_p), C.c_size_t]
    lib.sbmf_cuda_host_free.argtypes = [C.c_void_p]
    lib.sbmf_cuda_plan_shards.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    lib.sbmf_cuda_plan_exchange.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.sbmf_cuda_plan_exchange_device.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    lib.sbmf_cuda_write_libfm_xt.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64] + [C.c_void_p] * 4
    lib.sbmf_cuda_write_libfm_xt_last_error.restype = C.c_char_p
    lib.sbmf_cuda_synth_generate.argtypes = [P(SynthSpec), P(C.c_uint64), P(C.c_uint64)] + [C.c_void_p] * 6
    lib.sbmf_cuda_synth_last_error.restype = C.c_char_p
    lib.sbmf_cuda_synth_host_generate.argtypes = [P(SynthSpec), C.c_int, P(C.c_uint64), P(C.c_uint64)] + [P(C.c_void_p)] * 6
    lib.sbmf_cuda_synth_host_free.argtypes = [C.c_void_p]
    lib.sbmf_cuda_synth_host_free.restype = None
    lib.sbmf_cuda_synth_host_last_error.restype = C.c_char_p
    _bind_fm(lib)
    if path is None:
        _lib = lib
    return lib


def default_config(**kw):
    cfg = Config()
    load_library().sbmf_cuda_config_default(C.byref(cfg))
    for k, v in kw.items():
        if k == "nccl_id":
            C.memmove(cfg.nccl_id, v, 128)
        else:
            setattr(cfg, k, v)
    return cfg


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class SbmfModel:
    """One handle of the C ABI.  Method names follow the sbmf_cuda_* entry points."""

    def __init__(self, cfg=None, options=None, **kw):
        self.lib = load_library()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.h = C.c_void_p()
        rc = self.lib.sbmf_cuda_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise SbmfError(rc, self.lib.sbmf_cuda_last_error(None).decode())
        self.K = self.cfg.K
        self.N = self.Nt = self.I = self.J = 0
        for name, value in (options or {}).items():
            self.set_option(name, value)

    def _ck(self, rc):
        if rc != 0:
            raise SbmfError(rc, self.lib.sbmf_cuda_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.sbmf_cuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_train(self, user, item, rating, num_users, num_items):
        user, item, rating = _u32(user), _u32(item), _f32(rating)
        assert user.shape == item.shape == rating.shape
        self._ck(self.lib.sbmf_cuda_set_train(self.h, user.size, _ptr(user), _ptr(item), _ptr(rating), num_users, num_items))
        self.N, self.I, self.J = user.size, num_users, num_items

    def set_test(self, user, item, rating):
        user, item, rating = _u32(user), _u32(item), _f32(rating)
        self._ck(self.lib.sbmf_cuda_set_test(self.h, user.size, _ptr(user), _ptr(item), _ptr(rating)))
        self.Nt = user.size

    def get_layout(self):
        N, I, J = self.N, self.I, self.J
        out = {"row_ptr": np.empty(I + 1, np.int64), "col": np.empty(N, np.uint32), "csr_id": np.empty(N, np.uint64),
               "col_ptr": np.empty(J + 1, np.int64), "row": np.empty(N, np.uint32), "csc_id": np.empty(N, np.uint64),
               "perm": np.empty(N, np.uint64)}
        self._ck(self.lib.sbmf_cuda_get_layout(self.h, *[_ptr(out[k]) for k in ("row_ptr", "col", "csr_id", "col_ptr", "row", "csc_id", "perm")]))
        return out

    def get_storage_layout(self):
        """The arrays the kernels run on (position space when option relabel is on) + the row positions of the caller's ids."""
        N, I, J = self.N, self.I, self.J
        out = {"row_ptr": np.empty(I + 1, np.int64), "col": np.empty(N, np.uint32), "csr_id": np.empty(N, np.uint64),
               "col_ptr": np.empty(J + 1, np.int64), "row": np.empty(N, np.uint32), "csc_id": np.empty(N, np.uint64),
               "perm": np.empty(N, np.uint64), "user_pos": np.empty(I, np.uint32), "item_pos": np.empty(J, np.uint32)}
        self._ck(self.lib.sbmf_cuda_get_storage_layout(self.h, *[_ptr(out[k]) for k in ("row_ptr", "col", "csr_id", "col_ptr", "row", "csc_id", "perm")]))
        self._ck(self.lib.sbmf_cuda_get_row_positions(self.h, _ptr(out["user_pos"]), _ptr(out["item_pos"])))
        return out

    def init_factors(self, U0=None, V0=None):
        u = _f32(U0) if U0 is not None else None
        v = _f32(V0) if V0 is not None else None
        if u is not None:
            assert u.shape == (self.I, self.K)
        if v is not None:
            assert v.shape == (self.K, self.J)
        self._ck(self.lib.sbmf_cuda_init_factors(self.h, _ptr(u), _ptr(v)))

    def sweep(self, n=1):
        self._ck(self.lib.sbmf_cuda_sweep(self.h, n))

    def eval(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.sbmf_cuda_eval(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def rmse_history(self, first, count):
        a, b = np.empty(count, np.float64), np.empty(count, np.float64)
        self._ck(self.lib.sbmf_cuda_get_rmse_history(self.h, first, count, _ptr(a), _ptr(b)))
        return a, b

    def get_pred(self):
        p = np.empty(self.Nt, np.float32)
        self._ck(self.lib.sbmf_cuda_get_pred(self.h, _ptr(p)))
        return p

    def get_state(self, with_E=True):
        I, J, K, N = self.I, self.J, self.K, self.N
        arr = {"U": np.empty((I, K), np.float32), "V": np.empty((K, J), np.float32), "b_i": np.empty(I, np.float32),
               "b_j": np.empty(J, np.float32), "mu_b_i": np.empty(I, np.float32), "sigma_b_i": np.empty(I, np.float32),
               "mu_b_j": np.empty(J, np.float32), "sigma_b_j": np.empty(J, np.float32), "sigma_u": np.empty(K, np.float64),
               "mu_u": np.empty(K, np.float64), "sigma_v": np.empty(K, np.float64), "mu_v": np.empty(K, np.float64)}
        if with_E:
            arr["E"] = np.empty(N, np.float32)
        st = State()
        for k, v in arr.items():
            setattr(st, k, v.ctypes.data)
        self._ck(self.lib.sbmf_cuda_get_state(self.h, C.byref(st)))
        for k in ("b_0", "alpha", "mu_b_0", "sigma_b_0", "sum_e", "sum_e2", "sweeps_done"):
            arr[k] = getattr(st, k)
        return arr

    def set_state(self, arr):
        """sbmf_cuda_set_state from a dict shaped like get_state()'s (missing / None arrays are passed as NULL)"""
        st, keep = _state_struct(arr, self.I, self.J, self.K, self.N)
        self._ck(self.lib.sbmf_cuda_set_state(self.h, C.byref(st)))

    def get_pred_sum(self):
        s = np.empty(self.Nt, np.float64)
        self._ck(self.lib.sbmf_cuda_get_pred_sum(self.h, _ptr(s)))
        return s

    def set_pred_sum(self, s):
        s = np.ascontiguousarray(s, np.float64)
        assert s.shape == (self.Nt,)
        self._ck(self.lib.sbmf_cuda_set_pred_sum(self.h, _ptr(s)))

    def timing(self):
        t = Timing()
        self._ck(self.lib.sbmf_cuda_get_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def reset_timing(self):
        self._ck(self.lib.sbmf_cuda_reset_timing(self.h))

    def set_option(self, name, value):
        """sbmf_cuda_set_option: tuning / developer options by name (include/sbmf_cuda.h lists them)"""
        self._ck(self.lib.sbmf_cuda_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64()
        self._ck(self.lib.sbmf_cuda_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_timing_enabled(self, on):
        """0 = off (sweeps only enqueue work), 1 = per-phase events, 2 = also per-launch events of the dominant kernel"""
        self._ck(self.lib.sbmf_cuda_set_timing_enabled(self.h, int(on)))

    def last_sweep_call_ms(self):
        ms = C.c_double()
        self._ck(self.lib.sbmf_cuda_last_sweep_call_ms(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._ck(self.lib.sbmf_cuda_synchronize(self.h))


class FmModel:
    """One handle of the general FM Gibbs sampler (include/sbmf_fm_cuda.h): libFM's fm_learn interface -- init / learn / predict
    (src/libfm/src/fm_learn.h:80, 150, 191) -- on a design matrix in row form: dict(row_ptr int64, attr uint32, x float32, y float32)."""

    def __init__(self, num_attr, K, attr_group=None, **kw):
        self.lib = load_fm_library()
        self.cfg = FmConfig()
        self.lib.sbmf_fm_config_default(C.byref(self.cfg))
        self.group = None if attr_group is None else _u32(attr_group)
        self.cfg.num_attr, self.cfg.K = int(num_attr), int(K)
        self.cfg.num_groups = 1 if self.group is None else int(self.group.max()) + 1
        for k, v in kw.items():
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = self.lib.sbmf_fm_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise SbmfError(rc, self.lib.sbmf_fm_last_error(None).decode())
        self.p, self.K, self.G = self.cfg.num_attr, self.cfg.K, self.cfg.num_groups
        self.n = self.nt = self.nnz = 0
        if self.group is not None:
            self._ck(self.lib.sbmf_fm_set_groups(self.h, _ptr(self.group)))

    def _ck(self, rc):
        if rc != 0:
            raise SbmfError(rc, self.lib.sbmf_fm_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.sbmf_fm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _mat(m):
        return (np.ascontiguousarray(m["row_ptr"], dtype=np.int64), _u32(m["attr"]), _f32(m["x"]), _f32(m["y"]))

    def set_train(self, m):
        rp, at, x, y = self._mat(m)
        self._ck(self.lib.sbmf_fm_set_train(self.h, y.size, _ptr(rp), _ptr(at), _ptr(x), _ptr(y)))
        self.n, self.nnz = int(y.size), int(rp[-1])

    def set_test(self, m):
        rp, at, x, y = self._mat(m)
        self._ck(self.lib.sbmf_fm_set_test(self.h, y.size, _ptr(rp), _ptr(at), _ptr(x), _ptr(y)))
        self.nt = int(y.size)

    def init(self, w=None, v=None):
        w = None if w is None else _f32(w)
        v = None if v is None else _f32(v)
        assert w is None or w.shape == (self.p,)
        assert v is None or v.shape == (self.K, self.p)
        self._ck(self.lib.sbmf_fm_init(self.h, _ptr(w), _ptr(v)))

    def learn(self, iters=1):
        self._ck(self.lib.sbmf_fm_learn(self.h, iters))

    def rmse_history(self, first, count):
        a, b = np.zeros(count), np.zeros(count)
        self._ck(self.lib.sbmf_fm_rmse_history(self.h, first, count, _ptr(a), _ptr(b)))
        return a, b

    def predict(self):
        out = np.zeros(self.nt, dtype=np.float32)
        self._ck(self.lib.sbmf_fm_predict(self.h, _ptr(out)))
        return out

    def get_state(self):
        arr = {"w": np.zeros(self.p, dtype=np.float32), "v": np.zeros((self.K, self.p), dtype=np.float32), "w_mu": np.zeros(self.G),
               "w_lambda": np.zeros(self.G), "v_mu": np.zeros((self.G, self.K)), "v_lambda": np.zeros((self.G, self.K)),
               "e": np.zeros(self.n, dtype=np.float32), "pred_sum": np.zeros(max(self.nt, 1))}
        st = FmState()
        for k, a in arr.items():
            setattr(st, k, a.ctypes.data)
        self._ck(self.lib.sbmf_fm_get_state(self.h, C.byref(st)))
        arr["pred_sum"] = arr["pred_sum"][:self.nt]
        arr.update(w0=st.w0, alpha=st.alpha, iterations=st.iterations)
        return arr

    def get_columns(self):
        cp, ci, x = np.zeros(self.p + 1, dtype=np.int64), np.zeros(max(self.nnz, 1), dtype=np.uint32), np.zeros(max(self.nnz, 1), dtype=np.float32)
        self._ck(self.lib.sbmf_fm_get_columns(self.h, _ptr(cp), _ptr(ci), _ptr(x)))
        return {"col_ptr": cp, "case": ci[:self.nnz], "x": x[:self.nnz]}

    def get_runs(self):
        n, rb = C.c_uint32(0), np.zeros(self.p + 1, dtype=np.uint32)
        self._ck(self.lib.sbmf_fm_get_runs(self.h, C.byref(n), _ptr(rb)))
        return rb[:n.value + 1]


def fm_plan_runs(next_attr):
    next_attr = _u32(next_attr)
    out = np.zeros(next_attr.size + 1, dtype=np.uint32)
    n = load_fm_library().sbmf_fm_plan_runs(next_attr.size, _ptr(next_attr), _ptr(out))
    return out[:n + 1]


def _state_shapes(I, J, K, N):
    return {"U": ((I, K), np.float32), "V": ((K, J), np.float32), "b_i": ((I,), np.float32), "b_j": ((J,), np.float32),
            "mu_b_i": ((I,), np.float32), "sigma_b_i": ((I,), np.float32), "mu_b_j": ((J,), np.float32),
            "sigma_b_j": ((J,), np.float32), "sigma_u": ((K,), np.float64), "mu_u": ((K,), np.float64),
            "sigma_v": ((K,), np.float64), "mu_v": ((K,), np.float64), "E": ((N,), np.float32)}


def _state_struct(arr, I, J, K, N):
    """sbmf_state over the arrays of a get_state()-shaped dict; returns (struct, list keeping the buffers alive)"""
    st, keep = State(), []
    for k, (shape, dt) in _state_shapes(I, J, K, N).items():
        v = arr.get(k)
        if v is None:
            continue
        v = np.ascontiguousarray(v, dt)
        assert v.shape == shape, (k, v.shape, shape)
        keep.append(v)
        setattr(st, k, v.ctypes.data)
    for k in STATE_SCALARS:
        if k in arr:
            setattr(st, k, arr[k])
    return st, keep


def checkpoint_write(path, arr, num_users, num_items, K, n_train, n_test, hyper_mode=0, pred_sum=None, seed=1, sample_mode=0, burn_in=0,
                     residual_mode=0, rebuild_every=1):
    """sbmf_cuda_checkpoint_write: a get_state()-shaped dict (+ the running prediction sums) to a file; host only.  seed and the mode
    flags are those of the chain that produced the state (a resume continues it only under the same values)."""
    lib = load_library()
    st, keep = _state_struct(arr, num_users, num_items, K, n_train)
    ps = None if pred_sum is None else np.ascontiguousarray(pred_sum, np.float64)
    assert ps is None or ps.shape == (n_test,)
    dims = CheckpointDims(num_users, num_items, K, hyper_mode, n_train, n_test, 0, 0, seed, sample_mode, burn_in, residual_mode, rebuild_every)
    rc = lib.sbmf_cuda_checkpoint_write(os.fsencode(path), C.byref(dims), C.byref(st), _ptr(ps))
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_checkpoint_last_error().decode())


def checkpoint_read(path):
    """sbmf_cuda_checkpoint_read_dims + _read: returns (dims dict, get_state()-shaped dict, pred_sum or None); host only"""
    lib = load_library()
    dims = CheckpointDims()
    rc = lib.sbmf_cuda_checkpoint_read_dims(os.fsencode(path), C.byref(dims))
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_checkpoint_last_error().decode())
    I, J, K, N, Nt = dims.num_users, dims.num_items, dims.K, dims.n_train, dims.n_test
    arr = {k: np.empty(shape, dt) for k, (shape, dt) in _state_shapes(I, J, K, N).items() if (dims.present >> STATE_ARRAYS.index(k)) & 1}
    st = State()
    for k, v in arr.items():
        setattr(st, k, v.ctypes.data)
    ps = np.empty(Nt, np.float64) if (dims.present >> 13) & 1 else None
    have = C.c_int(0)
    rc = lib.sbmf_cuda_checkpoint_read(os.fsencode(path), C.byref(dims), C.byref(st), _ptr(ps), C.byref(have))
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_checkpoint_last_error().decode())
    for k in STATE_SCALARS:
        arr[k] = getattr(st, k)
    d = {k: getattr(dims, k) for k, _ in CheckpointDims._fields_}
    return d, arr, (ps if have.value else None)


def probe(device=0, table_bytes=0, n_gathers=0):
    """sbmf_cuda_probe: HBM streaming bandwidth and the L2 -> SM gather rates of this device, measured now (csrc/probe.cu)."""
    lib = load_library()
    lib.sbmf_cuda_probe.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(ProbeResult)]
    lib.sbmf_cuda_probe_last_error.restype = C.c_char_p
    r = ProbeResult()
    rc = lib.sbmf_cuda_probe(int(device), int(table_bytes), int(n_gathers), C.byref(r))
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_probe_last_error().decode())
    return {"hbm_copy_gbs": r.hbm_copy_gbs, "hbm_read_gbs": r.hbm_read_gbs, "sm_count": r.sm_count, "sm_clock_mhz_max": r.sm_clock_mhz_max,
            "table_bytes": r.table_bytes, "n_gathers": r.n_gathers,
            "gather_sectors_per_s": {name: r.gather_sectors_per_s[i] for i, name in enumerate(PROBE_FORMS)}}


def nccl_unique_id():
    """128-byte ncclUniqueId (rank 0 creates it, the host program ships it to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    lib = load_library()
    if lib.sbmf_cuda_nccl_unique_id(buf) != 0:
        raise SbmfError(-5, lib.sbmf_cuda_last_error(None).decode())
    return bytes(buf)


def plan_shards(ptr, world):
    ptr = np.ascontiguousarray(ptr, np.int64)
    b = np.empty(world + 1, np.uint32)
    rc = load_library().sbmf_cuda_plan_shards(_ptr(ptr), ptr.size - 1, world, _ptr(b))
    assert rc == 0
    return b


def plan_exchange(perm, world, rank, csr_bounds, csc_bounds):
    perm = np.ascontiguousarray(perm, np.uint32)
    cb, tb = np.ascontiguousarray(csr_bounds, np.int64), np.ascontiguousarray(csc_bounds, np.int64)
    send_idx = np.empty(int(cb[rank + 1] - cb[rank]), np.uint32)
    recv_pos = np.empty(int(tb[rank + 1] - tb[rank]), np.uint32)
    sc, rc_ = np.empty(world, np.int64), np.empty(world, np.int64)
    rc = load_library().sbmf_cuda_plan_exchange(perm.size, _ptr(perm), world, rank, _ptr(cb), _ptr(tb), _ptr(send_idx), _ptr(sc), _ptr(recv_pos), _ptr(rc_))
    assert rc == 0, rc
    return send_idx, sc, recv_pos, rc_


def plan_exchange_device(perm, world, rank, csr_bounds, csc_bounds, device=0):
    """the device-side planner on host arrays (needs a GPU): returns (send_idx, recv_pos, pair_counts[world, world])"""
    perm = np.ascontiguousarray(perm, np.uint32)
    cb, tb = np.ascontiguousarray(csr_bounds, np.int64), np.ascontiguousarray(csc_bounds, np.int64)
    send_idx = np.empty(int(cb[rank + 1] - cb[rank]), np.uint32)
    recv_pos = np.empty(int(tb[rank + 1] - tb[rank]), np.uint32)
    pc = np.empty((world, world), np.int64)
    rc = load_library().sbmf_cuda_plan_exchange_device(perm.size, _ptr(perm), world, rank, _ptr(cb), _ptr(tb), _ptr(send_idx), _ptr(recv_pos), _ptr(pc), device)
    if rc != 0:
        raise SbmfError(rc, "sbmf_cuda_plan_exchange_device failed")
    return send_idx, recv_pos, pc


def write_libfm_xt(path, layout, num_features, num_users, num_items, item_offset):
    """sbmf_cuda_write_libfm_xt from a get_layout()-shaped dict (row_ptr, csr_id, col_ptr, csc_id); host only"""
    lib = load_library()
    rp, cp = np.ascontiguousarray(layout["row_ptr"], np.int64), np.ascontiguousarray(layout["col_ptr"], np.int64)
    rid, cid = np.ascontiguousarray(layout["csr_id"], np.uint64), np.ascontiguousarray(layout["csc_id"], np.uint64)
    rc = lib.sbmf_cuda_write_libfm_xt(os.fsencode(path), num_features, num_users, num_items, item_offset, rid.size, _ptr(rp), _ptr(rid), _ptr(cp), _ptr(cid))
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_write_libfm_xt_last_error().decode())


def pinned_empty(n, dtype):
    """numpy array over cudaMallocHost memory (sbmf_cuda_host_alloc); keeps itself alive via a finalizer."""
    import weakref
    lib = load_library()
    dt = np.dtype(dtype)
    p = C.c_void_p()
    if lib.sbmf_cuda_host_alloc(C.byref(p), int(n) * dt.itemsize) != 0:
        raise MemoryError("sbmf_cuda_host_alloc failed")
    buf = (C.c_char * (int(n) * dt.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    weakref.finalize(buf, lib.sbmf_cuda_host_free, p)
    return arr


def synth_generate(num_users, num_items, n_ratings, s_user=0.8, s_item=1.0, test_frac=0.1, seed=20151001, device=0):
    """Device-side synthetic rating matrix (csrc/synth.cu).  Returns dict of host arrays."""
    lib = load_library()
    spec = SynthSpec(num_users, num_items, n_ratings, s_user, s_item, test_frac, seed, device, 0)
    ntr, nte = C.c_uint64(), C.c_uint64()
    rc = lib.sbmf_cuda_synth_generate(C.byref(spec), C.byref(ntr), C.byref(nte), None, None, None, None, None, None)
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_synth_last_error().decode())
    out = {"train_user": np.empty(ntr.value, np.uint32), "train_item": np.empty(ntr.value, np.uint32),
           "train_rating": np.empty(ntr.value, np.float32), "test_user": np.empty(nte.value, np.uint32),
           "test_item": np.empty(nte.value, np.uint32), "test_rating": np.empty(nte.value, np.float32)}
    rc = lib.sbmf_cuda_synth_generate(C.byref(spec), C.byref(ntr), C.byref(nte), *[_ptr(out[k]) for k in
                                      ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating")])
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_synth_last_error().decode())
    out["num_users"], out["num_items"] = num_users, num_items
    return out


def synth_generate_host(num_users, num_items, n_ratings, s_user=0.8, s_item=1.0, test_frac=0.1, seed=20151001, threads=0):
    """Sparse host-side sampler of the same matrix family (csrc/synth_host.cpp), for shapes whose pair grid is too large
    for synth_generate (10M x 1M).  Returns dict of numpy arrays over the library's buffers (freed with the arrays)."""
    import weakref
    lib = load_library()
    spec = SynthSpec(num_users, num_items, n_ratings, s_user, s_item, test_frac, seed, 0, 0)
    ntr, nte = C.c_uint64(), C.c_uint64()
    ptrs = [C.c_void_p() for _ in range(6)]
    rc = lib.sbmf_cuda_synth_host_generate(C.byref(spec), int(threads), C.byref(ntr), C.byref(nte), *[C.byref(p) for p in ptrs])
    if rc != 0:
        raise SbmfError(rc, lib.sbmf_cuda_synth_host_last_error().decode())
    out = {}
    names = ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating")
    for k, p in zip(names, ptrs):
        n = ntr.value if k.startswith("train") else nte.value
        buf = (C.c_char * (n * 4)).from_address(p.value)
        out[k] = np.frombuffer(buf, dtype=np.float32 if k.endswith("rating") else np.uint32, count=n)
        weakref.finalize(buf, lib.sbmf_cuda_synth_host_free, p)
    out["num_users"], out["num_items"] = num_users, num_items
    return out


def read_triples(path):
    """`user SEP item SEP rating` triples, 0-based ids ([T]:35-73 semantics for well-formed files)."""
    a = np.loadtxt(path, dtype=np.float64, ndmin=2)
    return a[:, 0].astype(np.uint32), a[:, 1].astype(np.uint32), a[:, 2].astype(np.float32)
