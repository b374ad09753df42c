"""ctypes binding of libcorrla_b200.so (include/corrla_b200.h).  No compute happens in Python and
there is no CPU fallback: if the shared library is missing, loading fails loudly; if no CUDA device
is usable the C ABI returns CORRLA_ERR_NO_DEVICE and `check()` raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libcorrla_b200.so"

c_double_p = C.POINTER(C.c_double)


class RsvdOpts(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("omega", C.c_void_p),
        ("omega_rs", C.c_int64),
        ("omega_cs", C.c_int64),
        ("omega_on_device", C.c_int),
        ("schedule", C.c_int),
        ("a_on_device", C.c_int),
        ("out_on_device", C.c_int),
        ("device", C.c_int),
        ("stream", C.c_void_p),
        ("ctx", C.c_void_p),
        ("comm", C.c_void_p),
        ("global_rows", C.c_int64),
        ("center", C.c_int),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("device_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("gpu_launches", C.c_int),
        ("passes_over_a", C.c_int),
        ("qr_third_passes", C.c_int),
        ("qr_refills", C.c_int),
        ("jacobi_sweeps", C.c_int),
        ("live_columns", C.c_int),
        ("pass_launches", C.c_int),
        ("pass_ms", C.c_double),
        ("pass_flops", C.c_double),
        ("p2p_exchanges", C.c_int),
        ("streamed_chunks", C.c_int),
        ("jacobi_converged", C.c_int),
        ("fused_small", C.c_int),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/corrla_b200.h declares
SYMBOLS = {
    "corrla_rsvd_opts_default": (None, [C.POINTER(RsvdOpts)]),
    "corrla_rsvd_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t, C.c_size_t,
                                  C.c_size_t, C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(Timings)]),
    "corrla_power_iter_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t,
                                        C.c_size_t, C.POINTER(RsvdOpts), C.c_void_p, C.POINTER(Timings)]),
    "corrla_rsvd_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t, C.c_size_t,
                                  C.c_size_t, C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(Timings)]),
    "corrla_power_iter_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t,
                                        C.c_size_t, C.POINTER(RsvdOpts), C.c_void_p, C.POINTER(Timings)]),
    "corrla_rpca_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t,
                                  C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Timings)]),
    "corrla_par_matmul_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_double, C.c_int, C.POINTER(RsvdOpts)]),
    "corrla_random_mat_normal_f64": (C.c_int, [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                               C.POINTER(RsvdOpts)]),
    "corrla_dmdc_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int64, C.c_size_t, C.c_size_t, C.POINTER(RsvdOpts), C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Timings)]),
    "corrla_pod_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_size_t,
                                 C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Timings)]),
    "corrla_cov_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double,
                                 C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "corrla_active_ss_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                       C.c_int, C.c_int, C.POINTER(RsvdOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int)]),
    "corrla_poly_grad_at_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                          C.POINTER(RsvdOpts), C.c_void_p, C.POINTER(C.c_int)]),
    "corrla_thin_q_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                    C.POINTER(RsvdOpts), C.c_void_p, C.POINTER(C.c_int)]),
    "corrla_host_alloc": (C.c_void_p, [C.c_size_t]),
    "corrla_host_free": (None, [C.c_void_p, C.c_size_t]),
    "corrla_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "corrla_ctx_destroy": (None, [C.c_void_p]),
    "corrla_ctx_trim": (C.c_size_t, [C.c_void_p]),
    "corrla_comm_unique_id": (C.c_int, [C.c_char_p]),
    "corrla_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "corrla_comm_destroy": (None, [C.c_void_p]),
    "corrla_comm_rank": (C.c_int, [C.c_void_p]),
    "corrla_comm_size": (C.c_int, [C.c_void_p]),
    "corrla_status_str": (C.c_char_p, [C.c_int]),
    "corrla_last_error": (C.c_char_p, []),
    "corrla_version": (C.c_char_p, []),
}

CORRLA_OK = 0
CORRLA_ERR_INVALID = -1
CORRLA_ERR_RANK = -2
CORRLA_ERR_CUDA = -3
CORRLA_ERR_UNSUPPORTED = -4
CORRLA_ERR_ALLOC = -5
CORRLA_ERR_COMM = -6
CORRLA_ERR_NO_DEVICE = -7


class CorrlaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"corrla_b200 status {status}: {message}")
        self.status = status


class RankPanic(IndexError, CorrlaError):
    """n_rank > min(n_rank + n_oversamples, ncols): the reference panics with an out-of-range `get`
    (random_svd.rs:98-107), surfaced by pyo3 as PanicException; here it is an IndexError."""

    def __init__(self, status: int, message: str):
        CorrlaError.__init__(self, status, message)


_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("CORRLA_B200_LIB", _LIB_PATH))


def load():
    """dlopen the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make`. "
            "corrla_rs_b200 has no CPU fallback.")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL if hasattr(C, "RTLD_GLOBAL") else 0)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int):
    if status == CORRLA_OK:
        return
    lib = load()
    detail = lib.corrla_last_error().decode(errors="replace")
    text = lib.corrla_status_str(status).decode()
    msg = f"{text}: {detail}" if detail else text
    if status == CORRLA_ERR_RANK:
        raise RankPanic(status, msg)
    raise CorrlaError(status, msg)
