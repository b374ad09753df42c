"""corrla_rs_b200 -- host-side mirror of CORRLA-RS's RSVD interface over libcorrla_b200.so.

The functions here have the names, argument meaning and error behaviour of the reference's pyo3 module
(`corrla_rs.rsvd`, src/lib_math_utils_py.rs:21-36) and of the Rust functions underneath it
(`random_svd`, `power_iter`: src/lib_math_utils/random_svd.rs:15-110; `par_matmul_helper`,
`random_mat_normal`: src/lib_math_utils/mat_utils.rs:20-33, :161-175).  All arithmetic runs in the
CUDA library; this module only marshals pointers, strides and sizes.  There is no CPU fallback.

Inputs may be numpy arrays (host path: the library copies A to the GPU and the results back) or
torch CUDA tensors (device path: nothing crosses PCIe).  torch is imported lazily and only for the
device path and for the multi-GPU communicator bootstrap.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _ffi
from ._ffi import CorrlaError, RankPanic, Timings  # noqa: F401  (re-exported)

__all__ = ["rsvd", "random_svd", "rsvd_f32", "rpca", "power_iter", "par_matmul", "par_matmul_helper", "random_mat_normal", "thin_q",
           "Context", "ShardComm", "CorrlaError", "RankPanic", "version", "last_timings", "release_buffers",
           "DMDc", "PodI", "RbfInterp", "dmdc_operators", "pod_modes_weights",
           "cov", "mat_cov_centered", "pearson_corr", "ActiveSsRsvd", "FittedActiveSsRsvd", "PolyGradientEstimator", "active_ss", "active_ss_fit"]

_SCHEDULES = {"reference": 0, "stabilised": 1, "stabilized": 1, 0: 0, 1: 1}
_tls = threading.local()


def version() -> str:
    return _ffi.load().corrla_version().decode()


def last_timings() -> dict | None:
    """corrla_timings of the most recent call made by this thread (as a dict)."""
    return getattr(_tls, "timings", None)


# --------------------------------------------------------------------------------------------
# contexts (persistent device buffers) and communicators
# --------------------------------------------------------------------------------------------
class Context:
    """Per-GPU handle that keeps the engine's device buffers alive between calls."""

    def __init__(self, device: int | None = None):
        lib = _ffi.load()
        h = C.c_void_p()
        _ffi.check(lib.corrla_ctx_create(-1 if device is None else int(device), C.byref(h)))
        self._h = h
        self.device = device

    @property
    def handle(self):
        return self._h

    def release_buffers(self) -> int:
        """Give the cached device buffers back to the driver (corrla_ctx_trim); returns the bytes released.  After a
        host-path call on a large matrix the device copy of A, Y and the outputs otherwise stay allocated for the life
        of the context, which starves e.g. torch's caching allocator in the same process."""
        if getattr(self, "_h", None):
            return int(_ffi.load().corrla_ctx_trim(self._h))
        return 0

    def close(self):
        if getattr(self, "_h", None):
            _ffi.load().corrla_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}
_ctx_lock = threading.Lock()


def release_buffers() -> int:
    """Trim every hidden per-device default context (see Context.release_buffers)."""
    with _ctx_lock:
        return sum(c.release_buffers() for c in _default_ctx.values())


def _context_for(device: int | None) -> Context:
    key = -1 if device is None else int(device)
    with _ctx_lock:
        ctx = _default_ctx.get(key)
        if ctx is None or ctx.handle is None:
            ctx = Context(device)
            _default_ctx[key] = ctx
        return ctx


class ShardComm:
    """NCCL communicator over the ranks that each hold a block of rows of one tall matrix.
    Bootstrapped through torch.distributed (any backend): rank 0 draws the ncclUniqueId and broadcasts it."""

    def __init__(self, device: int | None = None, group=None):
        import torch
        import torch.distributed as dist
        lib = _ffi.load()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before ShardComm()")
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _ffi.check(lib.corrla_comm_unique_id(buf))
        # gloo and nccl both broadcast CPU/GPU tensors respectively; use an object broadcast to stay backend-neutral
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        h = C.c_void_p()
        _ffi.check(lib.corrla_comm_init(box[0], self.rank, self.size, int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            _ffi.load().corrla_comm_destroy(self._h)
            self._h = None


# --------------------------------------------------------------------------------------------
# marshalling helpers
# --------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


class _Mat:
    """(pointer, shape, element strides, residency) of a 2-D f64 array, numpy or torch."""

    def __init__(self, x, name="a_mat"):
        if not _is_torch(x) and not isinstance(x, np.ndarray) and hasattr(x, "__cuda_array_interface__"):
            import torch                      # cupy / numba device arrays: zero-copy view
            x = torch.as_tensor(x, device="cuda")
        if _is_torch(x):
            import torch
            if x.dtype != torch.float64 or x.dim() != 2:
                raise TypeError(f"{name} must be a 2-D float64 array")   # pyo3 extraction error
            self.on_device = x.is_cuda
            if not self.on_device:
                x = x.numpy()
            else:
                self.keep = x
                self.ptr = x.data_ptr()
                self.shape = tuple(x.shape)
                self.strides = tuple(x.stride())
                self.device = x.device.index
                return
        if not isinstance(x, np.ndarray):
            raise TypeError(f"{name} must be a numpy.ndarray or a torch tensor of dtype float64")
        if x.dtype != np.float64 or x.ndim != 2:
            raise TypeError(f"{name} must be a 2-D float64 array")       # pyo3: PyReadonlyArray2<f64>
        self.on_device = False
        self.keep = x
        self.ptr = x.ctypes.data
        self.shape = x.shape
        self.strides = tuple(s // 8 for s in x.strides)
        if any(s * 8 != b for s, b in zip(self.strides, x.strides)) or any(s < 0 for s in self.strides):
            # exotic (negative / unaligned) strides: take a contiguous copy, as PyReadonlyArray would refuse them
            x = np.ascontiguousarray(x)
            self.keep = x
            self.ptr = x.ctypes.data
            self.strides = tuple(s // 8 for s in x.strides)
        self.device = None


def _make_opts(*, ctx, on_device, out_on_device, omega, seed, schedule, comm, global_rows, stream, device,
               center=False):
    lib = _ffi.load()
    o = _ffi.RsvdOpts()
    lib.corrla_rsvd_opts_default(C.byref(o))
    if schedule not in _SCHEDULES:
        raise ValueError(f"schedule must be 'reference' or 'stabilised', got {schedule!r}")
    o.schedule = _SCHEDULES[schedule]
    o.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2**64 - 1)
    o.a_on_device = 1 if on_device else 0
    o.out_on_device = 1 if out_on_device else 0
    o.device = -1 if device is None else int(device)
    o.ctx = ctx.handle
    keep = None
    if omega is not None:
        om = _Mat(omega, "omega")
        o.omega = om.ptr
        o.omega_rs, o.omega_cs = om.strides
        o.omega_on_device = 1 if om.on_device else 0
        keep = om
    if comm is not None:
        o.comm = comm.handle
        o.global_rows = int(global_rows or 0)
    if stream is not None:
        o.stream = int(stream)
    o.center = 1 if center else 0
    return o, keep


def _current_stream(device):
    """Handle of torch's current stream on `device`.  The legacy default stream has handle 0, which the C ABI reads as
    "no stream given" (it would then use the context's own non-blocking stream, unordered with the producer of the
    tensor): it is passed as cudaStreamLegacy (0x1) instead."""
    import torch
    h = torch.cuda.current_stream(device).cuda_stream
    return h if h != 0 else 1


def _colmajor_empty_like(a: _Mat, rows: int, cols: int):
    """Column-major rows x cols output living where `a` lives."""
    if a.on_device:
        import torch
        return torch.empty((cols, rows), dtype=torch.float64, device=a.keep.device).t()
    n = rows * cols
    if n * 8 >= _HUGE_OUTPUT_BYTES:
        big = _huge_empty(n)
        if big is not None:
            return big.reshape((rows, cols), order="F")
    return np.empty((rows, cols), dtype=np.float64, order="F")


_HUGE_OUTPUT_BYTES = 64 << 20


def _huge_empty(n: int):
    """n float64 in a buffer from corrla_host_alloc (2 MiB pages when the kernel allows): large outputs such as U are
    filled by a threaded device->host copy whose speed is otherwise set by 4 KiB first-touch page faults."""
    import weakref
    lib = _ffi.load()
    nbytes = n * 8
    p = lib.corrla_host_alloc(nbytes)
    if not p:
        return None
    buf = (C.c_double * n).from_address(p)
    weakref.finalize(buf, lib.corrla_host_free, p, nbytes)
    return np.frombuffer(buf, dtype=np.float64)


def _ptr(x):
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def _check_out(x, rows: int, cols: int, on_device: bool, name: str):
    """A caller-provided output: float64, rows x cols, column-major and contiguous (the layout the C ABI writes),
    living where the input lives.  Reusing outputs across calls keeps their pages faulted in (or pinned), which is
    what makes the device->host copy of a multi-gigabyte U run at DMA speed."""
    if _is_torch(x):
        import torch
        shape, strides = tuple(x.shape), tuple(x.stride())
        ok = x.dtype == torch.float64 and x.is_cuda == on_device
    elif isinstance(x, np.ndarray):
        shape, strides = x.shape, tuple(st // 8 for st in x.strides)
        ok = x.dtype == np.float64 and not on_device and x.flags.writeable
    else:
        raise TypeError(f"out[{name}] must be a numpy array or a torch tensor")
    if len(shape) == 1 and cols == 1:
        shape, strides = (shape[0], 1), (strides[0], shape[0])
    colmajor = len(shape) == 2 and (strides[0] == 1 or shape[0] == 1) and (strides[1] == shape[0] or shape[1] == 1)
    if not ok or tuple(shape) != (rows, cols) or not colmajor:
        raise ValueError(f"out[{name}] must be a column-major float64 array of shape ({rows}, {cols}) "
                         f"{'on the device' if on_device else 'on the host'}")
    return x


# --------------------------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------------------------
def rsvd(a_mat, n_rank: int, n_iters: int, n_oversamples: int, *, omega=None, seed: int | None = None,
         schedule="reference", ctx: Context | None = None, comm: ShardComm | None = None,
         global_rows: int | None = None, center: bool = False, out=None):
    """Randomized SVD, drop-in for `corrla_rs.rsvd(a_mat, n_rank, n_iters, n_oversamples)`
    (src/lib_math_utils_py.rs:21-36 -> random_svd, src/lib_math_utils/random_svd.rs:63-110).

    Returns (ur, sr, vr) with shapes (nrows, n_rank), (n_rank, 1), (n_rank, ncols), like the reference.
    Keyword extras (not in the reference): `omega` injects the Gaussian test matrix (ncols_thin x l),
    `seed` makes the on-device Philox generator reproducible (the reference is unseeded),
    `schedule="stabilised"` re-orthonormalises in every power iteration, `comm` runs row-sharded over
    several GPUs (a_mat is then this rank's rows of the thin matrix), `center=True` decomposes a_mat minus its
    column means (center_mat_col, mat_utils.rs:482-502) without forming the centred copy when a_mat is tall,
    `out=(u, s, vt)` writes into caller-owned column-major arrays (reused across calls: no fresh multi-gigabyte
    allocation and no page faults in the device->host copy; pinned host arrays are copied by direct DMA)."""
    for name, v in (("n_rank", n_rank), ("n_iters", n_iters), ("n_oversamples", n_oversamples)):
        if not isinstance(v, (int, np.integer)) or isinstance(v, bool):
            raise TypeError(f"{name} must be an int")
        if v < 0:
            raise OverflowError(f"can't convert negative int to unsigned ({name})")   # pyo3 usize extraction
    a = _Mat(a_mat)
    lib = _ffi.load()
    nrows, ncols = a.shape
    device = a.device if a.on_device else (comm.device if comm is not None else None)
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if a.on_device else None
    o, keep = _make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=omega, seed=seed,
                         schedule=schedule, comm=comm, global_rows=global_rows, stream=stream, device=device,
                         center=center)
    k = int(n_rank)
    if out is not None:
        u, s, vt = out
        _check_out(u, nrows, max(k, 1), a.on_device, "u")
        _check_out(s, max(k, 1), 1, a.on_device, "s")
        _check_out(vt, max(k, 1), ncols, a.on_device, "vt")
    else:
        u = _colmajor_empty_like(a, nrows, max(k, 1))
        vt = _colmajor_empty_like(a, max(k, 1), ncols)
        s = _colmajor_empty_like(a, max(k, 1), 1)
    t = Timings()
    st = lib.corrla_rsvd_f64(a.ptr, nrows, ncols, a.strides[0], a.strides[1], k, int(n_iters), int(n_oversamples),
                             C.byref(o), _ptr(u), _ptr(s), _ptr(vt), C.byref(t))
    del keep
    _ffi.check(st)
    _tls.timings = t.as_dict()
    return u, s, vt


random_svd = rsvd


def rsvd_f32(a_mat, n_rank: int, n_iters: int, n_oversamples: int, *, omega=None, seed: int | None = None,
             schedule="reference", ctx: Context | None = None):
    """`random_svd::<f32>` (random_svd.rs:63-66 is generic over T: RealField + Float): a_mat is a 2-D float32 numpy array
    or torch CUDA tensor, the factors come back as float32 with the reference's shapes.  (The pyo3 `rsvd` itself takes
    float64 only; this is the instantiation a Rust caller reaches with f32 matrices.)  The arithmetic runs on the FP64
    tensor pipe: the data are widened once on the device and the factors rounded on the way out."""
    lib = _ffi.load()
    for name, v in (("n_rank", n_rank), ("n_iters", n_iters), ("n_oversamples", n_oversamples)):
        if not isinstance(v, (int, np.integer)) or isinstance(v, bool):
            raise TypeError(f"{name} must be an int")
        if v < 0:
            raise OverflowError(f"can't convert negative int to unsigned ({name})")
    k = max(int(n_rank), 1)
    if _is_torch(a_mat):
        import torch
        if a_mat.dtype != torch.float32 or a_mat.dim() != 2 or not a_mat.is_cuda:
            raise TypeError("a_mat must be a 2-D float32 CUDA tensor or numpy array")
        on_device, ptr, shape, strides, device = True, a_mat.data_ptr(), tuple(a_mat.shape), tuple(a_mat.stride()), a_mat.device.index
        mk = lambda r, c: torch.empty((c, r), dtype=torch.float32, device=a_mat.device).t()
    else:
        if not isinstance(a_mat, np.ndarray) or a_mat.dtype != np.float32 or a_mat.ndim != 2:
            raise TypeError("a_mat must be a 2-D float32 CUDA tensor or numpy array")
        if any(st % 4 or st < 0 for st in a_mat.strides):
            a_mat = np.ascontiguousarray(a_mat)
        on_device, ptr, shape, strides, device = False, a_mat.ctypes.data, a_mat.shape, tuple(st // 4 for st in a_mat.strides), None
        mk = lambda r, c: np.empty((r, c), dtype=np.float32, order="F")
    nrows, ncols = shape
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if on_device else None
    o, keep = _make_opts(ctx=ctx, on_device=on_device, out_on_device=on_device, omega=omega, seed=seed, schedule=schedule,
                         comm=None, global_rows=None, stream=stream, device=device)
    u, s, vt = mk(nrows, k), mk(k, 1), mk(k, ncols)
    t = Timings()
    st = lib.corrla_rsvd_f32(ptr, nrows, ncols, strides[0], strides[1], int(n_rank), int(n_iters), int(n_oversamples),
                             C.byref(o), _ptr(u), _ptr(s), _ptr(vt), C.byref(t))
    del keep
    _ffi.check(st)
    _tls.timings = t.as_dict()
    return u, s, vt


def rpca(a_mat, n_rank: int, n_iters: int = 0, n_oversamples: int = 0, *, omega=None, seed: int | None = None,
         ctx: Context | None = None, comm: ShardComm | None = None, global_rows: int | None = None,
         return_means: bool = False):
    """PCA by RSVD, drop-in for `corrla_rs.rpca(a_mat, n_rank, n_iters, n_oversamples)` (lib_math_utils_py.rs:38-55 ->
    PcaRsvd::new, pca_rsvd.rs:56-82).  Returns (singular_vals (n_rank, 1), components (n_rank, n_dim)).
    Like the reference, `n_iters` and `n_oversamples` are accepted and IGNORED: PcaRsvd hard-codes 20 power
    iterations and min(n_dim, 10) oversamples (pca_rsvd.rs:65-66).  The centred copy of a tall a_mat is never formed."""
    a = _Mat(a_mat)
    lib = _ffi.load()
    nrows, ncols = a.shape
    device = a.device if a.on_device else (comm.device if comm is not None else None)
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if a.on_device else None
    o, keep = _make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=omega, seed=seed,
                         schedule="reference", comm=comm, global_rows=global_rows, stream=stream, device=device,
                         center=True)
    k = int(n_rank)
    s = _colmajor_empty_like(a, max(k, 1), 1)
    comps = _colmajor_empty_like(a, max(k, 1), ncols)
    means = _colmajor_empty_like(a, 1, ncols)
    t = Timings()
    st = lib.corrla_rpca_f64(a.ptr, nrows, ncols, a.strides[0], a.strides[1], k, C.byref(o), _ptr(s), _ptr(comps),
                             _ptr(means), C.byref(t))
    del keep
    _ffi.check(st)
    _tls.timings = t.as_dict()
    return (s, comps, means) if return_means else (s, comps)


def power_iter(a_mat, omega_rank: int, n_iter: int, *, omega=None, seed: int | None = None, schedule="reference",
               ctx: Context | None = None, comm: ShardComm | None = None, global_rows: int | None = None):
    """Q = power_iter(a, omega_rank, n_iter) (random_svd.rs:15-59): orthonormal basis (nrows x omega_rank)."""
    a = _Mat(a_mat)
    lib = _ffi.load()
    nrows, ncols = a.shape
    device = a.device if a.on_device else (comm.device if comm is not None else None)
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if a.on_device else None
    o, keep = _make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=omega, seed=seed,
                         schedule=schedule, comm=comm, global_rows=global_rows, stream=stream, device=device)
    q = _colmajor_empty_like(a, nrows, int(omega_rank))
    t = Timings()
    st = lib.corrla_power_iter_f64(a.ptr, nrows, ncols, a.strides[0], a.strides[1], int(omega_rank), int(n_iter),
                                   C.byref(o), _ptr(q), C.byref(t))
    del keep
    _ffi.check(st)
    _tls.timings = t.as_dict()
    return q


def par_matmul(lhs, rhs, beta: float = 1.0, *, ctx: Context | None = None):
    """res = beta * lhs @ rhs with a skinny rhs (128-column panels above that): par_matmul_helper with alpha=None
    (mat_utils.rs:20-33).  Returns a new array where the inputs live."""
    a = _Mat(lhs, "lhs")
    b = _Mat(rhs, "rhs")
    if a.on_device != b.on_device:
        raise ValueError("lhs and rhs must both be on the host or both on the device")
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"dimension mismatch: lhs is {a.shape}, rhs is {b.shape}")   # faer asserts
    lib = _ffi.load()
    device = a.device if a.on_device else None
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if a.on_device else None
    o, _ = _make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=None, seed=0,
                      schedule="reference", comm=None, global_rows=None, stream=stream, device=device)
    res = _colmajor_empty_like(a, a.shape[0], b.shape[1])
    st = lib.corrla_par_matmul_f64(_ptr(res), 1, a.shape[0], a.ptr, a.shape[0], a.shape[1], a.strides[0], a.strides[1],
                                   b.ptr, b.shape[1], b.strides[0], b.strides[1], float(beta),
                                   1 if a.on_device else 0, C.byref(o))
    _ffi.check(st)
    return res


par_matmul_helper = par_matmul


def random_mat_normal(n_rows: int, n_cols: int, seed: int | None = None, *, device: int | None = None,
                      as_torch: bool = False, ctx: Context | None = None):
    """n_rows x n_cols i.i.d. N(0,1) (mat_utils.rs:161-175), drawn on the GPU with Philox4x32-10."""
    lib = _ffi.load()
    ctx = ctx or _context_for(device)
    o, _ = _make_opts(ctx=ctx, on_device=False, out_on_device=as_torch, omega=None, seed=seed, schedule="reference",
                      comm=None, global_rows=None, stream=None, device=device)
    if as_torch:
        import torch
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        out = torch.empty((n_cols, n_rows), dtype=torch.float64, device=dev).t()
        torch.cuda.synchronize(dev)
    else:
        out = np.empty((n_rows, n_cols), dtype=np.float64, order="F")
    st = lib.corrla_random_mat_normal_f64(o.seed, int(n_rows), int(n_cols), _ptr(out), 1 if as_torch else 0, C.byref(o))
    _ffi.check(st)
    return out


def thin_q(a_mat, *, ctx: Context | None = None, comm: ShardComm | None = None, global_rows: int | None = None,
           return_rank: bool = False):
    """Thin Q of a tall matrix by adaptive CholeskyQR2/3 (column panels above 128 columns): the engine's stand-in for
    faer's `qr().compute_thin_q()` (random_svd.rs:38, :57).  Numerically dependent columns are replaced by an
    orthonormal completion, as a Householder QR would."""
    a = _Mat(a_mat)
    lib = _ffi.load()
    device = a.device if a.on_device else (comm.device if comm is not None else None)
    ctx = ctx or _context_for(device)
    stream = _current_stream(device) if a.on_device else None
    o, _ = _make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=None, seed=0,
                      schedule="reference", comm=comm, global_rows=global_rows, stream=stream, device=device)
    q = _colmajor_empty_like(a, a.shape[0], a.shape[1])
    rank = C.c_int(0)
    st = lib.corrla_thin_q_f64(a.ptr, a.shape[0], a.shape[1], a.strides[0], a.strides[1], 1 if a.on_device else 0,
                               C.byref(o), _ptr(q), C.byref(rank))
    _ffi.check(st)
    return (q, rank.value) if return_rank else q


from .rom import DMDc, PodI, RbfInterp, dmdc_operators, pod_modes_weights  # noqa: E402,F401
from .stats import (ActiveSsRsvd, FittedActiveSsRsvd, PolyGradientEstimator, active_ss, active_ss_fit, cov, mat_cov_centered,  # noqa: E402,F401
                    pearson_corr)
