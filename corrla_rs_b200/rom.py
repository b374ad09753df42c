"""Reduced-order models on top of the B200 RSVD: host-side mirrors of the reference's `DMDc` (src/lib_math_utils/
dmd_rom.rs, pyo3 `PyDMDc` lib_math_utils_py.rs:255-283) and `PodI` (src/lib_math_utils/pod_rom.rs, pyo3 `PyPodI`
lib_math_utils_py.rs:223-250), with the same constructor arguments, method names and outputs.

Everything that touches the tall snapshot matrix (the RSVDs, the products with the snapshots, the mode lift) runs in
libcorrla_b200.so (`corrla_dmdc_f64`, `corrla_pod_f64`, `corrla_par_matmul_f64`).  What stays here is the part the
reference also does on small dense matrices: the r x r eigendecomposition of the reduced operator, the complex
pseudo-inverse of the n_x x r mode matrix used by `predict`, and the n_snap x n_snap RBF systems of the POD weight
interpolants.  No CPU fallback: without the CUDA library the constructors raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import Timings

__all__ = ["dmdc_operators", "pod_modes_weights", "DMDc", "PodI", "RbfInterp", "PyDMDc", "PyPodI", "PyRbfInterp"]


def _api():
    import corrla_rs_b200 as api
    return api


# --------------------------------------------------------------------------------------------
# thin wrappers over the two C entry points
# --------------------------------------------------------------------------------------------
def dmdc_operators(x_data, u_data, n_modes: int, n_iters: int, *, omegas=(None, None), seed=None, ctx=None, comm=None,
                   global_rows=None):
    """corrla_dmdc_f64: returns dict(a_til (r, r), b (n_x, n_u), modes_scale (n_x, r), s_til (r, 1), u_hat (n_x, r)).
    x_data: n_x x n_snapshots, u_data: n_u x n_snapshots; numpy (host) or torch CUDA tensors (device).
    With `comm` (one process per GPU) x_data is this rank's block of state rows and u_data the whole control matrix;
    a_til and s_til come back replicated, b / modes_scale / u_hat for the local rows; `global_rows` = total state rows."""
    api = _api()
    lib = _ffi.load()
    x = api._Mat(x_data, "x_data")
    u = api._Mat(u_data, "u_data")
    if x.on_device != u.on_device:
        raise ValueError("x_data and u_data must both be on the host or both on the device")
    if x.shape[1] != u.shape[1]:
        raise ValueError(f"x_data has {x.shape[1]} snapshots, u_data has {u.shape[1]}")
    n_x, n_snap = x.shape
    n_u = u.shape[0]
    r = int(n_modes)
    device = x.device if x.on_device else (comm.device if comm is not None else None)
    ctx = ctx or api._context_for(device)
    stream = api._current_stream(device) if x.on_device else None
    o, keep = api._make_opts(ctx=ctx, on_device=x.on_device, out_on_device=x.on_device, omega=omegas[0], seed=seed,
                             schedule="reference", comm=comm, global_rows=global_rows, stream=stream, device=device)
    oy = None
    if omegas[1] is not None:
        oy = api._Mat(omegas[1], "omega_y")
        if omegas[0] is None:
            raise ValueError("inject both sketch matrices or neither")
        if oy.on_device != bool(o.omega_on_device) or oy.strides != (o.omega_rs, o.omega_cs):
            raise ValueError("the two injected sketch matrices must share residency and strides")
    a_til = api._colmajor_empty_like(x, max(r, 1), max(r, 1))
    b = api._colmajor_empty_like(x, n_x, max(n_u, 1))
    modes_scale = api._colmajor_empty_like(x, n_x, max(r, 1))
    s_til = api._colmajor_empty_like(x, max(r, 1), 1)
    u_hat = api._colmajor_empty_like(x, n_x, max(r, 1))
    t = Timings()
    st = lib.corrla_dmdc_f64(x.ptr, n_x, n_snap, x.strides[0], x.strides[1], u.ptr if n_u else None, n_u,
                             u.strides[0], u.strides[1], r, int(n_iters), C.byref(o), oy.ptr if oy else None,
                             api._ptr(a_til), api._ptr(b) if n_u else None, api._ptr(modes_scale), api._ptr(s_til),
                             api._ptr(u_hat), C.byref(t))
    del keep
    _ffi.check(st)
    api._tls.timings = t.as_dict()
    return {"a_til": a_til, "b": b[:, :n_u], "modes_scale": modes_scale, "s_til": s_til, "u_hat": u_hat}


def pod_modes_weights(x_data, n_modes: int, *, omega=None, seed=None, ctx=None, comm=None, global_rows=None):
    """corrla_pod_f64: (modes (n_points, r), weights (n_snapshots, r), s (r, 1)); x_data is n_snapshots x n_points.
    With `comm` x_data is this rank's block of POINTS (columns); modes come back for those points, weights and s
    replicated; `global_rows` = total number of points."""
    api = _api()
    lib = _ffi.load()
    x = api._Mat(x_data, "x_data")
    n_snap, n_points = x.shape
    r = int(n_modes)
    device = x.device if x.on_device else (comm.device if comm is not None else None)
    ctx = ctx or api._context_for(device)
    stream = api._current_stream(device) if x.on_device else None
    o, keep = api._make_opts(ctx=ctx, on_device=x.on_device, out_on_device=x.on_device, omega=omega, seed=seed,
                             schedule="reference", comm=comm, global_rows=global_rows, stream=stream, device=device)
    modes = api._colmajor_empty_like(x, n_points, max(r, 1))
    weights = api._colmajor_empty_like(x, n_snap, max(r, 1))
    s = api._colmajor_empty_like(x, max(r, 1), 1)
    t = Timings()
    st = lib.corrla_pod_f64(x.ptr, n_snap, n_points, x.strides[0], x.strides[1], r, C.byref(o), api._ptr(modes),
                            api._ptr(weights), api._ptr(s), C.byref(t))
    del keep
    _ffi.check(st)
    api._tls.timings = t.as_dict()
    return modes, weights, s


def _to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


# --------------------------------------------------------------------------------------------
# DMDc
# --------------------------------------------------------------------------------------------
class DMDc:
    """x_{t+1} = A x_t + B u_t fitted from snapshots (dmd_rom.rs:46-146).

    DMDc(x_data (n_x, n_snapshots), u_data (n_u, n_snapshots), dt, n_modes, n_iters) like `DMDc::new` (:46);
    attributes `lambdas` (n_modes, 1) complex, `modes_re`, `modes_im` (n_x, n_modes), `_A` (reduced operator),
    `_B` (n_x, n_u); methods `est_a_til`, `est_b_til`, `predict`, `predict_multiple` (:165-232)."""

    def __init__(self, x_data, u_data, dt: float, n_modes: int, n_iters: int, *, omegas=(None, None), seed=None,
                 ctx=None):
        api = _api()
        self.n_x, self.n_snapshots = (int(v) for v in x_data.shape)
        self.n_u = int(u_data.shape[0])
        self.n_modes, self.dt_snapshots = int(n_modes), float(dt)
        ops = dmdc_operators(x_data, u_data, n_modes, n_iters, omegas=omegas, seed=seed, ctx=ctx)
        self.timings = api.last_timings()
        self._A = _to_numpy(ops["a_til"])
        self._B = _to_numpy(ops["b"])
        self.s_til = _to_numpy(ops["s_til"])
        # _calc_eigs (:112-126): r x r, on the host like the reference's faer call
        lambdas, w = np.linalg.eig(self._A)
        self.lambdas = lambdas.reshape(-1, 1)
        # _calc_modes (:128-146): lift with two more skinny products on the device
        scale = ops["modes_scale"]
        if hasattr(scale, "detach"):
            import torch
            w_re = torch.from_numpy(np.ascontiguousarray(w.real)).to(scale.device)
            w_im = torch.from_numpy(np.ascontiguousarray(w.imag)).to(scale.device)
        else:
            w_re, w_im = np.ascontiguousarray(w.real), np.ascontiguousarray(w.imag)
        self.modes_re = _to_numpy(api.par_matmul(scale, w_re, ctx=ctx))
        self.modes_im = _to_numpy(api.par_matmul(scale, w_im, ctx=ctx))
        self._modes_pinv = None

    def _pinv_modes(self):
        # mat_pinv_comp (mat_utils.rs:56-71): eps = 1e-16 + 1e-16 i added to every singular value
        if self._modes_pinv is None:
            modes = self.modes_re + 1j * self.modes_im
            u, s, vh = np.linalg.svd(modes, full_matrices=False)
            s_inv = 1.0 / (s.astype(np.complex128) + complex(1.0e-16, 1.0e-16))
            self._modes_pinv = (vh.conj().T * s_inv) @ u.conj().T
        return self._modes_pinv

    def est_a_til(self) -> np.ndarray:
        """Re(Phi Lambda Phi^+), n_x x n_x (:165-176).  Dense: only sensible for small n_x; `predict*` never forms it."""
        modes = self.modes_re + 1j * self.modes_im
        return ((modes * self.lambdas.ravel()) @ self._pinv_modes()).real

    def est_b_til(self) -> np.ndarray:
        return self._B

    def _step(self, x_cur, u_col):
        modes = self.modes_re + 1j * self.modes_im
        z = self._pinv_modes() @ x_cur                       # r x 1
        return (modes @ (self.lambdas * z)).real + self._B @ u_col

    def predict(self, x_0, u_input) -> np.ndarray:
        x_0, u_input = np.asarray(x_0, dtype=np.float64), np.asarray(u_input, dtype=np.float64)
        assert x_0.shape == (self.n_x, 1) and u_input.shape == (self.n_u, 1)     # :185-188
        return self._step(x_0, u_input)

    def predict_multiple(self, x_0, u_seq) -> np.ndarray:
        x_0, u_seq = np.asarray(x_0, dtype=np.float64), np.asarray(u_seq, dtype=np.float64)
        assert x_0.shape == (self.n_x, 1) and u_seq.shape[0] == self.n_u          # :202-204
        out = np.zeros((self.n_x, u_seq.shape[1]))
        x_cur = x_0
        for j in range(u_seq.shape[1]):
            x_cur = self._step(x_cur, u_seq[:, j:j + 1])
            out[:, j] = x_cur[:, 0]
        return out


class PyDMDc:
    """pyo3 signature (lib_math_utils_py.rs:262-282): PyDMDc(x_np, u_np, n_modes, n_iters); predict(x0_np, u_np)."""

    def __init__(self, x_np, u_np, n_modes: int, n_iters: int):
        self.dmd = DMDc(x_np, u_np, 1.0, n_modes, n_iters)

    def predict(self, x0_np, u_np) -> np.ndarray:
        return self.dmd.predict_multiple(x0_np, u_np)


# --------------------------------------------------------------------------------------------
# RBF interpolation of the POD weights (interp_utils.rs) -- small dense host work
# --------------------------------------------------------------------------------------------
def _pinv_eps(x: np.ndarray, eps: float = 1.0e-14) -> np.ndarray:
    u, s, vt = np.linalg.svd(x, full_matrices=False)           # mat_pinv, mat_utils.rs:37-53
    return (vt.T * (1.0 / (s + eps))) @ u.T


class RbfInterp:
    """interp_utils.rs:84-155.  kernel_type: 1 linear, 2 multiquadric, 3 cubic, anything else Gaussian
    (lib_math_utils_py.rs:191-196); augmenting polynomial [x | 1] for poly_degree < 2 (stats_corr.rs:183-190)."""

    def __init__(self, kernel_type: int, kernel_param: float, dim: int, poly_degree: int):
        if poly_degree >= 2:
            raise NotImplementedError("only the affine augmenting polynomial used by PodI is provided")
        self.kernel_type, self.eps, self.dim = int(kernel_type), float(kernel_param), int(dim)
        self.x_known = None
        self.coeffs = None

    def _phi(self, r):
        if self.kernel_type == 1:
            return r
        if self.kernel_type == 2:
            return np.sqrt(1.0 + (self.eps * r) ** 2)
        if self.kernel_type == 3:
            return r * r * r
        return np.exp(-((r * self.eps) ** 2))

    def _build_kp(self, x_in, full):
        k = self._phi(np.linalg.norm(x_in[:, None, :] - self.x_known[None, :, :], axis=2))
        p = np.hstack([x_in, np.ones((x_in.shape[0], 1))])
        upper = np.hstack([k, p])
        if not full:
            return upper
        return np.vstack([upper, np.hstack([p.T, np.zeros((p.shape[1], p.shape[1]))])])

    def fit(self, x_in, y_in) -> None:
        x_in = np.asarray(x_in, dtype=np.float64)
        y = np.asarray(y_in, dtype=np.float64).reshape(x_in.shape[0], -1)
        assert x_in.shape[1] == self.dim
        self.x_known = x_in.copy()
        kp_inv = _pinv_eps(self._build_kp(x_in, True))
        self.coeffs = kp_inv @ np.vstack([y, np.zeros((kp_inv.shape[1] - y.shape[0], y.shape[1]))])

    def predict(self, x_query) -> np.ndarray:
        x_query = np.asarray(x_query, dtype=np.float64)
        assert x_query.shape[1] == self.dim
        return self._build_kp(x_query, False) @ self.coeffs


PyRbfInterp = RbfInterp


# --------------------------------------------------------------------------------------------
# POD with interpolated weights
# --------------------------------------------------------------------------------------------
class PodI:
    """y(x, t) = sum_i w_i(t) phi_i(x) (pod_rom.rs:36-120).  PodI(x_data (n_snapshots, n_points), t (n_snapshots, dim),
    n_modes); attributes `modes` (n_points, n_modes), `mode_weights` (n_snapshots, n_modes); `predict(t_query (1, dim))`
    returns (n_points, 1)."""

    def __init__(self, x_data, t, n_modes: int, *, omega=None, seed=None, ctx=None):
        t = np.asarray(_to_numpy(t), dtype=np.float64)
        assert t.shape[0] == x_data.shape[0]                                       # :38
        modes, weights, s = pod_modes_weights(x_data, n_modes, omega=omega, seed=seed, ctx=ctx)
        self.timings = _api().last_timings()
        self.modes = modes
        self.mode_weights = _to_numpy(weights)
        self.singular_vals = _to_numpy(s)
        self.n_modes, self.n_snapshots, self.t_abscissa = int(n_modes), int(x_data.shape[0]), t.copy()
        # one linear-kernel interpolant per weight (:78-96); all share the kernel matrix, so solve them together
        self._interp = RbfInterp(1, 0.0, t.shape[1], 1)
        self._interp.fit(t, self.mode_weights)

    def fit(self, x_data, t, n_modes: int) -> None:
        self.__init__(x_data, t, n_modes)                                          # :98-102

    def weights_at(self, t_query) -> np.ndarray:
        t_query = np.asarray(t_query, dtype=np.float64)
        assert t_query.shape[0] == 1                                               # :109
        return self._interp.predict(t_query).reshape(self.n_modes, 1)

    def predict(self, t_query):
        w = self.weights_at(t_query)
        if hasattr(self.modes, "detach"):
            import torch
            return self.modes @ torch.from_numpy(w).to(self.modes.device)
        return self.modes @ w                                                      # :117


class PyPodI:
    """pyo3 signature (lib_math_utils_py.rs:230-249): PyPodI(x_np, t_np, n_modes); predict(t_np)."""

    def __init__(self, x_np, t_np, n_modes: int):
        self.pod = PodI(x_np, t_np, n_modes)

    def predict(self, t_np) -> np.ndarray:
        return _to_numpy(self.pod.predict(t_np))
