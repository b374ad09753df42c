"""Second-moment statistics and the active-subspace fit on the Gram kernel: host-side mirrors of `mat_cov_centered`,
`pearson_corr` (src/lib_math_utils/stats_corr.rs:14-43) and of `ActiveSsRsvd::fit` / `fit_svd` / `FittedActiveSsRsvd`
(src/lib_math_utils/active_subspaces.rs:147-278) over `corrla_cov_f64` and `corrla_rsvd_f64`.

The gradient matrix of the reference's `PolyGradientEstimator` (kd-tree neighbours + one local polynomial fit per sample,
active_subspaces.rs:66-141, :226-238) is built on the device by `corrla_active_ss_f64` (exact brute-force neighbour
search + batched least squares): `active_ss_fit` / `active_ss`.  `ActiveSsRsvd` additionally takes a precomputed
gradient matrix or any estimator object with the reference's `grad_at(x0)` interface."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

__all__ = ["cov", "mat_cov_centered", "pearson_corr", "ActiveSsRsvd", "FittedActiveSsRsvd", "PolyGradientEstimator",
           "active_ss_fit", "active_ss"]

_KINDS = {"gram": 0, "centered": 1, "pearson": 2}


def _api():
    import corrla_rs_b200 as api
    return api


def cov(x, kind: str = "centered", *, scale: float = 1.0, evd: bool = False, ctx=None, comm=None, global_rows=None):
    """corrla_cov_f64 on a samples-by-features matrix (numpy or torch CUDA tensor, <= 128 features).
    Returns `out` (d, d), or (out, means (1, d), evals (d, 1), evecs (d, d)) with evd=True."""
    api = _api()
    lib = _ffi.load()
    a = api._Mat(x, "x")
    n, d = a.shape
    device = a.device if a.on_device else (comm.device if comm is not None else None)
    ctx = ctx or api._context_for(device)
    stream = api._current_stream(device) if a.on_device else None
    o, _ = api._make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=None, seed=0,
                          schedule="reference", comm=comm, global_rows=global_rows, stream=stream, device=device)
    out = api._colmajor_empty_like(a, d, d)
    means = api._colmajor_empty_like(a, 1, d)
    evals = api._colmajor_empty_like(a, d, 1) if evd else None
    evecs = api._colmajor_empty_like(a, d, d) if evd else None
    st = lib.corrla_cov_f64(a.ptr, n, d, a.strides[0], a.strides[1], _KINDS[kind], float(scale), C.byref(o),
                            api._ptr(out), api._ptr(means), api._ptr(evals) if evd else None,
                            api._ptr(evecs) if evd else None)
    _ffi.check(st)
    if a.on_device:
        import torch
        torch.cuda.current_stream(device).synchronize()
    return (out, means, evals, evecs) if evd else out


def mat_cov_centered(x, **kw):
    """(x - mean)^T (x - mean) / (N - 1)  (stats_corr.rs:32-43)."""
    return cov(x, "centered", **kw)


def pearson_corr(x, **kw):
    """Linear correlation coefficients between the columns of x (stats_corr.rs:14-28)."""
    return cov(x, "pearson", **kw)


def active_ss_fit(x, y, order: int, n_nbr: int, n_comps: int, *, return_gradients: bool = False, ctx=None):
    """Gradient matrix from the samples (x: (N, k), y: (N,) or (N, 1)) and the fit of `ActiveSsRsvd::fit`, in one device
    call.  Returns a FittedActiveSsRsvd (plus the k x N gradient matrix with return_gradients=True); its attribute
    `n_deficient` counts samples whose neighbourhood gave a rank-deficient local fit."""
    api = _api()
    lib = _ffi.load()
    a = api._Mat(x, "a_mat")
    if api._is_torch(y):
        yy = y.reshape(-1, 1)
    else:
        yy = np.asarray(y)
        if yy.dtype != np.float64:
            raise TypeError("y must be a float64 array")
        yy = yy.reshape(-1, 1)
    b = api._Mat(yy, "y")
    if a.on_device != b.on_device:
        raise ValueError("a_mat and y must both be on the host or both on the device")
    n, k = a.shape
    if b.shape[0] != n:
        raise ValueError(f"a_mat has {n} samples, y has {b.shape[0]}")
    for name, v in (("order", order), ("n_nbr", n_nbr), ("n_comps", n_comps)):
        if not isinstance(v, (int, np.integer)) or isinstance(v, bool):
            raise TypeError(f"{name} must be an int")
        if v < 0:
            raise OverflowError(f"can't convert negative int to unsigned ({name})")
    device = a.device if a.on_device else None
    ctx = ctx or api._context_for(device)
    stream = api._current_stream(device) if a.on_device else None
    o, _ = api._make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=None, seed=0,
                          schedule="reference", comm=None, global_rows=None, stream=stream, device=device)
    evals = api._colmajor_empty_like(a, k, 1)
    evecs = api._colmajor_empty_like(a, k, k)
    grads = api._colmajor_empty_like(a, k, n) if return_gradients else None
    ndef = C.c_int(0)
    st = lib.corrla_active_ss_f64(a.ptr, n, k, a.strides[0], a.strides[1], b.ptr, b.strides[0], int(order), int(n_nbr),
                                  C.byref(o), api._ptr(evals), api._ptr(evecs),
                                  api._ptr(grads) if return_gradients else None, C.byref(ndef))
    _ffi.check(st)
    ev, vec = (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v) for v in (evals, evecs))
    fit = FittedActiveSsRsvd(vec, np.diag(ev.ravel()), n_comps)
    fit.n_deficient = int(ndef.value)
    return (fit, grads) if return_gradients else fit


def active_ss(a_mat, y, order: int, n_nbr: int, n_comps: int):
    """Drop-in for the pyo3 `corrla_rs.active_ss(a_mat, y, order, n_nbr, n_comps)` (lib_math_utils_py.rs:56-86):
    returns (components (k, n_comps), singular_vals (k, n_comps) -- the leading columns of the diagonal eigenvalue
    matrix, as the reference slices it --, var_sensi (k,))."""
    fit = active_ss_fit(a_mat, y, order, n_nbr, n_comps)
    return (np.ascontiguousarray(fit.components()), np.ascontiguousarray(fit.singular_vals()),
            np.asarray(fit.var_diag_evd_sensi()))


class PolyGradientEstimator:
    """`PolyGradientEstimator` of active_subspaces.rs:23-141 over `corrla_poly_grad_at_f64`: holds the samples
    (x_mat: (N, k), y: (N,) or (N, 1)), and `grad_at(x0)` fits the order-1 or order-2 polynomial through the n_nbrs nearest
    samples of x0 and returns its gradient there as a (1, k) row, like the reference.  `grad_at_many(xq)` does the same
    for every row of xq in one device call ((n_query, k) out); `ActiveSsRsvd.create_grad_mat` uses it instead of the
    reference's loop.  The reference asserts on construction that N and n_nbrs exceed the number of fitted
    coefficients (:115-116, :127-128); here the same condition raises ValueError on the first call."""

    def __init__(self, x_mat, y, est_order: int, n_nbrs: int, *, ctx=None):
        api = _api()
        self._x = api._Mat(x_mat, "x_mat")
        if api._is_torch(y):
            yy = y.reshape(-1, 1)
        else:
            yy = np.asarray(y)
            if yy.dtype != np.float64:
                raise TypeError("y must be a float64 array")
            yy = yy.reshape(-1, 1)
        self._y = api._Mat(yy, "y")
        if self._x.on_device != self._y.on_device:
            raise ValueError("x_mat and y must both be on the host or both on the device")
        if self._y.shape[0] != self._x.shape[0]:
            raise ValueError(f"x_mat has {self._x.shape[0]} samples, y has {self._y.shape[0]}")
        for name, v in (("est_order", est_order), ("n_nbrs", n_nbrs)):
            if not isinstance(v, (int, np.integer)) or isinstance(v, bool):
                raise TypeError(f"{name} must be an int")
            if v < 0:
                raise OverflowError(f"can't convert negative int to unsigned ({name})")
        self.est_order, self.n_nbrs, self.k = int(est_order), int(n_nbrs), self._x.shape[1]
        self._ctx = ctx
        self.n_deficient = 0

    def grad_at_many(self, xq):
        api = _api()
        lib = _ffi.load()
        a, b = self._x, self._y
        if a.on_device:
            import torch
            if not api._is_torch(xq):
                xq = torch.as_tensor(np.asarray(xq, dtype=np.float64), device=a.device)
            q = api._Mat(xq.reshape(-1, self.k), "xq")
        else:
            q = api._Mat(np.asarray(xq, dtype=np.float64).reshape(-1, self.k), "xq")
        nq = q.shape[0]
        device = a.device if a.on_device else None
        ctx = self._ctx or api._context_for(device)
        stream = api._current_stream(device) if a.on_device else None
        o, _ = api._make_opts(ctx=ctx, on_device=a.on_device, out_on_device=a.on_device, omega=None, seed=0,
                              schedule="reference", comm=None, global_rows=None, stream=stream, device=device)
        out = api._colmajor_empty_like(a, self.k, nq)          # k x nq column-major == nq x k row-major
        ndef = C.c_int(0)
        st = lib.corrla_poly_grad_at_f64(a.ptr, a.shape[0], self.k, a.strides[0], a.strides[1], b.ptr, b.strides[0],
                                         self.est_order, self.n_nbrs, q.ptr, nq, q.strides[0], q.strides[1], C.byref(o),
                                         api._ptr(out), C.byref(ndef))
        _ffi.check(st)
        self.n_deficient = int(ndef.value)
        return out.t() if hasattr(out, "detach") else out.T

    def grad_at(self, x0):
        """Gradient [dy/dx_1 .. dy/dx_k] at x0 (sequence of k values) as a (1, k) row (:57-66)."""
        return self.grad_at_many(np.asarray(x0, dtype=np.float64).reshape(1, -1) if not hasattr(x0, "detach") else x0)


class FittedActiveSsRsvd:
    """active_subspaces.rs:147-212: `components_` (k, k) one direction per column, `singular_vals_` (k, k) diagonal."""

    def __init__(self, components, singular_vals, n_comps: int):
        self.components_ = np.asarray(components)
        self.singular_vals_ = np.asarray(singular_vals)
        self.n_comps = int(n_comps)

    def var_diag_evd_sensi(self):
        m = self.components_.T @ self.singular_vals_ @ self.components_          # :163-170
        return [float(m[i, i]) for i in range(self.singular_vals_.shape[0])]

    def components(self):
        return self.components_[:, :self.n_comps]

    def singular_vals(self):
        return self.singular_vals_[:, :self.n_comps]

    def transform(self, x_mat):
        return np.asarray(x_mat) @ self.components()                             # :175-181

    def inv_transform(self, x_mat):
        x_mat = np.asarray(x_mat)
        assert x_mat.shape[1] == self.n_comps                                    # :187
        return x_mat @ self.components().T


class ActiveSsRsvd:
    """active_subspaces.rs:215-278.  `grad_est` is an object with `grad_at(x0) -> (1, k)` like the reference's `GradEst`
    trait; the `*_gradients` methods take the k x N gradient matrix directly (torch CUDA tensors stay on the device)."""

    def __init__(self, grad_est=None, n_comps: int = 1):
        self.grad_est, self.n_comps = grad_est, int(n_comps)

    def create_grad_mat(self, x_mat):
        if isinstance(self.grad_est, PolyGradientEstimator):                     # one device call for all rows
            g = self.grad_est.grad_at_many(x_mat)
            return g.t() if hasattr(g, "detach") else np.ascontiguousarray(g.T)
        x_mat = np.asarray(x_mat, dtype=np.float64)                              # :226-238 (host loop over the estimator)
        g = np.zeros((x_mat.shape[1], x_mat.shape[0]))
        for i in range(x_mat.shape[0]):
            g[:, i] = np.asarray(self.grad_est.grad_at(list(x_mat[i]))).reshape(-1)
        return g

    def fit(self, x_mat):
        return self.fit_gradients(self.create_grad_mat(x_mat))

    def fit_svd(self, x_mat, n_iter=None, n_oversamples=None):
        return self.fit_svd_gradients(self.create_grad_mat(x_mat), n_iter, n_oversamples)

    def fit_gradients(self, grad_mat, **kw):
        """Sorted eigendecomposition of grad_mat grad_mat^T / N (:248-278): Gram and Jacobi on the device."""
        k, n = grad_mat.shape
        gt = grad_mat.t() if hasattr(grad_mat, "detach") else np.asarray(grad_mat).T    # N x k view, no copy
        _out, _mu, evals, evecs = cov(gt, "gram", scale=1.0 / n, evd=True, **kw)
        evals, evecs = (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v) for v in (evals, evecs))
        return FittedActiveSsRsvd(evecs, np.diag(evals.ravel()), self.n_comps)

    def fit_svd_gradients(self, grad_mat, n_iter=None, n_oversamples=None, **kw):
        """RSVD of grad_mat / sqrt(N) (:231-246); the 1/sqrt(N) is applied to the k singular values, not to a copy."""
        api = _api()
        k, n = grad_mat.shape
        ur, sr, _vr = api.rsvd(grad_mat, min(k, self.n_comps), 8 if n_iter is None else n_iter,
                               10 if n_oversamples is None else n_oversamples, **kw)
        ur, sr = (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v) for v in (ur, sr))
        return FittedActiveSsRsvd(ur, np.diag(sr.ravel() / np.sqrt(n)), self.n_comps)
