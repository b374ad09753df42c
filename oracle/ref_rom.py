"""CPU oracle for the two reduced-order models that sit directly on the RSVD hot path of wgurecky/CORRLA_RS
  --  TEST INFRASTRUCTURE ONLY, like oracle/ref_rsvd.py.

What it follows (paths relative to the reference checkout):
  * src/lib_math_utils/dmd_rom.rs:46-146     DMDc::new, _calc_dmdc_modes, _calc_eigs, _calc_modes
  * src/lib_math_utils/dmd_rom.rs:148-162    _X / _Y sub-views of the stacked [x; u] snapshot matrix
  * src/lib_math_utils/dmd_rom.rs:165-232    est_a_til, est_b_til, predict, predict_multiple
  * src/lib_math_utils/pod_rom.rs:36-120     PodI::new, _modes, _weights, _mode_interp, predict
  * src/lib_math_utils/interp_utils.rs:84-155  RbfInterp (fit / predict) with the linear kernel (:32-41)
  * src/lib_math_utils/mat_utils.rs:37-71    mat_pinv (1/(s + 1e-14)), mat_pinv_comp (1/(s + (1e-16 + 1e-16 i)))
  * src/lib_math_utils/mat_utils.rs:386-402  mat_pinv_diag (0 if |s| < 1e-20 else 1/(s + 1e-20))
  * src/lib_math_utils/mat_utils.rs:600-610  mat_linspace (i * (end - start) / n: `start` is NOT added, end excluded)
  * src/lib_math_utils/stats_corr.rs:183-198 build_full_vandermonde (degree < 2: [x | 1])

faer's eigendecomposition / SVD / pinv arithmetic is restated with LAPACK through numpy.

Pinning status: the reference cannot be compiled here.  The oracle is pinned against the reference's own DMDc test
(dmd_rom.rs:245-309: 19-step prediction within 5e-2 of the true snapshot for nx in {20, 50, 500}) and runs the POD
test's generator (pod_rom.rs:130-146, which asserts nothing).  Beyond those: "parity unpinned".
"""
from __future__ import annotations

import numpy as np

from . import ref_rsvd

__all__ = ["mat_linspace", "mat_pinv", "mat_pinv_comp", "mat_pinv_diag", "DMDc", "RbfInterpLin", "PodI",
           "dmdc_test_snapshots", "pod_test_snapshots"]


def mat_linspace(start: float, end: float, n_steps: int) -> np.ndarray:
    """mat_utils.rs:600-610 -- note the values are i*delta, the start point is not added."""
    delta = (end - start) / float(n_steps)
    return (np.arange(n_steps, dtype=np.float64) * delta).reshape(n_steps, 1)


def mat_pinv(x: np.ndarray) -> np.ndarray:
    """mat_utils.rs:37-53: V diag(1/(s + 1e-14)) U^T from the full SVD."""
    u, s, vt = np.linalg.svd(x, full_matrices=False)
    return (vt.T * (1.0 / (s + 1.0e-14))) @ u.T


def mat_pinv_comp(x: np.ndarray) -> np.ndarray:
    """mat_utils.rs:56-71: complex pseudo-inverse with eps = 1e-16 + 1e-16 i added to every singular value."""
    u, s, vh = np.linalg.svd(x, full_matrices=False)
    s_inv = 1.0 / (s.astype(np.complex128) + complex(1.0e-16, 1.0e-16))
    return (vh.conj().T * s_inv) @ u.conj().T


def mat_pinv_diag(s: np.ndarray) -> np.ndarray:
    """mat_utils.rs:386-402 on the diagonal entries."""
    s = np.asarray(s, dtype=np.float64).ravel()
    out = np.zeros_like(s)
    live = ~((s < 1.0e-20) & (s > -1.0e-20))
    out[live] = 1.0 / (s[live] + 1.0e-20)
    return out


class DMDc:
    """dmd_rom.rs:20-232.  x_data: n_x x n_snapshots, u_data: n_u x n_snapshots.
    `omegas` = (Omega for the RSVD of the input space, Omega for the RSVD of the output space), each shaped like
    ref_rsvd.random_svd's `omega` for that view (the reference draws them unseeded)."""

    N_OVERSAMPLES = 12                                                   # dmd_rom.rs:72,82

    def __init__(self, x_data, u_data, dt: float, n_modes: int, n_iters: int, omegas=(None, None), rng=None):
        x_data = np.asarray(x_data, dtype=np.float64)
        u_data = np.asarray(u_data, dtype=np.float64)
        self.n_snapshots = x_data.shape[1]
        self.n_x, self.n_u = x_data.shape[0], u_data.shape[0]
        self.n_modes, self.dt_snapshots = n_modes, dt
        omega = np.vstack([x_data, u_data])                              # :66
        xv, yv = omega[:, :-1], omega[:self.n_x, 1:]                     # :148-162
        u_til, s_til, v_til_t = ref_rsvd.random_svd(xv, n_modes, n_iters, self.N_OVERSAMPLES, omega=omegas[0], rng=rng)  # :72
        v_til = v_til_t.T
        u_til_1, u_til_2 = u_til[:self.n_x], u_til[self.n_x:]            # :75-79
        u_hat, _s, _v = ref_rsvd.random_svd(yv, n_modes, n_iters, self.N_OVERSAMPLES, omega=omegas[1], rng=rng)  # :82
        s_inv = mat_pinv_diag(s_til)                                     # :86-87
        tmp_op_scale = ((u_hat.T @ yv) @ v_til) * s_inv                  # :90-94
        self.a_til = (tmp_op_scale @ u_til_1.T) @ u_hat                  # :95-97   (self._A)
        b_til = tmp_op_scale @ u_til_2.T                                 # :100-102
        self.b = u_hat @ b_til                                           # :106     (self._B)
        self.u_hat, self.s_til = u_hat, s_til
        # _calc_eigs / _calc_modes (:112-146)
        lambdas, w = np.linalg.eig(self.a_til)
        self.lambdas = lambdas.reshape(-1, 1)
        self.modes_scale = yv @ (v_til @ (s_inv[:, None] * (u_til_1.T @ u_hat)))   # :133-139
        self.modes_re = self.modes_scale @ w.real
        self.modes_im = self.modes_scale @ w.imag

    def est_a_til(self) -> np.ndarray:
        modes = self.modes_re + 1j * self.modes_im                       # :167-174
        return ((modes * self.lambdas.ravel()) @ mat_pinv_comp(modes)).real

    def est_b_til(self) -> np.ndarray:
        return self.b

    def predict(self, x_0, u_input) -> np.ndarray:
        return self.est_a_til() @ x_0 + self.b @ u_input                 # :185-196

    def predict_multiple(self, x_0, u_seq) -> np.ndarray:
        a_til = self.est_a_til()                                         # :201-231
        x_cur = np.asarray(x_0, dtype=np.float64).reshape(self.n_x, 1)
        out = np.zeros((self.n_x, u_seq.shape[1]))
        for j in range(u_seq.shape[1]):
            x_cur = a_til @ x_cur + self.b @ u_seq[:, j:j + 1]
            out[:, j] = x_cur[:, 0]
        return out


class RbfInterpLin:
    """interp_utils.rs:84-155 with RbfKernelLin (phi(r) = r) and poly_degree 1, as pod_rom.rs:86-88 builds it."""

    def __init__(self, dim: int):
        self.dim = dim
        self.x_known = None
        self.coeffs = None

    def _build_kp(self, x_in: np.ndarray, full: bool) -> np.ndarray:
        k = np.linalg.norm(x_in[:, None, :] - self.x_known[None, :, :], axis=2)      # :97-107
        p = np.hstack([x_in, np.ones((x_in.shape[0], 1))])                           # stats_corr.rs:185-189
        upper = np.hstack([k, p])
        if not full:
            return upper
        lower = np.hstack([p.T, np.zeros((p.shape[1], p.shape[1]))])
        return np.vstack([upper, lower])

    def fit(self, x_in, y_in) -> None:
        x_in = np.asarray(x_in, dtype=np.float64)
        assert x_in.shape[1] == self.dim
        self.x_known = x_in.copy()
        kp_inv = mat_pinv(self._build_kp(x_in, True))                                # :137-139
        y = np.asarray(y_in, dtype=np.float64).reshape(-1, 1)
        y_pad = np.zeros((kp_inv.shape[1] - y.shape[0], 1))
        self.coeffs = kp_inv @ np.vstack([y, y_pad])                                 # :140-144

    def predict(self, x_query) -> np.ndarray:
        x_query = np.asarray(x_query, dtype=np.float64)
        return self._build_kp(x_query, False) @ self.coeffs                          # :150-154


class PodI:
    """pod_rom.rs:20-120.  x_data: n_snapshots x n_points (one snapshot per ROW), t: n_snapshots x dim."""

    def __init__(self, x_data, t, n_modes: int, omega=None, rng=None):
        x_data = np.asarray(x_data, dtype=np.float64)
        t = np.asarray(t, dtype=np.float64)
        assert t.shape[0] == x_data.shape[0]                                         # :38
        _u, _s, vt = ref_rsvd.random_svd(x_data, n_modes, 10, 10, omega=omega, rng=rng)   # :56
        self.modes = vt.T.copy()                                                     # :57  n_points x n_modes
        modes_inv = mat_pinv(self.modes)                                             # :64
        self.mode_weights = (modes_inv @ x_data.T).T                                 # :66-73, one row at a time there
        self.n_modes, self.n_snapshots, self.t_abscissa = n_modes, x_data.shape[0], t.copy()
        self.interps = []
        for j in range(n_modes):                                                     # :84-93
            f = RbfInterpLin(t.shape[1])
            f.fit(t, self.mode_weights[:, j])
            self.interps.append(f)

    def weights_at(self, t_query) -> np.ndarray:
        t_query = np.asarray(t_query, dtype=np.float64)
        assert t_query.shape[0] == 1                                                 # :109
        return np.array([[f.predict(t_query)[0, 0]] for f in self.interps])          # :110-113

    def predict(self, t_query) -> np.ndarray:
        return self.modes @ self.weights_at(t_query)                                 # :117


def dmdc_test_snapshots(nx: int, nt: int):
    """The generator of the reference's DMDc test (dmd_rom.rs:245-267): returns (p_snapshots nx x nt, u 1 x nt)."""
    x = mat_linspace(0.0, 10.0, nx)[:, 0]
    t = mat_linspace(0.0, 10.0, nt)[:, 0]
    u = np.exp(0.2 * t)
    p = np.sin(x[:, None] + 0.2 * t[None, :]) * u[None, :]
    return p, u.reshape(1, nt)


def pod_test_snapshots(nx: int = 100, n_snapshots: int = 20, sigma: float = 0.25):
    """The generator of the reference's POD test (pod_rom.rs:125-146): returns (snapshots n_snapshots x nx, t)."""
    x = mat_linspace(0.0, 10.0, nx)[:, 0]
    t = mat_linspace(1.0, 9.0, n_snapshots)
    p = (0.5 * t[:, 0])[None, :] * np.exp(-(x[:, None] - t[:, 0][None, :]) ** 2 / sigma ** 2)
    return p.T.copy(), t
