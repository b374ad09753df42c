"""CPU oracle for the second-moment statistics and the active-subspace fit
--  TEST INFRASTRUCTURE ONLY, like the other modules in oracle/.

What it follows (paths relative to the reference checkout):
  * src/lib_math_utils/stats_corr.rs:14-43      pearson_corr, mat_cov_centered
  * src/lib_math_utils/mat_utils.rs:87-160      mat_mean, mat_std (N - 1)
  * src/lib_math_utils/mat_utils.rs:482-520     center_mat_col, zcenter_mat_col
  * src/lib_math_utils/stats_corr.rs:112-249    mat_col_interactions, linear_fit, jac_from_lin, build_vandermonde,
                                                quad_fit, quad_eval, jac_from_quad (forward difference, eps = 1e-10)
  * src/lib_math_utils/active_subspaces.rs:66-141   PolyGradientEstimator (the kd-tree is restated as an exact
                                                brute-force k-nearest-neighbour search: same neighbours)
  * src/lib_math_utils/active_subspaces.rs:147-278  FittedActiveSsRsvd, ActiveSsRsvd::{create_grad_mat, fit, fit_svd}

Pinning status: the reference cannot be compiled here.  Pinned against the reference's own tests, which are statistical
(test_pearson / test_cov: identity within 1e-1 on 10 000 x 5 Gaussian samples, stats_corr.rs:259-298; test_grad_est /
test_active_ss: gradients within 1e-2 / 1e-1 and orderings, active_subspaces.rs:281-384).  Beyond those: "parity unpinned".
"""
from __future__ import annotations

import numpy as np

from . import ref_rsvd
from .ref_rom import mat_pinv

__all__ = ["mat_cov_centered", "pearson_corr", "PolyGradientEstimator", "ActiveSsRsvd", "FittedActiveSs"]


def mat_cov_centered(x: np.ndarray) -> np.ndarray:
    xc = x - x.mean(axis=0)                                   # center_mat_col
    return (xc.T @ xc) * (1.0 / (x.shape[0] - 1.0))           # stats_corr.rs:34-42


def pearson_corr(x: np.ndarray) -> np.ndarray:
    mu = x.mean(axis=0)
    sd = np.sqrt(((x - mu) ** 2).sum(axis=0) / (x.shape[0] - 1.0))     # mat_std axis 1
    z = (x - mu) / sd                                                   # zcenter_mat_col
    return (z.T @ z) * (1.0 / (x.shape[0] - 1.0))                       # stats_corr.rs:18-27


def _col_interactions(x: np.ndarray) -> np.ndarray:
    cols = [x[:, a] * x[:, b] for a in range(x.shape[1]) for b in range(a, x.shape[1])]     # self interactions included
    return np.stack(cols, axis=1)


def _vandermonde(x: np.ndarray) -> np.ndarray:
    return np.hstack([x, _col_interactions(x), np.ones((x.shape[0], 1))])                    # stats_corr.rs:201-209


class PolyGradientEstimator:
    def __init__(self, x_mat, y, est_order: int, n_nbrs: int):
        self.x, self.y = np.asarray(x_mat, dtype=np.float64), np.asarray(y, dtype=np.float64).reshape(-1, 1)
        self.est_order, self.n_nbrs, self.k = est_order, n_nbrs, self.x.shape[1]

    def _nearest(self, x0):
        d2 = ((self.x - np.asarray(x0)[None, :]) ** 2).sum(axis=1)
        idx = np.argsort(d2, kind="stable")[:self.n_nbrs]
        return self.x[idx], self.y[idx]

    def grad_at(self, x0):
        xn, yn = self._nearest(x0)
        if self.est_order == 1:                                                              # :112-121
            coeffs = mat_pinv(np.hstack([xn, np.ones((xn.shape[0], 1))])) @ yn
            return coeffs[:self.k].T
        if self.est_order != 2:
            raise NotImplementedError(self.est_order)
        coeffs = mat_pinv(_vandermonde(xn)) @ yn                                             # quad_fit
        x0 = np.asarray(x0, dtype=np.float64).reshape(1, -1)
        eps = 1.0e-10                                                                        # jac_from_quad
        y0 = _vandermonde(x0) @ coeffs
        out = np.zeros((1, self.k))
        for j in range(self.k):
            xp = x0.copy(); xp[0, j] += eps
            out[0, j] = ((_vandermonde(xp) @ coeffs - y0) / eps)[0, 0]
        return out


class FittedActiveSs:
    def __init__(self, components, singular_vals, n_comps):
        self.components_, self.singular_vals_, self.n_comps = components, singular_vals, n_comps

    def var_diag_evd_sensi(self):
        m = self.components_.T @ self.singular_vals_ @ self.components_
        return [float(m[i, i]) for i in range(self.singular_vals_.shape[0])]

    def components(self):
        return self.components_[:, :self.n_comps]

    def singular_vals(self):
        return self.singular_vals_[:, :self.n_comps]


class ActiveSsRsvd:
    def __init__(self, grad_est, n_comps: int):
        self.grad_est, self.n_comps = grad_est, n_comps

    def create_grad_mat(self, x_mat):
        x_mat = np.asarray(x_mat, dtype=np.float64)
        g = np.zeros((x_mat.shape[1], x_mat.shape[0]))
        for i in range(x_mat.shape[0]):
            g[:, i] = self.grad_est.grad_at(x_mat[i]).ravel()
        return g

    def fit_gradients(self, g):
        c = (g @ g.T) * (1.0 / g.shape[1])                                                   # :253
        w, v = np.linalg.eigh(c)
        order = np.argsort(-w, kind="stable")                                                # sort_evd
        return FittedActiveSs(v[:, order], np.diag(w[order]), self.n_comps)

    def fit_svd_gradients(self, g, n_iter=8, n_oversamples=10, omega=None):
        gs = g * (1.0 / np.sqrt(g.shape[1]))                                                 # :236
        ur, sr, _ = ref_rsvd.random_svd(gs, min(g.shape[0], self.n_comps), n_iter, n_oversamples, omega=omega)
        return FittedActiveSs(ur, np.diag(sr.ravel()), self.n_comps)

    def fit(self, x_mat):
        return self.fit_gradients(self.create_grad_mat(x_mat))
