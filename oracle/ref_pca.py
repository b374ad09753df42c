"""CPU oracle for the first "next" row of SURVEY section 8(f): PCA by RSVD  --  TEST INFRASTRUCTURE ONLY.

Restates, with numpy, what the reference does around the hot path (paths relative to the reference checkout):
  * src/lib_math_utils/mat_utils.rs:87-119   mat_mean(axis = 1): column means, 1 x ncols
  * src/lib_math_utils/mat_utils.rs:482-502  center_mat_col: owned copy with the column means subtracted
  * src/lib_math_utils/pca_rsvd.rs:56-82     PcaRsvd::new: random_svd(centred x, rank, 20, min(n_dim, 10)); keeps S and Vt
  * src/lib_math_utils/pca_rsvd.rs:91-99     explained_var = S^2 / (n_samples - 1)
  * src/lib_math_utils_py.rs:38-55           pyo3 rpca(a_mat, n_rank, n_iters, n_oversamples) -> (singular_vals, components);
                                             n_iters and n_oversamples are accepted and IGNORED
Pinning: the reference's test_pca (pca_rsvd.rs:120-134) only prints; parity unpinned beyond oracle/ref_rsvd.py's vectors.
"""
from __future__ import annotations

import numpy as np

from . import ref_rsvd

PCA_N_ITER = 20            # pca_rsvd.rs:66


def mat_mean_cols(x: np.ndarray) -> np.ndarray:
    """mat_mean(x, 1): 1 x ncols, serial accumulation down each column (mat_utils.rs:107-117)."""
    return (np.add.reduce(x, axis=0) / x.shape[0]).reshape(1, -1)


def center_mat_col(x: np.ndarray) -> np.ndarray:
    return x - mat_mean_cols(x)


def pca_rsvd_new(x: np.ndarray, rank: int, omega: np.ndarray | None = None, rng=None):
    """PcaRsvd::new -> dict(means 1 x n_dim, singular_values k x 1, components k x n_dim, n_samples)."""
    x = np.asarray(x, dtype=np.float64)
    means = mat_mean_cols(x)
    n_samples, n_dim = x.shape
    cx = center_mat_col(x)
    _u, s, vt = ref_rsvd.random_svd(cx, rank, PCA_N_ITER, min(n_dim, 10), omega=omega, rng=rng)
    return {"means": means, "singular_values": s, "components": vt, "n_samples": n_samples}


def explained_var(pca: dict) -> np.ndarray:
    return pca["singular_values"] ** 2 / (pca["n_samples"] - 1.0)


def rpca(a_mat, n_rank: int, n_iters: int = 0, n_oversamples: int = 0, omega=None):
    """The pyo3 call shape (lib_math_utils_py.rs:38-55): returns (singular_vals k x 1, components k x n_dim)."""
    p = pca_rsvd_new(a_mat, n_rank, omega=omega)
    return p["singular_values"], p["components"]
