"""CPU oracle for the RSVD hot path of wgurecky/CORRLA_RS  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm, step for step.  It is the
checker (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg);
the product path (corrla_rs_b200/) never imports it and has no CPU fallback.

What it follows (paths relative to the reference checkout):
  * src/lib_math_utils/random_svd.rs:15-59   power_iter  (QR only when i > 2, Frobenius scaling each trip)
  * src/lib_math_utils/random_svd.rs:63-110  random_svd  (fat -> transposed view, l clamp, output slicing)
  * src/lib_math_utils/mat_utils.rs:20-33    par_matmul_helper (alpha=None => res = beta*lhs*rhs, overwrite)
  * src/lib_math_utils/mat_utils.rs:161-175  random_mat_normal (n x l i.i.d. N(0,1); unseeded in the
                                             reference, so the oracle takes Omega as an argument)

The arithmetic the reference delegates to the un-vendored crate faer 0.19.x (Cargo.toml:26:
GEMM, blocked Householder QR + compute_thin_q, SVD, norm_l2) is restated with LAPACK through
numpy: numpy.linalg.qr(mode="reduced") and numpy.linalg.svd(full_matrices=False).  faer's SVD
is a full SVD (random_svd.rs:89) but only the leading k vectors are used (:98-107), which a
thin SVD reproduces exactly.

Pinning status: the reference cannot be compiled here (no Rust toolchain, crates un-vendored).
The oracle is pinned against the only known-answer vector the reference's tests hold for this
path, test_rsvd_lowrank (random_svd.rs:153-196, sigma = {3, 2.2360679, 2, 0, 0}, abs tol 1e-3),
and cross-checked against the reference's own numpy statement of the algorithm
(examples/benchmark_rsvd.py:16-54).  At the 1e-10 / 1e-8 tolerance of the GPU parity tests the
reference's tests pin nothing: "parity unpinned" beyond that vector.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "par_matmul_helper", "power_iter", "random_svd", "rsvd",
    "philox4x32_10", "philox_normal", "subspace_sine", "sigma_rel_err",
]


# --------------------------------------------------------------------------------------
# mat_utils.rs:20-33
# --------------------------------------------------------------------------------------
def par_matmul_helper(lhs: np.ndarray, rhs: np.ndarray, beta: float = 1.0) -> np.ndarray:
    """res = beta * lhs @ rhs (alpha=None overwrites the destination; n_threads is ignored by
    the reference, mat_utils.rs:29-31)."""
    out = lhs @ rhs
    if beta != 1.0:
        out *= beta
    return out


# --------------------------------------------------------------------------------------
# random_svd.rs:15-59
# --------------------------------------------------------------------------------------
def power_iter(a: np.ndarray, omega_rank: int, n_iter: int, omega: np.ndarray | None = None,
               rng: np.random.Generator | None = None) -> np.ndarray:
    """Range finder + power iteration.  `a` is the thin view (nrows >= ncols is NOT checked here,
    as in the reference).  `omega` (ncols x omega_rank) replaces random_mat_normal (:24)."""
    a_ncols = a.shape[1]
    if omega is None:
        rng = rng or np.random.default_rng()
        omega = rng.standard_normal((a_ncols, omega_rank))
    assert omega.shape == (a_ncols, omega_rank), (omega.shape, (a_ncols, omega_rank))
    y = a @ omega                                            # :31
    for i in range(n_iter):                                  # :35
        if i > 2:                                            # :37
            y, _ = np.linalg.qr(y, mode="reduced")           # :38
        o = par_matmul_helper(a.T, y, 1.0)                   # :42-46
        y = par_matmul_helper(a, o, 1.0)                     # :47-51
        y = y * (1.0 / np.linalg.norm(y))                    # :53-55 (norm_l2 == Frobenius)
    q, _ = np.linalg.qr(y, mode="reduced")                   # :57
    return q


# --------------------------------------------------------------------------------------
# random_svd.rs:63-110
# --------------------------------------------------------------------------------------
def random_svd(a: np.ndarray, omega_rank: int, n_iter: int, n_oversamples: int,
               omega: np.ndarray | None = None, rng: np.random.Generator | None = None):
    """Returns (U m x k, S k x 1, Vt k x n) exactly like the reference, including the
    fat-matrix role swap (:96-102).  Raises IndexError where the reference panics (k > l)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise TypeError("a must be 2-D")
    fat = a.shape[0] < a.shape[1]                            # :71
    aa = a.T if fat else a                                   # :73 (a view, no copy)
    l = min(omega_rank + n_oversamples, aa.shape[1])         # :77
    if omega_rank > l:                                       # out-of-range get at :98-107
        raise IndexError("n_rank exceeds min(n_rank + n_oversamples, ncols(thin a))")
    q = power_iter(aa, l, n_iter, omega=omega, rng=rng)      # :76-77
    b = q.T @ aa                                             # :80
    ub, s, vbt = np.linalg.svd(b, full_matrices=False)       # :89 (thin == leading part of full)
    u = q @ ub                                               # :92
    k = omega_rank
    if fat:                                                  # :96-102
        return vbt.T[:, :k].copy(), s[:k].reshape(k, 1).copy(), u.T[:k, :].copy()
    return u[:, :k].copy(), s[:k].reshape(k, 1).copy(), vbt[:k, :].copy()   # :103-109


def rsvd(a_mat, n_rank: int, n_iters: int, n_oversamples: int, omega=None):
    """The pyo3 binding's call shape (src/lib_math_utils_py.rs:21-36)."""
    return random_svd(a_mat, n_rank, n_iters, n_oversamples, omega=omega)


# --------------------------------------------------------------------------------------
# Omega generator of the B200 engine, restated (there is nothing to follow in the reference:
# its generator is unseeded thread_rng, mat_utils.rs:166-173).  Philox4x32-10 (Salmon et al.,
# SC'11) keyed by the seed; counter j yields the normals for flat elements 2j and 2j+1 of the
# row-major n x l matrix through Box-Muller on two 53-bit uniforms.
# --------------------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter_lo: np.ndarray, seed: int) -> np.ndarray:
    """Vectorised Philox4x32-10.  counter = (lo32(j), hi32(j), 0, 0), key = (lo32(seed), hi32(seed)).
    Returns uint32 array of shape (len, 4)."""
    j = np.asarray(counter_lo, dtype=np.uint64)
    c0 = (j & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (j >> np.uint64(32)).astype(np.uint32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return np.stack([c0, c1, c2, c3], axis=1)


def philox_normal(n_rows: int, n_cols: int, seed: int) -> np.ndarray:
    """n_rows x n_cols standard normals, element (i, j) <- flat index e = i*n_cols + j,
    pair p = e >> 1; Box-Muller: r = sqrt(-2 ln u1), z0 = r cos(2 pi u2), z1 = r sin(2 pi u2);
    u1 = ((x0 | x1<<32) >> 11 + 1) * 2^-53 in (0, 1], u2 = ((x2 | x3<<32) >> 11) * 2^-53 in [0, 1)."""
    total = n_rows * n_cols
    npairs = (total + 1) // 2
    w = philox4x32_10(np.arange(npairs, dtype=np.uint64), seed).astype(np.uint64)
    a = (w[:, 0] | (w[:, 1] << np.uint64(32))) >> np.uint64(11)
    b = (w[:, 2] | (w[:, 3] << np.uint64(32))) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 1.0) * 2.0 ** -53
    u2 = b.astype(np.float64) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * npairs, dtype=np.float64)
    z[0::2] = r * np.cos(2.0 * np.pi * u2)
    z[1::2] = r * np.sin(2.0 * np.pi * u2)
    return z[:total].reshape(n_rows, n_cols)


# --------------------------------------------------------------------------------------
# Comparison metrics
# --------------------------------------------------------------------------------------
def subspace_sine(u_ref: np.ndarray, u_hat: np.ndarray) -> float:
    """Sine of the largest principal angle between span(u_ref) and span(u_hat), both with
    orthonormal columns: || u_hat - u_ref (u_ref^T u_hat) ||_2."""
    r = u_hat - u_ref @ (u_ref.T @ u_hat)
    return float(np.linalg.norm(r, 2))


def sigma_rel_err(s_ref: np.ndarray, s_hat: np.ndarray) -> float:
    s_ref = np.asarray(s_ref, dtype=np.float64).ravel()
    s_hat = np.asarray(s_hat, dtype=np.float64).ravel()
    return float(np.max(np.abs(s_ref - s_hat) / np.abs(s_ref)))
