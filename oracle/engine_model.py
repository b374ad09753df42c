"""CPU model of the B200 engine's *algorithm* (not of the reference) -- TEST INFRASTRUCTURE ONLY.

corrla_rs_b200/csrc/engine.cu reorganises random_svd.rs:15-110 so that it maps onto skinny GEMMs and
small replicated factors: sketch-preconditioned CholeskyQR (sparse sign sketch -> Householder QR of the small sketch ->
one CholeskyQR pass) with a deflated triangular inverse instead of Householder QR of the tall matrix,
R^-1 folded into the small side, the Frobenius scaling deferred into the next product, QR-preconditioned
one-sided Jacobi for the SVD of B, and row sharding with all-reduces of Z = A^T Y and of the Gram matrices.
This file restates that reorganisation in numpy so that

  * its equivalence with the reference restatement (oracle/ref_rsvd.py) can be checked on the CPU, and
  * the multi-rank data flow can be exercised with torch.distributed/gloo (tests/test_sharded_gloo.py)
    by plugging an all-reduce callable in.

Only tests/ import this module.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps
TOL_DEAD_PER_COL = 8.0 * EPS
TAU_SHIFT = 1e-10


def _identity_allreduce(x: np.ndarray) -> np.ndarray:
    return x


def chol_factor(g: np.ndarray, d0: np.ndarray, shifted: bool, tol_dead: float):
    """Upper Cholesky with per-column deflation; mirrors chol_factor() in small_kernels.cu."""
    l = g.shape[0]
    s = np.array(g, dtype=np.float64, copy=True)
    dead = np.zeros(l, dtype=bool)
    minratio = np.inf
    for j in range(l):
        piv, dj = s[j, j], d0[j]
        if shifted:
            is_dead = (not dj > 0.0) or (not piv > 0.0)
        else:
            is_dead = (not dj > 0.0) or (not piv > tol_dead * dj)
        ratio = piv / dj if (dj > 0.0 and piv > 0.0) else 0.0
        minratio = min(minratio, ratio)
        dead[j] = is_dead
        if is_dead:
            s[j, j:] = 0.0
            continue
        rjj = np.sqrt(piv)
        s[j, j + 1:] /= rjj
        s[j, j] = rjj
        r = s[j, j + 1:]
        s[j + 1:, j + 1:] -= np.outer(r, r)
    return np.triu(s), dead, minratio


def chol_inv(g: np.ndarray, mode_auto: bool, global_rows: float):
    """Returns (T, shifted, live): T = deflated inverse of the Cholesky factor of g."""
    l = g.shape[0]
    d0 = np.diag(g).copy()
    tol_dead = TOL_DEAD_PER_COL * l
    r, dead, minratio = chol_factor(g, d0, False, tol_dead)
    shifted = False
    if mode_auto and minratio < TAU_SHIFT:
        trace = float(np.sum(d0[d0 > 0.0]))
        shift = 11.0 * (global_rows * l + l * (l + 1.0)) * (0.5 * EPS) * trace
        gs = np.array(g, copy=True)
        idx = np.arange(l)
        pos = d0 > 0.0
        gs[idx[pos], idx[pos]] += shift
        r, dead, minratio = chol_factor(gs, d0, True, tol_dead)
        shifted = True
    rt = r.copy()
    rt[dead, dead] = 1.0
    t = np.linalg.solve(rt, np.eye(l)) if l else rt
    t = np.triu(t)
    t[:, dead] = 0.0
    return t, shifted, dead


SKETCH_ZETA = 8
SKETCH_TILE = 64
SKETCH_MUL = (1, 17, 19, 23, 29, 31, 37, 41)


def sketch_rows(lc: int) -> int:
    s = max(2 * lc, 64)
    cap = (220 * 1024) // (lc * 8)
    if s > cap:
        s = cap // 16 * 16
    return s


def sparse_sign_sketch(x: np.ndarray, s: int, rng: np.random.Generator) -> np.ndarray:
    """Structured sparse sign embedding of sketch_kernel (small_kernels.cu): rows in tiles of 64; in a tile, hash t sends
    row j to bucket (a_t * j + offset) mod s (an injection) with a random sign.  The hash bits differ from the CUDA
    kernel's (splitmix64 there), the structure and the statistics are the same."""
    m, l = x.shape
    out = np.zeros((s, l))
    for b in range((m + SKETCH_TILE - 1) // SKETCH_TILE):
        rows = np.arange(b * SKETCH_TILE, min(m, (b + 1) * SKETCH_TILE))
        j = rows - b * SKETCH_TILE
        for t in range(SKETCH_ZETA):
            bucket = (SKETCH_MUL[t] * j + int(rng.integers(0, s))) % s
            sign = rng.integers(0, 2, size=len(rows)) * 2.0 - 1.0
            out[bucket] += sign[:, None] * x[rows]          # buckets are distinct within one (tile, t) step
    return out


def hqr_inv(sk: np.ndarray):
    """Householder R of the sketch with a column-relative rank test, then the deflated inverse (hqr_inv_kernel)."""
    l = sk.shape[1]
    r0 = np.linalg.qr(sk, mode="r") if sk.shape[0] >= l else np.linalg.qr(np.vstack([sk, np.zeros((l - sk.shape[0], l))]), mode="r")
    cn = np.linalg.norm(sk, axis=0)
    dead = ~(np.abs(np.diag(r0)) > TOL_DEAD_PER_COL * l * cn) | ~(cn > 0.0)
    live = np.flatnonzero(~dead)
    t = np.zeros((l, l))
    if live.size:
        # staircase: the live columns restricted to their own rows form an upper-triangular block
        rl = np.linalg.qr(sk[:, live], mode="r")
        t[np.ix_(live, live)] = np.triu(np.linalg.solve(rl, np.eye(live.size)))
    return t, dead


N_REFILL = [0]
N_ROBUST = [0]


def qr_fold(x: np.ndarray, distributed_allreduce, global_rows: float, refill_rng=None, sketch_rng=None,
            refill_a: np.ndarray | None = None, complete: bool = False, basis_only: bool = False):
    """CholeskyQR2 when a Cholesky probe says cond(x) is below ~1e4, else sketch-preconditioned CholeskyQR with refill
    of numerically dependent columns (Core::qr_inplace in engine.cu).
    Returns (x_last, t_fold, second_pass, live): the orthonormal factor is x_last @ t_fold."""
    # probe: Gram + Cholesky pivots; well-conditioned matrices take plain CholeskyQR2
    g = distributed_allreduce(x.T @ x)
    _, _, probe = chol_factor(g, np.diag(g).copy(), False, TOL_DEAD_PER_COL * x.shape[1])
    if probe >= 1e-8:
        t1, _, _ = chol_inv(g, False, global_rows)
        if basis_only:
            # in-loop QR: a basis of range(x) of condition 1 + cond(x)^2 eps is all the next product needs -- one pass
            return x, t1, False, x.shape[1]
        x = x @ t1
        g = distributed_allreduce(x.T @ x)
        e = g - np.eye(g.shape[0])
        if np.sum(e * e) <= 4e-16:
            tf = np.eye(g.shape[0]) - 0.5 * e          # first-order (Loewdin) factor instead of the second Cholesky
        else:
            tf, _, _ = chol_inv(g, False, global_rows)
        return x, tf, False, x.shape[1]
    N_ROBUST[0] += 1                                    # Core::n_robust
    lc = (x.shape[1] + 7) // 8 * 8
    sk = distributed_allreduce(sparse_sign_sketch(x, sketch_rows(lc), sketch_rng or np.random.default_rng(777)))
    t1, dead = hqr_inv(sk)
    x = x @ t1
    g = distributed_allreduce(x.T @ x)
    d0 = np.diag(g).copy()
    _, dead_c, minratio = chol_factor(g, d0, False, TOL_DEAD_PER_COL * x.shape[1])
    tf, _, dead2 = chol_inv(g, False, global_rows)
    second = minratio < 1e-3
    if second:
        x = x @ tf
        g = distributed_allreduce(x.T @ x)
        tf, _, dead2 = chol_inv(g, False, global_rows)
    dead = dead | dead2
    if dead.any():
        N_REFILL[0] += 1                                # Core::n_refill: Wide::block_qr re-projects a refilled panel
        rng = refill_rng or np.random.default_rng(12345)
        x = x @ tf
        nd = int(dead.sum())
        if refill_a is not None:
            # fresh directions from range(A): the same Omega' on every rank (engine: Philox stream 0)
            x[:, dead] = refill_a @ rng.standard_normal((refill_a.shape[1], nd))
        else:
            x[:, dead] = rng.standard_normal((x.shape[0], nd))
        # the refill stage is the robust (sketch) stage again
        lc = (x.shape[1] + 7) // 8 * 8
        sk = distributed_allreduce(sparse_sign_sketch(x, sketch_rows(lc), np.random.default_rng(778)))
        t1, dead_r = hqr_inv(sk)
        x = x @ t1
        g = distributed_allreduce(x.T @ x)
        tf, _, dead = chol_inv(g, False, global_rows)
        if refill_a is not None and complete and dead_r.any():
            # range(A) has fewer than l dimensions: arbitrary vectors in the columns that are still dead, made
            # orthogonal to the live ones (Householder's completion; the final Q keeps l orthonormal columns)
            x = x @ tf
            x[:, dead_r] = rng.standard_normal((x.shape[0], int(dead_r.sum())))
            sk = distributed_allreduce(sparse_sign_sketch(x, sketch_rows(lc), np.random.default_rng(779)))
            t1, _ = hqr_inv(sk)
            x = x @ t1
            g = distributed_allreduce(x.T @ x)
            tf, _, dead = chol_inv(g, False, global_rows)
    return x, tf, second, int(np.sum(~dead))


def jacobi_svd(w: np.ndarray, max_sweeps: int = 60):
    """One-sided Hestenes Jacobi with the round-robin ordering of jacobi_svd_kernel.
    Returns (ur, sigma, vr) with w = ur diag(sigma) vr^T, sigma descending."""
    l = w.shape[0]
    wc = np.array(w, dtype=np.float64, copy=True)
    vc = np.eye(l)
    h = (l + 1) // 2
    top = list(range(h))
    bot = list(range(h, 2 * h))
    tol = np.sqrt(l) * EPS
    for _ in range(max_sweeps):
        rotations = 0
        for _step in range(max(2 * h - 1, 1)):
            for p, q in zip(top, bot):
                if p > q:
                    p, q = q, p
                if q >= l:
                    continue
                x, y = wc[:, p], wc[:, q]
                a, b, c = x @ x, y @ y, x @ y
                if c != 0.0 and abs(c) > tol * np.sqrt(a) * np.sqrt(b):
                    zeta = (b - a) / (2.0 * c)
                    t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                    cs = 1.0 / np.sqrt(1.0 + t * t)
                    sn = cs * t
                    wc[:, p], wc[:, q] = cs * x - sn * y, sn * x + cs * y
                    vx, vy = vc[:, p].copy(), vc[:, q].copy()
                    vc[:, p], vc[:, q] = cs * vx - sn * vy, sn * vx + cs * vy
                    rotations += 1
            if h > 1:
                t_last = top[h - 1]
                top = [top[0], bot[0]] + top[1:h - 1]
                bot = bot[1:] + [t_last]
        if rotations == 0:
            break
    sig = np.sqrt(np.sum(wc * wc, axis=0))
    order = np.argsort(-sig, kind="stable")
    ur = np.zeros_like(wc)
    nz = sig > 0.0
    ur[:, nz] = wc[:, nz] / sig[nz]
    return ur[:, order], sig[order], vc[:, order]


def engine_rsvd(a_local: np.ndarray, n_rank: int, n_iter: int, n_oversamples: int, omega: np.ndarray,
                allreduce=_identity_allreduce, global_rows: float | None = None, schedule: int = 0,
                use_numpy_svd_for_core: bool = False):
    """The engine's algorithm on one row shard of a THIN matrix.  `allreduce(x)` must return the
    element-wise sum of x over all ranks.  Returns (u_local m x k, s k x 1, vt k x n)."""
    a = np.asarray(a_local, dtype=np.float64)
    m, n = a.shape
    l = min(n_rank + n_oversamples, n)
    if n_rank > l:
        raise IndexError("n_rank exceeds l")
    grows = float(m if global_rows is None else global_rows)
    y = a @ omega
    nu2 = float(allreduce(np.array([np.sum(y * y)]))[0])
    for i in range(n_iter):
        if schedule == 1 or i > 2:
            y, tf, _, _ = qr_fold(y, allreduce, grows, refill_a=a, basis_only=True)
            z = allreduce(a.T @ y) @ tf
            y = a @ z
        else:
            z = allreduce(a.T @ y)
            y = (a @ z) * (1.0 / np.sqrt(nu2))
        nu2 = float(allreduce(np.array([np.sum(y * y)]))[0])
    y, tf, _, _ = qr_fold(y, allreduce, grows, refill_a=a, complete=True)
    zb = allreduce(a.T @ y) @ tf                      # B^T, replicated
    qz, tzf, _, _ = qr_fold(zb.copy(), _identity_allreduce, float(n))
    qz = qz @ tzf
    w = qz.T @ zb
    if use_numpy_svd_for_core:
        ur, sig, vrt = np.linalg.svd(w)
        vr = vrt.T
    else:
        ur, sig, vr = jacobi_svd(w)
    k = n_rank
    u = y @ (tf @ vr[:, :k])
    v = qz @ ur[:, :k]
    return u, sig[:k].reshape(k, 1).copy(), v.T.copy()


# --------------------------------------------------------------------------------------
# row-sharded data flow of the consumers in csrc/rom.cu (for the gloo tests)
# --------------------------------------------------------------------------------------
def dmdc_sharded(x_local: np.ndarray, u: np.ndarray, n_modes: int, n_iter: int, omegas, rank: int, nranks: int,
                 allreduce=_identity_allreduce, n_x_global: float | None = None):
    """corrla_dmdc_f64 with a communicator: every rank holds a block of state rows, the control rows are stacked under the
    LAST rank's block; the r x r cross products are all-reduced, u_til_2 travels as an all-reduce of a buffer that is zero
    everywhere but on the last rank.  Returns (a_til replicated, b local rows, modes_scale local rows, s_til, u_hat local)."""
    n_x, n_u = x_local.shape[0], u.shape[0]
    nxg = float(n_x if n_x_global is None else n_x_global)
    last = rank == nranks - 1
    stack = np.vstack([x_local, u]) if last else x_local
    xv, yv = stack[:, :-1], x_local[:, 1:]
    r = n_modes
    u_til, s_til, vt_til = engine_rsvd(xv, r, n_iter, 12, omegas[0], allreduce=allreduce, global_rows=nxg + n_u)
    u_hat, _s, _vt = engine_rsvd(yv, r, n_iter, 12, omegas[1], allreduce=allreduce, global_rows=nxg)
    s = s_til.ravel()
    s_inv = np.where(np.abs(s) < 1e-20, 0.0, 1.0 / (s + 1e-20))
    vs = vt_til.T * s_inv
    u1 = u_til[:n_x]
    p1 = yv @ vs
    tmp = allreduce(u_hat.T @ p1)
    c1 = allreduce(u1.T @ u_hat)
    a_til = tmp @ c1
    u2t = allreduce(u_til[n_x:].T.copy() if last else np.zeros((r, n_u)))        # broadcast from the last rank
    b = (u_hat @ tmp) @ u2t
    modes_scale = yv @ (vs @ c1)
    return a_til, b, modes_scale, s_til, u_hat


def pod_sharded(x_local_cols: np.ndarray, n_modes: int, omega, allreduce=_identity_allreduce,
                n_points_global: float | None = None):
    """corrla_pod_f64 with a communicator: every rank holds a block of POINTS (columns of x); the thin matrix of the
    RSVD is x_local^T, its left vectors are the local modes; weights = sum over ranks of x_local * modes_local."""
    thin = x_local_cols.T
    u, s, _vt = engine_rsvd(thin, n_modes, 10, 10, omega, allreduce=allreduce,
                            global_rows=float(thin.shape[0] if n_points_global is None else n_points_global))
    weights = allreduce(x_local_cols @ u)
    return u, weights, s


# --------------------------------------------------------------------------------------
# sketches wider than one 128-column panel (csrc/wide.cuh)
# --------------------------------------------------------------------------------------
def wide_plan(l: int):
    """Panel geometry of Wide::plan: P panels of equal padded width w (multiple of 8, <= 128)."""
    lc = (l + 7) // 8 * 8
    p = (lc + 127) // 128
    w = ((l + p - 1) // p + 7) // 8 * 8
    return p, w


def block_qr(x: np.ndarray, w: int, allreduce, global_rows: float, refill_a=None, complete=False, basis_only=False):
    """Wide::block_qr: block classical Gram-Schmidt with two projection sweeps against the finished panels, then the
    adaptive CholeskyQR of the single-panel path (qr_fold) on the panel; Q is formed explicitly.  basis_only (in-loop
    QR): one sweep and one Cholesky pass -- a basis of condition ~1 of the same range."""
    q = np.array(x, dtype=np.float64, copy=True)
    l = q.shape[1]
    for j0 in range(0, l, w):
        j1 = min(l, j0 + w)
        for _attempt in range(3):
            for _rep in range(1 if basis_only else 2):
                for i0 in range(0, j0, w):
                    qi = q[:, i0:i0 + w]
                    q[:, j0:j1] -= qi @ allreduce(qi.T @ q[:, j0:j1])
            refills, robusts = N_REFILL[0], N_ROBUST[0]
            xj, tf, _, _ = qr_fold(q[:, j0:j1].copy(), allreduce, global_rows, refill_a=refill_a, complete=complete,
                                   basis_only=basis_only)
            q[:, j0:j1] = xj @ tf
            # Columns refilled inside the panel QR are not orthogonal to the earlier panels yet; and a panel that needed the
            # robust stage may hold columns that collapsed in the projection (what is left of them is rounding noise, which
            # the normalisation blows up with its components along the earlier panels): project and factor again.
            if (N_REFILL[0] == refills and N_ROBUST[0] == robusts) or j0 == 0:
                break
    return q


def wide_rsvd(a_local: np.ndarray, n_rank: int, n_iter: int, n_oversamples: int, omega: np.ndarray,
              allreduce=_identity_allreduce, global_rows: float | None = None, schedule: int = 0):
    """The panel path of the engine (l > 128) on one row shard of a thin matrix: same outputs as engine_rsvd."""
    a = np.asarray(a_local, dtype=np.float64)
    m, n = a.shape
    l = min(n_rank + n_oversamples, n)
    _p, w = wide_plan(l)
    grows = float(m if global_rows is None else global_rows)
    y = a @ omega
    nu2 = float(allreduce(np.array([np.sum(y * y)]))[0])
    for i in range(n_iter):
        do_qr = schedule == 1 or i > 2
        if do_qr:
            y = block_qr(y, w, allreduce, grows, refill_a=a, basis_only=True)
        z = allreduce(a.T @ y)
        y = a @ z if do_qr else (a @ z) * (1.0 / np.sqrt(nu2))
        nu2 = float(allreduce(np.array([np.sum(y * y)]))[0])
    q = block_qr(y, w, allreduce, grows, refill_a=a, complete=True)
    zb = allreduce(a.T @ q)
    qz = block_qr(zb, w, _identity_allreduce, float(n))
    ur, sig, vr = jacobi_svd(qz.T @ zb)
    k = n_rank
    return q @ vr[:, :k], sig[:k].reshape(k, 1).copy(), (qz @ ur[:, :k]).T.copy()
