"""GPU parity of the single-kernel path for tiny matrices (csrc/fused_small.cu): the whole reference schedule
(random_svd.rs:15-110) in one CTA.  Checked against the oracle on the same A and Omega at the north-star tolerances, and
against the multi-kernel engine on the same Philox seed."""
import os

import numpy as np
import pytest

from oracle import ref_rsvd

pytestmark = pytest.mark.gpu
TOL_SIGMA, TOL_ANGLE = 1e-10, 1e-8


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


def check(out, ref, k):
    u, s, vt = (np.asarray(x) for x in out)
    u0, s0, vt0 = ref
    assert u.shape == u0.shape and s.shape == (k, 1) and vt.shape == vt0.shape
    assert ref_rsvd.sigma_rel_err(s0, s) < TOL_SIGMA
    assert ref_rsvd.subspace_sine(u0, u) < TOL_ANGLE
    assert ref_rsvd.subspace_sine(vt0.T, vt.T) < TOL_ANGLE
    assert np.max(np.abs(u.T @ u - np.eye(k))) < 1e-12
    assert np.max(np.abs(vt @ vt.T - np.eye(k))) < 1e-12


CASES = [
    # m, n, (k, q, p), layout
    (100, 100, (10, 12, 8), "C"),       # BASELINE config C1 (README example)
    (100, 100, (10, 12, 8), "F"),
    (60, 200, (7, 5, 6), "C"),          # fat: transposed view, roles of U and V swap
    (203, 37, (12, 4, 10), "C"),        # ragged, l = 22
    (100, 64, (32, 3, 0), "C"),         # l = k = 32: the widest sketch the path takes
    (40, 9, (4, 20, 10), "C"),          # l clamps to n = 9 (PCA-like q = 20)
    (3, 2, (1, 2, 1), "C"),
    (129, 65, (5, 0, 3), "F"),          # no power iteration at all
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}_{c[3]}_k{c[2][0]}q{c[2][1]}p{c[2][2]}" for c in CASES])
def test_fused_small_parity_host_and_device(cb, case):
    import torch
    m, n, (k, q, p), order = case
    rng = np.random.default_rng(m * 1000 + n)
    a = rng.standard_normal((m, n))
    if order == "F":
        a = np.asfortranarray(a)
    l = min(k + p, min(m, n))
    omega = rng.standard_normal((min(m, n), l))
    ref = ref_rsvd.random_svd(np.ascontiguousarray(a), k, q, p, omega=omega)
    out = cb.rsvd(a, k, q, p, omega=omega)
    t = cb.last_timings()
    assert t["fused_small"] == 1 and t["gpu_launches"] == 1, t
    check(out, ref, k)
    ad = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if order == "F":
        ad = ad.t().contiguous().t()
    outd = cb.rsvd(ad, k, q, p, omega=torch.from_numpy(omega).cuda())
    torch.cuda.synchronize()
    assert cb.last_timings()["fused_small"] == 1
    check(tuple(x.cpu().numpy() for x in outd), ref, k)
    for x, y in zip(out, outd):
        assert np.array_equal(np.asarray(x), y.cpu().numpy())          # host and device entry: the same kernel, same bits


def test_fused_small_matches_the_multi_kernel_engine_on_the_same_seed(cb):
    """Same Philox Omega (element (i, j) is draw i*l + j in both paths), two different QR algorithms (Householder here,
    CholeskyQR2 there): same sigma and subspaces."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal((100, 100))
    k, q, p = 10, 12, 8
    fused = cb.rsvd(a, k, q, p, seed=42)
    assert cb.last_timings()["fused_small"] == 1
    os.environ["CORRLA_B200_NO_FUSED_SMALL"] = "1"
    try:
        general = cb.rsvd(a, k, q, p, seed=42)
        assert cb.last_timings()["fused_small"] == 0 and cb.last_timings()["gpu_launches"] > 20
    finally:
        os.environ.pop("CORRLA_B200_NO_FUSED_SMALL", None)
    assert ref_rsvd.sigma_rel_err(general[1], fused[1]) < 1e-11
    assert ref_rsvd.subspace_sine(np.asarray(general[0]), np.asarray(fused[0])) < 1e-9
    assert ref_rsvd.subspace_sine(np.asarray(general[2]).T, np.asarray(fused[2]).T) < 1e-9
    omega = ref_rsvd.philox_normal(100, 18, 42)
    check(fused, ref_rsvd.random_svd(a, k, q, p, omega=omega), k)


def test_fused_small_stabilised_schedule_and_power_iter(cb):
    rng = np.random.default_rng(6)
    a = rng.standard_normal((150, 60))
    omega = rng.standard_normal((60, 20))
    ref = ref_rsvd.random_svd(a, 12, 5, 8, omega=omega)
    out = cb.rsvd(a, 12, 5, 8, omega=omega, schedule="stabilised")
    assert cb.last_timings()["fused_small"] == 1
    check(out, ref, 12)                                       # same subspace in exact arithmetic; benign input
    q = cb.power_iter(a, 20, 5, omega=omega)
    assert cb.last_timings()["fused_small"] == 1
    q0 = ref_rsvd.power_iter(a, 20, 5, omega=omega)
    assert q.shape == (150, 20)
    assert np.max(np.abs(q.T @ q - np.eye(20))) < 1e-13
    assert ref_rsvd.subspace_sine(q0, np.asarray(q)) < TOL_ANGLE


def test_fused_small_rank_deficient_keeps_orthonormal_factors(cb):
    """Rank 3 in a 40 x 30 matrix with l = 12: Y is rank deficient at every QR.  Householder completes the basis; the
    singular values beyond the rank are at rounding level and U, V stay orthonormal."""
    rng = np.random.default_rng(7)
    a = rng.standard_normal((40, 3)) @ rng.standard_normal((3, 30))
    u, s, vt = cb.rsvd(a, 6, 6, 6, seed=3)
    assert cb.last_timings()["fused_small"] in (0, 1)         # an exactly zero sigma hands over to the general path
    s_true = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s.ravel()[:3] - s_true[:3]) / s_true[:3]) < 1e-12
    assert np.max(np.abs(s.ravel()[3:])) < 1e-12 * s_true[0]
    assert np.max(np.abs(u.T @ u - np.eye(6))) < 1e-12 and np.max(np.abs(vt @ vt.T - np.eye(6))) < 1e-12
    assert np.max(np.abs((u * s.ravel()) @ vt - a)) < 1e-12 * s_true[0]


def test_fused_small_is_not_taken_outside_its_limits(cb):
    rng = np.random.default_rng(8)
    a = rng.standard_normal((300, 300))                       # 90 000 elements: over the shared-memory budget
    cb.rsvd(a, 10, 2, 8, seed=1)
    assert cb.last_timings()["fused_small"] == 0
    b = rng.standard_normal((200, 60))
    cb.rsvd(b, 30, 2, 10, seed=1)                             # l = 40 > 32
    assert cb.last_timings()["fused_small"] == 0
    cb.rsvd(b, 10, 2, 8, seed=1, center=True)                 # centring: the general path owns the rank-1 corrections
    assert cb.last_timings()["fused_small"] == 0


def test_fused_small_choleskyqr2_and_householder_paths_agree(cb):
    """Inside the fused kernel thin-Q is CholeskyQR2 when its Cholesky probe passes and Householder otherwise
    (CORRLA_B200_FUSED_NO_CHOL=1 forces Householder): same sigma and subspaces on a benign input."""
    rng = np.random.default_rng(9)
    a = rng.standard_normal((100, 100))
    omega = rng.standard_normal((100, 18))
    ref = ref_rsvd.random_svd(a, 10, 12, 8, omega=omega)
    fast = cb.rsvd(a, 10, 12, 8, omega=omega)
    os.environ["CORRLA_B200_FUSED_NO_CHOL"] = "1"
    try:
        house = cb.rsvd(a, 10, 12, 8, omega=omega)
        assert cb.last_timings()["fused_small"] == 1
    finally:
        os.environ.pop("CORRLA_B200_FUSED_NO_CHOL", None)
    check(fast, ref, 10)
    check(house, ref, 10)
    assert ref_rsvd.sigma_rel_err(house[1], fast[1]) < 1e-12


def test_fused_small_ill_conditioned_falls_back_to_householder_inside_the_kernel(cb):
    """sigma_j = 2^-j: after three raw power iterations cond(Y) is ~1e16 at the first QR, the Cholesky probe fails and
    the kernel's Householder path must take over.  Leading singular values against the exact ones (the reference
    schedule itself is not reproducible to 1e-10 on such spectra)."""
    rng = np.random.default_rng(10)
    m, n = 160, 48
    u0, _ = np.linalg.qr(rng.standard_normal((m, n)))
    v0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = 2.0 ** -np.arange(n)
    a = (u0 * sig) @ v0.T
    u, s, vt = cb.rsvd(a, 8, 6, 8, seed=5)
    assert cb.last_timings()["fused_small"] == 1
    assert np.max(np.abs(s.ravel() - sig[:8]) / sig[:8]) < 1e-9
    assert np.max(np.abs(u.T @ u - np.eye(8))) < 1e-12 and np.max(np.abs(vt @ vt.T - np.eye(8))) < 1e-12
    assert ref_rsvd.subspace_sine(u0[:, :4], np.asarray(u)[:, :4]) < 1e-7
