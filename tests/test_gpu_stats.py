"""GPU parity (-m gpu) of the Gram-kernel statistics and the active-subspace fit (corrla_cov_f64)
against oracle/ref_stats.py."""
import numpy as np
import pytest

from oracle import ref_rsvd, ref_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


@pytest.mark.parametrize("shape,order", [((10000, 5), "C"), ((4097, 64), "F"), ((50, 128), "C"), ((3, 2), "C")])
def test_cov_and_pearson_match_oracle(cb, shape, order):
    rng = np.random.default_rng(shape[0])
    mix = rng.standard_normal((shape[1], shape[1]))
    x = rng.standard_normal(shape) @ mix + 50.0 * rng.standard_normal((1, shape[1]))       # large means: centring matters
    x = np.asfortranarray(x) if order == "F" else x
    c0, p0 = ref_stats.mat_cov_centered(x), ref_stats.pearson_corr(x)
    c1, p1 = cb.mat_cov_centered(x), cb.pearson_corr(x)
    assert c1.shape == c0.shape and np.max(np.abs(c1 - c0)) < 1e-11 * np.max(np.abs(c0))
    assert np.max(np.abs(p1 - p0)) < 1e-11 and np.max(np.abs(np.diag(p1) - 1.0)) < 1e-14
    assert np.array_equal(c1, c1.T) and np.array_equal(p1, p1.T)
    out, mu, evals, evecs = cb.cov(x, "centered", evd=True)
    assert np.max(np.abs(mu.ravel() - x.mean(axis=0))) < 1e-12 * np.max(np.abs(x))
    w = np.linalg.eigvalsh(c0)[::-1]
    assert np.max(np.abs(evals.ravel() - w)) < 1e-11 * w[0]
    assert np.max(np.abs(evecs.T @ evecs - np.eye(shape[1]))) < 1e-12
    assert np.max(np.abs((evecs * evals.ravel()) @ evecs.T - c0)) < 1e-10 * w[0]


def test_reference_statistical_tests_and_device_input(cb):
    import torch
    x = cb.random_mat_normal(10000, 5, seed=3)                       # stats_corr.rs:259-298 with the engine's own generator
    assert np.max(np.abs(cb.pearson_corr(x) - np.eye(5))) < 1e-1
    assert np.max(np.abs(cb.mat_cov_centered(x) - np.eye(5))) < 1e-1
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    cd = cb.mat_cov_centered(xd)
    assert cd.is_cuda and np.max(np.abs(cd.cpu().numpy() - ref_stats.mat_cov_centered(x))) < 1e-12
    with pytest.raises(cb.CorrlaError):
        cb.pearson_corr(np.zeros((10, 200)))                         # more than 128 features


def test_active_subspace_fit_from_gradients(cb):
    """active_subspaces.rs:326-384 with the gradients from the oracle's estimator; `fit` (EVD of G G^T / N on the
    Gram + Jacobi kernels) and `fit_svd` (RSVD) against the oracle on the same gradient matrix."""
    rng = np.random.default_rng(1)
    cov3 = np.full((3, 3), 0.5) + 0.4 * np.eye(3)
    x3 = (cov3 @ rng.standard_normal((3, 100))).T
    y3 = 0.2 * x3[:, 0] + 0.5 * x3[:, 1] ** 2 + 0.10 * x3[:, 2] * x3[:, 0]
    est = ref_stats.PolyGradientEstimator(x3, y3, 2, 14)
    act = cb.ActiveSsRsvd(est, 2)
    fit = act.fit(x3)                                                # host loop over the estimator, device Gram + EVD
    assert abs(fit.components()[0, 0]) < abs(fit.components()[1, 0])
    assert fit.singular_vals()[0, 0] > fit.singular_vals()[1, 1]
    sens = fit.var_diag_evd_sensi()
    assert len(sens) == 3 and sens[1] > sens[0] and sens[1] > sens[2]
    tr = fit.transform(x3)
    assert tr.shape == (100, 2) and fit.inv_transform(tr).shape == (100, 3)
    g = act.create_grad_mat(x3)
    ref = ref_stats.ActiveSsRsvd(est, 2).fit_gradients(g)
    assert np.max(np.abs(np.diag(fit.singular_vals_) - np.diag(ref.singular_vals_))) < 1e-12 * ref.singular_vals_[0, 0]
    assert ref_rsvd.subspace_sine(ref.components(), fit.components()) < 1e-10
    # a larger gradient matrix (k = 64 features, 20 000 samples, dominant 8-dimensional active subspace)
    k, n = 64, 20000
    basis, _ = np.linalg.qr(rng.standard_normal((k, 8)))
    g = basis @ (rng.standard_normal((8, n)) * (3.0 * 0.7 ** np.arange(8))[:, None]) + 1e-3 * rng.standard_normal((k, n))
    ref = ref_stats.ActiveSsRsvd(None, 8).fit_gradients(g)
    fit = cb.ActiveSsRsvd(None, 8).fit_gradients(g)
    assert np.max(np.abs(np.diag(fit.singular_vals_) - np.diag(ref.singular_vals_))) < 1e-11 * ref.singular_vals_[0, 0]
    assert ref_rsvd.subspace_sine(ref.components(), fit.components()) < 1e-9
    omega = rng.standard_normal((k, 18))
    refs = ref_stats.ActiveSsRsvd(None, 8).fit_svd_gradients(g, omega=omega)
    fits = cb.ActiveSsRsvd(None, 8).fit_svd_gradients(g, omega=omega)
    assert ref_rsvd.sigma_rel_err(np.diag(refs.singular_vals_)[:, None], np.diag(fits.singular_vals_)[:, None]) < 1e-10
    assert ref_rsvd.subspace_sine(refs.components(), fits.components()) < 1e-8
    # the two routes agree: sigma(G / sqrt(N))^2 are the eigenvalues of G G^T / N
    assert np.allclose(np.diag(fits.singular_vals_)[:8] ** 2, np.diag(fit.singular_vals_)[:8], rtol=1e-8)


# ------------------------------------------------------------------ gradient matrix + active subspace from samples
def oracle_grad_mat(x, y, order, n_nbr, rows=None):
    est = ref_stats.PolyGradientEstimator(x, y, order, n_nbr)
    rows = range(x.shape[0]) if rows is None else rows
    return np.stack([est.grad_at(x[i]).ravel() for i in rows], axis=1)


def test_active_ss_readme_example(cb):
    """readme.md:100-107: x 1000 x 10, y 1000 x 1 Gaussian, linear fits through 30 neighbours, 8 components."""
    import corrla_rs
    rng = np.random.default_rng(11)
    x, y = rng.standard_normal((1000, 10)), rng.standard_normal((1000, 1))
    comps, vals, sensi = corrla_rs.active_ss(x, y, 1, 30, 8)
    assert comps.shape == (10, 8) and vals.shape == (10, 8) and sensi.shape == (10,)
    fit, g = cb.active_ss_fit(x, y, 1, 30, 8, return_gradients=True)
    g0 = oracle_grad_mat(x, y, 1, 30)
    assert g.shape == (10, 1000) and fit.n_deficient == 0
    assert np.max(np.abs(g - g0)) < 1e-10 * np.max(np.abs(g0))               # same neighbours, same least-squares fits
    ref = ref_stats.ActiveSsRsvd(None, 8).fit_gradients(g0)
    assert np.max(np.abs(np.diag(vals[:8, :8]) - np.diag(ref.singular_vals_)[:8])) < 1e-10 * ref.singular_vals_[0, 0]
    assert ref_rsvd.subspace_sine(ref.components()[:, :3], comps[:, :3]) < 1e-8
    assert np.allclose(sensi, ref.var_diag_evd_sensi(), rtol=1e-8, atol=1e-12)


def test_active_ss_reference_test_quadratic_fit(cb):
    """active_subspaces.rs:326-384 end to end on the device (quadratic local fits through 14 neighbours in 3-D).  The
    reference differentiates the fitted quadratic by a forward difference with eps = 1e-10 (noise ~1e-6); the kernel
    differentiates it analytically, hence the 1e-4 tolerance against the oracle."""
    rng = np.random.default_rng(1)
    cov3 = np.full((3, 3), 0.5) + 0.4 * np.eye(3)
    x3 = (cov3 @ rng.standard_normal((3, 100))).T.copy()
    y3 = 0.2 * x3[:, 0] + 0.5 * x3[:, 1] ** 2 + 0.10 * x3[:, 2] * x3[:, 0]
    fit, g = cb.active_ss_fit(x3, y3, 2, 14, 2, return_gradients=True)
    exact = np.stack([0.2 + 0.1 * x3[:, 2], x3[:, 1], 0.1 * x3[:, 0]])    # the function IS a quadratic: fits are exact
    assert np.max(np.abs(g - exact)) < 1e-9
    g0 = oracle_grad_mat(x3, y3, 2, 14)
    assert np.max(np.abs(g - g0)) < 1e-4
    assert abs(fit.components()[0, 0]) < abs(fit.components()[1, 0])
    assert fit.singular_vals()[0, 0] > fit.singular_vals()[1, 1]
    sens = fit.var_diag_evd_sensi()
    assert len(sens) == 3 and sens[1] > sens[0] and sens[1] > sens[2]


def test_gradient_matrix_larger_and_device(cb):
    import torch
    rng = np.random.default_rng(12)
    n, d = 20000, 16
    x = rng.standard_normal((n, d))
    w = rng.standard_normal((d, 3))
    y = np.sin(x @ w[:, 0]) + 0.5 * (x @ w[:, 1]) ** 2 + x @ w[:, 2]
    fit, g = cb.active_ss_fit(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), 1, 40, 3, return_gradients=True)
    g = g.cpu().numpy()
    rows = list(range(0, n, 97))
    g0 = oracle_grad_mat(x, y, 1, 40, rows)
    assert np.max(np.abs(g[:, rows] - g0)) < 1e-9 * np.max(np.abs(g0))
    ref = ref_stats.ActiveSsRsvd(None, 3).fit_gradients(g)           # the fit itself, on the same gradient matrix
    ev, ev0 = np.diag(fit.singular_vals_), np.diag(ref.singular_vals_)
    assert np.max(np.abs(ev - ev0)) < 1e-11 * ev0[0]
    assert ref_rsvd.subspace_sine(ref.components(), fit.components()) < 1e-8


def test_active_ss_asserts_ties_and_limits(cb):
    rng = np.random.default_rng(13)
    x, y = rng.standard_normal((200, 4)), rng.standard_normal(200)
    with pytest.raises(cb.CorrlaError) as ei:
        cb.active_ss(x, y, 1, 5, 2)                      # n_nbrs > k + 1 is asserted by the reference (:116)
    assert ei.value.status == -1
    with pytest.raises(cb.CorrlaError):
        cb.active_ss(x, y, 3, 20, 2)                     # "Not implemented est order"
    with pytest.raises(cb.CorrlaError) as ei:
        cb.active_ss(x, y, 1, 150, 2)                    # more neighbours than the selection lists hold
    assert ei.value.status == -4
    # duplicated samples: zero distances and ties, neighbourhoods of identical points give rank-deficient fits
    xd = np.repeat(x[:20], 10, axis=0)
    yd = np.repeat(y[:20], 10)
    fit = cb.active_ss_fit(xd, yd, 1, 8, 2)
    assert fit.n_deficient > 0 and np.all(np.isfinite(fit.components_)) and np.all(np.isfinite(fit.singular_vals_))
    # every neighbour (n_nbr >= N): the fit is the global regression, identical at every sample
    xs, ys = x[:40], x[:40] @ np.array([1.0, -2.0, 0.5, 3.0]) + 0.7
    fit, g = cb.active_ss_fit(xs, ys, 1, 64, 1, return_gradients=True)
    assert np.max(np.abs(g - np.array([[1.0], [-2.0], [0.5], [3.0]]))) < 1e-11
    assert abs(np.diag(fit.singular_vals_)[0] - 14.25) < 1e-9 and np.max(np.abs(np.diag(fit.singular_vals_)[1:])) < 1e-12


def test_knn_gemm_form_gives_the_exact_neighbour_lists(cb):
    """The tensor-pipe nearest-neighbour search (shortlist by |q|^2 + |c|^2 - 2 q.c, exact re-rank, certificate) must
    return the SAME neighbour lists as the exact sum-of-(a-b)^2 kernel: the gradient matrices (one local fit per sample
    through exactly those neighbours) are compared bit for bit -- on continuous data, on data with a large common offset
    (cancellation in the distance identity), and on lattice data full of exact ties (certificates fail: exact fallback)."""
    import os
    rng = np.random.default_rng(31)
    cases = {
        "gauss_d64": (rng.standard_normal((6000, 64)), 72),
        "offset_d7": (1e3 + rng.standard_normal((4100, 7)), 12),
        "lattice_d3": (rng.integers(0, 6, size=(3000, 3)).astype(np.float64), 10),
        "wide_d100_k104": (rng.standard_normal((2500, 100)), 104),     # 32-candidate tiles
    }
    cases["gauss_d64_20k"] = (rng.standard_normal((20000, 64)), 72)
    cases["scaled_d16"] = (rng.standard_normal((5000, 16)) * np.logspace(-3, 3, 16), 30)      # features of very different scale
    cases["clusters_d8"] = (np.repeat(rng.standard_normal((40, 8)), 100, axis=0) + 1e-6 * rng.standard_normal((4000, 8)), 20)
    for name, (x, k) in cases.items():
        y = np.sin(x[:, 0]) + x[:, 1] * x[:, 2] + 0.1 * x.sum(axis=1)
        os.environ["CORRLA_B200_KNN_EXACT"] = "1"
        try:
            _, g_exact = cb.active_ss_fit(x, y, 1, k, 2, return_gradients=True)
        finally:
            os.environ.pop("CORRLA_B200_KNN_EXACT", None)
        # shortlist on TF32 mma.sync over the centred float copy (default), and on the FP64 tensor pipe
        for tf32 in ("1", "0"):
            os.environ["CORRLA_B200_KNN_TF32"] = tf32
            try:
                _, g_fast = cb.active_ss_fit(x, y, 1, k, 2, return_gradients=True)
            finally:
                os.environ.pop("CORRLA_B200_KNN_TF32", None)
            assert np.array_equal(np.asarray(g_fast), np.asarray(g_exact)), (name, tf32)


def test_poly_gradient_estimator_grad_at_arbitrary_points(cb):
    """PolyGradientEstimator::grad_at (active_subspaces.rs:57-141) at points that are NOT samples, one at a time (the
    reference's call) and as a batch (one device call), orders 1 and 2, host and device samples; and ActiveSsRsvd over
    the estimator object (create_grad_mat, :226-238) against the oracle's loop."""
    import torch
    rng = np.random.default_rng(21)
    n, d = 3000, 6
    x = rng.standard_normal((n, d))
    w = rng.standard_normal((d, 2))
    y = np.sin(x @ w[:, 0]) + 0.3 * (x @ w[:, 1]) ** 2
    xq = 0.8 * rng.standard_normal((70, d))                                # 70 > 64: more than one query block
    for order, k, tol in ((1, 25, 1e-9), (2, 40, 2e-4)):                   # order 2: forward difference in the oracle
        est0 = ref_stats.PolyGradientEstimator(x, y, order, k)
        g0 = np.stack([est0.grad_at(q).ravel() for q in xq])
        est = cb.PolyGradientEstimator(x, y.reshape(-1, 1), order, k)
        g = np.asarray(est.grad_at_many(xq))
        assert g.shape == (70, d) and est.n_deficient == 0
        assert np.max(np.abs(g - g0)) < tol * max(1.0, np.max(np.abs(g0)))
        one = np.asarray(est.grad_at(list(xq[3])))
        assert one.shape == (1, d) and np.array_equal(one.ravel(), g[3])
        # device-resident samples and queries
        estd = cb.PolyGradientEstimator(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), order, k)
        gd = estd.grad_at_many(torch.from_numpy(xq).cuda())
        assert gd.is_cuda and np.array_equal(gd.cpu().numpy(), g)
        # at the samples themselves it is the gradient matrix of active_ss_fit
        _fit, gm = cb.active_ss_fit(x, y, order, k, 2, return_gradients=True)
        gs = np.asarray(est.grad_at_many(x[:130]))
        assert np.array_equal(gs.T, gm[:, :130])
    # an exact quadratic is recovered exactly anywhere inside the cloud
    yq = 0.2 * x[:, 0] + 0.5 * x[:, 1] ** 2 + 0.1 * x[:, 2] * x[:, 0]
    est = cb.PolyGradientEstimator(x, yq, 2, 60)
    g = np.asarray(est.grad_at_many(xq))
    exact = np.zeros_like(xq)
    exact[:, 0], exact[:, 1], exact[:, 2] = 0.2 + 0.1 * xq[:, 2], xq[:, 1], 0.1 * xq[:, 0]
    assert np.max(np.abs(g - exact)) < 1e-8
    # the reference's object interface end to end
    fit = cb.ActiveSsRsvd(cb.PolyGradientEstimator(x, y, 1, 25), 2).fit(x[:500])
    ref = ref_stats.ActiveSsRsvd(ref_stats.PolyGradientEstimator(x, y, 1, 25), 2).fit(x[:500])
    assert np.max(np.abs(np.diag(fit.singular_vals_) - np.diag(ref.singular_vals_))) < 1e-10 * ref.singular_vals_[0, 0]
    assert ref_rsvd.subspace_sine(ref.components(), fit.components()) < 1e-8
    with pytest.raises(cb.CorrlaError):
        cb.PolyGradientEstimator(x, y, 1, 5).grad_at(list(xq[0]))          # n_nbrs > k + 1 (:116)
    with pytest.raises(cb.CorrlaError):
        cb.PolyGradientEstimator(x, y, 3, 30).grad_at(list(xq[0]))         # "Not implemented est order"
