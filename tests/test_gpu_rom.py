"""GPU parity (-m gpu) of the reduced-order models built directly on the RSVD: DMDc
(corrla_dmdc_f64, dmd_rom.rs:46-146) and POD modes/weights (corrla_pod_f64, pod_rom.rs:53-75), through the C ABI and
the host mirrors corrla_rs.PyDMDc / corrla_rs.PyPodI, against oracle/ref_rom.py on the same inputs and the same
injected sketch matrices.  The reduced operator a_til lives in the basis u_hat, which is fixed only up to the sign of
each singular vector: comparisons are made on basis-independent quantities (eigenvalues, u_hat a_til u_hat^T, B,
the lifted modes' span, predictions)."""
import numpy as np
import pytest

from oracle import ref_rom, ref_rsvd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


def planted_system(rng, n_x, n_u, nt, r, noise=0.0):
    q, _ = np.linalg.qr(rng.standard_normal((n_x, r)))
    lam = np.linspace(0.95, 0.45, r) * np.where(np.arange(r) % 3 == 2, -1.0, 1.0)
    a = (q * lam) @ q.T
    bmat = q @ rng.standard_normal((r, n_u))
    u = rng.standard_normal((n_u, nt))
    x = np.zeros((n_x, nt))
    x[:, 0] = q @ rng.standard_normal(r)
    for t in range(nt - 1):
        x[:, t + 1] = a @ x[:, t] + bmat @ u[:, t]
    if noise:
        x = x + noise * rng.standard_normal(x.shape)
    return x, u, a, bmat, lam


def dmd_omegas(rng, n_x, n_u, nt, r):
    """Sketch matrices shaped like random_svd wants them for the two views (thin columns x l)."""
    out = []
    for rows in (n_x + n_u, n_x):
        n_thin = min(rows, nt - 1)
        l = min(r + 12, n_thin)
        out.append(rng.standard_normal((n_thin, l)))
    return tuple(out)


def full_operator(u_hat, a_til):
    return u_hat @ a_til @ u_hat.T


@pytest.mark.parametrize("n_x,n_u,nt,order", [(300, 2, 60, "C"), (301, 2, 61, "F"), (40, 1, 90, "C"), (1000, 3, 33, "F")])
def test_dmdc_operators_match_oracle(cb, n_x, n_u, nt, order):
    rng = np.random.default_rng(n_x + nt)
    r_true = 6
    x, u, _a, _b, _lam = planted_system(rng, n_x, n_u, nt, r_true)
    r = r_true + n_u        # rank of the input space [x; u]; the output space has rank r_true (u_hat's tail is arbitrary)
    omegas = dmd_omegas(rng, n_x, n_u, nt, r)
    ref = ref_rom.DMDc(x, u, 1.0, r, 5, omegas=omegas)
    xs = np.asfortranarray(x) if order == "F" else np.ascontiguousarray(x)
    ops = cb.dmdc_operators(xs, u, r, 5, omegas=omegas)
    t = cb.last_timings()
    assert t["passes_over_a"] == 2 * (2 + 2 * 5) + 2 and t["gpu_launches"] > 0
    assert ops["a_til"].shape == (r, r) and ops["b"].shape == (n_x, n_u) and ops["modes_scale"].shape == (n_x, r)
    assert ref_rsvd.sigma_rel_err(ref.s_til, ops["s_til"]) < 1e-10
    assert ref_rsvd.subspace_sine(ref.u_hat[:, :r_true], ops["u_hat"][:, :r_true]) < 1e-8
    scale = np.max(np.abs(ref.a_til))
    assert np.max(np.abs(full_operator(ops["u_hat"], ops["a_til"]) - full_operator(ref.u_hat, ref.a_til))) < 1e-8 * scale
    assert np.max(np.abs(ops["b"] - ref.b)) < 1e-8 * max(1.0, np.max(np.abs(ref.b)))
    ev, ev0 = np.linalg.eigvals(ops["a_til"]), np.linalg.eigvals(ref.a_til)
    assert np.max(np.abs(np.sort_complex(ev) - np.sort_complex(ev0))) < 1e-8
    # modes_scale = Y V S^-1 U1^T u_hat carries u_hat's column signs on the right: compare modes_scale * u_hat^T
    lift, lift0 = ops["modes_scale"] @ ops["u_hat"].T, ref.modes_scale @ ref.u_hat.T
    assert np.max(np.abs(lift - lift0)) < 1e-8 * np.max(np.abs(lift0))


@pytest.mark.parametrize("nx", [20, 50, 500])
def test_pydmdc_passes_the_reference_test(nx):
    """dmd_rom.rs:236-309 through the drop-in class: rank-3 data, 14 modes, 40 iterations (every QR is rank deficient,
    the trailing singular values are roundoff and get inverted): prediction of snapshot 20 within 5e-2."""
    import corrla_rs
    p, u = ref_rom.dmdc_test_snapshots(nx, 40)
    model = corrla_rs.PyDMDc(p, u, 14, 40)
    pred = model.predict(p[:, 0:1], u)
    assert pred.shape == (nx, 40)
    assert model.dmd.lambdas.shape == (14, 1)
    assert model.dmd.est_a_til().shape == (nx, nx) and model.dmd.est_b_til().shape == (nx, 1)
    assert np.max(np.abs(pred[:, 19] - p[:, 20])) < 5e-2


def test_dmdc_class_matches_oracle_predictions(cb):
    rng = np.random.default_rng(9)
    n_x, n_u, nt, r_true = 400, 2, 80, 5
    x, u, _a, bmat, lam = planted_system(rng, n_x, n_u, nt, r_true)
    r = r_true + n_u
    omegas = dmd_omegas(rng, n_x, n_u, nt, r)
    ref = ref_rom.DMDc(x, u, 1.0, r, 6, omegas=omegas)
    dm = cb.DMDc(x, u, 1.0, r, 6, omegas=omegas)
    assert np.allclose(np.sort(dm.lambdas.real.ravel()), np.sort(np.concatenate([lam, np.zeros(n_u)])), atol=1e-8)
    assert np.max(np.abs(dm.est_b_til() - bmat)) < 1e-8
    # est_a_til = Re(Phi Lambda Phi^+) inverts the roundoff-sized singular values of Phi that belong to the n_u zero
    # eigenvalues (mat_pinv_comp adds 1e-16 to them): outside the data it is noise in the reference too, so the two
    # are compared through their action on the snapshots
    assert np.max(np.abs(dm.est_a_til() @ x[:, :7] - ref.est_a_til() @ x[:, :7])) < 1e-8 * np.max(np.abs(x))
    p0, p1 = ref.predict_multiple(x[:, 0:1], u[:, :-1]), dm.predict_multiple(x[:, 0:1], u[:, :-1])
    assert np.max(np.abs(p1 - p0)) < 1e-8 and np.max(np.abs(p1 - x[:, 1:])) < 1e-7
    assert np.allclose(dm.predict(x[:, 3:4], u[:, 3:4]), x[:, 4:5], atol=1e-7)


def test_dmdc_device_resident_and_wide_control(cb):
    """Snapshots as torch CUDA tensors (nothing crosses PCIe; row- and column-major), and a control matrix wider than
    the GEMM column block (n_u = 19 > Lc = 8: the B panel loop).  The 19 controls are mixtures of 2 signals, so the
    input space has exact rank 4 + 2 = n_modes."""
    import torch
    rng = np.random.default_rng(21)
    n_x, nt, r_true = 20_000, 70, 4
    mix = rng.standard_normal((19, 2))
    x, z, _a, b2, lam = planted_system(rng, n_x, 2, nt, r_true)       # driven by z (2 x nt) through b2 (n_x x 2)
    u = mix @ z                                                        # what the model is given: 19 x nt
    r = r_true + 2
    omegas = dmd_omegas(rng, n_x, 19, nt, r)
    ref = ref_rom.DMDc(x, u, 1.0, r, 4, omegas=omegas)
    xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    for xin in (xd, xd.t().contiguous().t()):
        ops = cb.dmdc_operators(xin, ud, r, 4, omegas=omegas)
        torch.cuda.synchronize()
        assert ops["b"].is_cuda and tuple(ops["b"].shape) == (n_x, 19)
        bb = ops["b"].cpu().numpy()
        assert np.max(np.abs(bb - ref.b)) < 1e-8 * np.max(np.abs(ref.b))
        assert np.max(np.abs(bb @ u - b2 @ z)) < 1e-8 * np.max(np.abs(b2 @ z))     # the forcing it must reproduce
        ev = np.sort(np.linalg.eigvals(ops["a_til"].cpu().numpy()).real)
        assert np.allclose(ev, np.sort(np.concatenate([lam, np.zeros(2)])), atol=1e-8)
    seeded = cb.dmdc_operators(xd, ud, r, 4, seed=3)                   # Philox sketches drawn on the device
    ev = np.sort(np.linalg.eigvals(seeded["a_til"].cpu().numpy()).real)
    assert np.allclose(ev, np.sort(np.concatenate([lam, np.zeros(2)])), atol=1e-8)


def test_dmdc_without_control_and_errors(cb):
    rng = np.random.default_rng(5)
    x, _u, _a, _b, lam = planted_system(rng, 500, 0, 50, 5)
    ops = cb.dmdc_operators(x, np.zeros((0, 50)), 5, 4, seed=1)
    assert np.allclose(np.sort(np.linalg.eigvals(ops["a_til"]).real), np.sort(lam), atol=1e-8)
    with pytest.raises(cb.RankPanic):
        cb.dmdc_operators(x[:10], np.zeros((1, 50)), 11, 2)
    with pytest.raises(ValueError):
        cb.dmdc_operators(x, np.zeros((1, 49)), 3, 2)


# ------------------------------------------------------------------ POD
@pytest.mark.parametrize("shape", [(20, 100), (64, 5000), (300, 40)])
def test_pod_modes_and_weights_match_oracle(cb, shape):
    n_snap, n_points = shape
    rng = np.random.default_rng(n_points)
    r = 4
    base = rng.standard_normal((n_snap, 6)) * (5.0 * 0.6 ** np.arange(6))
    x = base @ np.linalg.qr(rng.standard_normal((n_points, 6)))[0].T + 1e-5 * rng.standard_normal(shape)
    n_thin = min(shape)
    omega = rng.standard_normal((n_thin, min(r + 10, n_thin)))
    t = np.linspace(0.0, 1.0, n_snap).reshape(-1, 1)
    ref = ref_rom.PodI(x, t, r, omega=omega)
    modes, weights, s = cb.pod_modes_weights(x, r, omega=omega)
    assert modes.shape == (n_points, r) and weights.shape == (n_snap, r) and s.shape == (r, 1)
    assert ref_rsvd.subspace_sine(ref.modes, modes) < 1e-8
    assert np.max(np.abs(modes.T @ modes - np.eye(r))) < 1e-12
    assert np.max(np.abs(weights - x @ modes)) < 1e-11 * np.max(np.abs(x))
    # reconstruction is basis independent
    assert np.max(np.abs(weights @ modes.T - ref.mode_weights @ ref.modes.T)) < 1e-8 * np.max(np.abs(x))
    pod = cb.PodI(x, t, r, omega=omega)
    tq = np.array([[0.37]])
    assert np.max(np.abs(pod.predict(tq) - ref.predict(tq))) < 1e-8 * np.max(np.abs(x))


def test_pypodi_reference_generator_and_device_input(cb):
    import corrla_rs
    import torch
    x, t = ref_rom.pod_test_snapshots()
    pod = corrla_rs.PyPodI(x, t, 4)                                   # pod_rom.rs:148-154
    y = pod.predict(np.array([[5.2]]))
    assert y.shape == (100, 1) and np.all(np.isfinite(y))
    full = corrla_rs.PyPodI(x, t, 20)
    assert np.max(np.abs(full.predict(t[7:8]).ravel() - x[7])) < 1e-8
    xd = torch.from_numpy(x).cuda()
    modes, weights, _s = cb.pod_modes_weights(xd, 4, seed=2)
    assert modes.is_cuda and tuple(modes.shape) == (100, 4) and tuple(weights.shape) == (20, 4)
    assert float((weights - xd @ modes).abs().max()) < 1e-11


# ------------------------------------------------------------------ randomised shapes / layouts / residency
def test_rom_fuzz(cb):
    """Random sizes, layouts (C / Fortran / strided views / device tensors) and mode counts for DMDc, POD, covariance
    and the active-subspace gradients, each checked against the oracle on basis-independent quantities."""
    import torch
    from oracle import ref_stats
    rng = np.random.default_rng(2024)

    def relayout(a, kind):
        if kind == 1:
            return np.asfortranarray(a)
        if kind == 2:
            wide = np.zeros((a.shape[0], 2 * a.shape[1])); wide[:, ::2] = a
            return wide[:, ::2]
        if kind == 3:
            wide = np.zeros((a.shape[0] + 1, a.shape[1] + 3)); wide[1:, 1:a.shape[1] + 1] = a
            return wide[1:, 1:a.shape[1] + 1]
        return a

    for trial in range(10):
        # --- DMDc on an exactly low-rank controlled system
        n_x, n_u, nt, r_true = int(rng.integers(30, 900)), int(rng.integers(1, 4)), int(rng.integers(24, 80)), int(rng.integers(2, 6))
        x, u, _a, bmat, lam = planted_system(rng, n_x, n_u, nt, r_true)
        r = r_true + n_u
        omegas = dmd_omegas(rng, n_x, n_u, nt, r)
        ref = ref_rom.DMDc(x, u, 1.0, r, 4, omegas=omegas)
        kind = int(rng.integers(0, 5))
        if kind == 4:
            ops = cb.dmdc_operators(torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda(), r, 4, omegas=omegas)
            ops = {k_: v.cpu().numpy() for k_, v in ops.items()}
        else:
            ops = cb.dmdc_operators(relayout(x, kind), u, r, 4, omegas=omegas)
        tag = (trial, n_x, n_u, nt, r_true, kind)
        assert np.max(np.abs(ops["b"] - ref.b)) < 1e-7 * max(1.0, np.max(np.abs(ref.b))), tag
        ev, ev0 = np.sort_complex(np.linalg.eigvals(ops["a_til"])), np.sort_complex(np.linalg.eigvals(ref.a_til))
        assert np.max(np.abs(ev - ev0)) < 1e-7, tag
        # --- POD
        n_snap, n_points, r = int(rng.integers(8, 60)), int(rng.integers(40, 3000)), int(rng.integers(1, 6))
        base = rng.standard_normal((n_snap, 7)) * (4.0 * 0.5 ** np.arange(7))
        xs = base @ np.linalg.qr(rng.standard_normal((n_points, 7)))[0].T + 1e-7 * rng.standard_normal((n_snap, n_points))
        n_thin = min(n_snap, n_points)
        omega = rng.standard_normal((n_thin, min(r + 10, n_thin)))
        refp = ref_rom.PodI(xs, np.arange(n_snap, dtype=np.float64).reshape(-1, 1), r, omega=omega)
        modes, weights, _s = cb.pod_modes_weights(relayout(xs, int(rng.integers(0, 4))), r, omega=omega)
        assert np.max(np.abs(weights @ modes.T - refp.mode_weights @ refp.modes.T)) < 1e-8 * np.max(np.abs(xs)), (trial, n_snap, n_points, r)
        # --- covariance / correlation
        n, d = int(rng.integers(3, 5000)), int(rng.integers(1, 129))
        xc = rng.standard_normal((n, d)) * rng.uniform(0.1, 10.0, d) + rng.uniform(-100.0, 100.0, d)
        lay = relayout(xc, int(rng.integers(0, 4)))
        c0 = ref_stats.mat_cov_centered(xc)
        assert np.max(np.abs(cb.mat_cov_centered(lay) - c0)) < 1e-10 * np.max(np.abs(c0)), (trial, n, d)
        if n > 3:
            assert np.max(np.abs(cb.pearson_corr(lay) - ref_stats.pearson_corr(xc))) < 1e-9, (trial, n, d)
        # --- gradients of local linear fits
        n, d = int(rng.integers(60, 1500)), int(rng.integers(1, 12))
        k = int(rng.integers(d + 2, min(n, 60) + 1))
        xg = rng.standard_normal((n, d))
        yg = np.sin(xg @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
        _fit, g = cb.active_ss_fit(relayout(xg, int(rng.integers(0, 4))), yg, 1, k, 1, return_gradients=True)
        est = ref_stats.PolyGradientEstimator(xg, yg, 1, k)
        rows = rng.choice(n, size=12, replace=False)
        g0 = np.stack([est.grad_at(xg[i]).ravel() for i in rows], axis=1)
        assert np.max(np.abs(g[:, rows] - g0)) < 1e-8 * max(1.0, np.max(np.abs(g0))), (trial, n, d, k)


# ------------------------------------------------------------------ more modes than one 128-column panel holds
def _wide_dmd_data(rng, n_x, n_u, nt, r_true):
    """Snapshots with r_true dominant directions of gently graded strength over a noise floor (low rank + noise, the
    parity class of SURVEY F9; the planted linear systems above lose rank when a couple of inputs drive hundreds of
    modes).  DMDc of arbitrary data is still a well-defined least-squares fit."""
    q, _ = np.linalg.qr(rng.standard_normal((n_x, r_true)))
    x = (q * (10.0 * 0.99 ** np.arange(r_true))) @ rng.standard_normal((r_true, nt)) + 1e-3 * rng.standard_normal((n_x, nt))
    u = rng.standard_normal((n_u, nt))
    return x, u


@pytest.mark.parametrize("n_x,n_u,nt,r", [(1200, 2, 400, 152), (1400, 2, 600, 258)])
def test_dmdc_more_modes_than_one_panel(cb, n_x, n_u, nt, r):
    """n_modes = 152 (sketch l = 164: two column panels) and 258 (l = 270: three) -- the reference has no cap on n_modes
    (dmd_rom.rs:45-61).  The data have r + 8 graded directions over a noise floor, so the truncation at r falls between
    two well separated singular values in both spaces.  Same comparisons as the single-panel test, on basis-independent
    quantities."""
    rng = np.random.default_rng(n_x)
    x, u = _wide_dmd_data(rng, n_x, n_u, nt, r + 8)
    omegas = dmd_omegas(rng, n_x, n_u, nt, r)
    ref = ref_rom.DMDc(x, u, 1.0, r, 4, omegas=omegas)
    ops = cb.dmdc_operators(x, u, r, 4, omegas=omegas)
    assert ops["a_til"].shape == (r, r) and ops["b"].shape == (n_x, n_u) and ops["modes_scale"].shape == (n_x, r)
    assert ref_rsvd.sigma_rel_err(ref.s_til, ops["s_til"]) < 1e-10
    assert ref_rsvd.subspace_sine(ref.u_hat, ops["u_hat"]) < 1e-8
    assert np.max(np.abs(ops["u_hat"].T @ ops["u_hat"] - np.eye(r))) < 1e-11
    scale = np.max(np.abs(full_operator(ref.u_hat, ref.a_til)))
    assert np.max(np.abs(full_operator(ops["u_hat"], ops["a_til"]) - full_operator(ref.u_hat, ref.a_til))) < 1e-8 * scale
    assert np.max(np.abs(ops["b"] - ref.b)) < 1e-8 * max(1.0, np.max(np.abs(ref.b)))
    lift, lift0 = ops["modes_scale"] @ ops["u_hat"].T, ref.modes_scale @ ref.u_hat.T
    assert np.max(np.abs(lift - lift0)) < 1e-8 * np.max(np.abs(lift0))
    assert abs(np.trace(ops["a_til"]) - np.trace(ref.a_til)) < 1e-8 * max(1.0, abs(np.trace(ref.a_til)))


@pytest.mark.parametrize("r", [150, 260])
def test_pod_more_modes_than_one_panel(cb, r):
    import torch
    rng = np.random.default_rng(r)
    n_snap, n_points, rank = 320, 5000, 280
    # gently graded: (sigma_1 / sigma_l)^7 stays far below 1/eps at the reference's first QR (SURVEY F9 parity class)
    base = rng.standard_normal((n_snap, rank)) * (10.0 * 0.995 ** np.arange(rank))
    x = base @ np.linalg.qr(rng.standard_normal((n_points, rank)))[0].T
    omega = rng.standard_normal((n_snap, min(r + 10, n_snap)))
    t = np.linspace(0.0, 1.0, n_snap).reshape(-1, 1)
    ref = ref_rom.PodI(x, t, r, omega=omega)
    modes, weights, s = cb.pod_modes_weights(x, r, omega=omega)
    assert modes.shape == (n_points, r) and weights.shape == (n_snap, r) and s.shape == (r, 1)
    assert ref_rsvd.subspace_sine(ref.modes, modes) < 1e-8
    assert np.max(np.abs(modes.T @ modes - np.eye(r))) < 1e-11
    assert np.max(np.abs(weights - x @ modes)) < 1e-11 * np.max(np.abs(x))
    assert np.max(np.abs(weights @ modes.T - ref.mode_weights @ ref.modes.T)) < 1e-8 * np.max(np.abs(x))
    md, wd, _ = cb.pod_modes_weights(torch.from_numpy(x).cuda(), r, omega=torch.from_numpy(omega).cuda())
    md, wd = md.cpu().numpy(), wd.cpu().numpy()            # device-resident snapshots: same factors up to the sign of a mode
    assert ref_rsvd.subspace_sine(np.asarray(modes), md) < 1e-9
    assert np.max(np.abs(wd @ md.T - np.asarray(weights) @ np.asarray(modes).T)) < 1e-9 * np.max(np.abs(x))
