"""GPU parity for the first "next" row of SURVEY 8(f): PCA by RSVD with the centring fused into the passes
(pca_rsvd.rs:56-82, lib_math_utils_py.rs:38-55), against oracle/ref_pca.py on the same input and Omega."""
import numpy as np
import pytest

from oracle import ref_pca, ref_rsvd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


def make_data(rng, n_samples, n_dim):
    """correlated features with a large offset: centring matters (|mean| >> std)."""
    latent = rng.standard_normal((n_samples, n_dim)) * np.linspace(3.0, 0.5, n_dim)
    mix, _ = np.linalg.qr(rng.standard_normal((n_dim, n_dim)))
    return latent @ mix + 50.0 * rng.standard_normal(n_dim)


@pytest.mark.parametrize("shape,rank,order", [((10000, 12), 4, "C"), ((20000, 64), 8, "C"), ((6000, 200), 10, "F"),
                                               ((100, 10), 4, "C")])
def test_rpca_matches_oracle_tall(cb, shape, rank, order):
    rng = np.random.default_rng(31)
    x = make_data(rng, *shape)
    l = min(rank + min(shape[1], 10), shape[1])
    omega = rng.standard_normal((shape[1], l))
    ref = ref_pca.pca_rsvd_new(x, rank, omega=omega)
    xin = np.asfortranarray(x) if order == "F" else x
    s, comps, means = cb.rpca(xin, rank, 7, 3, omega=omega, return_means=True)      # n_iters / n_oversamples ignored
    assert s.shape == (rank, 1) and comps.shape == (rank, shape[1]) and means.shape == (1, shape[1])
    assert np.max(np.abs(means - ref["means"])) < 1e-12 * np.max(np.abs(ref["means"]))
    assert ref_rsvd.sigma_rel_err(ref["singular_values"], s) < 1e-10
    assert ref_rsvd.subspace_sine(ref["components"].T, np.asarray(comps).T) < 1e-8
    assert np.max(np.abs(comps @ comps.T - np.eye(rank))) < 1e-12
    t = cb.last_timings()
    assert t["passes_over_a"] == 42


def test_rpca_fat_uses_explicit_centred_copy(cb):
    """n_samples < n_dim: random_svd works on the transposed view; the column means become per-row constants."""
    rng = np.random.default_rng(32)
    x = make_data(rng, 40, 600)
    rank = 6
    omega = rng.standard_normal((40, min(rank + 10, 40)))
    ref = ref_pca.pca_rsvd_new(x, rank, omega=omega)
    s, comps, means = cb.rpca(x, rank, omega=omega, return_means=True)
    assert np.max(np.abs(means - ref["means"])) < 1e-12 * np.max(np.abs(ref["means"]))
    assert ref_rsvd.sigma_rel_err(ref["singular_values"], s) < 1e-10
    assert ref_rsvd.subspace_sine(ref["components"].T, np.asarray(comps).T) < 1e-8


def test_centred_rsvd_equals_rsvd_of_centred_copy(cb):
    """center=True on A must equal rsvd of the explicitly centred matrix (same Omega), U included."""
    rng = np.random.default_rng(33)
    x = make_data(rng, 5000, 96)
    omega = rng.standard_normal((96, 30))
    cx = x - x.mean(axis=0)
    u0, s0, v0 = ref_rsvd.random_svd(cx, 20, 5, 10, omega=omega)
    u, s, vt = cb.rsvd(x, 20, 5, 10, omega=omega, center=True)
    assert ref_rsvd.sigma_rel_err(s0, s) < 1e-10
    assert ref_rsvd.subspace_sine(u0, np.asarray(u)) < 1e-8
    assert ref_rsvd.subspace_sine(v0.T, np.asarray(vt).T) < 1e-8
    # and it must differ from the uncentred decomposition (the offset dominates otherwise)
    u2, s2, _ = cb.rsvd(x, 20, 5, 10, omega=omega)
    assert abs(float(s2[0, 0]) / float(s[0, 0])) > 5.0


def test_rpca_device_resident(cb):
    import torch
    rng = np.random.default_rng(34)
    x = make_data(rng, 30000, 48)
    omega = rng.standard_normal((48, 18))
    ref = ref_pca.pca_rsvd_new(x, 8, omega=omega)
    s, comps = cb.rpca(torch.from_numpy(x).cuda(), 8, omega=torch.from_numpy(omega).cuda())
    torch.cuda.synchronize()
    assert ref_rsvd.sigma_rel_err(ref["singular_values"], s.cpu().numpy()) < 1e-10
    assert ref_rsvd.subspace_sine(ref["components"].T, comps.cpu().numpy().T) < 1e-8
    ev = ref_pca.explained_var(ref)
    assert np.allclose((s.cpu().numpy() ** 2 / (30000 - 1.0)), ev, rtol=1e-9)
