"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs and the same injected Omega.  Tolerances are the north-star ones: singular values within 1e-10 relative,
sine of the largest principal angle under 1e-8 -- on inputs where the oracle agrees with itself 10x tighter
(tests/test_oracle.py::test_oracle_self_consistency_defines_parity_class)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_rsvd

pytestmark = pytest.mark.gpu

TOL_SIGMA = 1e-10     # relative, per singular value
TOL_ANGLE = 1e-8      # sine of the largest principal angle
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


def lowrank_noise(rng, m, n, r, noise):
    u, _ = np.linalg.qr(rng.standard_normal((m, r)))
    v, _ = np.linalg.qr(rng.standard_normal((n, r)))
    return (u * (10.0 * 0.95 ** np.arange(r))) @ v.T + noise * rng.standard_normal((m, n))


def assert_parity(out, ref, k, tol_sigma=TOL_SIGMA, tol_angle=TOL_ANGLE):
    u, s, vt = (np.asarray(x) for x in out)
    u0, s0, vt0 = ref
    assert u.shape == u0.shape and s.shape == s0.shape == (k, 1) and vt.shape == vt0.shape
    es = ref_rsvd.sigma_rel_err(s0, s)
    eu = ref_rsvd.subspace_sine(u0, u)
    ev = ref_rsvd.subspace_sine(vt0.T, vt.T)
    assert es < tol_sigma, f"sigma rel err {es:.3e}"
    assert eu < tol_angle, f"sin(U angle) {eu:.3e}"
    assert ev < tol_angle, f"sin(V angle) {ev:.3e}"
    assert np.max(np.abs(u.T @ u - np.eye(k))) < 1e-12
    assert np.max(np.abs(vt @ vt.T - np.eye(k))) < 1e-12
    return es, eu, ev


# ------------------------------------------------------------------ fixed vectors
def test_known_answer_lowrank_5x5(cb):
    """The reference's only known-answer test (random_svd.rs:153-196): rank-3 5x5 matrix, l = min(15, 5) = 5, so Y is
    rank deficient at every QR (the CholeskyQR deflation + refill path)."""
    g = np.load(GOLD / "known_answer_5x5.npz")
    a, sigma = g["a"], g["sigma"]
    u, s, vt = cb.rsvd(a, 5, 12, 10, seed=1)
    assert u.shape == (5, 5) and s.shape == (5, 1) and vt.shape == (5, 5)
    assert np.max(np.abs(s.ravel() - sigma)) < 1e-3                 # the reference's own tolerance (:182)
    assert np.max(np.abs(s.ravel()[:3] - np.array([3.0, np.sqrt(5.0), 2.0]))) < 1e-12
    assert np.max(np.abs(u @ np.diag(s.ravel()) @ vt - a)) < 1e-12
    # like a Householder QR, the engine completes the basis beyond the numerical rank: U and V stay orthonormal
    assert np.max(np.abs(u.T @ u - np.eye(5))) < 1e-12 and np.max(np.abs(vt @ vt.T - np.eye(5))) < 1e-12
    u, s, vt = cb.rsvd(a, 3, 12, 10, seed=2)
    assert s.shape == (3, 1)
    assert np.max(np.abs(s.ravel() - sigma[:3])) < 1e-3             # :195
    assert np.max(np.abs(u @ np.diag(s.ravel()) @ vt - a)) < 1e-12


@pytest.mark.parametrize("name", ["c1", "tall", "fat", "clamp"])
def test_committed_oracle_vectors(cb, name):
    g = np.load(GOLD / "oracle_cases.npz")
    a, omega = g[f"{name}_a"], g[f"{name}_omega"]
    k, q, p = (int(x) for x in g[f"{name}_kqp"])
    out = cb.rsvd(a, k, q, p, omega=omega)
    assert_parity(out, (g[f"{name}_u"], g[f"{name}_s"], g[f"{name}_vt"]), k)


@pytest.mark.parametrize("name", ["tall", "fat"])
def test_reference_python_statement_vectors(cb, name):
    """Outputs of the reference's own examples/benchmark_rsvd.py:16-54 (no in-loop QR, no scaling: same subspace
    in exact arithmetic, so the tolerance is the oracle-vs-statement one, tests/test_oracle.py)."""
    g = np.load(GOLD / f"ref_examples_rsvd_{name}.npz")
    k, p, q = int(g["k"]), int(g["p"]), int(g["q"])
    u, s, vt = cb.rsvd(g["a"], k, q, p, omega=g["omega"])
    assert ref_rsvd.sigma_rel_err(g["s"], s) < 1e-9
    assert ref_rsvd.subspace_sine(g["u"], u) < 1e-7
    assert ref_rsvd.subspace_sine(g["vt"].T, vt.T) < 1e-7


# ------------------------------------------------------------------ seeded parity, every layout the callers produce
CASES = [
    # name, m, n, (k, q, p), generator
    ("readme_100x100", 100, 100, (10, 12, 8), "gauss"),
    ("tall_gauss", 4096, 512, (20, 4, 10), "gauss"),
    ("tall_l110", 6000, 700, (100, 4, 10), "gauss"),
    ("pca_like_q20", 3000, 12, (4, 20, 10), "gauss"),          # pca_rsvd.rs:65-66 (q = 20, p = min(n, 10))
    ("lowrank_noise", 5000, 640, (30, 4, 10), "lowrank"),
    ("active_ss_like", 20000, 64, (8, 8, 10), "gauss"),        # active_subspaces.rs:241-243
    ("ragged", 1037, 131, (17, 5, 6), "gauss"),
]


def make(case, rng):
    name, m, n, kqp, gen = case
    a = rng.standard_normal((m, n)) if gen == "gauss" else lowrank_noise(rng, m, n, 60, 1e-2)
    k, q, p = kqp
    omega = rng.standard_normal((n, min(k + p, n)))
    return a, omega, kqp


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_parity_row_major_host(cb, case):
    a, omega, (k, q, p) = make(case, np.random.default_rng(100))
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    assert_parity(cb.rsvd(a, k, q, p, omega=omega), ref, k)


@pytest.mark.parametrize("case", CASES[1:5], ids=[c[0] for c in CASES[1:5]])
def test_parity_column_major_host(cb, case):
    """faer Mat / numpy F-order: unit row stride (the layout POD and the Rust callers hand in)."""
    a, omega, (k, q, p) = make(case, np.random.default_rng(101))
    af = np.asfortranarray(a)
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    assert_parity(cb.rsvd(af, k, q, p, omega=np.asfortranarray(omega)), ref, k)


def test_parity_fat_pod_like(cb):
    """pod_rom.rs:56: snapshots x space (fat) => the path runs on the transposed view."""
    rng = np.random.default_rng(102)
    a = rng.standard_normal((20, 5000))
    k, q, p = 8, 10, 10
    omega = rng.standard_normal((20, 18))
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    out = cb.rsvd(a, k, q, p, omega=omega)
    assert out[0].shape == (20, 8) and out[2].shape == (8, 5000)
    assert_parity(out, ref, k)


def test_parity_dmd_like_subviews(cb):
    """dmd_rom.rs:149-162: _X = all rows, cols 0..N-1 and _Y = first nx rows, cols 1..N of one column-major buffer
    with an odd column stride (unaligned for TMA => repack / pitched-copy path)."""
    rng = np.random.default_rng(103)
    nx, nu, nt = 3001, 2, 41
    buf = np.asfortranarray(rng.standard_normal((nx + nu, nt)))
    x = buf[:, : nt - 1]
    y = buf[:nx, 1:]
    assert x.strides == (8, 8 * (nx + nu)) and (nx + nu) % 2 == 1
    for view in (x, y):
        k, q, p = 14, 6, 12
        omega = rng.standard_normal((view.shape[1], min(k + p, view.shape[1])))
        ref = ref_rsvd.random_svd(view, k, q, p, omega=omega)
        assert_parity(cb.rsvd(view, k, q, p, omega=omega), ref, k)


def test_parity_device_resident_torch(cb):
    """Device path: torch CUDA tensors in, torch CUDA tensors out, row-major, transposed view and odd-offset slice."""
    import torch
    rng = np.random.default_rng(104)
    m, n, k, q, p = 8192, 384, 24, 4, 10
    a = rng.standard_normal((m, n))
    omega = rng.standard_normal((n, k + p))
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    ad = torch.from_numpy(a).cuda()
    od = torch.from_numpy(omega).cuda()
    out = cb.rsvd(ad, k, q, p, omega=od)
    assert all(t.is_cuda for t in out)
    torch.cuda.synchronize()
    assert_parity([t.cpu().numpy() for t in out], ref, k)
    # column-major device view of the same matrix
    adt = torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
    assert adt.stride() == (1, m)
    out = cb.rsvd(adt, k, q, p, omega=od)
    torch.cuda.synchronize()
    assert_parity([t.cpu().numpy() for t in out], ref, k)
    # slice with an odd element offset and odd pitch: not TMA-compatible => device repack
    big = torch.zeros((m, n + 3), dtype=torch.float64, device="cuda")
    big[:, 1:n + 1] = ad
    out = cb.rsvd(big[:, 1:n + 1], k, q, p, omega=od)
    torch.cuda.synchronize()
    assert_parity([t.cpu().numpy() for t in out], ref, k)


def test_schedules_agree_on_benign_input(cb):
    rng = np.random.default_rng(105)
    a = rng.standard_normal((3000, 200))
    omega = rng.standard_normal((200, 30))
    r = cb.rsvd(a, 20, 5, 10, omega=omega, schedule="reference")
    s = cb.rsvd(a, 20, 5, 10, omega=omega, schedule="stabilised")
    assert ref_rsvd.sigma_rel_err(r[1], s[1]) < 1e-10
    assert ref_rsvd.subspace_sine(np.asarray(r[0]), np.asarray(s[0])) < 1e-8


def test_power_iter_parity(cb):
    rng = np.random.default_rng(106)
    a = rng.standard_normal((2500, 150))
    omega = rng.standard_normal((150, 22))
    q0 = ref_rsvd.power_iter(a, 22, 5, omega=omega)
    q = cb.power_iter(a, 22, 5, omega=omega)
    assert q.shape == (2500, 22)
    assert np.max(np.abs(q.T @ q - np.eye(22))) < 1e-12
    assert ref_rsvd.subspace_sine(q0, q) < TOL_ANGLE


# ------------------------------------------------------------------ error behaviour
def test_rank_panic_and_limits(cb):
    a = np.random.default_rng(107).standard_normal((64, 9))
    with pytest.raises(IndexError):                      # k > l = min(k + p, 9): the reference panics
        cb.rsvd(a, 10, 2, 10)
    out = cb.rsvd(a, 6, 5, 10, seed=3)                   # l clamps to 9
    assert out[0].shape == (64, 6) and out[2].shape == (6, 9)
    with pytest.raises(cb.CorrlaError) as ei:            # sketch width limit of the panel path (2048 columns)
        cb.rsvd(np.zeros((3000, 2500)), 2040, 1, 10)
    assert ei.value.status == -4
    # all-zero input without power iterations (with them the reference divides 0 by ||Y|| = 0): sigma = 0, U and V are
    # whatever orthonormal completion the QR returns -- single-panel and panel path
    for k in (20, 120):
        u, s, vt = cb.rsvd(np.zeros((400, 300)), k, 0, 10, seed=1)
        assert np.all(s == 0.0) and np.all(np.isfinite(u)) and np.all(np.isfinite(vt))
        assert np.max(np.abs(u.T @ u - np.eye(k))) < 1e-12


def test_input_is_not_modified_and_deterministic(cb):
    rng = np.random.default_rng(108)
    a = rng.standard_normal((2000, 100))
    keep = a.copy()
    o1 = cb.rsvd(a, 10, 4, 10, seed=77)
    o2 = cb.rsvd(a, 10, 4, 10, seed=77)
    o3 = cb.rsvd(a, 10, 4, 10, seed=78)
    assert np.array_equal(a, keep)
    for x, y in zip(o1, o2):
        assert np.array_equal(np.asarray(x), np.asarray(y))     # bit-reproducible (no atomics on the data path)
    assert not np.array_equal(np.asarray(o1[0]), np.asarray(o3[0]))
    assert ref_rsvd.sigma_rel_err(o1[1], o3[1]) < 5e-2           # different Omega, same (flat) spectrum estimate


# ------------------------------------------------------------------ the other drop-in rows
@pytest.mark.parametrize("shape,cols,beta", [((300, 200), 110, 1.0), ((257, 129), 7, -0.5), ((64, 1000), 128, 2.0),
                                              ((5, 5), 3, 1.0), ((4000, 18), 18, 1.0)])
def test_par_matmul(cb, shape, cols, beta):
    """par_matmul_helper (mat_utils.rs:20-33): res = beta * lhs * rhs, every layout, 1e-13 relative."""
    rng = np.random.default_rng(109)
    lhs = rng.standard_normal(shape)
    rhs = rng.standard_normal((shape[1], cols))
    ref = beta * (lhs @ rhs)
    scale = np.max(np.abs(ref))
    for l_, r_ in ((lhs, rhs), (np.asfortranarray(lhs), rhs), (lhs, np.asfortranarray(rhs)),
                   (np.ascontiguousarray(lhs.T).T, rhs)):
        res = cb.par_matmul(l_, r_, beta)
        assert res.shape == ref.shape
        assert np.max(np.abs(res - ref)) < 1e-13 * scale * np.sqrt(shape[1])


def test_par_matmul_transposed_operand_on_device(cb):
    import torch
    rng = np.random.default_rng(110)
    a = rng.standard_normal((5000, 300))
    y = rng.standard_normal((5000, 40))
    ad, yd = torch.from_numpy(a).cuda(), torch.from_numpy(y).cuda()
    z = cb.par_matmul(ad.t(), yd)                      # A^T * Y: random_svd.rs:42-46
    torch.cuda.synchronize()
    ref = a.T @ y
    assert np.max(np.abs(z.cpu().numpy() - ref)) < 1e-12 * np.max(np.abs(ref))


def test_random_mat_normal_matches_philox_restatement(cb):
    for rows, cols, seed in ((1024, 110, 42), (7, 3, 1), (1, 1, 5), (333, 18, 2**40 + 17)):
        z = cb.random_mat_normal(rows, cols, seed)
        z0 = ref_rsvd.philox_normal(rows, cols, seed)
        assert z.shape == (rows, cols)
        assert np.max(np.abs(z - z0)) < 1e-13
    assert np.array_equal(cb.random_mat_normal(64, 8, 9), cb.random_mat_normal(64, 8, 9))
    assert not np.array_equal(cb.random_mat_normal(64, 8, 9), cb.random_mat_normal(64, 8, 10))


def test_rsvd_with_engine_omega_equals_injected(cb):
    """seed path == injecting the same Omega fetched from the generator."""
    rng = np.random.default_rng(111)
    a = rng.standard_normal((1500, 90))
    omega = cb.random_mat_normal(90, 22, 1234)
    o1 = cb.rsvd(a, 12, 4, 10, seed=1234)
    o2 = cb.rsvd(a, 12, 4, 10, omega=omega, seed=1234)     # same seed: it also keys the QR sketch
    for x, y in zip(o1, o2):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    assert_parity(o1, ref_rsvd.random_svd(a, 12, 4, 10, omega=omega), 12)


# ------------------------------------------------------------------ CholeskyQR (thin Q) robustness
def test_thin_q_well_conditioned(cb):
    rng = np.random.default_rng(112)
    a = rng.standard_normal((5000, 110))
    q, rank = cb.thin_q(a, return_rank=True)
    assert rank == 110
    assert np.max(np.abs(q.T @ q - np.eye(110))) < 1e-13
    assert np.linalg.norm(a - q @ (q.T @ a)) < 1e-12 * np.linalg.norm(a)
    q0, _ = np.linalg.qr(a)
    assert ref_rsvd.subspace_sine(q0, q) < 1e-12


@pytest.mark.parametrize("kappa", [1e4, 1e8, 1e12])
def test_thin_q_ill_conditioned(cb, kappa):
    """cond up to 1e12: the shifted third pass must kick in and still deliver orthonormality at machine precision."""
    rng = np.random.default_rng(113)
    m, l = 4000, 40
    u, _ = np.linalg.qr(rng.standard_normal((m, l)))
    v, _ = np.linalg.qr(rng.standard_normal((l, l)))
    a = (u * np.logspace(0, -np.log10(kappa), l)) @ v.T
    q = cb.thin_q(a)
    assert np.max(np.abs(q.T @ q - np.eye(l))) < 1e-12
    assert np.linalg.norm(a - q @ (q.T @ a)) < 1e-11 * np.linalg.norm(a)


@pytest.mark.parametrize("kappa", [30.0, 1e3, 3e3, 8e3, 2e4])
def test_thin_q_near_the_fast_path_threshold(cb, kappa):
    """cond(Y) around the probe threshold (~1e4) on a tall matrix: the CholeskyQR fast path replaces the second Cholesky
    by its first-order expansion only while the measured ||G2 - I||_F is below 2e-8; above it the real factorisation runs
    behind a device flag.  Orthonormality must stay at machine precision on both sides of that switch."""
    rng = np.random.default_rng(int(kappa))
    m, l = 300_000, 96
    u, _ = np.linalg.qr(rng.standard_normal((m, l)))
    v, _ = np.linalg.qr(rng.standard_normal((l, l)))
    a = (u * np.logspace(0, -np.log10(kappa), l)) @ v.T
    q = cb.thin_q(a)
    assert np.max(np.abs(q.T @ q - np.eye(l))) < 1e-13
    assert np.linalg.norm(a - q @ (q.T @ a)) < 1e-12 * np.linalg.norm(a)


def test_thin_q_rank_deficient_is_completed(cb):
    rng = np.random.default_rng(114)
    base = rng.standard_normal((3000, 10))
    a = np.hstack([base, base @ rng.standard_normal((10, 6)), np.zeros((3000, 2))])     # rank 10 of 18 columns
    q, rank = cb.thin_q(a, return_rank=True)
    assert np.max(np.abs(q.T @ q - np.eye(18))) < 1e-12        # completed like a Householder QR would
    assert np.linalg.norm(a - q @ (q.T @ a)) < 1e-11 * np.linalg.norm(a)


# ------------------------------------------------------------------ size-independent properties at larger sizes
def test_large_device_properties(cb):
    """1M x 1024 (8 GB) generated on the device: orthonormality, A^T U = V S, and agreement with a planted spectrum."""
    import torch
    torch.manual_seed(0)
    m, n, r, k, q, p = 1 << 20, 1024, 56, 48, 4, 10     # rank 56 <= l = 58: exact recovery, 2 dependent columns
    u0, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device="cuda"))
    v0, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    sig = 100.0 * 0.97 ** torch.arange(r, dtype=torch.float64, device="cuda")
    a = (u0 * sig) @ v0.T
    u, s, vt = cb.rsvd(a, k, q, p, seed=5)
    torch.cuda.synchronize()
    t = cb.last_timings()
    assert t["passes_over_a"] == 10 and t["gpu_launches"] > 0
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert float((u.T @ u - eye).abs().max()) < 1e-12
    assert float((vt @ vt.T - eye).abs().max()) < 1e-12
    assert float(((s.ravel() - sig[:k]).abs() / sig[:k]).max()) < 1e-10      # planted singular values
    resid = a.T @ u - vt.T * s.ravel()
    assert float(resid.abs().max()) < 1e-9 * float(sig[0])
    pu = u0[:, :k] - u @ (u.T @ u0[:, :k])
    assert float(torch.linalg.matrix_norm(pu, 2)) < 1e-8


# ------------------------------------------------------------------ beyond the parity class: robustness
def test_thin_q_cond_1e14_sketch_preconditioned(cb):
    """cond 1e14: the Cholesky probe must hand over to the sketch-preconditioned stage (Householder QR of a sparse sign
    sketch), which keeps orthonormality at machine precision where CholeskyQR2 cannot."""
    rng = np.random.default_rng(115)
    m, l = 20000, 60
    u, _ = np.linalg.qr(rng.standard_normal((m, l)))
    v, _ = np.linalg.qr(rng.standard_normal((l, l)))
    a = (u * np.logspace(0, -14, l)) @ v.T
    q, rank = cb.thin_q(a, return_rank=True)
    assert rank == l
    assert np.max(np.abs(q.T @ q - np.eye(l))) < 1e-12
    assert np.linalg.norm(a - q @ (q.T @ a)) < 1e-12 * np.linalg.norm(a)


@pytest.mark.parametrize("decay", [1.2, 2.0])
def test_fast_decay_reference_schedule_stays_accurate(cb, decay):
    """sigma_j = decay^-j: after the reference's three raw power iterations cond(Y) is 1e16 and beyond.  The reference
    (Householder) still returns the leading singular values to ~1e-13; the engine must not fall apart either."""
    rng = np.random.default_rng(116)
    m, n, k, q, p = 4000, 300, 20, 4, 10
    uu, _ = np.linalg.qr(rng.standard_normal((m, n)))
    vv, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = decay ** -np.arange(n, dtype=np.float64)
    a = (uu * sig) @ vv.T
    omega = rng.standard_normal((n, k + p))
    u, s, vt = cb.rsvd(a, k, q, p, omega=omega, seed=4)
    assert np.max(np.abs(s.ravel() - sig[:k]) / sig[:k]) < 1e-9
    assert np.max(np.abs(np.asarray(u).T @ np.asarray(u) - np.eye(k))) < 1e-12
    t = cb.last_timings()
    assert t["live_columns"] <= k + p


def test_plain_c_caller_of_the_abi(tmp_path):
    """The C ABI without Python in the way: tools/c_abi_example.c compiled with gcc against include/corrla_b200.h."""
    import shutil
    import subprocess
    root = Path(__file__).resolve().parents[1]
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    exe = tmp_path / "c_abi_example"
    libdir = root / "corrla_rs_b200" / "lib"
    subprocess.run(["gcc", "-O2", f"-I{root / 'include'}", "-o", str(exe), str(root / "tools" / "c_abi_example.c"),
                    f"-L{libdir}", "-lcorrla_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI example OK" in r.stdout


def test_c_replay_of_the_rust_wrapper_call_sequences(tmp_path):
    """tools/c_abi_replay.c makes the calls rust/corrla-b200/src/lib.rs makes (default opts, NULL timings, NULL optional
    outputs, faer column-major strides): the Rust crate cannot be compiled in this image, its call shapes can be run."""
    import shutil
    import subprocess
    root = Path(__file__).resolve().parents[1]
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    exe = tmp_path / "c_abi_replay"
    libdir = root / "corrla_rs_b200" / "lib"
    subprocess.run(["gcc", "-O2", f"-I{root / 'include'}", "-o", str(exe), str(root / "tools" / "c_abi_replay.c"),
                    f"-L{libdir}", "-lcorrla_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI replay OK" in r.stdout


# ------------------------------------------------------------------ degenerate shapes and parameters
@pytest.mark.parametrize("shape,kqp", [((1, 1), (1, 2, 0)), ((2, 1), (1, 0, 0)), ((7, 3), (3, 1, 5)), ((3, 7), (2, 2, 1)),
                                        ((40, 40), (40, 3, 0)), ((500, 17), (1, 0, 0)), ((129, 128), (118, 2, 10)),
                                        ((33, 16), (16, 12, 16))])
def test_degenerate_shapes(cb, shape, kqp):
    """1x1, single column, n_iters = 0, k = l = n, l = 128 (the kernel limit), rows not a multiple of anything."""
    rng = np.random.default_rng(117)
    a = rng.standard_normal(shape)
    k, q, p = kqp
    thin_cols = min(shape)
    l = min(k + p, thin_cols)
    omega = rng.standard_normal((thin_cols, l))
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    out = cb.rsvd(a, k, q, p, omega=omega, seed=1)
    u, s, vt = (np.asarray(x) for x in out)
    assert u.shape == ref[0].shape and s.shape == (k, 1) and vt.shape == ref[2].shape
    assert np.max(np.abs(s - ref[1])) < 1e-10 * max(1.0, float(ref[1][0, 0]))
    # reconstruction of the rank-k part agrees (vectors themselves may differ by sign / rotation inside clusters)
    rec = u @ np.diag(s.ravel()) @ vt
    rec0 = ref[0] @ np.diag(ref[1].ravel()) @ ref[2]
    assert np.max(np.abs(rec - rec0)) < 1e-8 * max(1.0, float(ref[1][0, 0]))


def test_context_reuse_across_shapes(cb):
    """One context, shrinking and growing problems back to back: cached buffers must be re-zeroed / regrown correctly."""
    rng = np.random.default_rng(118)
    ctx = cb.Context()
    for shape, (k, q, p) in [((3000, 200), (20, 4, 10)), ((100, 50), (5, 2, 3)), ((5000, 300), (60, 5, 10)),
                             ((64, 9), (6, 5, 10)), ((3000, 200), (20, 4, 10))]:
        a = rng.standard_normal(shape)
        omega = rng.standard_normal((shape[1], min(k + p, shape[1])))
        assert_parity(cb.rsvd(a, k, q, p, omega=omega, ctx=ctx, seed=2), ref_rsvd.random_svd(a, k, q, p, omega=omega), k)
    ctx.close()


def test_bench_line_contract_small(tmp_path):
    """bench.py on a reduced row count: one JSON line on stdout carrying every key of the contract."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--rows", "65536", "--steps", "2", "--warmup", "3",
                        "--cpu-sample-rows", "4096", "--e2e-steps", "1"], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] == 65536 * 1024 * 8 and d["e2e"]["d2h_bytes_per_step"] > 0


def test_full_size_c3_properties(cb):
    """BASELINE config C3 at full size (4 194 304 x 1024, k 100, q 4, p 10) on the device, through size-independent
    properties: planted spectrum recovered, U and V orthonormal, A^T U = V S.  Needs ~45 GB of HBM."""
    import torch
    free, _total = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("not enough free device memory for the full-size case")
    torch.manual_seed(1)
    m, n, r, k, q, p = 4_194_304, 1024, 104, 100, 4, 10          # rank 104 <= l = 110: exact recovery
    u0, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device="cuda"))
    v0, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    sig = 50.0 * 0.985 ** torch.arange(r, dtype=torch.float64, device="cuda")
    a = torch.empty((m, n), dtype=torch.float64, device="cuda")
    step = 1 << 19
    for r0 in range(0, m, step):                                  # build A in row blocks to bound temporaries
        a[r0:r0 + step] = (u0[r0:r0 + step] * sig) @ v0.T
    u, s, vt = cb.rsvd(a, k, q, p, seed=11)
    torch.cuda.synchronize()
    t = cb.last_timings()
    assert t["passes_over_a"] == 10 and t["pass_launches"] == 10
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert float((u.T @ u - eye).abs().max()) < 1e-12
    assert float((vt @ vt.T - eye).abs().max()) < 1e-12
    assert float(((s.ravel() - sig[:k]).abs() / sig[:k]).max()) < 1e-10
    resid = a.T @ u - vt.T * s.ravel()
    assert float(resid.abs().max()) < 1e-9 * float(sig[0])
    del a, u0, resid
    torch.cuda.empty_cache()


def test_full_size_c2_properties(cb):
    """BASELINE config C2 (20 000 x 20 000, k 100, q 4, p 10): split-K on both products, Z of 18 MB."""
    import torch
    torch.manual_seed(2)
    m = n = 20000
    r, k, q, p = 104, 100, 4, 10
    u0, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device="cuda"))
    v0, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device="cuda"))
    sig = 50.0 * 0.985 ** torch.arange(r, dtype=torch.float64, device="cuda")
    a = (u0 * sig) @ v0.T
    u, s, vt = cb.rsvd(a, k, q, p, seed=12)
    torch.cuda.synchronize()
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert float((u.T @ u - eye).abs().max()) < 1e-12
    assert float((vt @ vt.T - eye).abs().max()) < 1e-12
    assert float(((s.ravel() - sig[:k]).abs() / sig[:k]).max()) < 1e-10
    assert float(torch.linalg.matrix_norm(v0[:, :k] - vt.T @ (vt @ v0[:, :k]), 2)) < 1e-8


def test_fuzz_random_shapes_layouts_parameters(cb):
    """40 seeded random problems: shapes from 1 to a few thousand, every layout, random (k, q, p); singular values and the
    rank-k reconstruction against the oracle (vectors can differ by sign / rotation inside clusters)."""
    rng = np.random.default_rng(2026)
    for case in range(40):
        m = int(rng.integers(1, 2500))
        n = int(rng.integers(1, 260))
        if rng.random() < 0.3:
            m, n = n, m                                        # fat
        thin_cols = min(m, n)
        k = int(rng.integers(1, min(thin_cols, 100) + 1))
        p = int(rng.integers(0, 20))
        if k + p > 128:
            p = 128 - k
        q = int(rng.integers(0, 7))
        kind = rng.integers(0, 3)
        if kind == 0:
            a = rng.standard_normal((m, n))
        elif kind == 1:                                        # low rank + noise
            r = int(rng.integers(1, thin_cols + 1))
            a = rng.standard_normal((m, r)) @ rng.standard_normal((r, n)) + 1e-3 * rng.standard_normal((m, n))
        else:                                                  # exactly rank deficient
            r = max(1, thin_cols // 3)
            a = rng.standard_normal((m, r)) @ rng.standard_normal((r, n))
        layout = rng.integers(0, 4)
        if layout == 1:
            a_in = np.asfortranarray(a)
        elif layout == 2:                                      # strided view (every other column of a wider buffer)
            wide = np.zeros((m, 2 * n)); wide[:, ::2] = a; a_in = wide[:, ::2]
        elif layout == 3:                                      # offset sub-view with odd pitch
            wide = np.zeros((m + 1, n + 3)); wide[1:, 1:n + 1] = a; a_in = wide[1:, 1:n + 1]
        else:
            a_in = a
        l = min(k + p, thin_cols)
        omega = rng.standard_normal((thin_cols, l))
        ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
        u, s, vt = (np.asarray(x) for x in cb.rsvd(a_in, k, q, p, omega=omega, seed=case))
        tag = f"case {case}: {m}x{n} k={k} q={q} p={p} kind={kind} layout={layout}"
        assert u.shape == (m, k) and s.shape == (k, 1) and vt.shape == (k, n), tag
        s0 = ref[1].ravel()
        scale = max(float(s0[0]), 1e-300)
        # parity class: direction j survives the reference's raw power iterations with relative weight
        # (sigma_j / sigma_1)^(2 min(q, 3) + 1) in Y; below ~1e-6 the reference itself is not reproducible, so only the
        # components above that are compared tightly (the rest must merely be finite and no larger than the reference's)
        expo = 2 * min(q, 3) + 1
        elig = (s0 / scale) ** expo > 1e-6
        ne = int(np.sum(elig))
        assert ne >= 1, tag
        assert np.max(np.abs(s.ravel()[:ne] - s0[:ne])) < 1e-9 * scale, tag
        rec = u[:, :ne] @ np.diag(s.ravel()[:ne]) @ vt[:ne]
        rec0 = ref[0][:, :ne] @ np.diag(s0[:ne]) @ ref[2][:ne]
        gap_ok = ne == k or (s0[ne - 1] - s0[ne]) > 1e-3 * scale      # a cluster cut in two is not comparable
        if gap_ok:
            assert np.max(np.abs(rec - rec0)) < 1e-6 * scale, tag
        assert np.all(np.isfinite(u)) and np.all(np.isfinite(vt)) and np.all(np.isfinite(s)), tag
        assert np.all(s.ravel() <= 1.001 * scale + 1e-300), tag


# ------------------------------------------------------------------ streamed host input
@pytest.mark.parametrize("layout", ["row", "col"])
@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("q,schedule", [(4, "reference"), (0, "reference"), (5, "reference"), (2, "stabilised")])
def test_streamed_host_input_matches_resident(cb, monkeypatch, layout, pinned, q, schedule):
    """Host matrices above the streaming threshold are copied in row chunks with the first product(s) running behind
    the copies (engine.cu stream_in).  Forced here with a 256-row chunk on a small matrix (ragged last chunk): same
    factors as the copy-then-compute path up to summation order, and the same parity with the oracle."""
    import torch
    rng = np.random.default_rng(77)
    m, n, k, p = 2000 + 72, 96, 12, 6
    a = lowrank_noise(rng, m, n, 20, 1e-3)
    omega = rng.standard_normal((n, k + p))
    if layout == "col":
        a = np.asfortranarray(a)
    if pinned:
        t = torch.from_numpy(a.T if layout == "col" else a).contiguous().pin_memory()
        host = t.numpy().T if layout == "col" else t.numpy()
        assert np.array_equal(host, a)
    else:
        host = a
    monkeypatch.setenv("CORRLA_B200_STREAM_ROWS", "0")
    base = cb.rsvd(host, k, q, p, omega=omega, schedule=schedule)
    assert cb.last_timings()["streamed_chunks"] == 0
    monkeypatch.setenv("CORRLA_B200_STREAM_ROWS", "256")
    out = cb.rsvd(host, k, q, p, omega=omega, schedule=schedule)
    t = cb.last_timings()
    assert t["streamed_chunks"] == 9
    assert t["pass_launches"] == 2 + 2 * q - (2 if (q >= 1 and schedule == "reference") else 1)
    assert_parity(out, base, k, tol_sigma=1e-12, tol_angle=1e-10)
    if schedule == "reference":
        assert_parity(out, ref_rsvd.random_svd(a, k, q, p, omega=omega), k)


def test_streamed_host_input_drawn_omega_and_power_iter(cb, monkeypatch):
    """Streaming with the engine's own Philox Omega (same seed => same Omega as the resident path) and through
    corrla_power_iter_f64."""
    rng = np.random.default_rng(78)
    m, n, k, p, q = 3000, 64, 10, 6, 3
    a = lowrank_noise(rng, m, n, 16, 1e-4)
    monkeypatch.setenv("CORRLA_B200_STREAM_ROWS", "0")
    base = cb.rsvd(a, k, q, p, seed=5)
    qb = cb.power_iter(a, 16, q, seed=5)
    monkeypatch.setenv("CORRLA_B200_STREAM_ROWS", "512")
    out = cb.rsvd(a, k, q, p, seed=5)
    assert cb.last_timings()["streamed_chunks"] == 6
    assert_parity(out, base, k, tol_sigma=1e-12, tol_angle=1e-10)
    qs = cb.power_iter(a, 16, q, seed=5)
    assert cb.last_timings()["streamed_chunks"] == 6
    assert ref_rsvd.subspace_sine(qb, qs) < 1e-10
    assert np.max(np.abs(qs.T @ qs - np.eye(16))) < 1e-12


# ------------------------------------------------------------------ sketches wider than one 128-column GEMM tile
@pytest.mark.parametrize("shape,kqp,kind", [
    ((3000, 400), (150, 4, 10), "lowrank"),       # l = 160: 2 panels of 80
    ((2000, 300), (250, 2, 20), "gauss"),         # l = 270: 3 panels of 96, the last one with 78 live columns
    ((300, 2500), (140, 3, 10), "lowrank"),       # fat input
    ((1500, 200), (190, 5, 30), "lowrank"),       # l clamps to n = 200: the sketch spans the whole row space
    ((5000, 520), (129, 1, 0), "gauss"),          # l = 129: just over the single-panel limit
])
def test_wide_sketch_matches_oracle(cb, shape, kqp, kind):
    """n_rank + n_oversamples > 128 runs in column panels (csrc/wide.cuh): same parity bar as the single-panel path."""
    m, n = shape
    k, q, p = kqp
    rng = np.random.default_rng(m + n + k)
    if kind == "gauss":
        a = rng.standard_normal((m, n))
    else:
        r = min(m, n)
        u, _ = np.linalg.qr(rng.standard_normal((m, r)))
        v, _ = np.linalg.qr(rng.standard_normal((n, r)))
        a = (u * (10.0 * 0.985 ** np.arange(r))) @ v.T
    l = min(k + p, min(m, n))
    omega = rng.standard_normal((min(m, n), l))
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    out = cb.rsvd(a, k, q, p, omega=omega)
    t = cb.last_timings()
    assert t["pass_launches"] == (2 + 2 * q) * (-(-(8 * (-(-l // 8))) // 128))
    assert_parity(out, ref, k)
    assert np.max(np.abs((out[0] * out[1].ravel()) @ out[2] - (ref[0] * ref[1].ravel()) @ ref[2])) < 1e-9 * ref[1][0, 0]


def test_wide_sketch_device_seeded_stabilised_and_power_iter(cb):
    import torch
    rng = np.random.default_rng(91)
    m, n, k, p, q = 6000, 384, 180, 12, 3
    uu, _ = np.linalg.qr(rng.standard_normal((m, n)))
    vv, _ = np.linalg.qr(rng.standard_normal((n, n)))
    a = (uu * (10.0 * 0.99 ** np.arange(n))) @ vv.T                     # slow decay: inside the parity class
    ad = torch.from_numpy(a).cuda()
    omega = cb.random_mat_normal(n, k + p, seed=17)                      # the Omega the engine draws for this seed
    ref = ref_rsvd.random_svd(a, k, q, p, omega=omega)
    out = cb.rsvd(ad, k, q, p, seed=17)
    torch.cuda.synchronize()
    assert_parity(tuple(x.cpu().numpy() for x in out), ref, k)
    stab = cb.rsvd(ad, k, q, p, seed=17, schedule="stabilised")
    assert ref_rsvd.sigma_rel_err(ref[1], stab[1].cpu().numpy()) < 1e-9
    qd = cb.power_iter(ad, k + p, q, seed=17).cpu().numpy()
    assert qd.shape == (m, k + p)
    assert np.max(np.abs(qd.T @ qd - np.eye(k + p))) < 1e-12
    qref = ref_rsvd.power_iter(a, k + p, q, omega=omega)
    assert ref_rsvd.subspace_sine(qref, qd) < 1e-8


def test_wide_sketch_rank_deficient_and_pca(cb):
    """Exactly rank-60 input with a 160-column sketch (whole panels are numerically dependent), and the fused centring
    through the panel path (rpca with n_rank = 130 => l = 140)."""
    from oracle import ref_pca
    rng = np.random.default_rng(92)
    m, n = 2500, 300
    a = rng.standard_normal((m, 60)) @ rng.standard_normal((60, n))
    u, s, vt = cb.rsvd(a, 150, 4, 10, seed=3)
    s0 = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s.ravel()[:60] - s0[:60])) < 1e-10 * s0[0]
    assert np.max(np.abs(s.ravel()[60:])) < 1e-9 * s0[0]
    assert np.max(np.abs((u * s.ravel()) @ vt - a)) < 1e-9 * s0[0]
    # the live singular subspaces, and a rank just above one panel (150 of a 160-column sketch: columns of the second panel
    # collapse in the projection against the first) -- 1e-6 before the re-projection of panels that took the robust stage
    ux, _, vxt = np.linalg.svd(a, full_matrices=False)
    assert ref_rsvd.subspace_sine(ux[:, :60], np.asarray(u)[:, :60]) < 1e-9
    a2 = rng.standard_normal((m, 150)) @ rng.standard_normal((150, n))
    u2, s2, vt2 = cb.rsvd(a2, 150, 4, 10, seed=5)
    ux2, sx2, vxt2 = np.linalg.svd(a2, full_matrices=False)
    assert np.max(np.abs(s2.ravel() - sx2[:150])) < 1e-10 * sx2[0]
    assert np.max(np.abs(np.asarray(u2).T @ np.asarray(u2) - np.eye(150))) < 1e-12
    assert ref_rsvd.subspace_sine(ux2[:, :150], np.asarray(u2)) < 1e-8
    assert ref_rsvd.subspace_sine(vxt2[:150].T, np.asarray(vt2).T) < 1e-8
    assert np.max(np.abs((np.asarray(u2) * s2.ravel()) @ np.asarray(vt2) - a2)) < 1e-10 * sx2[0]
    x = lowrank_noise(rng, 3000, 260, 180, 1e-6) + rng.standard_normal((1, 260))
    sv, comps = cb.rpca(x, 130, seed=4)
    xc = x - x.mean(axis=0)
    sc0 = np.linalg.svd(xc, compute_uv=False)
    assert np.max(np.abs(sv.ravel() - sc0[:130]) / sc0[:130]) < 1e-9
    _, vt0 = ref_pca.rpca(x, 130, omega=rng.standard_normal((260, 140)))
    assert ref_rsvd.subspace_sine(vt0.T, comps.T) < 1e-7


def test_wide_par_matmul_and_thin_q(cb):
    rng = np.random.default_rng(93)
    lhs, rhs = rng.standard_normal((1500, 77)), rng.standard_normal((77, 300))
    out = cb.par_matmul(lhs, rhs, beta=-1.5)
    assert out.shape == (1500, 300)
    assert np.max(np.abs(out - (-1.5) * (lhs @ rhs))) < 1e-11
    x = rng.standard_normal((4000, 200)) * (0.97 ** np.arange(200))
    q = cb.thin_q(x)
    assert q.shape == (4000, 200)
    assert np.max(np.abs(q.T @ q - np.eye(200))) < 1e-12
    assert ref_rsvd.subspace_sine(np.linalg.qr(x)[0], q) < 1e-9
    q0, _ = np.linalg.qr(x[:, :90])
    assert ref_rsvd.subspace_sine(q0, q[:, :90]) < 1e-9          # nested: leading columns span the leading columns


def test_device_input_is_ordered_after_its_producer(cb):
    """A tensor that is still being written by kernels queued on torch's current (legacy default) stream: the engine
    must enqueue behind them (cudaStreamLegacy), not on its own unordered stream."""
    import torch
    torch.manual_seed(5)
    base = torch.randn(300_000, 128, dtype=torch.float64, device="cuda")
    u0, s0, v0 = cb.rsvd(base * 3.0 + 1.0, 10, 2, 6, seed=2)
    torch.cuda.synchronize()
    for _ in range(3):
        a = base.clone()
        for _ in range(20):                       # a queue of cheap in-place kernels that nets out to 3*base + 1
            a.mul_(2.0).mul_(0.5)
        a.mul_(3.0).add_(1.0)
        u, s, vt = cb.rsvd(a, 10, 2, 6, seed=2)   # no synchronize in between
        assert torch.equal(s, s0) and torch.equal(u, u0)


def test_concurrent_callers(cb):
    """The boundary is synchronous and re-entrant: threads sharing the default context are serialised
    by its lock, threads with their own context run concurrently; every result equals the sequential one bit for bit."""
    import threading
    rng = np.random.default_rng(94)
    mats = [rng.standard_normal((3000 + 500 * i, 96 + 8 * i)) for i in range(4)]
    seq = [cb.rsvd(a, 12, 3, 6, seed=40 + i) for i, a in enumerate(mats)]
    for own_ctx in (False, True):
        res, errs = [None] * 4, []

        def work(i):
            try:
                ctx = cb.Context(0) if own_ctx else None
                for _ in range(3):
                    res[i] = cb.rsvd(mats[i], 12, 3, 6, seed=40 + i, ctx=ctx)
                if ctx is not None:
                    ctx.close()
            except Exception as exc:          # noqa: BLE001
                errs.append(exc)

        th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for (u0, s0, v0), (u, s, v) in zip(seq, res):
            assert np.array_equal(s0, s) and np.array_equal(u0, u) and np.array_equal(v0, v)


def test_cuda_array_interface_input(cb):
    """Device arrays from other libraries (cupy, numba) are accepted through __cuda_array_interface__ without a copy."""
    import torch

    class Foreign:
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    rng = np.random.default_rng(95)
    a = rng.standard_normal((4000, 64))
    omega = rng.standard_normal((64, 20))
    t = torch.from_numpy(a).cuda()
    u, s, vt = cb.rsvd(Foreign(t), 12, 3, 8, omega=omega)
    torch.cuda.synchronize()
    assert u.is_cuda
    assert_parity(tuple(x.cpu().numpy() for x in (u, s, vt)), ref_rsvd.random_svd(a, 12, 3, 8, omega=omega), 12)
