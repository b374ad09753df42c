"""CPU suite: the C-ABI shared library loads and exports exactly what include/corrla_b200.h declares; the host
logic (argument checks, error mapping) behaves like the reference binding; and the product path fails loudly
(never falls back to the CPU) when no GPU is present."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "corrla_b200.h").read_text()
    return sorted(set(re.findall(r"CORRLA_API[^;]*?\b(corrla_\w+)\s*\(", text)))


def test_header_symbols_all_exported(built_lib):
    from corrla_rs_b200 import _ffi
    names = declared_symbols()
    assert len(names) >= 19
    assert sorted(_ffi.SYMBOLS) == names            # the ctypes table binds the header 1:1
    for n in names:
        assert getattr(built_lib, n) is not None


def test_status_strings_and_version(built_lib):
    assert built_lib.corrla_status_str(0) == b"ok"
    assert b"panics" in built_lib.corrla_status_str(-2) or b"n_rank" in built_lib.corrla_status_str(-2)
    assert b"sm_100a" in built_lib.corrla_version()


def test_struct_layout_matches_header(built_lib):
    """opts_default() must leave device = -1 and everything else zero: catches field-order drift between the
    ctypes mirror and the C struct."""
    import ctypes as C
    from corrla_rs_b200 import _ffi
    o = _ffi.RsvdOpts()
    C.memset(C.byref(o), 0xFF, C.sizeof(o))
    built_lib.corrla_rsvd_opts_default(C.byref(o))
    assert o.device == -1 and o.seed == 0 and o.omega is None and o.schedule == 0
    assert o.a_on_device == 0 and o.out_on_device == 0 and o.ctx is None and o.comm is None and o.global_rows == 0
    assert o.center == 0
    assert C.sizeof(o) == 96


def test_type_errors_like_pyo3(built_lib):
    import corrla_rs
    with pytest.raises(TypeError):
        corrla_rs.rsvd(np.zeros((4, 4), dtype=np.float32), 1, 1, 1)       # PyReadonlyArray2<f64>
    with pytest.raises(TypeError):
        corrla_rs.rsvd(np.zeros(4), 1, 1, 1)
    with pytest.raises(TypeError):
        corrla_rs.rsvd([[1.0, 2.0]], 1, 1, 1)
    with pytest.raises(OverflowError):
        corrla_rs.rsvd(np.zeros((4, 4)), -1, 1, 1)                        # usize extraction
    with pytest.raises(NotImplementedError):
        corrla_rs.cs_dirichlet_sample
    assert callable(corrla_rs.PyDMDc) and callable(corrla_rs.PyPodI) and callable(corrla_rs.PyRbfInterp)
    assert callable(corrla_rs.active_ss)


def test_no_cpu_fallback_without_gpu(built_lib):
    """On a box without a GPU every compute entry point must fail with CORRLA_ERR_NO_DEVICE."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    import corrla_rs_b200 as cb
    a = np.random.default_rng(0).standard_normal((32, 8))
    for call in (lambda: cb.rsvd(a, 2, 2, 2), lambda: cb.power_iter(a, 3, 2), lambda: cb.par_matmul(a, a.T.copy()[:, :4]),
                 lambda: cb.random_mat_normal(4, 4, 1), lambda: cb.thin_q(a),
                 lambda: cb.DMDc(a, np.ones((1, 8)), 1.0, 2, 2), lambda: cb.PodI(a.T.copy(), np.arange(8.0).reshape(8, 1), 2),
                 lambda: cb.mat_cov_centered(a), lambda: cb.active_ss(a, a[:, :1].copy(), 1, 12, 2)):
        with pytest.raises(cb.CorrlaError) as ei:
            call()
        assert ei.value.status == -7


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under corrla_rs_b200/ or corrla_rs/ may reference it."""
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|#\s*include.*oracle)|oracle\.\w+\(|dlopen\([^)]*oracle", re.M)
    for pkg in ("corrla_rs_b200", "corrla_rs"):
        for f in (ROOT / pkg).rglob("*"):
            if f.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
                assert pat.search(f.read_text()) is None, f


def test_host_alloc_roundtrip(built_lib):
    """corrla_host_alloc / corrla_host_free and the numpy wrapper used for large outputs."""
    import gc
    import numpy as np
    import corrla_rs_b200 as cb
    arr = cb._huge_empty(3 * (1 << 20))
    assert arr is not None and arr.shape == (3 * (1 << 20),) and arr.dtype == np.float64
    arr[:] = 1.5
    view = arr.reshape((3, 1 << 20), order="F")
    assert float(view.sum()) == 1.5 * 3 * (1 << 20)
    del arr, view
    gc.collect()


def test_host_side_mirrors_without_gpu():
    """The parts of the DMDc / POD / active-subspace mirrors that are host work by design (small dense algebra) against
    the oracle's restatement -- no device needed."""
    from corrla_rs_b200.rom import RbfInterp
    from corrla_rs_b200.stats import FittedActiveSsRsvd
    from oracle import ref_rom, ref_stats
    rng = np.random.default_rng(21)
    t = np.sort(rng.uniform(0.0, 5.0, 17)).reshape(-1, 1)
    w = np.stack([np.sin(t[:, 0]), t[:, 0] ** 2, np.exp(-t[:, 0])], axis=1)
    ours = RbfInterp(1, 0.0, 1, 1)
    ours.fit(t, w)                                                     # all weight columns at once
    for j in range(3):
        ref = ref_rom.RbfInterpLin(1)
        ref.fit(t, w[:, j])
        tq = np.array([[1.234]])
        assert abs(ours.predict(tq)[0, j] - ref.predict(tq)[0, 0]) < 1e-12
        assert np.max(np.abs(ours.predict(t)[:, j] - w[:, j])) < 1e-8     # interpolates the nodes
    x2 = rng.standard_normal((25, 2))
    y2 = np.sin(x2[:, 0]) + x2[:, 1]
    for kernel_type, param in ((2, 1.0), (3, 0.0), (0, 0.7)):          # multiquadric, cubic, Gaussian
        f = RbfInterp(kernel_type, param, 2, 1)
        f.fit(x2, y2)
        assert np.max(np.abs(f.predict(x2).ravel() - y2)) < 1e-6
    with pytest.raises(NotImplementedError):
        RbfInterp(1, 0.0, 2, 2)
    g = rng.standard_normal((5, 400))
    ref = ref_stats.ActiveSsRsvd(None, 2).fit_gradients(g)
    fit = FittedActiveSsRsvd(ref.components_, ref.singular_vals_, 2)
    assert np.allclose(fit.var_diag_evd_sensi(), ref.var_diag_evd_sensi())
    xs = rng.standard_normal((7, 5))
    assert fit.transform(xs).shape == (7, 2) and fit.inv_transform(fit.transform(xs)).shape == (7, 5)
    assert np.allclose(fit.components(), ref.components()) and np.allclose(fit.singular_vals(), ref.singular_vals())
