"""GPU tests of the failure and housekeeping paths around the hot path: device-side failures must come back as a
status code whether or not the caller passed a timings struct (the Rust wrapper and the C example pass NULL),
cached buffers can be released, and caller-owned outputs (`out=`) are written in place."""
import ctypes as C
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def cb():
    import corrla_rs_b200
    corrla_rs_b200._ffi.load()
    return corrla_rs_b200


_RAW_CALL = textwrap.dedent("""
    import ctypes as C, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    from corrla_rs_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(3)
    m, n, k, p = 600, 96, 40, 8                      # l = 48 >= 32: the cluster Jacobi kernel runs
    a = rng.standard_normal((m, n))
    u = np.zeros((m, k), order="F"); s = np.zeros(k); vt = np.zeros((k, n), order="F")
    o = _ffi.RsvdOpts(); lib.corrla_rsvd_opts_default(C.byref(o)); o.seed = 1
    t = _ffi.Timings()
    st_null = lib.corrla_rsvd_f64(a.ctypes.data, m, n, n, 1, k, 2, p, C.byref(o), u.ctypes.data, s.ctypes.data,
                                  vt.ctypes.data, None)                  # timings = NULL, like the Rust wrapper
    st_tm = lib.corrla_rsvd_f64(a.ctypes.data, m, n, n, 1, k, 2, p, C.byref(o), u.ctypes.data, s.ctypes.data,
                                vt.ctypes.data, C.byref(t))
    print("STATUS", st_null, st_tm, lib.corrla_last_error().decode())
""")


def _run_raw(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", _RAW_CALL.format(root=str(ROOT))], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("STATUS")][-1].split(None, 3)
    return int(line[1]), int(line[2]), (line[3] if len(line) > 3 else "")


def test_jacobi_cluster_timeout_is_reported_without_timings():
    """CORRLA_B200_TEST_JACOBI_SPIN_LIMIT=0 makes every bounded wait of the cluster Jacobi kernel give up at once: the
    call must fail with CORRLA_ERR_CUDA even with timings == NULL."""
    st_null, st_tm, msg = _run_raw({"CORRLA_B200_TEST_JACOBI_SPIN_LIMIT": "0"})
    assert st_null == -3 and st_tm == -3, (st_null, st_tm, msg)
    assert "Jacobi" in msg
    ok_null, ok_tm, _ = _run_raw({})
    assert ok_null == 0 and ok_tm == 0


def test_jacobi_failure_in_the_middle_of_the_iteration_fails_cleanly():
    """CORRLA_B200_TEST_JACOBI_FAIL_SWEEP=1: one warp of the ring Jacobi kernel gives up at the start of the second sweep, as
    after a timed-out wait -- the failure path taken from INSIDE the iteration, with its neighbours still waiting for the
    columns it will never send.  The call must come back (no hang, the waits poll the failure flag) with an error status,
    never with CORRLA_OK and garbage."""
    st_null, st_tm, msg = _run_raw({"CORRLA_B200_TEST_JACOBI_FAIL_SWEEP": "1"})
    assert st_null == -3 and st_tm == -3, (st_null, st_tm, msg)
    assert "Jacobi" in msg


def test_ctx_trim_releases_and_recovers(cb):
    rng = np.random.default_rng(4)
    a = rng.standard_normal((20000, 256))
    omega = rng.standard_normal((256, 30))
    ctx = cb.Context()
    ref = cb.rsvd(a, 20, 3, 10, omega=omega, ctx=ctx)
    freed = ctx.release_buffers()
    assert freed >= a.nbytes                       # at least the device copy of A went back to the driver
    assert ctx.release_buffers() == 0              # nothing left to release
    again = cb.rsvd(a, 20, 3, 10, omega=omega, ctx=ctx)
    for x, y in zip(ref, again):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    ctx.close()


def test_out_arrays_are_written_in_place(cb):
    import torch
    rng = np.random.default_rng(5)
    m, n, k, q, p = 5000, 300, 12, 3, 8
    a = rng.standard_normal((m, n))
    omega = rng.standard_normal((n, k + p))
    u0, s0, vt0 = cb.rsvd(a, k, q, p, omega=omega)
    # host, pageable and pinned
    u = np.empty((m, k), order="F"); s = np.empty((k, 1), order="F"); vt = np.empty((k, n), order="F")
    r = cb.rsvd(a, k, q, p, omega=omega, out=(u, s, vt))
    assert r[0] is u and r[1] is s and r[2] is vt
    assert np.array_equal(u, u0) and np.array_equal(s, s0) and np.array_equal(vt, vt0)
    up = torch.empty((k, m), dtype=torch.float64, pin_memory=True).numpy().T
    sp = torch.empty((1, k), dtype=torch.float64, pin_memory=True).numpy().T
    vp = torch.empty((n, k), dtype=torch.float64, pin_memory=True).numpy().T
    cb.rsvd(a, k, q, p, omega=omega, out=(up, sp, vp))
    assert np.array_equal(up, u0) and np.array_equal(sp, s0) and np.array_equal(vp, vt0)
    # device
    ad, od = torch.from_numpy(a).cuda(), torch.from_numpy(omega).cuda()
    ud = torch.empty((k, m), dtype=torch.float64, device="cuda").t()
    sd = torch.empty((1, k), dtype=torch.float64, device="cuda").t()
    vd = torch.empty((n, k), dtype=torch.float64, device="cuda").t()
    cb.rsvd(ad, k, q, p, omega=od, out=(ud, sd, vd))
    torch.cuda.synchronize()
    assert np.array_equal(ud.cpu().numpy(), u0) and np.array_equal(vd.cpu().numpy(), vt0)
    # wrong layout / shape / residency are refused before anything runs
    with pytest.raises(ValueError):
        cb.rsvd(a, k, q, p, out=(np.empty((m, k)), s, vt))                 # row-major U
    with pytest.raises(ValueError):
        cb.rsvd(a, k, q, p, out=(u, s, np.empty((k, n + 1), order="F")))
    with pytest.raises(ValueError):
        cb.rsvd(a, k, q, p, out=(ud, sd, vd))                              # device outputs for a host input


def test_timings_report_convergence(cb):
    rng = np.random.default_rng(6)
    a = rng.standard_normal((3000, 200))
    cb.rsvd(a, 50, 2, 10, seed=1)
    t = cb.last_timings()
    assert t["jacobi_converged"] == 1 and 1 <= t["jacobi_sweeps"] < 30
    assert t["fused_small"] in (0, 1)


def test_f32_instantiation(cb):
    """random_svd::<f32>: float32 in, float32 out (host and device, row- and column-major, fat), against the f64 oracle
    on the widened data at single-precision tolerances."""
    import torch
    from oracle import ref_rsvd
    rng = np.random.default_rng(8)
    for shape, order in (((3000, 200), "C"), ((2500, 130), "F"), ((90, 1500), "C")):
        a32 = rng.standard_normal(shape).astype(np.float32)
        if order == "F":
            a32 = np.asfortranarray(a32)
        k, q, p = 12, 3, 8
        omega = rng.standard_normal((min(shape), k + p))
        ref = ref_rsvd.random_svd(np.ascontiguousarray(a32).astype(np.float64), k, q, p, omega=omega)
        u, s, vt = cb.rsvd_f32(a32, k, q, p, omega=omega)
        assert u.dtype == np.float32 and s.dtype == np.float32 and vt.dtype == np.float32
        assert u.shape == (shape[0], k) and s.shape == (k, 1) and vt.shape == (k, shape[1])
        assert ref_rsvd.sigma_rel_err(ref[1], s.astype(np.float64)) < 1e-6
        assert ref_rsvd.subspace_sine(ref[0], np.linalg.qr(u.astype(np.float64))[0]) < 1e-5
        ud, sd, vd = cb.rsvd_f32(torch.from_numpy(np.ascontiguousarray(a32)).cuda(), k, q, p, omega=torch.from_numpy(omega).cuda())
        assert ud.dtype == torch.float32
        assert np.max(np.abs(sd.cpu().numpy() - s)) < 1e-5 * s[0, 0]
        assert ref_rsvd.subspace_sine(ref[0], np.linalg.qr(ud.cpu().numpy().astype(np.float64))[0]) < 1e-5
    with pytest.raises(TypeError):
        cb.rsvd_f32(np.zeros((4, 4)), 1, 1, 1)
