"""Drop-in module name of the reference's pyo3 extension (`import corrla_rs`, Cargo.toml:7,
src/lib_math_utils_py.rs:17-18).  The RSVD hot path is provided: `corrla_rs.rsvd(a_mat, n_rank,
n_iters, n_oversamples)` with the reference's positional signature (src/lib_math_utils_py.rs:21-23,
examples/benchmark_rsvd.py:101), executed by the B200 engine, and its first consumer `corrla_rs.rpca` (:38-55, PCA with the centring fused into
the passes).  The other functions of the reference
module (active_ss, cs_*, PyRbfInterp, PyPodI, PyDMDc) are out of scope and raise on access."""
from corrla_rs_b200 import rpca, rsvd  # noqa: F401

_OUT_OF_SCOPE = ("active_ss", "cs_dirichlet_sample", "cs_mcmc_dirichlet_sample", "PyRbfInterp", "PyPodI",
                 "PyDMDc")


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"corrla_rs.{name} is outside the scope of the B200 engine (only the RSVD hot path is replaced); "
            "use the reference crate for it")
    raise AttributeError(name)
