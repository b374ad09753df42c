"""Drop-in module name of the reference's pyo3 extension (`import corrla_rs`, Cargo.toml:7,
src/lib_math_utils_py.rs:17-18), executed by the B200 engine:

  * `corrla_rs.rsvd(a_mat, n_rank, n_iters, n_oversamples)`  -- the hot path (lib_math_utils_py.rs:21-36,
    examples/benchmark_rsvd.py:101);
  * `corrla_rs.rpca(a_mat, n_rank, n_iters, n_oversamples)`  -- PCA with the centring fused into the passes (:38-55);
  * `corrla_rs.active_ss(a_mat, y, order, n_nbr, n_comps)`   -- active subspace from samples (:56-86);
  * `corrla_rs.PyDMDc(x, u, n_modes, n_iters).predict(x0, u)` -- DMD with control (:255-283);
  * `corrla_rs.PyPodI(x, t, n_modes).predict(t)` and `corrla_rs.PyRbfInterp` -- POD with interpolated weights (:172-250).

The constrained samplers of the reference module (cs_dirichlet_sample, cs_mcmc_dirichlet_sample) do not touch the RSVD
hot path, are out of scope and raise on access."""
from corrla_rs_b200 import active_ss, rpca, rsvd  # noqa: F401
from corrla_rs_b200.rom import PyDMDc, PyPodI, PyRbfInterp  # noqa: F401

_OUT_OF_SCOPE = ("cs_dirichlet_sample", "cs_mcmc_dirichlet_sample")


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"corrla_rs.{name} is outside the scope of the B200 engine (only the RSVD hot path and the routines built "
            "directly on it are replaced); use the reference crate for it")
    raise AttributeError(name)
