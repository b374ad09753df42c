"""Generate tests/golden/*.npz.  Run in the build container (it reads /root/reference, which does not exist
on the GPU box); the fixtures it writes are committed.

1. ref_examples_rsvd_*.npz -- outputs of the REFERENCE'S OWN numpy statement of the algorithm,
   examples/benchmark_rsvd.py:16-54 (`power_iteration`, `rsvd`), executed from the reference checkout with the
   top-level `import corrla_rs` line skipped (that module is the Rust extension, not buildable here).  The
   Omega it drew from numpy's global RNG is recorded by replaying the seed.
2. known_answer_5x5.npz -- the matrix and singular values of test_rsvd_lowrank (random_svd.rs:153-196).
3. oracle_cases.npz -- seeded inputs + oracle outputs used by the GPU parity tests as fixed vectors.
"""
import ast
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/examples/benchmark_rsvd.py")
OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def load_reference_functions():
    tree = ast.parse(REF.read_text())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("power_iteration", "rsvd")]
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, str(REF), "exec"), ns)
    return ns["rsvd"]


def main():
    ref_rsvd_py = load_reference_functions()
    for name, shape, k, p, q, seed in (("tall", (400, 60), 6, 10, 8, 11), ("fat", (50, 700), 5, 4, 3, 12)):
        rng = np.random.default_rng(seed)
        a = rng.standard_normal(shape)
        thin_cols = min(shape)
        np.random.seed(seed)
        omega = np.random.randn(thin_cols, k + p)       # what rsvd() will draw (benchmark_rsvd.py:46)
        np.random.seed(seed)
        u, s, vt = ref_rsvd_py(a, omega_rank=k, n_oversamples=p, power_iter=q)
        np.savez(OUT / f"ref_examples_rsvd_{name}.npz", a=a, omega=omega, u=u, s=s, vt=vt, k=k, p=p, q=q)
        print(name, a.shape, "->", u.shape, s.shape, vt.shape)

    a5 = np.array([[1, 0, 0, 0, 2], [0, 0, 3, 0, 0], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0, 2, 0, 0, 0]], dtype=np.float64)
    np.savez(OUT / "known_answer_5x5.npz", a=a5, sigma=np.array([3.0, 2.2360679, 2.0, 0.0, 0.0]), tol=1e-3)

    from oracle import ref_rsvd
    cases = {}
    rng = np.random.default_rng(2024)
    for name, (m, n, k, q, p) in {"c1": (100, 100, 10, 12, 8), "tall": (1500, 96, 12, 4, 10),
                                  "fat": (40, 900, 8, 10, 10), "clamp": (64, 9, 6, 5, 10)}.items():
        a = rng.standard_normal((m, n))
        l = min(k + p, min(m, n))
        omega = rng.standard_normal((min(m, n), l))
        u, s, vt = ref_rsvd.random_svd(a, k, q, p, omega=omega)
        for key, val in (("a", a), ("omega", omega), ("u", u), ("s", s), ("vt", vt), ("kqp", np.array([k, q, p]))):
            cases[f"{name}_{key}"] = val
    np.savez_compressed(OUT / "oracle_cases.npz", **cases)
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
