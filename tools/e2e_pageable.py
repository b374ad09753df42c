#!/usr/bin/env python
"""End-to-end RSVD from ordinary (pageable) numpy memory: 1 048 576 x 1024 f64 (8 GiB), k = 100, q = 4, p = 10."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import corrla_rs_b200 as cb  # noqa: E402

m, n = 1_048_576, 1024
rng = np.random.default_rng(3)
a = np.empty((m, n))
for r0 in range(0, m, 65536):
    a[r0:r0 + 65536] = rng.standard_normal((65536, n))
for i in range(3):
    t0 = time.perf_counter()
    u, s, vt = cb.rsvd(a, 100, 4, 10, seed=5)
    dt = time.perf_counter() - t0
    t = cb.last_timings()
    print(f"call {i}: {dt * 1e3:.1f} ms  h2d {t['h2d_ms']:.1f} device {t['device_ms']:.1f} d2h {t['d2h_ms']:.1f} chunks {t['streamed_chunks']}"
          f"  H2D {a.nbytes / t['h2d_ms'] / 1e6:.1f} GB/s  sigma0 {s[0, 0]:.6f}", flush=True)
print("orth", float(np.max(np.abs(u[:, :8].T @ u[:, :8] - np.eye(8)))))
