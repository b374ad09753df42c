"""Time the active-subspace gradient build at BASELINE config C5 scale (1 048 576 x 64 samples, 72 neighbours): the
GEMM-form nearest-neighbour search against the exact kernel (CORRLA_B200_KNN_EXACT=1) on a subsample."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import corrla_rs_b200 as cb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d, k = 64, 72
g = torch.Generator(device="cuda"); g.manual_seed(9)
x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
h = torch.randn((d, 8), dtype=torch.float64, device="cuda", generator=g)
y = 0.5 * ((x @ h) ** 2).sum(dim=1) + 1e-2 * torch.randn((n,), dtype=torch.float64, device="cuda", generator=g)
out = {}
for name, env in (("gemm_form", None), ("exact", "1")):
    if env and n > (1 << 18) and "--exact" not in sys.argv:
        continue
    if env:
        os.environ["CORRLA_B200_KNN_EXACT"] = env
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = cb.active_ss_fit(x, y, 1, k, 8)
    torch.cuda.synchronize()
    out[name] = {"seconds": time.perf_counter() - t0, "eig_head": np.diag(np.asarray(fit.singular_vals_))[:3].tolist()}
    os.environ.pop("CORRLA_B200_KNN_EXACT", None)
out["n"] = n; out["d"] = d; out["n_nbr"] = k
out["pair_feature_products"] = float(n) * n * d
if "gemm_form" in out:
    out["gemm_form"]["tflops_equiv"] = 2.0 * n * n * d / out["gemm_form"]["seconds"] * 1e-12
print(json.dumps(out))
