#!/usr/bin/env python
"""One active_ss_fit call at N = 65 536, d = 64, 72 neighbours, for ncu (knn_kernel, poly_grad_kernel)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import corrla_rs_b200 as cb  # noqa: E402

g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.randn((65536, 64), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn(65536, dtype=torch.float64, device="cuda", generator=g)
fit = cb.active_ss_fit(x, y, 1, 72, 8)
torch.cuda.synchronize()
print("ok", fit.n_deficient)
