"""Short program for ncu: two rsvd calls (one warm-up) on a device-resident Gaussian matrix.
usage: python tools/profile_run.py [rows] [cols] [k] [q] [p]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import corrla_rs_b200 as cb

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
q = int(sys.argv[4]) if len(sys.argv) > 4 else 4
p = int(sys.argv[5]) if len(sys.argv) > 5 else 10
torch.manual_seed(0)
a = torch.randn(rows, cols, dtype=torch.float64, device="cuda")
for i in range(2):
    u, s, vt = cb.rsvd(a, k, q, p, seed=3)
    torch.cuda.synchronize()
    t = cb.last_timings()
    print(i, "device_ms", t["device_ms"], "pass_ms", t["pass_ms"], "launches", t["gpu_launches"], "sweeps", t["jacobi_sweeps"], "sigma0", float(s[0, 0]))
