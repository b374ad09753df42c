"""CPU check of the error bound the nearest-neighbour certificate uses for the TF32 shortlist (csrc/gradients.cu,
knn_tf32_delta_rel): distances |q|^2 + |c|^2 - 2 q.c from TF32-rounded centred samples with FP32 accumulation, emulated in
numpy, must stay within delta = (1.5 * 2^-10 + (d8 + 24) * 2^-22) * (|q|^2 + max |c|^2) of the exact sum of (a - b)^2 --
also when the rounding errors cannot average out (one feature, all products rounded the same way)."""
import numpy as np


def tf32_round(x):
    """cvt.rna.tf32.f32: 10 mantissa bits, round to nearest, ties away from zero."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def delta_rel(d):
    d8 = (d + 7) // 8 * 8
    return 1.5 * 2.0 ** -10 + (d8 + 24) * 2.0 ** -22


def worst_ratio(x):
    xc = x - x.mean(axis=0)
    n2 = (xc * xc).sum(axis=1)
    xf = tf32_round(xc.astype(np.float32))
    dots = np.zeros((x.shape[0], x.shape[0]), dtype=np.float32)
    for c in range(x.shape[1]):                                    # FP32 accumulation of exact TF32 x TF32 products
        dots = (dots + (xf[:, c:c + 1] * xf[:, c][None, :]).astype(np.float32)).astype(np.float32)
    nf = n2.astype(np.float32)
    approx = (nf[:, None] + nf[None, :]).astype(np.float32) + np.float32(-2.0) * dots
    exact = ((xc[:, None, :] - xc[None, :, :]) ** 2).sum(axis=2)
    bound = delta_rel(x.shape[1]) * (n2[:, None] + n2.max())
    return float((np.abs(approx.astype(np.float64) - exact) / bound).max())


def test_tf32_shortlist_bound_holds_on_emulation():
    rng = np.random.default_rng(0)
    m = 600
    just_below_tie = 1.0 + 2.0 ** -11 * 0.999                     # every element rounds DOWN by almost half a TF32 ulp
    cases = {
        "gauss_d64": rng.standard_normal((m, 64)),
        "gauss_d100": rng.standard_normal((m, 100)),
        "scales_1e-3_1e3": rng.standard_normal((m, 16)) * np.logspace(-3, 3, 16),
        "offset_1e3": 1e3 + rng.standard_normal((m, 7)),
        "one_feature": rng.standard_normal((m, 1)),
        "aligned_rounding": np.sign(rng.standard_normal((m, 64))) * just_below_tie,
        "aligned_rounding_d1": np.concatenate([np.full((m // 2, 1), just_below_tie), np.full((m // 2, 1), -just_below_tie)]),
    }
    for name, x in cases.items():
        r = worst_ratio(x)
        assert r < 1.0, (name, r)
