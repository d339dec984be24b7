"""CPU suite: the mathematics behind the certified pre-filter (cmdiad_b200/csrc/score_tail.cu, refine_cert_kernel).

The GEMM works on fp16 roundings q_hi, b_hi of the power-of-two-scaled float32 operands and produces
v(q,r) = ||b_r||^2 - 2 q_hi.b_hi; the exact re-check computes f(q,r) = sum_i (q_i - b_ri)^2 in float32.  The certificate
needs |v + ||q||^2 - f| <= E(q) for EVERY bank row, with E built from norms only (Cauchy-Schwarz on the rounding
residuals + an accumulation model + float32 rounding terms).  Here the bound is evaluated exhaustively on small data,
with the products accumulated exactly (float64) and in float32, and the selection rule built on it is replayed: every row
above `min v + 2E` must be strictly farther than the row the exact scan returns."""
import numpy as np
import pytest

from cmdiad_b200 import synth


def _fp16_round_scaled(x, axis_scale):
    """x * 2^e -> fp16 -> back, e chosen like the library: max|x| * 2^e in [2^12, 2^13) per row (queries) or per bank"""
    amax = np.abs(x).max(axis=1, keepdims=True) if axis_scale == "row" else np.abs(x).max()
    e = 13 - np.frexp(amax)[1]
    scale = np.ldexp(np.float32(1.0), e).astype(np.float32)
    return ((x * scale).astype(np.float16).astype(np.float64)) / scale.astype(np.float64)


def _bound(q, bank, q_hi, b_hi):
    D = q.shape[1]
    qn = np.linalg.norm(q.astype(np.float64), axis=1)
    qe = np.linalg.norm(q.astype(np.float64) - q_hi, axis=1)
    bmax = np.linalg.norm(bank.astype(np.float64), axis=1).max()
    eb = np.linalg.norm(bank.astype(np.float64) - b_hi, axis=1).max()
    bh = bmax + eb
    acc_model = (D // 16 + 1) * 17 * 2.0 ** -23
    return 2 * (qe * bh + qn * eb + acc_model * (qn + qe) * bh) + (D + 16) * 2.0 ** -24 * (qn + bmax) ** 2


@pytest.mark.parametrize("dist,D,scale", [("C", 256, 1.0), ("G", 256, 1.0), ("C", 768, 250.0), ("G", 64, 1e-3)])
def test_error_bound_and_selection_rule(dist, D, scale):
    R, P = 1500, 160
    cent = synth.centroids(D, 32) if dist == "C" else None
    bank = synth.patches(R, D, seed=5, dist=dist, cent=cent) * np.float32(scale)
    q = synth.patches(P, D, seed=6, dist=dist, anomalous_frac=0.05, cent=cent) * np.float32(scale)
    b_hi, q_hi = _fp16_round_scaled(bank, "bank"), _fp16_round_scaled(q, "row")
    bn = (bank.astype(np.float64) ** 2).sum(1).astype(np.float32).astype(np.float64)   # ||b||^2 as the library stores it
    E = _bound(q, bank, q_hi, b_hi)
    # exact re-check value f: float32 direct form
    f = ((q[:, None, :] - bank[None, :, :]) ** 2).sum(2, dtype=np.float32).astype(np.float64)
    qq = (q.astype(np.float64) ** 2).sum(1)
    for acc in ("exact", "float32"):
        if acc == "exact":
            dot = q_hi @ b_hi.T
        else:
            dot = (q_hi.astype(np.float32) @ b_hi.astype(np.float32).T).astype(np.float64)
        v = bn[None, :] - 2.0 * dot
        err = np.abs(v + qq[:, None] - f)
        assert (err <= E[:, None]).all(), (acc, float((err / E[:, None]).max()))
        assert (err / E[:, None]).max() < 0.5   # Cauchy-Schwarz is far from tight on real roundings
        # selection rule: rows outside the band can never be the exact nearest neighbour (ties included)
        thr = v.min(1) + 2 * E
        exact_best = f.min(1)
        outside = v > thr[:, None]
        assert (np.where(outside, f, np.inf).min(1) > exact_best).all()
        # ... and the same for the three nearest rows (re-weighting / neighbour table): band = third smallest + 2E
        thr3 = np.sort(v, 1)[:, 2] + 2 * E
        third = np.sort(f, 1)[:, 2]
        assert (np.where(v > thr3[:, None], f, np.inf).min(1) > third).all()
    # the band is narrow: on this data only a handful of rows per query need the exact re-check
    assert (v <= thr[:, None]).sum(1).mean() < 12
