"""
CPU emulations (float32 NumPy) of arithmetic rewrites the CUDA kernels use, against the plain float64 formulas of the oracle.
They pin the ALGEBRA of a rewrite independently of a GPU; the kernels themselves are held to the oracle by tests/test_gpu_parity.py.

* the LSTM cell with merged quotients (gordo_components_b200/csrc/lstm_infer_tc.cu `tanh_cell`): five ex2 and two rcp instead of five
  and five, exponentials capped at 2^30;
* the TF32 head/tail split of the measured mma.sync training variant: what three TF32 products recover of an
  fp32 product.
"""
import numpy as np

f32 = np.float32
CAP = f32(2.0 ** 30)
LOG2E = f32(1.4426950408889634)


def _ex2_capped(t):
    """ex2.approx followed by min.NaN with 2^30 (NaN stays NaN)."""
    with np.errstate(over="ignore"):
        e = np.exp2(t.astype(f32)).astype(f32)
    return np.where(np.isnan(e), e, np.minimum(e, CAP)).astype(f32)


def merged_quotient_cell(zi, zf, zg, zo, c_prev):
    """float32 restatement of `tanh_cell`: c' = [c(1+ei)(1+eg) + (eg-1)(1+ef)] / [(1+ef)(1+ei)(1+eg)],  h = (ec-1) / [(1+eo)(1+ec)]."""
    zi, zf, zg, zo, c_prev = (np.asarray(v, dtype=f32) for v in (zi, zf, zg, zo, c_prev))
    pi = f32(1) + _ex2_capped(-LOG2E * zi)
    pf = f32(1) + _ex2_capped(-LOG2E * zf)
    eg = _ex2_capped(f32(2) * LOG2E * zg)
    po = f32(1) + _ex2_capped(-LOG2E * zo)
    pig = (pi * (eg + f32(1))).astype(f32)
    r1 = (f32(1) / (pig * pf)).astype(f32)
    c_new = ((c_prev * pig + (eg - f32(1)) * pf) * r1).astype(f32)
    ec = _ex2_capped(f32(2) * LOG2E * c_new)
    r2 = (f32(1) / (po * (ec + f32(1)))).astype(f32)
    return c_new, ((ec - f32(1)) * r2).astype(f32)


def reference_cell(zi, zf, zg, zo, c_prev):
    """The Keras LSTM cell in float64 (oracle/keras_math.py `lstm_predict`: sigmoid gates, tanh candidate and output)."""
    zi, zf, zg, zo, c_prev = (np.asarray(v, dtype=np.float64) for v in (zi, zf, zg, zo, c_prev))
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    c_new = sig(zf) * c_prev + sig(zi) * np.tanh(zg)
    return c_new, sig(zo) * np.tanh(c_new)


def test_merged_quotient_cell_equals_the_lstm_cell():
    rng = np.random.default_rng(0)
    n = 200_000
    z = rng.normal(0, 4, (4, n))
    c_prev = rng.normal(0, 2, n)
    c, h = merged_quotient_cell(*z, c_prev)
    c_ref, h_ref = reference_cell(*z, c_prev)
    assert np.abs(c - c_ref).max() < 2e-6 * (1 + np.abs(c_ref).max())
    assert np.abs(h - h_ref).max() < 2e-6
    # the whole float range: the products stay finite because the exponentials are capped, and the limits are the cell's limits
    edge = np.array([-1e30, -200.0, -50.0, -21.0, -20.0, 0.0, 20.0, 21.0, 50.0, 200.0, 1e30])
    grids = np.meshgrid(edge, edge, edge, edge, np.array([-100.0, -1.0, 0.0, 1.0, 100.0]), indexing="ij")
    flat = [gr.ravel() for gr in grids]
    c, h = merged_quotient_cell(*flat)
    c_ref, h_ref = reference_cell(*[np.clip(v, -700, 700) for v in flat])
    assert np.isfinite(c).all() and np.isfinite(h).all()
    np.testing.assert_allclose(c, c_ref, rtol=3e-6, atol=3e-6)
    np.testing.assert_allclose(h, h_ref, rtol=0, atol=3e-6)
    # a NaN pre-activation or state stays a NaN, as in the reference
    for k in range(5):
        args = [np.array([0.3], dtype=f32) for _ in range(5)]
        args[k] = np.array([np.nan], dtype=f32)
        c, h = merged_quotient_cell(*args)
        assert np.isnan(h).all() and (np.isnan(c).all() or k == 3)  # the output gate does not enter c'


def _tf32_round(x):
    """cvt.rna.tf32.f32: round to 10 explicit mantissa bits, ties away from zero (add half an ulp of TF32 to the magnitude, truncate)."""
    bits = np.asarray(x, dtype=f32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)


def test_tf32_head_tail_split_recovers_fp32_products():
    rng = np.random.default_rng(1)
    a = (rng.normal(0, 1, 100_000) * 10.0 ** rng.uniform(-6, 3, 100_000)).astype(f32)
    w = rng.normal(0, 0.2, 100_000).astype(f32)
    a_hi, w_hi = _tf32_round(a), _tf32_round(w)
    a_lo, w_lo = _tf32_round(a - a_hi), _tf32_round(w - w_hi)
    # head + tail represents the operand to 2^-22 of its magnitude (the tail of a round-to-nearest head has 12 significant bits, 11 are kept)
    assert (np.abs((a_hi.astype(np.float64) + a_lo) - a) <= 2.0 ** -22 * np.abs(a)).all()
    exact = a.astype(np.float64) * w.astype(np.float64)
    three = a_lo.astype(np.float64) * w_hi + a_hi.astype(np.float64) * w_lo + a_hi.astype(np.float64) * w_hi
    one = a_hi.astype(np.float64) * w_hi
    rel3 = np.abs(three - exact) / np.abs(exact)
    rel1 = np.abs(one - exact) / np.abs(exact)
    assert rel3.max() < 2.0 ** -20 and np.median(rel3) < 2.0 ** -23  # fp32 class (an fp32 product rounds at 2^-24)
    assert np.median(rel1) > 2.0 ** -13                              # a single TF32 product is three decimal digits
