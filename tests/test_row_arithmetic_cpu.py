"""The per-element arithmetic of K1 (csrc/row_math.cuh: RowDiv) restated exactly on the CPU.

K1 divides a whole row by one divisor with a per-row reciprocal and one FMA correction instead of an IEEE division per
element:   r = RN(1/d);  q0 = RN(x*r);  rem = fma(-q0, d, x);  q = fma(rem, r, q0)   (Markstein).
The claims that rest on it -- "K1 mirrors numpy's x / (norm + 1e-8) op for op" (utils/cv_evaluator.py:95-105) and
"deferred fp32 rows re-create the value K1 would have stored bit for bit" -- need q to be the CORRECTLY ROUNDED quotient,
i.e. exactly what numpy's float32 division returns.  Here the sequence is evaluated in exact rational arithmetic with
explicit round-to-nearest-even to float32 after every operation and compared with numpy, and the whole deferred chain
x -> RN(x / n_seg) -> * w -> RN(. / n_row) with numpy's own float32 chain."""
from fractions import Fraction

import numpy as np


def _rn32(v: Fraction) -> np.float32:
    """Round an exact rational to the nearest float32 (ties to even), without going through float64."""
    if v == 0:
        return np.float32(0.0)
    sign = -1 if v < 0 else 1
    a = abs(v)
    e = a.numerator.bit_length() - a.denominator.bit_length() - 24
    while a / Fraction(2) ** e >= 2 ** 24:
        e += 1
    while a / Fraction(2) ** e < 2 ** 23:
        e -= 1
    e = max(e, -149)
    m = a / Fraction(2) ** e
    fl = m.numerator // m.denominator
    rem = m - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return np.float32(sign * float(Fraction(fl) * Fraction(2) ** e))


def _f(x) -> Fraction:
    return Fraction(float(x))


def _fma(a, b, c) -> np.float32:
    return _rn32(_f(a) * _f(b) + _f(c))


def _row_div(x: np.float32, d: np.float32) -> np.float32:
    r = _rn32(1 / _f(d))                         # __frcp_rn
    q0 = _rn32(_f(x) * _f(r))
    rem = _fma(-q0, d, x)
    return _fma(rem, r, q0)


def test_markstein_sequence_is_the_correctly_rounded_quotient():
    rng = np.random.default_rng(0)
    for _ in range(4000):
        x = np.float32(rng.standard_normal() * 10.0 ** rng.integers(-6, 4))
        d = np.float32(abs(rng.standard_normal()) * 10.0 ** rng.integers(-3, 3) + 1e-8)
        assert _row_div(x, d) == np.float32(x) / np.float32(d), (x, d)
    # unit-norm rows (the common case): divisors near 1 and near sqrt(2), elements of a few 1e-2
    for _ in range(2000):
        x = np.float32(rng.standard_normal() * 0.03)
        d = np.float32(rng.choice([1.0, np.sqrt(2.0), 22.6]) * (1 + rng.standard_normal() * 1e-3) + 1e-8)
        assert _row_div(x, d) == x / d


def test_deferred_row_chain_equals_numpy_float32_chain():
    """x -> RN(x / n_seg) -> * w -> RN(. / n_row): the value a deferred row re-creates == what numpy computes for the
    reference's `unit_rows` + weighted concat + `unit_rows` given the same divisors."""
    rng = np.random.default_rng(1)
    for _ in range(1500):
        x = np.float32(rng.standard_normal() * 3.0)
        n_seg = np.float32(abs(rng.standard_normal()) * 60 + 1e-8)
        w = np.float32(rng.choice([1.0, 0.4, 0.6, 0.25]))
        n_row = np.float32(abs(rng.standard_normal()) + 0.5)
        ours = _row_div(_rn32(_f(_row_div(x, n_seg)) * _f(w)), n_row)
        ref = ((x / n_seg) * w) / n_row
        assert ours == ref and ref.dtype == np.float32


def test_standardize_sequence_equals_sklearn_float32_ops():
    """K1 with EMR2A_NF_STANDARDIZE / emr2a_standardize: t = RN(x - mean); t / scale by the same sequence with the
    per-column reciprocal -- sklearn's in-place fp32 `X -= mean; X /= scale` (utils/cv_evaluator.py:78-80)."""
    rng = np.random.default_rng(2)
    for _ in range(1500):
        x = np.float32(rng.standard_normal() * 2 + 0.3)
        m = np.float32(rng.standard_normal() * 0.5)
        s = np.float32(abs(rng.standard_normal()) + 0.05)
        t = _rn32(_f(x) - _f(m))
        assert _row_div(t, s) == (x - m) / s
