#!/usr/bin/env python
"""Chebyshev coefficients for exp(-x) I0(x) and exp(-x) I1(x), x >= 0, as used by ig_uq.cuh (Rician objective).

    [0, 8]:     ascending power series in u = x^2 / 4 (coefficients 1 / (k!)^2 and 1 / (k! (k+1)!), 16 terms, written out in ig_uq.cuh)
    (8, inf):   sqrt(x) i0e = sum_k b_k T_k(16 / x - 1)           x (1 - I1(x) / I0(x)) = sum_k d_k T_k(16 / x - 1)

(the ratio I1 / I0 tends to 1 like 1 - 1/(2x): its distance from 1 is fitted directly so that no cancellation is left)

Fitted here against scipy.special (float64) and truncated where the tail drops below 2e-9 of the leading term; the
script prints C arrays and the measured maximum relative error of the float32 evaluation.  python tools/gen_bessel_coeffs.py
"""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy import special as sp


def fit(fn, deg):
    # Chebyshev interpolation at deg + 1 Chebyshev nodes of t in [-1, 1]
    k = np.arange(deg + 1)
    t = np.cos(np.pi * (k + 0.5) / (deg + 1))
    return C.chebfit(t, fn(t), deg)


def trim(c, tol=2e-9):
    n = len(c)
    while n > 1 and abs(c[n - 1]) < tol * abs(c[0]):
        n -= 1
    return c[:n]


def horner32(c, t):
    c = c.astype(np.float32)
    t = t.astype(np.float32)
    acc = np.full_like(t, c[-1])
    for a in c[-2::-1]:
        acc = acc * t + a
    return acc


def clenshaw32(c, t):
    c = c.astype(np.float32)
    t = t.astype(np.float32)
    b1 = np.zeros_like(t)
    b2 = np.zeros_like(t)
    two_t = np.float32(2) * t
    for a in c[:0:-1]:
        b1, b2 = two_t * b1 - b2 + a, b1
    return t * b1 - b2 + c[0]


def _om(x):
    """x (1 - I1(x)/I0(x)) in extended precision where float64 cancels (large x: asymptotic series)."""
    import mpmath as mp
    out = np.empty_like(x)
    for i, v in enumerate(x):
        if np.isinf(v) or v > 1e6:
            out[i] = 0.5
        else:
            mp.mp.dps = 40
            out[i] = float(mp.mpf(v) * (1 - mp.besseli(1, v) / mp.besseli(0, v)))
    return out


def main():
    small = lambda t: 4.0 * (t + 1.0)            # x in [0, 8]
    large = lambda t: 16.0 / (t + 1.0 + 1e-300)  # x in [8, inf)
    sets = {
        "kI0B": trim(fit(lambda t: np.sqrt(large(t)) * sp.i0e(large(t)), 30)),
        "kOMB": trim(fit(lambda t: large(t) * (1.0 - sp.i1e(large(t)) / sp.i0e(large(t))) if False else _om(large(t)), 30)),
    }
    # the two short large-argument series are re-expressed in the monomial basis (Horner, one FMA per term; harmless at
    # degree <= 9 on [-1, 1]); the degree-18 small-argument series stay in the Chebyshev basis (Clenshaw)
    for key in ("kI0B", "kOMB"):
        sets[key + "m"] = C.cheb2poly(sets.pop(key))
    for name, c in sets.items():
        body = ", ".join(f"{v:.9e}f" for v in c)
        print(f"__device__ constexpr float {name}[{len(c)}] = {{{body}}};")
    x = np.concatenate([np.linspace(0, 8, 20001), 8 + np.logspace(-6, 5, 20001)])
    xs, xl = x[x <= 8], x[x > 8]
    import math
    u = (np.float32(0.25) * xs.astype(np.float32) ** 2)
    s0 = horner32(np.array([1.0 / math.factorial(k) ** 2 for k in range(16)]), u)
    s1 = horner32(np.array([1.0 / (math.factorial(k) * math.factorial(k + 1)) for k in range(16)]), u) * np.float32(0.5) * xs.astype(np.float32)
    i0 = np.concatenate([s0 * np.exp(-xs), horner32(sets["kI0Bm"], 16 / xl - 1) / np.sqrt(xl).astype(np.float32)])
    e0 = np.max(np.abs(i0 - sp.i0e(x)) / sp.i0e(x))
    e1 = np.max(np.abs(s1[1:] / s0[1:] - sp.i1e(xs[1:]) / sp.i0e(xs[1:])) / (sp.i1e(xs[1:]) / sp.i0e(xs[1:])))
    sub = xl[::200]
    om = horner32(sets["kOMBm"], 16 / sub - 1) / sub.astype(np.float32)
    e2 = np.max(np.abs(om - _om(sub) / sub) / (_om(sub) / sub))
    print(f"// float32 evaluation (power series on [0, 8], Horner of the fits above): max rel err i0e {e0:.2e} on [0, 1e5], I1/I0 {e1:.2e} on (0, 8], 1 - I1/I0 {e2:.2e} on (8, 1e5]")


if __name__ == "__main__":
    main()
