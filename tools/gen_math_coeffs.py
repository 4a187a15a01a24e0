#!/usr/bin/env python
"""Generates the polynomial coefficients of gokalman_b200/csrc/fastmath.cuh (Box-Muller pieces on
restricted domains) with mpmath and prints their worst-case errors.  Run: python tools/gen_math_coeffs.py"""
import mpmath as mp

mp.mp.dps = 60


def fit(f, lo, hi, deg):
    c, err = mp.chebyfit(f, [lo, hi], deg + 1, error=True)  # highest power first
    return [mp.mpf(x) for x in c][::-1], err  # lowest power first


def show(name, coeffs):
    print("// %s" % name)
    for i, c in enumerate(coeffs):
        print("  %s,  // t^%d" % (float.hex(float(c)), i))


# sin(f*pi/4) = f * S(t), cos(f*pi/4) = C(t), t = f^2, f in [0, 1]
q = mp.pi / 4
S, es = fit(lambda t: mp.sin(mp.sqrt(t) * q) / mp.sqrt(t) if t > 0 else q, mp.mpf(0), mp.mpf(1), 6)
Cc, ec = fit(lambda t: mp.cos(mp.sqrt(t) * q), mp.mpf(0), mp.mpf(1), 6)
print("sin fit err", mp.nstr(es, 5), "cos fit err", mp.nstr(ec, 5))
show("S", S)
show("C", Cc)

# ln(m) = 2 s (1 + w L(w)), s = (m-1)/(m+1), w = s^2, m in [sqrt(1/2), sqrt(2)]
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
wmax = smax ** 2
Lc, el = fit(lambda w: (mp.atanh(mp.sqrt(w)) / mp.sqrt(w) - 1) / w if w > 0 else mp.mpf(1) / 3, mp.mpf(0), wmax * mp.mpf("1.0001"), 6)
print("log fit err", mp.nstr(el, 5), "wmax", mp.nstr(wmax, 8))
show("L", Lc)
ln2 = mp.log(2)
hi = mp.mpf(int(ln2 * 2 ** 44)) / 2 ** 44
print("ln2_hi", float.hex(float(hi)), "ln2_lo", float.hex(float(ln2 - hi)))
