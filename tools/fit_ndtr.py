"""Generates and validates the coefficients of fpt_fast.cu's ndtr_tail():
    Phi(-t) = exp(-t^2/2) * F(u),   u = (t - 5)/(t + 5),   F(u) = 0.5 * erfcx(t / sqrt 2),  0 <= t < 26.5
F is fitted by Chebyshev interpolation (degree 16) against mpmath (40 digits) and converted to the
monomial basis in u in high precision; the script then replays the float64 evaluation order of the
kernel (Horner, Cody-Waite exp with a degree-11 Taylor polynomial) and prints the worst relative
error against mpmath. Run: python tools/fit_ndtr.py   (needs mpmath; build container only)."""
import math
from math import comb

import mpmath as mp
import numpy as np
from numpy.polynomial import chebyshev as Ch

mp.mp.dps = 40
C, TM, N = 5.0, 26.5, 16


def F(t):
    x = mp.mpf(t) / mp.sqrt(2)
    return 0.5 * mp.exp(x * x) * mp.erfc(x)


def horner(c, x):
    a = np.full_like(x, c[-1])
    for ck in c[-2::-1]:
        a = a * x + ck
    return a


def main():
    umax = (TM - C) / (TM + C)
    k = np.arange(N + 1)
    v = np.cos(np.pi * (k + 0.5) / (N + 1))
    u = (v + 1) / 2 * (umax + 1) - 1
    f = np.array([float(F(C * (1 + ui) / (1 - ui))) for ui in u])
    mono_v = Ch.cheb2poly(Ch.chebfit(v, f, N))
    A = 2 / (umax + 1)
    B = A - 1  # v = A*u + B
    coef_u = [mp.mpf(0)] * (N + 1)
    for kk, c in enumerate(mono_v):
        for j in range(kk + 1):
            coef_u[j] += mp.mpf(c) * comb(kk, j) * (mp.mpf(A) ** j) * (mp.mpf(B) ** (kk - j))
    cu = np.array([float(c) for c in coef_u])
    print("kNdF (descending powers of u):")
    print(", ".join("%.17e" % c for c in cu[::-1]))
    tt = np.concatenate([np.linspace(0, 26, 20001), np.random.default_rng(0).uniform(0, 26, 20000)])
    ut = 1.0 - 10.0 / (tt + C)
    Fu = horner(cu, ut)
    ln2hi, ln2lo, l2e = 6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.4426950408889634
    tay = np.array([1.0 / math.factorial(i) for i in range(12)])
    y = -0.5 * (tt * tt)
    nf = np.rint(y * l2e)
    r = (y - nf * ln2hi) - nf * ln2lo
    E = horner(tay, r) * np.exp2(nf)
    p = E * Fu
    pref = np.array([float(0.5 * mp.erfc(mp.mpf(t) / mp.sqrt(2))) for t in tt])
    print("max relative error of Phi(-t), 0 <= t <= 26: %.3e" % np.max(np.abs(p - pref) / pref))


if __name__ == "__main__":
    main()
