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


# ---- the mixed-precision evaluation of ndtr4 (fpt_tile.cuh) -------------------------------------------
#   Phi(-t) = exp(-t^2/2) * G(u) / (t + 5),  G(u) = (t + 5) * 0.5 * erfcx(t / sqrt 2),  degree 13;
#   the 7 highest-order coefficients of G and the terms q^3/3! .. q^6/6! of exp(q) are summed in float32.
def fit_g(n=13, c=5.0):
    umax = (TM - c) / (TM + c)
    k = np.arange(n + 1)
    v = np.cos(np.pi * (k + 0.5) / (n + 1))
    u = (v + 1) / 2 * (umax + 1) - 1
    t = c * (1 + u) / (1 - u)
    f = np.array([float(F(ti) * (ti + c)) for ti in t])
    mono_v = Ch.cheb2poly(Ch.chebfit(v, f, n))
    a, b = 2 / (umax + 1), 2 / (umax + 1) - 1
    coef_u = [mp.mpf(0)] * (n + 1)
    for kk, ck in enumerate(mono_v):
        for j in range(kk + 1):
            coef_u[j] += mp.mpf(ck) * comb(kk, j) * (mp.mpf(a) ** j) * (mp.mpf(b) ** (kk - j))
    return np.array([float(x) for x in coef_u])[::-1]  # descending powers


def replay_mixed(tt, cg, k32=7, div=4, ke64=3, ne=6, c=5.0):
    f32 = np.float32
    lmp = mp.log(2) / div
    bits = np.array([float(lmp)], dtype=np.float64).view(np.int64)
    lhi = (bits & np.int64(~((1 << 21) - 1))).view(np.float64)[0]
    llo = float(lmp - mp.mpf(lhi))
    r = (1.0 / (tt + c)) * (1 + 9e-13)  # one Newton step on MUFU.RCP64H
    u = 1.0 - 2 * c * r
    uf = u.astype(f32)
    a = np.full(uf.shape, f32(cg[0]), dtype=f32)
    for ck in cg[1:k32]:
        a = (a.astype(np.float64) * uf.astype(np.float64) + np.float64(f32(ck))).astype(f32)  # FFMA
    g = a.astype(np.float64)
    for ck in cg[k32:]:
        g = g * u + ck
    y = -0.5 * (tt * tt)
    nf = np.rint(y * (div / math.log(2)))
    q = (y - nf * lhi) - nf * llo
    qf = q.astype(f32)
    a = np.full(qf.shape, f32(1.0 / math.factorial(ne)), dtype=f32)
    for k in range(ne - 1, ke64 - 1, -1):
        a = (a.astype(np.float64) * qf.astype(np.float64) + np.float64(f32(1.0 / math.factorial(k)))).astype(f32)
    pe = a.astype(np.float64)
    for k in range(ke64 - 1, -1, -1):
        pe = pe * q + 1.0 / math.factorial(k)
    n = nf.astype(np.int64)
    j = n & (div - 1)
    s = np.array([float(mp.mpf(2) ** (mp.mpf(i) / div)) for i in range(div)])[j]
    return ((pe * s * np.exp2((n - j) // div)) * r) * g, lhi, llo


def main_mixed():
    cg = fit_g()
    print("G coefficients (descending powers of u; the first 7 are used as float32):")
    print(", ".join("%.17e" % x for x in cg))
    tt = np.concatenate([np.linspace(0, 26, 20001), np.random.default_rng(0).uniform(0, 26, 20000)])
    p, lhi, llo = replay_mixed(tt, cg)
    print("ln2/4 = %.20e + %.20e" % (lhi, llo))
    pref = np.array([float(0.5 * mp.erfc(mp.mpf(t) / mp.sqrt(2))) for t in tt])
    print("max relative error of Phi(-t), 0 <= t <= 26 (mixed FP32/FP64): %.3e" % np.max(np.abs(p - pref) / pref))


if __name__ == "__main__":
    import sys

    if "--mixed" in sys.argv:
        main_mixed()
    else:
        main()
