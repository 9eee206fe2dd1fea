"""Feasibility of a table + Taylor normal tail for the window kernel: Phi(-t) for t in [0, 8] from
(Phi(-t0), phi(t0)) on a 1/64 grid and a Hermite-polynomial step of order 4 / 5 / 6 in d = t0 - t, against scipy's ndtr
on 2e6 random points. Upper tail by symmetry (1 - Phi(-t)), as the kernel does today. CPU only; nothing here ships.

    python tools/fit_ndtr_table.py
"""
import numpy as np
from scipy.special import ndtr


def main():
    h = 1.0 / 64.0
    x0s = np.arange(-8, 0 + h / 2, h)
    Phi0 = ndtr(x0s)
    phi0 = np.exp(-x0s * x0s / 2) / np.sqrt(2 * np.pi)
    rng = np.random.default_rng(0)
    x = -rng.uniform(0, 8, 2_000_000)
    i = np.clip(np.rint((x + 8) / h).astype(int), 0, len(x0s) - 1)
    x0, d = x0s[i], x - x0s[i]
    ref = ndtr(x)
    # Phi(x0 + d) = Phi(x0) + phi(x0) d [1 - x0 d/2 + (x0^2-1) d^2/6 - (x0^3-3x0) d^3/24 + He4 d^4/120 - He5 d^5/720]
    c = [np.ones_like(x0), -x0 / 2, (x0 ** 2 - 1) / 6, -(x0 ** 3 - 3 * x0) / 24, (x0 ** 4 - 6 * x0 ** 2 + 3) / 120,
         -(x0 ** 5 - 10 * x0 ** 3 + 15 * x0) / 720]
    for order in (4, 5, 6):
        acc = np.zeros_like(x)
        for k in range(order - 1, -1, -1):
            acc = acc * d + c[k]
        a = Phi0[i] + phi0[i] * d * acc
        rel = np.abs(a - ref) / ref
        print("order %d: max relative error of the tail %.2e (FP64 FMAs per value: %d + index/offset 3)" % (order, rel.max(), order + 1))
    print("table: %d entries; with the %d step coefficients stored per entry: %d KB of shared memory"
          % (len(x0s), 5, len(x0s) * 7 * 8 // 1024))


if __name__ == "__main__":
    main()
