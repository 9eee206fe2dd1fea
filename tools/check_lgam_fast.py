"""Accuracy of lgam_pos_fast (fpt_ops.cu, the fused posterior kernel's log-gamma) against mpmath: the same formula in
numpy — Stirling's series with seven correction terms at an argument >= 7, reached through
Gamma(x) = Gamma(x + 7) / (x (x+1) ... (x+6)) below 7. Prints the largest absolute error on (1e-6, 5000) next to
scipy.special.gammaln's (both are at the rounding of the result: ~1.5e-11 at lgam(5000) ~ 3.8e4)."""
import mpmath as mp
import numpy as np
from scipy.special import gammaln


def lgam_pos_fast(x):
    x = np.asarray(x, dtype=np.float64)
    small = x < 7.0
    z = np.where(small, ((x * (x + 1)) * ((x + 2) * (x + 3))) * (((x + 4) * (x + 5)) * (x + 6)), 1.0)
    xs = np.where(small, x + 7.0, x)
    inv = 1.0 / xs
    w = inv * inv
    c = [1 / 12., -1 / 360., 1 / 1260., -1 / 1680., 1 / 1188., -691 / 360360., 1 / 156.]
    s = c[6]
    for k in range(5, -1, -1):
        s = s * w + c[k]
    return (xs - 0.5) * np.log(xs) + (s * inv + (0.91893853320467274178 - xs)) - np.where(small, np.log(z), 0.0)


rng = np.random.default_rng(0)
xs = np.concatenate([10 ** rng.uniform(-6, 0, 2000), rng.uniform(0, 7, 4000), rng.uniform(7, 60, 4000), rng.uniform(60, 5000, 2000)])
ref = np.array([float(mp.loggamma(mp.mpf(float(v)))) for v in xs])
for name, got in (("lgam_pos_fast", lgam_pos_fast(xs)), ("scipy gammaln", gammaln(xs))):
    err = np.abs(got - ref)
    print("%-14s max abs err %.3g (at x = %.6g), below 60: %.3g" % (name, err.max(), xs[err.argmax()], err[xs < 60].max()))
