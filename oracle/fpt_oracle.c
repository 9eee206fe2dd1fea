/* fpt_oracle.c — CPU ORACLE for the footprint-tools per-nucleotide scoring path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity checker, not the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product path (footprint-tools_b200/) never links, imports or calls it and
 * fails loudly when its CUDA library is missing.
 *
 * What it is: a plain-C restatement, written from scratch, of the reference's algorithm
 * for the hot path. Every function cites the reference file:line it
 * follows (paths relative to /root/reference). It is compiled with
 * -O2 -ffp-contract=off (no FMA), like the reference's own x86-64 baseline build.
 *
 * How it is pinned (see tests/test_oracle_*.py):
 *   - against oracle/_ref/libref.so = the reference's own C compiled in place, and
 *   - against tests/golden/*.npz = outputs of the reference's Cython/Python API generated
 *     in the build container by tests/golden/make_golden.py.
 * Parity status: PINNED (bit-exact vs libref.so on the special functions, fast_predict and
 * the window reducers; the Python-level geometry is pinned by the golden files).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* hcephes/include/hcephes.h:73-86 */
static const double kMachEp = 1.11022302462515654042E-16; /* 2^-53 */
static const double kMaxLog = 7.09782712893383996732E2;
static const double kMinLog = -7.451332191019412076235E2;
static const double kMaxGam = 171.624376956302725; /* incbet.c:3, gamma.c:18 */
static const double kPi = 3.14159265358979323846;
static const double kSqrtH = 7.07106781186547524401E-1;
static const double kBig = 4.503599627370496e15;       /* incbet.c:5, igam.c:3 */
static const double kBigInv = 2.22044604925031308085e-16; /* incbet.c:6, igam.c:4 */

/* ---- Horner evaluators: hcephes/src/polyn/polevl.c:3-32 --------------------------------- */
static double horner(double x, const double *c, int n) { /* polevl: degree n, n+1 coefs */
    double a = c[0];
    for (int i = 1; i <= n; ++i) a = a * x + c[i];
    return a;
}
static double horner1(double x, const double *c, int n) { /* p1evl: leading coef 1 implied */
    double a = x + c[0];
    for (int i = 1; i < n; ++i) a = a * x + c[i];
    return a;
}

/* ---- gamma / lgam: hcephes/src/cprob/gamma.c ---------------------------------------------- */
static const double gP[7] = {1.60119522476751861407E-4, 1.19135147006586384913E-3, 1.04213797561761569935E-2,
                             4.76367800457137231464E-2, 2.07448227648435975150E-1, 4.94214826801497100753E-1,
                             9.99999999999999996796E-1};
static const double gQ[8] = {-2.31581873324120129819E-5, 5.39605580493303397842E-4, -4.45641913851797240494E-3,
                             1.18139785222060435552E-2,  3.58236398605498653373E-2, -2.34591795718243348568E-1,
                             7.14304917030273074085E-2,  1.00000000000000000320E0};
static const double gStir[5] = {7.87311395793093628397E-4, -2.29549961613378126380E-4, -2.68132617805781232825E-3,
                                3.47222221605458667310E-3, 8.33333333333482257126E-2};
static const double kSqrt2Pi = 2.50662827463100050242E0;
static const double kLogPi = 1.14472988584940017414;
static const double kLogSqrt2Pi = 0.91893853320467274178;

/* gamma.c:35-49 */
static double stirling_gamma(double x) {
    double w = 1.0 / x;
    w = 1.0 + w * horner(w, gStir, 4);
    double y = exp(x);
    if (x > 143.01608) {
        double v = pow(x, 0.5 * x - 0.25);
        y = v * (v / y);
    } else {
        y = pow(x, x - 0.5) / y;
    }
    return kSqrt2Pi * y * w;
}

/* gamma.c:51-127 */
ORC_API double orc_gamma(double x) {
    int sgn = 1;
    if (isnan(x)) return x;
    if (x == HUGE_VAL) return x;
    if (x == -HUGE_VAL) return NAN;
    double q = fabs(x);
    if (q > 33.0) {
        double z;
        if (x < 0.0) {
            double p = floor(q);
            if (p == q) return NAN;
            if ((((int)p) & 1) == 0) sgn = -1;
            z = q - p;
            if (z > 0.5) {
                p += 1.0;
                z = q - p;
            }
            z = q * sin(kPi * z);
            if (z == 0.0) return sgn * HUGE_VAL;
            z = fabs(z);
            z = kPi / (z * stirling_gamma(q));
        } else {
            z = stirling_gamma(x);
        }
        return sgn * z;
    }
    double z = 1.0;
    while (x >= 3.0) {
        x -= 1.0;
        z *= x;
    }
    while (x < 0.0) {
        if (x > -1.E-9) goto tiny;
        z /= x;
        x += 1.0;
    }
    while (x < 2.0) {
        if (x < 1.e-9) goto tiny;
        z /= x;
        x += 1.0;
    }
    if (x == 2.0) return z;
    x -= 2.0;
    return z * horner(x, gP, 6) / horner(x, gQ, 7);
tiny:
    if (x == 0.0) return NAN;
    return z / ((1.0 + 0.5772156649015329 * x) * x);
}

static const double lgA[5] = {8.11614167470508450300E-4, -5.95061904284301438324E-4, 7.93650340457716943945E-4,
                              -2.77777777730099687205E-3, 8.33333333333331927722E-2};
static const double lgB[6] = {-1.37825152569120859100E3, -3.88016315134637840924E4, -3.31612992738871184744E5,
                              -1.16237097492762307383E6, -1.72173700820839662146E6, -8.53555664245765465627E5};
static const double lgC[6] = {-3.51815701436523470549E2, -1.70642106651881159223E4, -2.20528590553854454839E5,
                              -1.13933444367982507207E6, -2.53252307177582951285E6, -2.01889141433532773231E6};

/* gamma.c:147-235 (hcephes_lgam -> hcephes_lgam_sgn; the sign is dropped by hcephes_lgam) */
ORC_API double orc_lgam(double x) {
    if (isnan(x)) return x;
    if (!isfinite(x)) return HUGE_VAL;
    if (x < -34.0) {
        double q = -x;
        double w = orc_lgam(q);
        double p = floor(q);
        if (p == q) return HUGE_VAL;
        double z = q - p;
        if (z > 0.5) {
            p += 1.0;
            z = p - q;
        }
        z = q * sin(kPi * z);
        if (z == 0.0) return HUGE_VAL;
        return kLogPi - log(z) - w;
    }
    if (x < 13.0) {
        double z = 1.0, p = 0.0, u = x;
        while (u >= 3.0) {
            p -= 1.0;
            u = x + p;
            z *= u;
        }
        while (u < 2.0) {
            if (u == 0.0) return HUGE_VAL;
            z /= u;
            p += 1.0;
            u = x + p;
        }
        if (z < 0.0) z = -z;
        if (u == 2.0) return log(z);
        p -= 2.0;
        x = x + p;
        p = x * horner(x, lgB, 5) / horner1(x, lgC, 6);
        return log(z) + p;
    }
    if (x > 2.556348e305) return HUGE_VAL;
    double q = (x - 0.5) * log(x) - x + kLogSqrt2Pi;
    if (x > 1.0e8) return q;
    double p = 1.0 / (x * x);
    if (x >= 1000.0)
        q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
    else
        q += horner(p, lgA, 4) / x;
    return q;
}

/* ---- log1p: hcephes/src/cprob/unity.c:29-38 ---------------------------------------------- */
static const double l1P[7] = {4.5270000862445199635215E-5, 4.9854102823193375972212E-1, 6.5787325942061044846969E0,
                              2.9911919328553073277375E1,  6.0949667980987787057556E1,  5.7112963590585538103336E1,
                              2.0039553499201281259648E1};
static const double l1Q[6] = {1.5062909083469192043167E1, 8.3047565967967209469434E1, 2.2176239823732856465394E2,
                              3.0909872225312059774938E2, 2.1642788614495947685003E2, 6.0118660497603843919306E1};
ORC_API double orc_log1p(double x) {
    double z = 1.0 + x;
    if (z < 0.70710678118654752440 || z > 1.41421356237309504880) return log(z);
    z = x * x;
    z = -0.5 * z + x * (z * horner(x, l1P, 6) / horner1(x, l1Q, 6));
    return x + z;
}

/* ---- regularized incomplete beta: hcephes/src/cprob/incbet.c ------------------------------- */

/* incbet.c:266-299 */
static double beta_power_series(double a, double b, double x) {
    double ai = 1.0 / a;
    double u = (1.0 - b) * x;
    double v = u / (a + 1.0);
    double t1 = v, t = u, n = 2.0, s = 0.0;
    double z = kMachEp * ai;
    while (fabs(v) > z) {
        u = (n - b) * x / n;
        t *= u;
        v = t / (a + n);
        s += v;
        n += 1.0;
    }
    s += t1;
    s += ai;
    u = a * log(x);
    if ((a + b) < kMaxGam && fabs(u) < kMaxLog) {
        t = orc_gamma(a + b) / (orc_gamma(a) * orc_gamma(b));
        s = s * t * pow(x, a);
    } else {
        t = orc_lgam(a + b) - orc_lgam(a) - orc_lgam(b) + u + log(s);
        s = (t < kMinLog) ? 0.0 : exp(t);
    }
    return s;
}

/* incbet.c:100-177 (which==0, "incbcf") and :183-261 (which==1, "incbd"); the two continued
 * fractions share the recurrence and differ only in the k-coefficient schedule. */
static double beta_cfrac(double a, double b, double x, int which, int *iters) {
    double k1 = a, k3 = a, k4 = a + 1.0, k5 = 1.0, k7 = a + 1.0, k8 = a + 2.0;
    double k2, k6, d2, d6, zz;
    if (which == 0) {
        k2 = a + b; k6 = b - 1.0; d2 = 1.0; d6 = -1.0; zz = x;
    } else {
        k2 = b - 1.0; k6 = a + b; d2 = -1.0; d6 = 1.0; zz = x / (1.0 - x);
    }
    double pkm2 = 0.0, qkm2 = 1.0, pkm1 = 1.0, qkm1 = 1.0, ans = 1.0, r = 1.0, t;
    const double thresh = 3.0 * kMachEp;
    int n = 0;
    do {
        double xk = -(zz * k1 * k2) / (k3 * k4);
        double pk = pkm1 + pkm2 * xk;
        double qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        xk = (zz * k5 * k6) / (k7 * k8);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;

        if (qk != 0) r = pk / qk;
        if (r != 0) {
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;
        if (t < thresh) break;

        k1 += 1.0; k2 += d2; k3 += 2.0; k4 += 2.0;
        k5 += 1.0; k6 += d6; k7 += 2.0; k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > kBig) {
            pkm2 *= kBigInv; pkm1 *= kBigInv; qkm2 *= kBigInv; qkm1 *= kBigInv;
        }
        if ((fabs(qk) < kBigInv) || (fabs(pk) < kBigInv)) {
            pkm2 *= kBig; pkm1 *= kBig; qkm2 *= kBig; qkm1 *= kBig;
        }
    } while (++n < 300);
    if (iters) *iters = n + 1;
    return ans;
}

/* incbet.c:12-94. `iters` (optional) receives the continued-fraction iteration count (0 when
 * the power series was used) for the flop accounting of SURVEY.md §8d. */
static double incbet_impl(double aa, double bb, double xx, int *iters) {
    double a, b, t, x, xc, w, y;
    int flag = 0;
    if (iters) *iters = 0;
    if (aa <= 0.0 || bb <= 0.0) return 0.0;
    if (xx <= 0.0 || xx >= 1.0) {
        if (xx == 0.0) return 0.0;
        if (xx == 1.0) return 1.0;
        return 0.0;
    }
    if ((bb * xx) <= 1.0 && xx <= 0.95) {
        t = beta_power_series(aa, bb, xx);
        goto done;
    }
    w = 1.0 - xx;
    if (xx > (aa / (aa + bb))) {
        flag = 1; a = bb; b = aa; xc = xx; x = w;
    } else {
        a = aa; b = bb; xc = w; x = xx;
    }
    if (flag == 1 && (b * x) <= 1.0 && x <= 0.95) {
        t = beta_power_series(a, b, x);
        goto done;
    }
    y = x * (a + b - 2.0) - (a - 1.0);
    if (y < 0.0)
        w = beta_cfrac(a, b, x, 0, iters);
    else
        w = beta_cfrac(a, b, x, 1, iters) / xc;

    y = a * log(x);
    t = b * log(xc);
    if ((a + b) < kMaxGam && fabs(y) < kMaxLog && fabs(t) < kMaxLog) {
        t = pow(xc, b);
        t *= pow(x, a);
        t /= a;
        t *= w;
        t *= orc_gamma(a + b) / (orc_gamma(a) * orc_gamma(b));
        goto done;
    }
    y += t + orc_lgam(a + b) - orc_lgam(a) - orc_lgam(b);
    y += log(w / a);
    t = (y < kMinLog) ? 0.0 : exp(y);
done:
    if (flag == 1) {
        if (t <= kMachEp)
            t = 1.0 - kMachEp;
        else
            t = 1.0 - t;
    }
    return t;
}
ORC_API double orc_incbet(double a, double b, double x) { return incbet_impl(a, b, x, NULL); }
ORC_API double orc_incbet_iters(double a, double b, double x, int *iters) { return incbet_impl(a, b, x, iters); }

/* ---- incomplete gamma / chi-square tail: hcephes/src/cprob/igam.c, chdtr.c ---------------- */
ORC_API double orc_igam(double a, double x);

/* igam.c:6-59 */
ORC_API double orc_igamc(double a, double x) {
    if (x <= 0 || a <= 0) return 1.0;
    if (x < 1.0 || x < a) return 1.0 - orc_igam(a, x);
    double ax = a * log(x) - x - orc_lgam(a);
    if (ax < -kMaxLog) return 0.0;
    ax = exp(ax);
    double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
    double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
    double ans = pkm1 / qkm1, t;
    do {
        c += 1.0;
        y += 1.0;
        z += 2.0;
        double yc = y * c;
        double pk = pkm1 * z - pkm2 * yc;
        double qk = qkm1 * z - qkm2 * yc;
        if (qk != 0) {
            double r = pk / qk;
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;
        pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
        if (fabs(pk) > kBig) {
            pkm2 *= kBigInv; pkm1 *= kBigInv; qkm2 *= kBigInv; qkm1 *= kBigInv;
        }
    } while (t > kMachEp);
    return ans * ax;
}

/* igam.c:71-100 */
ORC_API double orc_igam(double a, double x) {
    if (x <= 0 || a <= 0) return 0.0;
    if (x > 1.0 && x > a) return 1.0 - orc_igamc(a, x);
    double ax = a * log(x) - x - orc_lgam(a);
    if (ax < -kMaxLog) return 0.0;
    ax = exp(ax);
    double r = a, c = 1.0, ans = 1.0;
    do {
        r += 1.0;
        c *= x / r;
        ans += c;
    } while (c / ans > kMachEp);
    return ans * ax / a;
}

/* chdtr.c:3-10 */
ORC_API double orc_chdtrc(double df, double x) {
    if (x < 0.0 || df < 1.0) return 0.0;
    return orc_igamc(df / 2.0, x / 2.0);
}

/* ---- normal CDF and its inverse: hcephes/src/cprob/ndtr.c, expx2.c, ndtri.c ---------------- */
static const double nP[9] = {2.46196981473530512524E-10, 5.64189564831068821977E-1, 7.46321056442269912687E0,
                             4.86371970985681366614E1,   1.96520832956077098242E2,  5.26445194995477358631E2,
                             9.34528527171957607540E2,   1.02755188689515710272E3,  5.57535335369399327526E2};
static const double nQ[8] = {1.32281951154744992508E1, 8.67072140885989742329E1, 3.54937778887819891062E2,
                             9.75708501743205489753E2, 1.82390916687909736289E3, 2.24633760818710981792E3,
                             1.65666309194161350182E3, 5.57535340817727675546E2};
static const double nR[6] = {5.64189583547755073984E-1, 1.27536670759978104416E0, 5.01905042251180477414E0,
                             6.16021097993053585195E0,  7.40974269950448939160E0, 2.97886665372100240670E0};
static const double nS[6] = {2.26052863220117276590E0, 9.39603524938001434673E0, 1.20489539808096656605E1,
                             1.70814450747565897222E1, 9.60896809063285878198E0, 3.36907645100081516050E0};
static const double nT[5] = {9.60497373987051638749E0, 9.00260197203842689217E1, 2.23200534594684319226E3,
                             7.00332514112805075473E3, 5.55923013010394962768E4};
static const double nU[5] = {3.35617141647503099647E1, 5.21357949780152679795E2, 4.59432382970980127987E3,
                             2.26290000613890934246E4, 4.92673942608635921086E4};

/* expx2.c:6-34 */
static double exp_x2(double x, int sign) {
    x = fabs(x);
    if (sign < 0) x = -x;
    double m = .0078125 * floor(128.0 * x + 0.5);
    double f = x - m;
    double u = m * m;
    double u1 = 2 * m * f + f * f;
    if (sign < 0) {
        u = -u;
        u1 = -u1;
    }
    if ((u + u1) > kMaxLog) return HUGE_VAL;
    return exp(u) * exp(u1);
}

/* ndtr.c:65-76 */
static double erfc_scaled(double x) {
    if (x < 8.0) return horner(x, nP, 8) / horner1(x, nQ, 8);
    return horner(x, nR, 5) / horner1(x, nS, 6);
}

static double erfc_full(double a);
/* ndtr.c:78-86 */
static double erf_full(double x) {
    if (fabs(x) > 1.0) return 1.0 - erfc_full(x);
    double z = x * x;
    return x * horner(z, nT, 4) / horner1(z, nU, 5);
}
/* ndtr.c:88-132 */
static double erfc_full(double a) {
    double x = (a < 0.0) ? -a : a;
    if (x < 1.0) return 1.0 - erf_full(a);
    double z = -a * a;
    if (z < -kMaxLog) return (a < 0) ? 2.0 : 0.0;
    z = exp_x2(a, -1);
    double p, q;
    if (x < 8.0) {
        p = horner(x, nP, 8);
        q = horner1(x, nQ, 8);
    } else {
        p = horner(x, nR, 5);
        q = horner1(x, nS, 6);
    }
    double y = (z * p) / q;
    if (a < 0) y = 2.0 - y;
    if (y == 0.0) return (a < 0) ? 2.0 : 0.0;
    return y;
}

/* ndtr.c:34-59 (USE_EXPXSQ branch) */
ORC_API double orc_ndtr(double a) {
    double x = a * kSqrtH;
    double z = fabs(x);
    double y;
    if (z < 1.0)
        y = 0.5 + 0.5 * erf_full(x);
    else {
        y = 0.5 * erfc_scaled(z);
        z = exp_x2(a, -1);
        y = y * sqrt(z);
        if (x > 0) y = 1.0 - y;
    }
    return y;
}

static const double iP0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                              1.39312609387279679503E1, -1.23916583867381258016E0};
static const double iQ0[8] = {1.95448858338141759834E0,  4.67627912898881538453E0, 8.63602421390890590575E1,
                              -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                              1.59056225126211695515E1,  -1.18331621121330003142E0};
static const double iP1[9] = {4.05544892305962419923E0,   3.15251094599893866154E1,   5.71628192246421288162E1,
                              4.40805073893200834700E1,   1.46849561928858024014E1,   2.18663306850790267539E0,
                              -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
static const double iQ1[8] = {1.57799883256466749731E1,   4.53907635128879210584E1,   4.13172038254672030440E1,
                              1.50425385692907503408E1,   2.50464946208309415979E0,   -1.42182922854787788574E-1,
                              -3.80806407691578277194E-2, -9.33259480895457427372E-4};
static const double iP2[9] = {3.23774891776946035970E0,  6.91522889068984211695E0,  3.93881025292474443415E0,
                              1.33303460815807542389E0,  2.01485389549179081538E-1, 1.23716634817820021358E-2,
                              3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
static const double iQ2[8] = {6.02427039364742014255E0,  3.67983563856160859403E0,  1.37702099489081330271E0,
                              2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                              2.89247864745380683936E-6, 6.79019408009981274425E-9};

/* ndtri.c:48-88 */
ORC_API double orc_ndtri(double y0) {
    const double em2 = 0.13533528323661269189; /* exp(-2) */
    if (y0 <= 0.0) return -HUGE_VAL;
    if (y0 >= 1.0) return HUGE_VAL;
    int negate = 1;
    double y = y0;
    if (y > 1.0 - em2) {
        y = 1.0 - y;
        negate = 0;
    }
    if (y > em2) {
        y = y - 0.5;
        double y2 = y * y;
        double x = y + y * (y2 * horner(y2, iP0, 4) / horner1(y2, iQ0, 8));
        return x * kSqrt2Pi;
    }
    double x = sqrt(-2.0 * log(y));
    double x0 = x - log(x) / x;
    double z = 1.0 / x, x1;
    if (x < 8.0)
        x1 = z * horner(z, iP1, 8) / horner1(z, iQ1, 8);
    else
        x1 = z * horner(z, iP2, 8) / horner1(z, iQ2, 8);
    x = x0 - x1;
    return negate ? -x : x;
}

/* ---- negative binomial: footprint_tools/stats/distributions/nbinom.pyx:82-138 ------------- */
ORC_API double orc_nb_logpmf(int k, double p, double r) {
    double coeff = orc_lgam(k + r) - orc_lgam(k + 1) - orc_lgam(r);
    return coeff + r * log(p) + k * orc_log1p(-p);
}
ORC_API double orc_nb_pmf(int k, double p, double r) { return exp(orc_nb_logpmf(k, p, r)); }
ORC_API double orc_nb_cdf(int k, double p, double r) { return orc_incbet(r, k + 1, p); }

/* ---- dispersion model: footprint_tools/modeling/dispersion.pyx:26-57, 127-163 -------------
 * Parameters are laid out as in the reference: [breaks..., intercepts..., slopes...].
 * Each segment is `y + k*x` (a multiply then an add, separately rounded); the unselected
 * segments contribute (bool 0)*(value) = +-0.0, which leaves the selected value unchanged. */
static double piecewise(const double *par, int nseg, double x) {
    const double *brk = par, *icpt = par + nseg, *slope = par + 2 * nseg;
    double acc = 0.0;
    for (int s = 0; s < nseg; ++s) {
        int sel;
        if (s == 0) sel = (x < brk[0]);
        else if (s == nseg - 1) sel = (x >= brk[nseg - 2]);
        else sel = (x >= brk[s - 1]) && (x < brk[s]);
        double term = (double)sel * (icpt[s] + slope[s] * x);
        acc = (s == 0) ? term : acc + term;
    }
    return acc;
}
ORC_API double orc_fit_mu(const double *mu_params, double x) { /* dispersion.pyx:127-144 */
    double v = piecewise(mu_params, 3, x);
    return v > 0.0 ? v : 0.1;
}
ORC_API double orc_fit_r(const double *r_params, double x) { /* dispersion.pyx:146-163 */
    double v = 1.0 / piecewise(r_params, 5, x);
    return v > 0.0 ? v : 1e-6;
}

/* dispersion.pyx:170-316: what = 0 cdf (p_values), 1 pmf, 2 logpmf */
ORC_API void orc_dm_values(const double *mu_params, const double *r_params, const double *expv, const double *obsv,
                           long n, int what, double *out) {
    for (long i = 0; i < n; ++i) {
        double r = orc_fit_r(r_params, expv[i]);
        double mu = orc_fit_mu(mu_params, expv[i]);
        int k = (int)obsv[i];
        double p = r / (r + mu);
        out[i] = what == 0 ? orc_nb_cdf(k, p, r) : what == 1 ? orc_nb_pmf(k, p, r) : orc_nb_logpmf(k, p, r);
    }
}

/* ---- k-mer bias: footprint_tools/modeling/bias.py:16-17,88-111; predict.pyx:47-61,151-153 --
 * `seq` has n characters; out has n-6 values. strand>0: out[u] = model[seq[u:u+6]];
 * strand<0: out[u] = model[revcomp(seq[u+1:u+7])] (= probs(revcomp(seq))[::-1], SURVEY hard
 * part 7). table is indexed A=0,C=1,G=2,T=3 base-4 big-endian; any other character in the
 * 6-mer gives `dflt` (1e-6 for k-mer models). With uniform!=0 the model ignores the sequence
 * (bias.py:121-122). */
static int base_code(char c) {
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}
ORC_API void orc_kmer_probs(const char *seq, long n, const double *table, double dflt, int strand, int uniform,
                            double *out) {
    for (long u = 0; u + 6 < n; ++u) { /* range(offset, len - offset): n-6 values */
        if (uniform) { out[u] = 1.0; continue; }
        int idx = 0, bad = 0;
        if (strand > 0) {
            for (int j = 0; j < 6; ++j) {
                int c = base_code(seq[u + j]);
                if (c < 0) bad = 1;
                idx = idx * 4 + (c & 3);
            }
        } else {
            for (int j = 0; j < 6; ++j) {
                int c = base_code(seq[u + 6 - j]);
                if (c < 0) bad = 1;
                idx = idx * 4 + (3 - (c & 3));
            }
        }
        out[u] = bad ? dflt : table[idx];
    }
}

/* ---- trimmed-mean smoothing: footprint_tools/modeling/smoothing.h -------------------------- */

/* smoothing.h:11-53 (Numerical-Recipes style select; permutes arr in place) */
static double nr_select(double *arr, unsigned int n, unsigned int k) {
    unsigned long lo = 0, hi = n - 1;
    for (;;) {
        if (hi <= lo + 1) {
            if (hi == lo + 1 && arr[hi] < arr[lo]) { double t = arr[lo]; arr[lo] = arr[hi]; arr[hi] = t; }
            return arr[k];
        }
        unsigned long mid = (lo + hi) >> 1;
        double t;
        t = arr[mid]; arr[mid] = arr[lo + 1]; arr[lo + 1] = t;
        if (arr[lo] > arr[hi]) { t = arr[lo]; arr[lo] = arr[hi]; arr[hi] = t; }
        if (arr[lo + 1] > arr[hi]) { t = arr[lo + 1]; arr[lo + 1] = arr[hi]; arr[hi] = t; }
        if (arr[lo] > arr[lo + 1]) { t = arr[lo]; arr[lo] = arr[lo + 1]; arr[lo + 1] = t; }
        unsigned long i = lo + 1, j = hi;
        double piv = arr[lo + 1];
        for (;;) {
            do i++; while (arr[i] < piv);
            do j--; while (arr[j] > piv);
            if (j < i) break;
            t = arr[i]; arr[i] = arr[j]; arr[j] = t;
        }
        arr[lo + 1] = arr[j];
        arr[j] = piv;
        if (j >= k) hi = j - 1;
        if (j <= k) lo = i;
    }
}

/* smoothing.h:59-104: tie-weighted trimmed sum over the (permuted) buffer, then /(n-2k) */
ORC_API double orc_trimmed_mean(double *x, int n, int k) {
    double os1 = nr_select(x, n, k);
    double os2 = nr_select(x, n, n - k - 1);
    double b = 0, d = 0, dm = 0, bm = 0;
    for (int i = 0; i < n; ++i) {
        double v = x[i];
        if (v < os1) bm += 1; else if (v == os1) b += 1;
        if (v < os2) dm += 1; else if (v == os2) d += 1;
    }
    double w1 = (b + bm - k) / b;
    double w2 = (n - k - dm) / d;
    double t = 0;
    for (int i = 0; i < n; ++i) {
        double v = x[i], c;
        if (v < os2 && v > os1) c = v;
        else if (v < os1) c = 0;
        else if (v > os2) c = 0;
        else if (v == os1) c = w1 * v;
        else c = w2 * v;
        t += c;
    }
    return t / (n - 2 * k);
}

/* ---- expected counts: footprint_tools/modeling/predict.h:23-74, smoothing.h:107-132 -------
 * exp_out and win_out have l entries (zero outside the computed range). */
ORC_API void orc_fast_predict(const double *obs, const double *probs, int l, int hw, int shw, double clip,
                              double *exp_out, double *win_out) {
    double *wc = (double *)calloc(l > 0 ? l : 1, sizeof(double));
    double *wp = (double *)calloc(l > 0 ? l : 1, sizeof(double));
    for (int i = hw; i < l - hw; ++i)
        for (int j = -hw; j < hw; ++j) {
            wc[i] += obs[i + j];
            wp[i] += probs[i + j];
        }
    if (shw > 0) {
        int w = 2 * shw + 1;
        int k = (int)((double)w * clip);
        double *tmp = (double *)malloc(w * sizeof(double));
        double *sm = (double *)calloc(l > 0 ? l : 1, sizeof(double));
        for (int i = shw; i < l - shw; ++i) {
            memcpy(tmp, wc + (i - shw), w * sizeof(double));
            sm[i] = orc_trimmed_mean(tmp, w, k);
        }
        free(tmp);
        free(wc);
        wc = sm;
    }
    memset(exp_out, 0, sizeof(double) * (l > 0 ? l : 0));
    for (int i = hw; i < l - hw; ++i) exp_out[i] = round((probs[i] / wp[i]) * wc[i]);
    for (int i = 0; i < l; ++i) win_out[i] = wc[i];
    free(wc);
    free(wp);
}

/* ---- window reducers: footprint_tools/stats/windowing.h:11-123, windowing.pyx:34-58,132-158
 * op: 0 sum, 1 product, 2 Fisher, 3 Stouffer, 4 weighted Stouffer (needs w). Positions
 * i<hw or i>=n-hw are 1.0. */
ORC_API void orc_window(const double *x, const double *wt, long n, int hw, int op, double *out) {
    int k = 2 * hw + 1;
    for (long i = 0; i < n; ++i) out[i] = 1.0;
    for (long i = hw; i < n - hw; ++i) {
        const double *v = x + (i - hw);
        double s = 0.0, res;
        switch (op) {
        case 0:
            for (int j = 0; j < k; ++j) s += v[j];
            res = s;
            break;
        case 1:
            s = 1.0;
            for (int j = 0; j < k; ++j) s *= v[j];
            res = s;
            break;
        case 2:
            for (int j = 0; j < k; ++j) s += log(v[j]);
            s *= -2.0;
            res = orc_chdtrc((double)2.0 * k, s);
            break;
        case 3:
            for (int j = 0; j < k; ++j) s += orc_ndtri(1.0 - v[j]);
            res = orc_ndtr(-(s / sqrt((double)k)));
            break;
        default: {
            const double *ww = wt + (i - hw);
            double sw = 0.0;
            for (int j = 0; j < k; ++j) {
                s += ww[j] * orc_ndtri(1.0 - v[j]);
                sw += ww[j] * ww[j];
            }
            res = orc_ndtr(-(s / sqrt(sw)));
        }
        }
        out[i] = res;
    }
}

/* ---- learn_dm histogram: footprint_tools/cli/learn_dm.py:276-287 -------------------------- */
ORC_API void orc_hist2d(const double *expv, const double *obsv, long n, int d0, int d1, int64_t *hist) {
    for (long i = 0; i < n; ++i) {
        long e = (long)expv[i], o = (long)obsv[i];
        /* python: negative indices wrap, out-of-range raises IndexError (ignored) */
        if (e < 0) e += d0;
        if (o < 0) o += d1;
        if (e < 0 || e >= d0 || o < 0 || o >= d1) continue;
        hist[e * (long)d1 + o] += 1;
    }
}

/* ---- one interval exactly as the callers do it ---------------------------------------------
 * footprint_tools/modeling/predict.pyx:116-163 (padding, per-strand predict, crop) followed by
 * footprint_tools/cli/detect.py:121-130 (strand combine, p_values, Stouffer) or
 * footprint_tools/cli/learn_dm.py:106-109 (strand combine only).
 *
 * Inputs for an interval of length len with pad = hw+shw, L = len + 2*pad + 1:
 *   seq        L+6 characters   (fasta.fetch(start-pad-1-3, end+pad+3))
 *   cuts_plus  L doubles, cuts_minus L doubles (read_func[padded interval])
 * Outputs (each len doubles unless NULL): exp, obs, pval, and winp[s*len..] for each of the
 * n_scales Stouffer half-widths in whw[].
 * The function table lets the same driver run on the reference's own compiled C
 * (oracle/_ref/libref.so) instead of this file's restatement: see orc_fn_table.
 */
typedef struct {
    double *exp;
    double *win;
} ref_result_t; /* layout of `result_t`, predict.h:10-14 */

typedef struct orc_fn_table {
    /* predict.h:23 / :16 */
    ref_result_t *(*fast_predict)(const double *, const double *, int, int, int, double);
    void (*free_result)(ref_result_t *);
    /* incbet.c:12 */
    double (*incbet)(double, double, double);
    /* windowing.h:69, :53 */
    double *(*windowing_func)(const double *, int, int, double (*)(const double *, int));
    double (*stouffers_z)(const double *, int);
} orc_fn_table;

typedef struct {
    int hw, shw;
    double clip;
    const double *table;
    double dflt;
    int uniform;
    const double *mu_params, *r_params; /* NULL => no p-values (learn_dm) */
    int n_scales;
    const int *whw;
    const orc_fn_table *fn; /* NULL => this file's restatement */
} orc_params;

static void score_interval(const orc_params *P, const char *seq, const double *cp, const double *cm, long len,
                           double *o_exp, double *o_obs, double *o_p, double *o_w, long w_stride) {
    const int pad = P->hw + P->shw;
    const int L = (int)(len + 2 * pad + 1);
    double *probs = (double *)malloc(sizeof(double) * L);
    double *e[2], *w[2];
    for (int s = 0; s < 2; ++s) {
        orc_kmer_probs(seq, L + 6, P->table, P->dflt, s == 0 ? 1 : -1, P->uniform, probs);
        const double *cuts = s == 0 ? cp : cm;
        if (P->fn) {
            ref_result_t *r = P->fn->fast_predict(cuts, probs, L, P->hw, P->shw, P->clip);
            e[s] = (double *)malloc(sizeof(double) * L);
            memcpy(e[s], r->exp, sizeof(double) * L);
            w[s] = NULL;
            P->fn->free_result(r);
        } else {
            e[s] = (double *)malloc(sizeof(double) * L);
            w[s] = (double *)malloc(sizeof(double) * L);
            orc_fast_predict(cuts, probs, L, P->hw, P->shw, P->clip, e[s], w[s]);
        }
    }
    /* crop [pad, L-pad) -> len+1; combine plus[1:] + minus[:-1] */
    for (long t = 0; t < len; ++t) {
        o_exp[t] = e[0][pad + t + 1] + e[1][pad + t];
        o_obs[t] = cp[pad + t + 1] + cm[pad + t];
    }
    for (int s = 0; s < 2; ++s) { free(e[s]); free(w[s]); }
    free(probs);
    if (!P->mu_params) return;
    for (long t = 0; t < len; ++t) {
        double r = orc_fit_r(P->r_params, o_exp[t]);
        double mu = orc_fit_mu(P->mu_params, o_exp[t]);
        int k = (int)o_obs[t];
        o_p[t] = P->fn ? P->fn->incbet(r, k + 1, r / (r + mu)) : orc_nb_cdf(k, r / (r + mu), r);
    }
    for (int s = 0; s < P->n_scales; ++s) {
        double *dst = o_w + s * w_stride;
        if (P->fn) {
            for (long t = 0; t < len; ++t) dst[t] = 1.0;
            double *res = P->fn->windowing_func(o_p, (int)len, P->whw[s], P->fn->stouffers_z);
            for (long t = P->whw[s]; t < len - P->whw[s]; ++t) dst[t] = res[t];
            free(res);
        } else {
            orc_window(o_p, NULL, len, P->whw[s], 3, dst);
        }
    }
}

typedef struct {
    const orc_params *P;
    const char *seq;
    const double *cp, *cm;
    const int64_t *in_off;  /* per interval: offset into cuts arrays (seq offset = in_off[k] + 6*k) */
    const int64_t *out_off; /* per interval: offset into outputs; out_off[n_iv] = total */
    long n_iv;
    double *o_exp, *o_obs, *o_p, *o_w;
    int tid, nthreads;
} job_t;

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    long total = J->out_off[J->n_iv];
    for (long k = J->tid; k < J->n_iv; k += J->nthreads) {
        long len = J->out_off[k + 1] - J->out_off[k];
        long oo = J->out_off[k];
        score_interval(J->P, J->seq + J->in_off[k] + 6 * k, J->cp + J->in_off[k], J->cm + J->in_off[k], len,
                       J->o_exp + oo, J->o_obs + oo, J->o_p ? J->o_p + oo : NULL, J->o_w ? J->o_w + oo : NULL, total);
    }
    return NULL;
}

/* Batch driver over intervals packed back to back (interval k: cuts at in_off[k], L_k values per
 * strand; sequence at in_off[k]+6k, L_k+6 characters). Threads split the interval list. */
ORC_API int orc_score_batch(const char *seq, const double *cuts_plus, const double *cuts_minus, const int64_t *in_off,
                            const int64_t *out_off, long n_iv, const double *table, double dflt, int uniform,
                            const double *mu_params, const double *r_params, int hw, int shw, double clip,
                            const int *whw, int n_scales, const orc_fn_table *fn, int nthreads, double *o_exp,
                            double *o_obs, double *o_p, double *o_w) {
    orc_params P = {hw, shw, clip, table, dflt, uniform, mu_params, r_params, n_scales, whw, fn};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        job_t j = {&P, seq, cuts_plus, cuts_minus, in_off, out_off, n_iv, o_exp, o_obs, o_p, o_w, t, nthreads};
        jobs[t] = j;
        if (nthreads > 1) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads == 1) worker(&jobs[0]);
    else for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}
