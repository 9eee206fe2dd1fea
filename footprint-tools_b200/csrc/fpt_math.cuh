// fpt_math.cuh — device FP64 special functions for the footprint scoring path (sm_100a).
//
// These follow the reference's vendored Cephes fork branch for branch (hcephes v0.4.1, paths
// relative to /root/reference) because parity is defined against it, including the regions
// where Cephes itself is inexact (SURVEY.md hard part 4):
//   incbet  hcephes/src/cprob/incbet.c:12-299     gamma/lgam hcephes/src/cprob/gamma.c:35-235
//   igam(c) hcephes/src/cprob/igam.c:6-100        chdtrc     hcephes/src/cprob/chdtr.c:3-10
//   ndtr    hcephes/src/cprob/ndtr.c:34-132       ndtri      hcephes/src/cprob/ndtri.c:48-88
//   expx2   hcephes/src/cprob/expx2.c:6-34        log1p      hcephes/src/cprob/unity.c:29-38
// The only tolerated deviations are CUDA libm (log/exp/pow/sin) vs glibc and FMA contraction
// inside these functions (both far below the 1e-9 tolerance on -log10 p). Everything that feeds
// an INTEGER result (expected counts) or the arguments (a,b,x) of incbet is computed with
// explicitly rounded __dmul_rn/__dadd_rn so that it is bit-identical to the x86-64 no-FMA build.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fpt {

#define FPT_DEV __device__ __forceinline__
#define FPT_DEV_NOINLINE static __device__ __noinline__

constexpr double kMachEp = 1.11022302462515654042E-16;
constexpr double kMaxLog = 7.09782712893383996732E2;
constexpr double kMinLog = -7.451332191019412076235E2;
constexpr double kMaxGam = 171.624376956302725;
constexpr double kPi = 3.14159265358979323846;
constexpr double kSqrtH = 7.07106781186547524401E-1;
constexpr double kBig = 4.503599627370496e15;
constexpr double kBigInv = 2.22044604925031308085e-16;
constexpr double kSqrt2Pi = 2.50662827463100050242E0;

// Horner with the coefficients as template-unrolled immediates (polyn/polevl.c:3-32).
template <int N>
FPT_DEV double poly(double x, const double (&c)[N]) {
    double a = c[0];
#pragma unroll
    for (int i = 1; i < N; ++i) a = fma(a, x, c[i]);
    return a;
}
template <int N>  // leading coefficient 1 implied (p1evl)
FPT_DEV double poly1(double x, const double (&c)[N]) {
    double a = x + c[0];
#pragma unroll
    for (int i = 1; i < N; ++i) a = fma(a, x, c[i]);
    return a;
}

// ---- gamma (gamma.c:35-127) ----------------------------------------------------------------
FPT_DEV double stirling_gamma(double x) {
    const double S[5] = {7.87311395793093628397E-4, -2.29549961613378126380E-4, -2.68132617805781232825E-3,
                         3.47222221605458667310E-3, 8.33333333333482257126E-2};
    double w = 1.0 / x;
    w = 1.0 + w * poly(w, S);
    double y = exp(x);
    if (x > 143.01608) {
        double v = pow(x, 0.5 * x - 0.25);
        y = v * (v / y);
    } else {
        y = pow(x, x - 0.5) / y;
    }
    return kSqrt2Pi * y * w;
}

FPT_DEV_NOINLINE double gamma_fn(double x) {
    const double P[7] = {1.60119522476751861407E-4, 1.19135147006586384913E-3, 1.04213797561761569935E-2,
                         4.76367800457137231464E-2, 2.07448227648435975150E-1, 4.94214826801497100753E-1,
                         9.99999999999999996796E-1};
    const double Q[8] = {-2.31581873324120129819E-5, 5.39605580493303397842E-4, -4.45641913851797240494E-3,
                         1.18139785222060435552E-2,  3.58236398605498653373E-2, -2.34591795718243348568E-1,
                         7.14304917030273074085E-2,  1.00000000000000000320E0};
    if (isnan(x)) return x;
    if (x == CUDART_INF) return x;
    if (x == -CUDART_INF) return CUDART_NAN;
    double q = fabs(x);
    if (q > 33.0) {
        double sgn = 1.0, z;
        if (x < 0.0) {
            double p = floor(q);
            if (p == q) return CUDART_NAN;
            if ((((int)p) & 1) == 0) sgn = -1.0;
            z = q - p;
            if (z > 0.5) {
                p += 1.0;
                z = q - p;
            }
            z = q * sin(kPi * z);
            if (z == 0.0) return sgn * CUDART_INF;
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
    bool tiny = false;
    while (x < 0.0) {
        if (x > -1.E-9) { tiny = true; break; }
        z /= x;
        x += 1.0;
    }
    if (!tiny) {
        while (x < 2.0) {
            if (x < 1.e-9) { tiny = true; break; }
            z /= x;
            x += 1.0;
        }
    }
    if (tiny) {
        if (x == 0.0) return CUDART_NAN;
        return z / ((1.0 + 0.5772156649015329 * x) * x);
    }
    if (x == 2.0) return z;
    x -= 2.0;
    return z * poly(x, P) / poly(x, Q);
}

// ---- lgam (gamma.c:147-235); the sign output of lgam_sgn is unused on this path -------------
FPT_DEV_NOINLINE double lgam_fn(double x) {
    const double A[5] = {8.11614167470508450300E-4, -5.95061904284301438324E-4, 7.93650340457716943945E-4,
                         -2.77777777730099687205E-3, 8.33333333333331927722E-2};
    const double B[6] = {-1.37825152569120859100E3, -3.88016315134637840924E4, -3.31612992738871184744E5,
                         -1.16237097492762307383E6, -1.72173700820839662146E6, -8.53555664245765465627E5};
    const double C[6] = {-3.51815701436523470549E2, -1.70642106651881159223E4, -2.20528590553854454839E5,
                         -1.13933444367982507207E6, -2.53252307177582951285E6, -2.01889141433532773231E6};
    if (isnan(x)) return x;
    if (isinf(x)) return CUDART_INF;
    double refl_q = 0.0;
    bool reflect = false;
    if (x < -34.0) {  // reflection: evaluate lgam(-x) below, then combine
        reflect = true;
        refl_q = -x;
        x = refl_q;
    }
    double res;
    if (x < 13.0) {
        double z = 1.0, p = 0.0, u = x;
        while (u >= 3.0) {
            p -= 1.0;
            u = x + p;
            z *= u;
        }
        bool sing = false;
        while (u < 2.0) {
            if (u == 0.0) { sing = true; break; }
            z /= u;
            p += 1.0;
            u = x + p;
        }
        if (sing) return CUDART_INF;
        if (z < 0.0) z = -z;
        if (u == 2.0) {
            res = log(z);
        } else {
            p -= 2.0;
            double xx = x + p;
            p = xx * poly(xx, B) / poly1(xx, C);
            res = log(z) + p;
        }
    } else if (x > 2.556348e305) {
        res = CUDART_INF;
    } else {
        double q = (x - 0.5) * log(x) - x + 0.91893853320467274178;
        if (x > 1.0e8) {
            res = q;
        } else {
            double p = 1.0 / (x * x);
            if (x >= 1000.0)
                q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
            else
                q += poly(p, A) / x;
            res = q;
        }
    }
    if (!reflect) return res;
    double q = refl_q, w = res;
    double p = floor(q);
    if (p == q) return CUDART_INF;
    double z = q - p;
    if (z > 0.5) {
        p += 1.0;
        z = p - q;
    }
    z = q * sin(kPi * z);
    if (z == 0.0) return CUDART_INF;
    return 1.14472988584940017414 - log(z) - w;
}

// ---- log1p (unity.c:29-38) -----------------------------------------------------------------
FPT_DEV double log1p_fn(double x) {
    const double LP[7] = {4.5270000862445199635215E-5, 4.9854102823193375972212E-1, 6.5787325942061044846969E0,
                          2.9911919328553073277375E1,  6.0949667980987787057556E1,  5.7112963590585538103336E1,
                          2.0039553499201281259648E1};
    const double LQ[6] = {1.5062909083469192043167E1, 8.3047565967967209469434E1, 2.2176239823732856465394E2,
                          3.0909872225312059774938E2, 2.1642788614495947685003E2, 6.0118660497603843919306E1};
    double z = 1.0 + x;
    if (z < 0.70710678118654752440 || z > 1.41421356237309504880) return log(z);
    z = x * x;
    z = -0.5 * z + x * (z * poly(x, LP) / poly1(x, LQ));
    return x + z;
}

// ---- incomplete beta (incbet.c) ------------------------------------------------------------
// gamma(a+b) / (gamma(a) gamma(b)), as written at incbet.c:74 and :287
FPT_DEV double gamma_ratio(double a, double b) { return gamma_fn(a + b) / (gamma_fn(a) * gamma_fn(b)); }

// incbet.c:266-299
FPT_DEV_NOINLINE double beta_power_series(double a, double b, double x) {
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
        t = gamma_ratio(a, b);
        s = s * t * pow(x, a);
    } else {
        t = lgam_fn(a + b) - lgam_fn(a) - lgam_fn(b) + u + log(s);
        s = (t < kMinLog) ? 0.0 : exp(t);
    }
    return s;
}

// incbet.c:100-177 (variant 0) and :183-261 (variant 1): same recurrence, different schedule of
// the k-coefficients and argument (x vs x/(1-x)).
FPT_DEV_NOINLINE double beta_cfrac(double a, double b, double x, int variant) {
    double k1 = a, k3 = a, k4 = a + 1.0, k5 = 1.0, k7 = a + 1.0, k8 = a + 2.0;
    double k2, k6, d2, zz;
    if (variant == 0) {
        k2 = a + b; k6 = b - 1.0; d2 = 1.0; zz = x;
    } else {
        k2 = b - 1.0; k6 = a + b; d2 = -1.0; zz = x / (1.0 - x);
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
        k5 += 1.0; k6 -= d2; k7 += 2.0; k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > kBig) {
            pkm2 *= kBigInv; pkm1 *= kBigInv; qkm2 *= kBigInv; qkm1 *= kBigInv;
        }
        if ((fabs(qk) < kBigInv) || (fabs(pk) < kBigInv)) {
            pkm2 *= kBig; pkm1 *= kBig; qkm2 *= kBig; qkm1 *= kBig;
        }
    } while (++n < 300);
    return ans;
}

// incbet.c:12-94
FPT_DEV_NOINLINE double incbet_fn(double aa, double bb, double xx) {
    if (aa <= 0.0 || bb <= 0.0) return 0.0;
    if (xx <= 0.0 || xx >= 1.0) {
        if (xx == 1.0) return 1.0;
        return 0.0;  // xx == 0, out of domain, or NaN-excluded cases all give 0
    }
    double a, b, t, x, xc, w, y;
    bool flag = false;
    if ((bb * xx) <= 1.0 && xx <= 0.95) {
        return beta_power_series(aa, bb, xx);
    }
    w = 1.0 - xx;
    if (xx > (aa / (aa + bb))) {
        flag = true; a = bb; b = aa; xc = xx; x = w;
    } else {
        a = aa; b = bb; xc = w; x = xx;
    }
    if (flag && (b * x) <= 1.0 && x <= 0.95) {
        t = beta_power_series(a, b, x);
    } else {
        y = x * (a + b - 2.0) - (a - 1.0);
        if (y < 0.0)
            w = beta_cfrac(a, b, x, 0);
        else
            w = beta_cfrac(a, b, x, 1) / xc;
        y = a * log(x);
        t = b * log(xc);
        if ((a + b) < kMaxGam && fabs(y) < kMaxLog && fabs(t) < kMaxLog) {
            t = pow(xc, b);
            t *= pow(x, a);
            t /= a;
            t *= w;
            t *= gamma_ratio(a, b);
        } else {
            y += t + lgam_fn(a + b) - lgam_fn(a) - lgam_fn(b);
            y += log(w / a);
            t = (y < kMinLog) ? 0.0 : exp(y);
        }
    }
    if (flag) {
        if (t <= kMachEp)
            t = 1.0 - kMachEp;
        else
            t = 1.0 - t;
    }
    return t;
}

// ---- incomplete gamma, chi-square tail (igam.c, chdtr.c) -----------------------------------
FPT_DEV double igam_series(double a, double x) {  // igam.c:80-99 (the x<=1 or x<=a branch)
    double ax = a * log(x) - x - lgam_fn(a);
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
FPT_DEV double igamc_cfrac(double a, double x) {  // igam.c:16-58 (the x>=1 and x>=a branch)
    double ax = a * log(x) - x - lgam_fn(a);
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
FPT_DEV_NOINLINE double igamc_fn(double a, double x) {  // igam.c:6-14 dispatch
    if (x <= 0 || a <= 0) return 1.0;
    if (x < 1.0 || x < a) return 1.0 - igam_series(a, x);  // igam() takes its series branch here
    return igamc_cfrac(a, x);
}
FPT_DEV double chdtrc_fn(double df, double x) {  // chdtr.c:3-10
    if (x < 0.0 || df < 1.0) return 0.0;
    return igamc_fn(df / 2.0, x / 2.0);
}

// ---- normal distribution (ndtr.c, expx2.c, ndtri.c) ----------------------------------------
FPT_DEV double exp_neg_x2(double x) {  // expx2.c:6-34 with sign = -1
    x = -fabs(x);
    double m = .0078125 * floor(128.0 * x + 0.5);
    double f = x - m;
    double u = -(m * m);
    double u1 = -(2 * m * f + f * f);
    if ((u + u1) > kMaxLog) return CUDART_INF;
    return exp(u) * exp(u1);
}

FPT_DEV double erfc_scaled(double x) {  // ndtr.c:65-76
    const double P[9] = {2.46196981473530512524E-10, 5.64189564831068821977E-1, 7.46321056442269912687E0,
                         4.86371970985681366614E1,   1.96520832956077098242E2,  5.26445194995477358631E2,
                         9.34528527171957607540E2,   1.02755188689515710272E3,  5.57535335369399327526E2};
    const double Q[8] = {1.32281951154744992508E1, 8.67072140885989742329E1, 3.54937778887819891062E2,
                         9.75708501743205489753E2, 1.82390916687909736289E3, 2.24633760818710981792E3,
                         1.65666309194161350182E3, 5.57535340817727675546E2};
    const double R[6] = {5.64189583547755073984E-1, 1.27536670759978104416E0, 5.01905042251180477414E0,
                         6.16021097993053585195E0,  7.40974269950448939160E0, 2.97886665372100240670E0};
    const double S[6] = {2.26052863220117276590E0, 9.39603524938001434673E0, 1.20489539808096656605E1,
                         1.70814450747565897222E1, 9.60896809063285878198E0, 3.36907645100081516050E0};
    if (x < 8.0) return poly(x, P) / poly1(x, Q);
    return poly(x, R) / poly1(x, S);
}

// ndtr.c:34-59. Inside ndtr, erf() is only reached with |x| < 1 (ndtr.c:78-86 polynomial branch).
FPT_DEV double ndtr_fn(double a) {
    const double T[5] = {9.60497373987051638749E0, 9.00260197203842689217E1, 2.23200534594684319226E3,
                         7.00332514112805075473E3, 5.55923013010394962768E4};
    const double U[5] = {3.35617141647503099647E1, 5.21357949780152679795E2, 4.59432382970980127987E3,
                         2.26290000613890934246E4, 4.92673942608635921086E4};
    double x = a * kSqrtH;
    double z = fabs(x);
    double y;
    if (z < 1.0) {
        double zz = x * x;
        double e = x * poly(zz, T) / poly1(zz, U);
        y = 0.5 + 0.5 * e;
    } else {
        y = 0.5 * erfc_scaled(z);
        z = exp_neg_x2(a);
        y = y * sqrt(z);
        if (x > 0) y = 1.0 - y;
    }
    return y;
}

// ndtri.c:48-88
FPT_DEV_NOINLINE double ndtri_fn(double y0) {
    const double P0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                          1.39312609387279679503E1, -1.23916583867381258016E0};
    const double Q0[8] = {1.95448858338141759834E0,  4.67627912898881538453E0, 8.63602421390890590575E1,
                          -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                          1.59056225126211695515E1,  -1.18331621121330003142E0};
    const double P1[9] = {4.05544892305962419923E0,   3.15251094599893866154E1,   5.71628192246421288162E1,
                          4.40805073893200834700E1,   1.46849561928858024014E1,   2.18663306850790267539E0,
                          -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
    const double Q1[8] = {1.57799883256466749731E1,   4.53907635128879210584E1,   4.13172038254672030440E1,
                          1.50425385692907503408E1,   2.50464946208309415979E0,   -1.42182922854787788574E-1,
                          -3.80806407691578277194E-2, -9.33259480895457427372E-4};
    const double P2[9] = {3.23774891776946035970E0,  6.91522889068984211695E0,  3.93881025292474443415E0,
                          1.33303460815807542389E0,  2.01485389549179081538E-1, 1.23716634817820021358E-2,
                          3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
    const double Q2[8] = {6.02427039364742014255E0,  3.67983563856160859403E0,  1.37702099489081330271E0,
                          2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                          2.89247864745380683936E-6, 6.79019408009981274425E-9};
    const double em2 = 0.13533528323661269189;
    if (y0 <= 0.0) return -CUDART_INF;
    if (y0 >= 1.0) return CUDART_INF;
    bool negate = true;
    double y = y0;
    if (y > 1.0 - em2) {
        y = 1.0 - y;
        negate = false;
    }
    if (y > em2) {
        y = y - 0.5;
        double y2 = y * y;
        double x = y + y * (y2 * poly(y2, P0) / poly1(y2, Q0));
        return x * kSqrt2Pi;
    }
    double x = sqrt(-2.0 * log(y));
    double x0 = x - log(x) / x;
    double z = 1.0 / x, x1;
    if (x < 8.0)
        x1 = z * poly(z, P1) / poly1(z, Q1);
    else
        x1 = z * poly(z, P2) / poly1(z, Q2);
    x = x0 - x1;
    return negate ? -x : x;
}

// ---- dispersion model (footprint_tools/modeling/dispersion.pyx:26-57, 127-163) ---------------
// Parameter layout as in the reference: [breaks (nseg), intercepts (nseg), slopes (nseg)]; the last
// break is unused. Exactly one segment is selected; its value is intercept + slope*x with the
// multiply and the add rounded separately (the reference evaluates it with Python float ops, no
// FMA), which makes r and mu — hence the arguments of incbet — bit-identical to the reference.
template <int NSEG>
FPT_DEV double piecewise_linear(const double *par, double x) {
    int s = NSEG - 1;
#pragma unroll
    for (int j = NSEG - 2; j >= 0; --j)
        if (x < par[j]) s = j;
    if (!(x == x)) return 0.0;  // NaN: every (bool) factor is 0 in the reference
    return __dadd_rn(par[NSEG + s], __dmul_rn(par[2 * NSEG + s], x));
}
FPT_DEV double fit_mu(const double *mu_params, double x) {
    double v = piecewise_linear<3>(mu_params, x);
    return v > 0.0 ? v : 0.1;
}
FPT_DEV double fit_r(const double *r_params, double x) {
    double v = __ddiv_rn(1.0, piecewise_linear<5>(r_params, x));
    return v > 0.0 ? v : 1e-6;
}

// ---- negative binomial (footprint_tools/stats/distributions/nbinom.pyx:82-138) ---------------
FPT_DEV double nb_cdf(int k, double p, double r) { return incbet_fn(r, (double)(k + 1), p); }
FPT_DEV double nb_logpmf(int k, double p, double r) {
    double coeff = lgam_fn((double)k + r) - lgam_fn((double)(k + 1)) - lgam_fn(r);
    return coeff + r * log(p) + (double)k * log1p_fn(-p);
}

// (r, mu) -> success probability exactly as dispersion.pyx:314 writes it: r/(r+mu)
FPT_DEV double nb_prob(double r, double mu) { return __ddiv_rn(r, __dadd_rn(r, mu)); }

}  // namespace fpt
