// fpt_text.cu — host-side text output of the scoring results: the bedGraph rows of
// write_stats_to_output and the BED rows of write_segments_to_output, formatted natively. No device code:
// once scoring runs at 1e10 bases/s, a Python f-string per value is the bottleneck of `ftd detect`.
//
// Reference behaviour (paths relative to /root/reference):
//   write_stats_to_output      footprint_tools/cli/utils.py:119-164   rows "chrom\tstart+i\tstart+i+1\tv0\t...\n",
//                                                                    every value as format(v, "0.4f")
//   write_segments_to_output   footprint_tools/cli/utils.py:167-214   utils.segment(stats, thr, 3, decreasing) then
//                                                                    "chrom\tstart+s\tstart+e\tname\tmin(stats[s:e])\n"
//   utils.segment              footprint_tools/stats/utils.pyx:15-50
//
// format(v, ".Df") is the correctly rounded decimal expansion of the exact binary value (ties to even). For
// |v| * 10^D < 2^52 it is produced here from t = v * 10^D and the exact residual e = fma(v, 10^D, -t): the
// scaled value is t + e exactly, so the rounding decision (fraction of t against 1/2, then the sign of e, then
// parity) is exact; everything else (huge values) goes through snprintf, non-finite values are spelled as
// Python spells them ("nan", "inf", "-inf").
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fpt_b200.h"

namespace {

const double kPow10[10] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};

inline char *put_u64(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char *put_i64(char *p, int64_t v) {
    if (v < 0) {
        *p++ = '-';
        return put_u64(p, (uint64_t)(-(v + 1)) + 1u);
    }
    return put_u64(p, (uint64_t)v);
}

// format(v, ".{prec}f"); writes at most 340 bytes
char *put_fixed(char *p, double v, int prec) {
    if (v != v) { memcpy(p, "nan", 3); return p + 3; }
    if (isinf(v)) {
        if (v < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    const double a = fabs(v), scale = kPow10[prec];
    const double t = a * scale;
    if (!(t < 4503599627370496.0)) return p + snprintf(p, 336, "%.*f", prec, v);  // 2^52: not exact below
    const double e = fma(a, scale, -t);  // a * scale == t + e exactly
    const double q = floor(t), f = t - q;
    uint64_t r = (uint64_t)q;
    // f and 1/2 are both multiples of ulp(t) and |e| <= ulp(t)/2, so e only decides when f is exactly 1/2
    if (f > 0.5 || (f == 0.5 && (e > 0.0 || (e == 0.0 && (r & 1u))))) r += 1;
    if (signbit(v)) *p++ = '-';
    if (prec == 0) return put_u64(p, r);
    const uint64_t s = (uint64_t)scale;
    p = put_u64(p, r / s);
    *p++ = '.';
    uint64_t frac = r % s;
    for (int i = prec - 1; i >= 0; --i) {
        p[i] = (char)('0' + frac % 10);
        frac /= 10;
    }
    return p + prec;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int64_t fpt_format_stats(const char *const *chroms, const int64_t *starts, const int64_t *out_off, int64_t n_iv,
                         const double *const *cols, int ncols, int precision, char delim, char *buf, int64_t cap,
                         int64_t *n_done) {
    if (n_done) *n_done = 0;
    if (n_iv < 0 || ncols < 0 || precision < 0 || precision > 9 || cap < 0 || (n_iv > 0 && (!chroms || !starts || !out_off)) ||
        (ncols > 0 && !cols) || (cap > 0 && !buf))
        return FPT_ERR_ARG;
    char *p = buf, *const end = buf + cap;
    int64_t k = 0;
    for (; k < n_iv; ++k) {
        char *const mark = p;
        const size_t cl = strlen(chroms[k]);
        const size_t row_bound = cl + 48 + (size_t)ncols * 342;
        bool fits = true;
        for (int64_t i = out_off[k]; i < out_off[k + 1]; ++i) {
            if ((size_t)(end - p) < row_bound) { fits = false; break; }
            const int64_t pos = starts[k] + (i - out_off[k]);
            memcpy(p, chroms[k], cl); p += cl;
            *p++ = delim; p = put_i64(p, pos);
            *p++ = delim; p = put_i64(p, pos + 1);
            *p++ = delim;  // cli/utils.py:160: the row is chrom, start, end, delim and then the joined values
            for (int c = 0; c < ncols; ++c) {
                if (c) *p++ = delim;
                p = put_fixed(p, cols[c][i], precision);
            }
            *p++ = '\n';
        }
        if (!fits) { p = mark; break; }  // whole intervals only: the caller flushes and calls again from interval k
    }
    if (n_done) *n_done = k;
    return (int64_t)(p - buf);
}

int64_t fpt_segment(const double *x, int64_t n, double threshold, int w, int decreasing, int64_t *pairs, int64_t cap_pairs) {
    if (n < 0 || (n > 0 && !x) || cap_pairs < 0 || (cap_pairs > 0 && !pairs)) return FPT_ERR_ARG;
    const double dir = decreasing ? -1.0 : 1.0;
    int64_t m = 0, curr = -1, last_end = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (curr < 0) {
            if (dir * x[i] >= dir * threshold) curr = i - w + 1;  // negative near the array start: not opened (utils.pyx:41-42)
        } else if (dir * x[i] < dir * threshold) {
            if (m > 0 && curr <= last_end) {
                last_end = i - 1 + w;
                if (m <= cap_pairs) pairs[2 * (m - 1) + 1] = last_end;
            } else {
                last_end = i - 1 + w;
                if (m < cap_pairs) { pairs[2 * m] = curr; pairs[2 * m + 1] = last_end; }
                ++m;
            }
            curr = -1;
        }
    }
    return m;  // number of segments found (may exceed cap_pairs: call again with a larger array)
}

int64_t fpt_format_segments(const char *const *chroms, const int64_t *starts, const int64_t *out_off, int64_t n_iv,
                            const double *stats, double threshold, int w, int decreasing, const char *name, int precision,
                            char delim, char *buf, int64_t cap, int64_t *n_done) {
    if (n_done) *n_done = 0;
    if (n_iv < 0 || precision < 0 || precision > 9 || cap < 0 || (n_iv > 0 && (!chroms || !starts || !out_off || !stats)) ||
        !name || (cap > 0 && !buf))
        return FPT_ERR_ARG;
    const double dir = decreasing ? -1.0 : 1.0;
    const size_t nl = strlen(name);
    char *p = buf, *const end = buf + cap;
    int64_t k = 0;
    for (; k < n_iv; ++k) {
        char *const mark = p;
        const double *x = stats + out_off[k];
        const int64_t n = out_off[k + 1] - out_off[k];
        const size_t cl = strlen(chroms[k]);
        const size_t row_bound = cl + nl + 48 + 342;
        bool fits = true;
        // utils.segment (utils.pyx:38-50); a segment is written when the next one turns out not to extend it
        int64_t curr = -1, seg_s = 0, seg_e = 0;
        bool have = false;
        auto emit = [&]() {
            if ((size_t)(end - p) < row_bound) { fits = false; return; }
            // score_fn = np.min over stats[s:e] (the slice is clipped to the array; NaN if any element is NaN)
            const int64_t a = seg_s < 0 ? 0 : seg_s, b = seg_e > n ? n : seg_e;
            double mn = INFINITY;
            bool nan = false;
            for (int64_t i = a; i < b; ++i) {
                if (x[i] != x[i]) nan = true;
                else if (x[i] < mn) mn = x[i];
            }
            if (nan) mn = NAN;
            memcpy(p, chroms[k], cl); p += cl;
            *p++ = delim; p = put_i64(p, starts[k] + seg_s);
            *p++ = delim; p = put_i64(p, starts[k] + seg_e);
            *p++ = delim; memcpy(p, name, nl); p += nl;
            *p++ = delim; p = put_fixed(p, mn, precision);
            *p++ = '\n';
        };
        for (int64_t i = 0; i < n && fits; ++i) {
            if (curr < 0) {
                if (dir * x[i] >= dir * threshold) curr = i - w + 1;
            } else if (dir * x[i] < dir * threshold) {
                if (have && curr <= seg_e) {
                    seg_e = i - 1 + w;
                } else {
                    if (have) emit();
                    seg_s = curr; seg_e = i - 1 + w; have = true;
                }
                curr = -1;
            }
        }
        if (have && fits) emit();
        if (!fits) { p = mark; break; }
    }
    if (n_done) *n_done = k;
    return (int64_t)(p - buf);
}

int64_t fpt_format_records(const char *const *chroms, const int64_t *starts, int64_t n_iv, const int64_t *seg_iv,
                           const int64_t *seg_start, const int64_t *seg_end, const double *seg_score, int64_t n_seg,
                           const char *name, int precision, char delim, char *buf, int64_t cap, int64_t *n_done) {
    if (n_done) *n_done = 0;
    if (n_seg < 0 || n_iv < 0 || precision < 0 || precision > 9 || cap < 0 || !name || (cap > 0 && !buf) ||
        (n_seg > 0 && (!chroms || !starts || !seg_iv || !seg_start || !seg_end || !seg_score)))
        return FPT_ERR_ARG;
    const size_t nl = strlen(name);
    char *p = buf, *const end = buf + cap;
    int64_t q = 0;
    for (; q < n_seg; ++q) {
        const int64_t k = seg_iv[q];
        if (k < 0 || k >= n_iv) return FPT_ERR_ARG;
        const size_t cl = strlen(chroms[k]);
        if ((size_t)(end - p) < cl + nl + 48 + 342) break;  // the caller flushes and calls again from record q
        memcpy(p, chroms[k], cl); p += cl;
        *p++ = delim; p = put_i64(p, starts[k] + seg_start[q]);
        *p++ = delim; p = put_i64(p, starts[k] + seg_end[q]);
        *p++ = delim; memcpy(p, name, nl); p += nl;
        *p++ = delim; p = put_fixed(p, seg_score[q], precision);
        *p++ = '\n';
    }
    if (n_done) *n_done = q;
    return (int64_t)(p - buf);
}

}  // extern "C"
#pragma GCC visibility pop
