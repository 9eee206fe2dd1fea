// fpt_segment.cu — footprint segmentation of a whole batch on the device.
//
// `ftd detect` turns the empirical-FDR column of every interval into footprints with utils.segment and scores each
// footprint with np.min (paths relative to /root/reference):
//   utils.segment               footprint_tools/stats/utils.pyx:15-50      single left-to-right pass, see below
//   write_segments_to_output    footprint_tools/cli/utils.py:165-214       segment(stats, threshold, 3, decreasing), np.min(stats[s:e])
//   call site                   footprint_tools/cli/detect.py:403-408      stats[:, -1] (the FDR column), one call per threshold
// Once the FDR column is device-resident (fpt_detect_fdr) only the footprints need to cross PCIe: 32 bytes per
// footprint instead of 8 bytes per base and threshold.
//
// The reference's loop, per interval: an element passing the threshold (dir*x >= dir*thr) opens a run at i-w+1 —
// unless that is negative, in which case the run is not opened and the next element is tested again; the first
// failing element (dir*x < dir*thr) after that closes it at i-1+w; a run that starts at or before the end of the
// previous segment extends that segment instead of starting a new one; NaN neither opens nor closes; a run still
// open at the end of the array is dropped. Here one warp walks one interval 32 elements at a time: two ballots give
// the pass / fail bit masks and every lane runs the same (uniform) state machine over them with ffs, so the
// sequential semantics are kept exactly while the loads stay coalesced. A first pass counts the segments of every
// interval, an exclusive scan turns counts into output slots, a second pass writes the records — ordered by
// (interval, start) like the reference's output, independent of scheduling.
#include "fpt_internal.h"

namespace fpt {

namespace {

constexpr int kSegWarps = 8;

template <bool WRITE>
__global__ void __launch_bounds__(32 * kSegWarps)
segment_kernel(const double *__restrict__ x, const long long *__restrict__ out_off, long long n_iv, double dthr, double dir,
               int w, const long long *__restrict__ first, long long *__restrict__ counts, long long cap,
               long long *__restrict__ seg_iv, long long *__restrict__ seg_start, long long *__restrict__ seg_end,
               double *__restrict__ seg_score) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kSegWarps;
    for (long long k = blockIdx.x * (long long)kSegWarps + (threadIdx.x >> 5); k < n_iv; k += nwarps) {
        const long long a = out_off[k], n = out_off[k + 1] - a;
        const double *xi = x + a;
        long long curr = -1, ss = 0, se = 0, m = 0;
        bool have = false;
        const long long slot0 = WRITE ? first[k] : 0;
        auto emit = [&]() {
            if (WRITE && slot0 + m < cap) {
                // np.min(stats[s:e]): the slice is clipped to the array, NaN if any element is NaN
                const long long b = se > n ? n : se;
                double mn = INFINITY;
                int nan = 0;
                for (long long i = ss + lane; i < b; i += 32) {
                    const double v = xi[i];
                    if (v != v) nan = 1;
                    else if (v < mn) mn = v;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double t = __shfl_xor_sync(0xFFFFFFFFu, mn, o);
                    mn = t < mn ? t : mn;
                }
                nan = __any_sync(0xFFFFFFFFu, nan);
                if (lane == 0) {
                    const long long q = slot0 + m;
                    seg_iv[q] = k;
                    seg_start[q] = ss;
                    seg_end[q] = se;
                    seg_score[q] = nan ? __longlong_as_double(0x7FF8000000000000LL) : mn;
                }
            }
            ++m;
        };
        for (long long base = 0; base < n; base += 32) {
            const long long i = base + lane;
            const double dv = i < n ? dir * xi[i] : __longlong_as_double(0x7FF8000000000000LL);
            unsigned P = __ballot_sync(0xFFFFFFFFu, dv >= dthr);
            const unsigned F = __ballot_sync(0xFFFFFFFFu, dv < dthr);
            // elements i < w-1 cannot open a run (i-w+1 < 0 reads as "still closed", utils.pyx:38-42)
            if (base < (long long)w - 1) {
                const long long nb = (long long)w - 1 - base;
                P = nb >= 32 ? 0u : (P & ~((1u << (int)nb) - 1u));
            }
            int pos = 0;
            while (pos < 32) {
                if (curr < 0) {
                    const unsigned mm = P >> pos;
                    if (!mm) break;
                    const int j = pos + __ffs(mm) - 1;
                    curr = base + j - w + 1;
                    pos = j + 1;
                } else {
                    const unsigned mm = F >> pos;
                    if (!mm) break;
                    const int j = pos + __ffs(mm) - 1;
                    const long long e = base + j - 1 + w;
                    if (have && curr <= se) {
                        se = e;
                    } else {
                        if (have) emit();
                        ss = curr;
                        se = e;
                        have = true;
                    }
                    curr = -1;
                    pos = j + 1;
                }
            }
        }
        if (have) emit();
        if (!WRITE && lane == 0) counts[k] = m;
    }
}

// first[i] = counts[0] + ... + counts[i-1], first[n] = total; one CTA, 4 consecutive elements per thread and round
// (250 k intervals: 62 rounds of three barriers each)
__global__ void __launch_bounds__(1024) segment_scan_kernel(const long long *__restrict__ counts, long long n,
                                                            long long *__restrict__ first) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 4096) {
        const long long i0 = base + 4 * (long long)tid;
        long long v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = i0 + e < n ? counts[i0 + e] : 0;
        const long long mine = (v[0] + v[1]) + (v[2] + v[3]);
        long long inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        long long run = carry + (warp ? wsum[warp - 1] : 0) + inc - mine;  // everything before this thread's elements
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (i0 + e < n) first[i0 + e] = run;
            run += v[e];
        }
        __syncthreads();
        if (tid == 0) carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) first[n] = carry;
}

}  // namespace

cudaError_t launch_segment_count(cudaStream_t st, const double *x, const long long *out_off, long long n_iv,
                                 double threshold, int w, int decreasing, long long *counts, long long *first,
                                 int sm_count) {
    if (n_iv <= 0) return cudaSuccess;
    const double dir = decreasing ? -1.0 : 1.0;
    long long blocks = (n_iv + kSegWarps - 1) / kSegWarps;
    const long long cap_blocks = (long long)sm_count * 8;
    if (blocks > cap_blocks) blocks = cap_blocks;
    segment_kernel<false><<<(unsigned)blocks, 32 * kSegWarps, 0, st>>>(x, out_off, n_iv, dir * threshold, dir, w, nullptr,
                                                                      counts, 0, nullptr, nullptr, nullptr, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    segment_scan_kernel<<<1, 1024, 0, st>>>(counts, n_iv, first);
    return cudaGetLastError();
}

cudaError_t launch_segment_write(cudaStream_t st, const double *x, const long long *out_off, long long n_iv,
                                 double threshold, int w, int decreasing, const long long *first, long long cap,
                                 long long *seg_iv, long long *seg_start, long long *seg_end, double *seg_score,
                                 int sm_count) {
    if (n_iv <= 0 || cap <= 0) return cudaSuccess;
    const double dir = decreasing ? -1.0 : 1.0;
    long long blocks = (n_iv + kSegWarps - 1) / kSegWarps;
    const long long cap_blocks = (long long)sm_count * 8;
    if (blocks > cap_blocks) blocks = cap_blocks;
    segment_kernel<true><<<(unsigned)blocks, 32 * kSegWarps, 0, st>>>(x, out_off, n_iv, dir * threshold, dir, w, first,
                                                                     nullptr, cap, seg_iv, seg_start, seg_end, seg_score);
    return cudaGetLastError();
}

}  // namespace fpt
