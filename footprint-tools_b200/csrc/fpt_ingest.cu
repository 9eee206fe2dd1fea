// fpt_ingest.cu — host-side ingest for the scoring path: turns decoded alignment columns into
// the per-strand uint32 cut-count tracks the kernels read, and unpacks the 2-bit + N sequence track back into
// characters. No device code and no scoring arithmetic: format conversion only, like fpt_pack_sequence.
//
// Reference behaviour (paths relative to /root/reference):
//   bamfile.validate_read        footprint_tools/cutcounts.py:118-146   QC-fail / duplicate / MAPQ filters
//   bamfile.read_pair_generator  footprint_tools/cutcounts.py:176-229   unpaired reads pass as they are; paired reads
//                                                                      must be proper pairs, primary, not supplementary
//   bamfile._add_read            footprint_tools/cutcounts.py:231-250   forward: cut at reference_start + offset[0];
//                                                                      reverse: cut at reference_end + offset[1]
//   bamfile.lookup               footprint_tools/cutcounts.py:276-313   per-interval arrays of those counts
//
// The reference pairs mates only to hand both to _add_read: every read that passes the filters is counted exactly
// once, whether or not its mate lies in the fetched window (the unpaired leftovers are flushed at :225-229). A
// whole-chromosome pass over the alignments therefore yields, at every position, the count `lookup` returns for any
// interval containing it (its fetch window is the interval widened by 10 bp, and |offset| <= 10 keeps a cut inside
// the interval within reach of its read). BAM/CRAM decoding itself stays with htslib on the host.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fpt_b200.h"
#include "fpt_internal.h"

namespace {
enum : unsigned {
    kPaired = 0x1, kProperPair = 0x2, kUnmapped = 0x4, kReverse = 0x10, kSecondary = 0x100, kQcFail = 0x200,
    kDuplicate = 0x400, kSupplementary = 0x800
};
}

#pragma GCC visibility push(default)
extern "C" {

int64_t fpt_cuts_from_alignments(const int64_t *ref_start, const int64_t *ref_end, const uint16_t *flag,
                                            const uint8_t *mapq, int64_t n, int min_qual, int remove_dups,
                                            int remove_qcfail, int offset_plus, int offset_minus, int64_t track_first,
                                            int64_t track_len, uint32_t *cuts_plus, uint32_t *cuts_minus) {
    if (n < 0 || track_len < 0 || (n > 0 && (!ref_start || !ref_end || !flag || !mapq)) ||
        (track_len > 0 && (!cuts_plus || !cuts_minus)))
        return fpt::set_error(FPT_ERR_ARG, "fpt_cuts_from_alignments: bad arguments");
    int64_t counted = 0;
    for (int64_t i = 0; i < n; ++i) {
        const unsigned f = flag[i];
        if (f & kUnmapped) continue;  // samfile.fetch yields placed reads only
        if (remove_qcfail && (f & kQcFail)) continue;
        if (remove_dups && (f & kDuplicate)) continue;
        if ((int)mapq[i] < min_qual) continue;
        if ((f & kPaired) && (!(f & kProperPair) || (f & (kSecondary | kSupplementary)))) continue;
        const bool rev = (f & kReverse) != 0;
        const int64_t a = (rev ? ref_end[i] + offset_minus : ref_start[i] + offset_plus) - track_first;
        if (a < 0 || a >= track_len) continue;
        uint32_t *t = rev ? cuts_minus : cuts_plus;
        if (t[a] == 0xFFFFFFFFu) return fpt::set_error(FPT_ERR_ARG, "fpt_cuts_from_alignments: cut count overflows uint32");
        ++t[a];
        ++counted;
    }
    return counted;
}

int fpt_unpack_sequence(const uint32_t *seq2, const uint32_t *nmask, int64_t first, int64_t n, char *out) {
    if (first < 0 || n < 0 || (n > 0 && (!seq2 || !nmask || !out)))
        return fpt::set_error(FPT_ERR_ARG, "fpt_unpack_sequence: bad arguments");
    static const char kBase[4] = {'A', 'C', 'G', 'T'};
    for (int64_t j = 0; j < n; ++j) {
        const int64_t i = first + j;
        if ((nmask[i >> 5] >> (i & 31)) & 1u)
            out[j] = 'N';
        else
            out[j] = kBase[(seq2[i >> 4] >> (2 * (i & 15))) & 3u];
    }
    return FPT_OK;
}

// posterior_stats._load_data (cli/post.py:59-87): rows of one sample's `ftd detect` bedGraph
//   chrom  start  end  exp  obs  -log p  -log win p  fdr
// are placed at column seg_off[k] + (start - iv_start[k]) of every interval k that contains `start` (the reference
// fetches each interval's rows through tabix: a row [s, s+1) overlaps [start, end) exactly when start <= s < end):
// exp = field 3, obs = field 4, fdr = field 7, w = 1. Positions without a row keep what the caller put there (the
// reference's zeros / ones / zeros). Comment lines ('#') and lines with fewer than 8 fields are skipped.
int64_t fpt_parse_stats_rows(const char *text, int64_t n_bytes, char delim, const char *const *iv_chroms,
                             const int64_t *iv_starts, const int64_t *iv_ends, const int64_t *seg_off, int64_t n_iv,
                             double *exp_row, double *obs_row, double *fdr_row, double *w_row) {
    if (n_bytes < 0 || n_iv < 0 || (n_bytes > 0 && !text) ||
        (n_iv > 0 && (!iv_chroms || !iv_starts || !iv_ends || !seg_off || !exp_row || !obs_row || !fdr_row || !w_row)))
        return fpt::set_error(FPT_ERR_ARG, "fpt_parse_stats_rows: bad arguments");
    // per chromosome: interval indices by start, with the running maximum of the ends (intervals may overlap)
    struct Group { std::vector<int64_t> idx; std::vector<int64_t> max_end; };
    std::unordered_map<std::string, Group> groups;
    for (int64_t k = 0; k < n_iv; ++k) groups[iv_chroms[k]].idx.push_back(k);
    for (auto &kv : groups) {
        Group &g = kv.second;
        std::stable_sort(g.idx.begin(), g.idx.end(), [&](int64_t a, int64_t b) { return iv_starts[a] < iv_starts[b]; });
        g.max_end.resize(g.idx.size());
        int64_t run = INT64_MIN;
        for (size_t i = 0; i < g.idx.size(); ++i) {
            run = std::max(run, iv_ends[g.idx[i]]);
            g.max_end[i] = run;
        }
    }
    int64_t placed = 0;
    const char *p = text, *const end = text + n_bytes;
    std::string chrom, last_chrom;
    const Group *grp = nullptr;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        const char *next = eol ? eol + 1 : end;
        if (le > p && le[-1] == '\r') --le;
        if (le > p && *p != '#') {
            const char *f[9];
            int nf = 0;
            f[nf++] = p;
            for (const char *q = p; q < le && nf < 9; ++q)
                if (*q == delim) f[nf++] = q + 1;
            if (nf >= 8) {
                chrom.assign(f[0], (size_t)(f[1] - 1 - f[0]));
                if (!grp || chrom != last_chrom) {
                    auto it = groups.find(chrom);
                    grp = it == groups.end() ? nullptr : &it->second;
                    last_chrom = chrom;
                    if (!grp) last_chrom.clear();
                }
                if (grp) {
                    // fields end at a delimiter or the line end: copy the numeric ones out so strtod / strtoll stop there
                    char num[4][64];
                    const int which[4] = {1, 3, 4, 7};
                    bool ok = true;
                    for (int t = 0; t < 4 && ok; ++t) {
                        const char *a = f[which[t]];
                        const char *b = which[t] + 1 < nf ? f[which[t] + 1] - 1 : le;
                        const size_t len = (size_t)(b - a);
                        if (len == 0 || len >= sizeof num[t]) { ok = false; break; }
                        memcpy(num[t], a, len);
                        num[t][len] = 0;
                    }
                    if (ok) {
                        char *stop = nullptr;
                        const long long pos = strtoll(num[0], &stop, 10);
                        if (stop != num[0] && *stop == 0) {
                            const double ve = strtod(num[1], nullptr), vo = strtod(num[2], nullptr), vf = strtod(num[3], nullptr);
                            // last interval (by start) that starts at or before pos, then back while one can still reach pos
                            const auto &idx = grp->idx;
                            size_t lo = 0, hi = idx.size();
                            while (lo < hi) {
                                const size_t m = (lo + hi) >> 1;
                                if (iv_starts[idx[m]] <= pos) lo = m + 1; else hi = m;
                            }
                            for (size_t i = lo; i-- > 0;) {
                                if (grp->max_end[i] <= pos) break;
                                const int64_t k = idx[i];
                                if (pos < iv_ends[k]) {
                                    const int64_t c = seg_off[k] + (pos - iv_starts[k]);
                                    exp_row[c] = ve; obs_row[c] = vo; fdr_row[c] = vf; w_row[c] = 1.0;
                                    ++placed;
                                }
                            }
                        }
                    }
                }
            }
        }
        p = next;
    }
    return placed;
}

}  // extern "C"
#pragma GCC visibility pop
