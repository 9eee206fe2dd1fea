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
#include <string.h>

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

}  // extern "C"
#pragma GCC visibility pop
