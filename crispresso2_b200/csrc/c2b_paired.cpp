// c2b_paired.cpp -- host-side native of the paired-end merge mode.
//
// c2b_consensus_from_pairs replaces get_consensus_alignment_from_pairs (CRISPRessoCORE.py:829-985; get_greater_qual_nuc :801-826):
// the two mates' alignments to one amplicon are merged column by column.  Walking both alignments with one cursor each:
//   * a column that is an insertion (amplicon gap) in both mates takes the base of the better quality; in one mate only, that
//     mate's base -- the other cursor waits;
//   * otherwise the amplicon position is shared: a read gap in both mates stays a gap while either mate is between its first and
//     last base (a deletion) and becomes 'N' outside (no mate covers it); one base -> that base; two bases -> the better quality
//     (equal qualities: the mate with the better alignment score), and a disagreement makes the pair unfit for caching;
//   * past the end of one alignment the other is copied; a read gap inside its covered range becomes 'N'.
// Quality characters are consumed exactly as the reference consumes them (including the one it takes for a leading / trailing gap
// of a lone mate), so a quality string that is too short fails here where the reference raises IndexError.
// Then leading / trailing amplicon-gap columns are dropped and the homology is matches / columns.
#include <cstdint>
#include <cstring>
#include <string>

#include "c2b200.h"

namespace {

struct Pick { char nuc; bool decided; char qual; };

inline Pick better(char n1, char q1, char n2, char q2, bool r1_best)
{
    const unsigned char a = (unsigned char)q1, b = (unsigned char)q2;
    if (n1 == n2) return {n1, false, a >= b ? q1 : q2};
    if (a == b) return {r1_best ? n1 : n2, true, q2};
    return a > b ? Pick{n1, true, q1} : Pick{n2, true, q2};
}

}  // namespace

extern "C" int c2b_consensus_from_pairs(const char *s1, int32_t ns1, const char *f1, int32_t n1, double score1, const char *q1, int32_t nq1,
                                        const char *s2, int32_t ns2, const char *f2, int32_t n2, double score2, const char *q2, int32_t nq2,
                                        char *out_aln, char *out_qual, char *out_ref, int32_t cap,
                                        int32_t *n_cols, int32_t *n_qual, int32_t *n_match, int32_t *caching_is_ok)
{
    if (!s1 || !f1 || !q1 || !s2 || !f2 || !q2 || !out_aln || !out_qual || !out_ref || !n_cols || !n_qual || !n_match || !caching_is_ok ||
        n1 < 0 || n2 < 0 || ns1 < 0 || ns2 < 0 || nq1 < 0 || nq2 < 0)
        return C2B_E_ARG;
    auto span = [](const char *s, int n, int &start, int &stop) {
        start = 0;
        while (start < n && s[start] == '-') start++;
        stop = n - 1;
        while (stop >= 0 && s[stop] == '-') stop--;
    };
    int start1, stop1, start2, stop2;
    span(s1, ns1, start1, stop1);                            // over the aligned read's own length: the reference walks the amplicon
    span(s2, ns2, start2, stop2);                            // strings' columns and indexes the read strings with the same cursor
    const bool r1_best = score1 >= score2;
    std::string aln, ref, qual;
    aln.reserve((size_t)n1 + n2); ref.reserve((size_t)n1 + n2); qual.reserve((size_t)n1 + n2);
    bool cache_ok = true, short_qual = false;
    int i1 = 0, i2 = 0, k1 = 0, k2 = 0;
    auto Q1 = [&]() -> char { if (k1 >= nq1) { short_qual = true; return '!'; } return q1[k1]; };
    auto Q2 = [&]() -> char { if (k2 >= nq2) { short_qual = true; return '!'; } return q2[k2]; };
    while ((i1 < n1 || i2 < n2) && !short_qual) {
        const bool in1 = i1 < n1, in2 = i2 < n2;
        if ((in1 && i1 >= ns1) || (in2 && i2 >= ns2)) { short_qual = true; break; }     // aligned read shorter than its amplicon string
        const bool ins1 = in1 && f1[i1] == '-', ins2 = in2 && f2[i2] == '-';
        if (ins1 && ins2) {
            const char a = Q1(), b = Q2();
            if (short_qual) break;
            const Pick p = better(s1[i1], a, s2[i2], b, r1_best);
            if (p.decided) cache_ok = false;
            aln.push_back(p.nuc); ref.push_back('-'); qual.push_back(p.qual);
            k1++; k2++; i1++; i2++;
            continue;
        }
        if (ins1) { const char a = Q1(); if (short_qual) break; aln.push_back(s1[i1]); ref.push_back('-'); qual.push_back(a); k1++; i1++; continue; }
        if (ins2) { const char b = Q2(); if (short_qual) break; aln.push_back(s2[i2]); ref.push_back('-'); qual.push_back(b); k2++; i2++; continue; }
        const bool gap1 = in1 && s1[i1] == '-', gap2 = in2 && s2[i2] == '-';
        if (in1 && in2) {
            if (gap1 && gap2) {
                const bool covered = (start1 <= i1 && i1 <= stop1) || (start2 <= i2 && i2 <= stop2);
                aln.push_back(covered ? '-' : 'N'); ref.push_back(f1[i1]);
            } else if (gap1) {
                const char b = Q2(); if (short_qual) break;
                aln.push_back(s2[i2]); ref.push_back(f2[i2]); qual.push_back(b); k2++;
            } else if (gap2) {
                const char a = Q1(); if (short_qual) break;
                aln.push_back(s1[i1]); ref.push_back(f1[i1]); qual.push_back(a); k1++;
            } else {
                const char a = Q1(), b = Q2();
                if (short_qual) break;
                const Pick p = better(s1[i1], a, s2[i2], b, r1_best);
                if (p.decided) cache_ok = false;
                aln.push_back(p.nuc); ref.push_back(f1[i1]); qual.push_back(p.qual);
                k1++; k2++;
            }
        } else if (in1) {
            const char a = Q1(); if (short_qual) break;
            aln.push_back((gap1 && start1 <= i1 && i1 <= stop1) ? 'N' : s1[i1]);
            qual.push_back(a); ref.push_back(f1[i1]); k1++;
        } else {
            const char b = Q2(); if (short_qual) break;
            aln.push_back((gap2 && start2 <= i2 && i2 <= stop2) ? 'N' : s2[i2]);
            qual.push_back(b); ref.push_back(f2[i2]); k2++;
        }
        i1++; i2++;
    }
    if (short_qual) return C2B_E_LIMIT;                      // IndexError in the reference
    // leading / trailing amplicon-gap columns go; aln follows column for column, qual loses one character per dropped column
    size_t lead = 0, n = ref.size();
    while (lead < n && ref[lead] == '-') lead++;
    if (lead == n) return C2B_E_STATE;                       // final_ref[0] on an empty string: IndexError in the reference
    size_t trail = 0;
    while (ref[n - 1 - trail] == '-') trail++;
    const size_t m = n - lead - trail;
    size_t ql = qual.size() > lead ? qual.size() - lead : 0;
    const size_t qoff = qual.size() > lead ? lead : qual.size();
    ql = ql > trail ? ql - trail : 0;
    if ((size_t)cap < m || (size_t)cap < ql) return C2B_E_LIMIT;
    memcpy(out_aln, aln.data() + lead, m);
    memcpy(out_ref, ref.data() + lead, m);
    if (ql) memcpy(out_qual, qual.data() + qoff, ql);
    int match = 0;
    for (size_t k = 0; k < m; k++) match += out_aln[k] == out_ref[k];
    *n_cols = (int32_t)m; *n_qual = (int32_t)ql; *n_match = match; *caching_is_ok = cache_ok ? 1 : 0;
    return C2B_OK;
}
