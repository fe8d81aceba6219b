// ============================================================================
// oracle/kbo_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// See kbo_oracle.hpp for the scope/pinning statement.  Every function cites the
// reference file:line it restates (paths relative to /root/reference).
// ============================================================================
#include "kbo_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace kbo_oracle {

static inline void ensure(bool cond, const char* what) {
    if (!cond) throw Panic{what};
}

// sbwt alphabet is upper-case ACGT only ("$ < A < C < G < T").
static inline int char_idx(uint8_t c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}
static const char ALPHABET[4] = {'A', 'C', 'G', 'T'};

static inline u128 top_mask(int nchars) {  // keeps the `nchars` most significant 2-bit characters
    if (nchars <= 0) return 0;
    if (nchars >= 64) return ~(u128)0;
    return (~(u128)0) << (128 - 2 * nchars);
}

static inline bool node_less(const Node& a, const Node& b) {
    // Colex order with '$' smallest: zero padding below `len` plus the (key,len)
    // tie-break realises "$ < A" (SURVEY 8c "Order").
    if (a.key != b.key) return a.key < b.key;
    return a.len < b.len;
}

static inline int clz128(u128 x) {
    uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
    if (hi) return __builtin_clzll(hi);
    if (lo) return 64 + __builtin_clzll(lo);
    return 128;
}

size_t Index::rank(int c, size_t p) const {
    // number of set bits of row c in [0, p); sampled every 512 bits like the
    // rank support the reference's sbwt crate builds over each row.
    size_t blk = p >> 9;
    size_t res = cum[c][blk];
    size_t w0 = blk << 3, w1 = p >> 6;
    for (size_t w = w0; w < w1; ++w) res += __builtin_popcountll(bits[c][w]);
    if (p & 63) res += __builtin_popcountll(bits[c][w1] & ((~0ULL) >> (64 - (p & 63))));
    return res;
}

// ---------------------------------------------------------------------------
// Index construction: reference src/index.rs:56-99 -> sbwt::SbwtIndexBuilder
// (k, add_rev_comp, build_lcs(true)); semantics per SURVEY 8c.
// ---------------------------------------------------------------------------
Index build_index(const std::vector<std::vector<uint8_t>>& seqs, int k, bool add_revcomp) {
    ensure(!seqs.empty(), "index.rs:60 assert!(!slices.is_empty())");
    ensure(k >= 1 && k <= 64, "oracle supports 1 <= k <= 64");
    Index ix;
    ix.k = k;
    const u128 kmask = top_mask(k);
    const int low_shift = 128 - 2 * k;  // bit position of the FIRST character of a k-mer

    // R: every length-k window of every maximal ACGT run (inputs kept separate).
    std::vector<u128> R;
    for (const auto& s : seqs) {
        u128 fwd = 0, rc = 0;
        size_t run = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            int c = char_idx(s[i]);
            if (c < 0) {
                run = 0;
                fwd = 0;
                rc = 0;
                continue;
            }
            fwd = ((fwd >> 2) | ((u128)c << 126)) & kmask;
            rc = ((rc << 2) | ((u128)(3 - c) << low_shift)) & kmask;
            if (++run >= (size_t)k) {
                R.push_back(fwd);
                if (add_revcomp) R.push_back(rc);
            }
        }
    }
    std::sort(R.begin(), R.end());
    R.erase(std::unique(R.begin(), R.end()), R.end());
    ix.n_kmers = R.size();

    // Dummies: '$'^k plus all '$'-padded proper prefixes of k-mers that have no
    // predecessor k-mer (no y in R with y[1..] == x[..k-1]).
    std::vector<Node> dummies;
    dummies.push_back(Node{0, 0});
    const u128 sufmask = top_mask(k - 1);
    for (u128 x : R) {
        u128 px = (k >= 2) ? (x << 2) : 0;  // x[..k-1] aligned as a (k-1)-suffix
        bool has_pred;
        if (k == 1) {
            has_pred = true;  // the empty string is a suffix of every k-mer
        } else {
            auto it = std::lower_bound(R.begin(), R.end(), px);
            has_pred = (it != R.end()) && ((*it & sufmask) == px);
        }
        if (!has_pred) {
            for (int j = 1; j < k; ++j) dummies.push_back(Node{x << (2 * (k - j)), (uint8_t)j});
        }
    }
    std::sort(dummies.begin(), dummies.end(), node_less);
    dummies.erase(std::unique(dummies.begin(), dummies.end(),
                              [](const Node& a, const Node& b) { return a.key == b.key && a.len == b.len; }),
                  dummies.end());

    // P = R u dummies, colex sorted.
    std::vector<Node>& P = ix.nodes;
    P.reserve(R.size() + dummies.size());
    {
        size_t a = 0, b = 0;
        while (a < R.size() || b < dummies.size()) {
            Node ra{a < R.size() ? R[a] : 0, (uint8_t)k};
            if (b >= dummies.size() || (a < R.size() && node_less(ra, dummies[b]))) {
                P.push_back(ra);
                ++a;
            } else {
                P.push_back(dummies[b]);
                ++b;
            }
        }
    }
    const size_t n = P.size();
    ix.n_sets = n;

    // LCS ('$' never matches).
    ix.lcs.assign(n, 0);
    for (size_t i = 1; i < n; ++i) {
        int common = clz128(P[i - 1].key ^ P[i].key) / 2;
        int lim = std::min<int>(P[i - 1].len, P[i].len);
        ix.lcs[i] = (uint8_t)std::min(common, lim);
    }

    // Subset rows: node i carries c iff P contains P[i][1..]+c and i is the
    // colex-smallest node with that (k-1)-suffix.
    const size_t nwords = (n + 63) / 64 + 1;
    for (int c = 0; c < 4; ++c) ix.bits[c].assign(nwords, 0);
    for (size_t i = 0; i < n; ++i) {
        bool group_first = (i == 0) || (ix.lcs[i] < k - 1);
        if (!group_first) continue;
        for (int c = 0; c < 4; ++c) {
            Node t;
            t.key = ((P[i].key >> 2) | ((u128)c << 126)) & kmask;
            t.len = (uint8_t)std::min<int>(P[i].len + 1, k);
            auto it = std::lower_bound(P.begin(), P.end(), t, node_less);
            if (it != P.end() && it->key == t.key && it->len == t.len) ix.bits[c][i >> 6] |= 1ULL << (i & 63);
        }
    }
    for (int c = 0; c < 4; ++c) {
        size_t nblk = (n >> 9) + 2;
        ix.cum[c].assign(nblk, 0);
        size_t acc = 0;
        for (size_t w = 0; w < nwords; ++w) {
            if ((w & 7) == 0) ix.cum[c][w >> 3] = acc;
            acc += __builtin_popcountll(ix.bits[c][w]);
        }
        for (size_t b = (nwords + 7) / 8; b < nblk; ++b) ix.cum[c][b] = acc;
    }
    size_t acc = 1;
    for (int c = 0; c < 4; ++c) {
        ix.C[c] = acc;
        acc += ix.rank(c, n);
    }
    return ix;
}

// ---------------------------------------------------------------------------
// sbwt::StreamingIndex::matching_statistics (call sites index.rs:251-252,
// variant_calling.rs:264-266,279-280); SURVEY 8a rows 2-4.
// ---------------------------------------------------------------------------
static inline bool extend_right(const Index& ix, size_t l, size_t r, uint8_t ch, size_t* nl, size_t* nr) {
    int c = char_idx(ch);
    if (c < 0) return false;  // empty interval for a byte outside ACGT
    *nl = ix.C[c] + ix.rank(c, l);
    *nr = ix.C[c] + ix.rank(c, r);
    return *nl < *nr;
}

static inline void contract_left(const Index& ix, size_t* l, size_t* r, size_t target) {
    if (target == 0) {  // the literal scans run to both ends of the array: identical result
        *l = 0;
        *r = ix.n_sets;
        return;
    }
    while (*l > 0 && ix.lcs[*l] >= target) --*l;
    while (*r < ix.n_sets && ix.lcs[*r] >= target) ++*r;
}

std::vector<MsEntry> matching_statistics(const Index& ix, const uint8_t* q, size_t len) {
    std::vector<MsEntry> out;
    out.reserve(len);
    size_t d = 0, l = 0, r = ix.n_sets;
    for (size_t i = 0; i < len; ++i) {
        size_t nl = 0, nr = 0;
        bool ok = extend_right(ix, l, r, q[i], &nl, &nr);
        while (d > 0 && !ok) {
            contract_left(ix, &l, &r, d - 1);
            d -= 1;
            ok = extend_right(ix, l, r, q[i], &nl, &nr);
        }
        if (ok) {
            l = nl;
            r = nr;
            d = std::min<size_t>(d + 1, (size_t)ix.k);
        }
        out.push_back(MsEntry{d, l, r});
    }
    return out;
}

std::vector<MsEntry> query_sbwt(const Index& ix, const uint8_t* q, size_t len) {
    ensure(len > 0, "index.rs:248 assert!(!query.is_empty())");
    return matching_statistics(ix, q, len);
}

// sbwt::SbwtIndex::search (call site gap_filling.rs:217): fold extend_right from [0,n).
bool search(const Index& ix, const uint8_t* pat, size_t len, size_t* l, size_t* r) {
    size_t a = 0, b = ix.n_sets;
    for (size_t i = 0; i < len; ++i) {
        size_t na, nb;
        if (!extend_right(ix, a, b, pat[i], &na, &nb)) return false;
        a = na;
        b = nb;
    }
    *l = a;
    *r = b;
    return true;
}

// sbwt::SbwtIndex::access_kmer / push_kmer_to_vec (call sites variant_calling.rs:276,
// gap_filling.rs:144): the string P[colex], dummies padded with '$'.
std::vector<uint8_t> access_kmer(const Index& ix, size_t colex) {
    ensure(colex < ix.n_sets, "access_kmer: colex rank out of range");
    const Node& nd = ix.nodes[colex];
    std::vector<uint8_t> s((size_t)ix.k, (uint8_t)'$');
    for (int t = 0; t < nd.len; ++t) {
        int code = (int)((nd.key >> (126 - 2 * t)) & 3);
        s[(size_t)ix.k - 1 - t] = (uint8_t)ALPHABET[code];
    }
    return s;
}

// ---------------------------------------------------------------------------
// derandomize.rs
// ---------------------------------------------------------------------------
double log_rm_max_cdf(size_t t, size_t alphabet_size, size_t n_kmers) {  // derandomize.rs:91-100
    ensure(n_kmers > 0, "derandomize.rs:96");
    ensure(alphabet_size > 0, "derandomize.rs:97");
    double q = std::exp(std::log(1.0) - std::log((double)alphabet_size));
    // f64::powi(t+1): repeated multiplication semantics
    double pw = __builtin_powi(q, (int)t + 1);
    return (double)n_kmers * std::log1p(-pw);
}

size_t random_match_threshold(size_t k, size_t n_kmers, size_t alphabet_size, double max_error_prob) {
    // derandomize.rs:127-145
    ensure(k > 0, "derandomize.rs:133");
    ensure(n_kmers > 0, "derandomize.rs:134");
    ensure(alphabet_size > 0, "derandomize.rs:135");
    ensure(max_error_prob <= 1.0, "derandomize.rs:136");
    ensure(max_error_prob > 0.0, "derandomize.rs:137");
    for (size_t i = 1; i < k; ++i) {
        if (log_rm_max_cdf(i, alphabet_size, n_kmers) > std::log1p(-max_error_prob)) return i;
    }
    return k;
}

int64_t derandomize_ms_val(size_t curr_noisy_ms, int64_t next_derand_ms, size_t threshold, size_t k) {
    // derandomize.rs:221-247
    ensure(k > 0, "derandomize.rs:227");
    ensure(threshold > 1, "derandomize.rs:228");
    ensure(curr_noisy_ms <= k, "derandomize.rs:229");
    ensure(next_derand_ms <= (int64_t)k, "derandomize.rs:230");
    int64_t run = next_derand_ms - 1;
    if (curr_noisy_ms == k) run = (int64_t)k;
    if (curr_noisy_ms > threshold && next_derand_ms < (int64_t)curr_noisy_ms) run = (int64_t)curr_noisy_ms;
    return run;
}

std::vector<int64_t> derandomize_ms_vec(const std::vector<size_t>& noisy_ms, size_t k, size_t threshold) {
    // derandomize.rs:269-288
    ensure(k > 0, "derandomize.rs:274");
    ensure(threshold > 1, "derandomize.rs:275");
    ensure(noisy_ms.size() > 2, "derandomize.rs:276");
    size_t len = noisy_ms.size();
    std::vector<int64_t> derand(len, 0);
    derand[len - 1] = noisy_ms[len - 1] > threshold ? (int64_t)noisy_ms[len - 1] : 0;
    for (size_t i = 2; i < len + 1; ++i)
        derand[len - i] = derandomize_ms_val(noisy_ms[len - i], derand[len - i + 1], threshold, k);
    return derand;
}

// ---------------------------------------------------------------------------
// translate.rs
// ---------------------------------------------------------------------------
std::pair<char, char> translate_ms_val(int64_t ms_curr, int64_t ms_next, int64_t ms_prev, size_t threshold) {
    // translate.rs:180-216
    ensure(threshold > 1, "translate.rs:186");
    char aln_curr, aln_next = ' ';
    if (ms_curr > (int64_t)threshold && ms_next > 0 && ms_next < (int64_t)threshold) {
        aln_curr = 'R';
        aln_next = 'R';
    } else if (ms_curr <= 0) {
        if (ms_next == 1 && ms_prev > 0)
            aln_curr = 'X';
        else
            aln_curr = '-';
    } else {
        aln_curr = 'M';
    }
    return {aln_curr, aln_next};
}

std::vector<char> translate_ms_vec(const std::vector<int64_t>& derand_ms, size_t k, size_t threshold) {
    // translate.rs:263-293
    ensure(k > 0, "translate.rs:268");
    ensure(threshold > 1, "translate.rs:269");
    ensure(derand_ms.size() > 2, "translate.rs:270");
    size_t len = derand_ms.size();
    std::vector<char> res(len, ' ');
    for (size_t pos = 0; pos < len; ++pos) {
        int64_t prev = pos > 1 ? derand_ms[pos - 1] : (int64_t)k;
        int64_t curr = derand_ms[pos];
        int64_t next = pos < len - 1 ? derand_ms[pos + 1] : derand_ms[pos];
        if (!(pos > 1 && res[pos - 1] == 'R' && res[pos] == 'R')) {
            auto pr = translate_ms_val(curr, next, prev, threshold);
            res[pos] = pr.first;
            if (pos + 1 < len - 1 && pr.second != ' ') res[pos + 1] = pr.second;
        }
    }
    return res;
}

std::vector<char> add_variants(const std::vector<char>& translation, const std::vector<Variant>& variants) {
    // translate.rs:350-386
    std::vector<char> refined = translation;
    auto at = [&](size_t i) -> char& {
        ensure(i < refined.size(), "translate.rs:350-386 index out of bounds");
        return refined[i];
    };
    for (const auto& var : variants) {
        size_t query_len = var.query_chars.size();
        size_t ref_len = var.ref_chars.size();
        if (query_len == ref_len) {
            for (size_t i = 0; i < ref_len; ++i) at(var.query_pos + i) = (char)var.ref_chars[i];
        } else if (query_len == 0) {
            ensure(var.query_pos >= 1, "translate.rs:366 usize underflow");
            at(var.query_pos - 1) = 'I';
            at(var.query_pos) = 'I';
        } else if (ref_len == 0) {
            for (size_t i = 0; i < query_len; ++i) at(var.query_pos + i) = 'D';
        } else {
            bool all_equal = true;
            for (uint8_t c : var.ref_chars) all_equal = all_equal && (c == var.ref_chars[0]);
            char fill = all_equal ? (char)var.ref_chars[0] : 'N';
            for (size_t i = 0; i < query_len; ++i) at(var.query_pos + i) = fill;
        }
    }
    return refined;
}

// ---------------------------------------------------------------------------
// format.rs
// ---------------------------------------------------------------------------
std::vector<RLE> run_lengths_gapped(const std::vector<char>& aln, size_t max_gap_len) {
    // format.rs:143-193
    std::vector<RLE> enc;
    size_t i = 0;
    bool match_start = false;
    while (i < aln.size()) {
        match_start = (aln[i] != '-' && aln[i] != ' ') && !match_start;
        if (match_start) {
            RLE rle;
            rle.start = i;
            size_t within_gap_bases = 0;
            bool within_gap_start = false;
            while (i < aln.size() && aln[i] != ' ') {
                bool is_true_gap = aln[i] == '-';
                if (is_true_gap && !within_gap_start) {
                    within_gap_start = true;
                    rle.gap_opens += 1;
                    within_gap_bases = 0;
                }
                if (!is_true_gap && within_gap_start) within_gap_start = false;
                bool is_match = aln[i] == 'M' || aln[i] == 'R' || aln[i] == 'I';
                bool is_gap = is_true_gap || aln[i] == 'D';
                rle.matches += is_match;
                rle.gap_bases += is_gap;
                rle.mismatches += (!is_match && !is_gap);
                rle.end = (is_match || !is_gap) ? i + 1 : rle.end;
                if (aln[i] == 'R') {
                    ensure(i >= 1, "format.rs:176 aln[i - 1] at i == 0");
                    rle.jumps += (aln[i - 1] == 'R');
                }
                within_gap_bases += (aln[i] == '-');
                i += 1;
                if (within_gap_bases > max_gap_len || (is_gap && i == aln.size() && rle.gap_opens > 0)) {
                    ensure(rle.gap_opens >= 1, "format.rs:181 usize underflow");
                    rle.gap_opens -= 1;
                    ensure(rle.gap_bases >= within_gap_bases, "format.rs:182 usize underflow");
                    rle.gap_bases -= within_gap_bases;
                    break;
                }
            }
            enc.push_back(rle);
            match_start = false;
        } else {
            i += 1;
        }
    }
    return enc;
}

std::vector<RLE> run_lengths(const std::vector<char>& aln) { return run_lengths_gapped(aln, 0); }  // format.rs:98-102

std::vector<uint8_t> relative_to_ref(const uint8_t* ref_seq, size_t len, const std::vector<char>& alignment) {
    // format.rs:266-287 (zip stops at the shorter input)
    size_t n = std::min(len, alignment.size());
    std::vector<uint8_t> out(n);
    for (size_t i = 0; i < n; ++i) {
        char a = alignment[i];
        if (a == 'M' || a == 'R' || a == 'I')
            out[i] = ref_seq[i];
        else if (a == 'X')
            out[i] = '-';
        else if (a == 'D')
            out[i] = '-';
        else if (a != '-')
            out[i] = (uint8_t)a;
        else
            out[i] = '-';
    }
    return out;
}

// ---------------------------------------------------------------------------
// variant_calling.rs
// ---------------------------------------------------------------------------
static std::vector<uint8_t> get_kmer_ending_at(const uint8_t* query, size_t end_pos, size_t k) {
    // variant_calling.rs:46-59
    std::vector<uint8_t> kmer;
    if (end_pos >= k - 1) {
        kmer.insert(kmer.end(), query + end_pos + 1 - k, query + end_pos + 1);
    } else {
        int64_t n_dollars = -((int64_t)end_pos - (int64_t)k + 1);
        ensure(n_dollars > 0, "variant_calling.rs:53");
        kmer.resize((size_t)n_dollars, (uint8_t)'$');
        kmer.insert(kmer.end(), query, query + end_pos + 1);
    }
    ensure(kmer.size() == k, "variant_calling.rs:57");
    return kmer;
}

static size_t longest_common_suffix(const std::vector<uint8_t>& x, const std::vector<uint8_t>& y) {
    // variant_calling.rs:61-72
    size_t len = 0;
    for (size_t i = 0; i < std::min(x.size(), y.size()); ++i) {
        if (x[x.size() - 1 - i] == y[y.size() - 1 - i])
            len += 1;
        else
            break;
    }
    return len;
}

static bool get_rightmost_significant_peak(const std::vector<MsEntry>& ms, size_t thr, size_t* peak) {
    // variant_calling.rs:74-83
    ensure(!ms.empty(), "variant_calling.rs:75");
    for (size_t ii = ms.size() - 1; ii-- > 0;) {
        size_t here = ms[ii].d, next = ms[ii + 1].d;
        if (here >= thr && here > next) {
            *peak = ii;
            return true;
        }
    }
    return false;
}

bool resolve_variant(const std::vector<uint8_t>& query_kmer, const std::vector<uint8_t>& ref_kmer,
                     const std::vector<MsEntry>& ms_vs_query, const std::vector<MsEntry>& ms_vs_ref,
                     size_t significant_match_threshold, std::vector<uint8_t>* query_chars,
                     std::vector<uint8_t>* ref_chars) {
    // variant_calling.rs:139-201; returns false where the reference returns Err.
    size_t k = query_kmer.size();
    ensure(ref_kmer.size() == k, "variant_calling.rs:148");
    ensure(ms_vs_query.size() == k, "variant_calling.rs:149");
    ensure(ms_vs_ref.size() == k, "variant_calling.rs:150");
    size_t common_suffix_len = longest_common_suffix(query_kmer, ref_kmer);
    ensure(common_suffix_len > 0, "variant_calling.rs:153");
    size_t query_ms_peak = 0, ref_ms_peak = 0;
    bool have_q = get_rightmost_significant_peak(ms_vs_ref, significant_match_threshold, &query_ms_peak);
    bool have_r = get_rightmost_significant_peak(ms_vs_query, significant_match_threshold, &ref_ms_peak);
    if (have_q && have_r) {
        size_t suffix_match_start = k - common_suffix_len;
        int64_t query_gap = (int64_t)suffix_match_start - (int64_t)query_ms_peak - 1;
        int64_t ref_gap = (int64_t)suffix_match_start - (int64_t)ref_ms_peak - 1;
        auto slice = [&](const std::vector<uint8_t>& v, size_t a, size_t b) {
            ensure(a <= b && b <= v.size(), "variant_calling.rs:139-201 slice out of range");
            return std::vector<uint8_t>(v.begin() + a, v.begin() + b);
        };
        if (query_gap > 0 && ref_gap > 0) {
            *query_chars = slice(query_kmer, query_ms_peak + 1, suffix_match_start);
            *ref_chars = slice(ref_kmer, ref_ms_peak + 1, suffix_match_start);
            return true;
        } else {
            int64_t query_overlap = -query_gap, ref_overlap = -ref_gap;
            if (query_overlap == ref_overlap) return false;
            size_t variant_len = (size_t)std::llabs(query_overlap - ref_overlap);
            if (query_overlap > ref_overlap) {
                query_chars->clear();
                *ref_chars = slice(ref_kmer, ref_ms_peak + 1, ref_ms_peak + 1 + variant_len);
                return true;
            } else {
                *query_chars = slice(query_kmer, query_ms_peak + 1, query_ms_peak + 1 + variant_len);
                ref_chars->clear();
                return true;
            }
        }
    }
    return false;
}

std::vector<Variant> call_variants(const Index& sbwt_ref, const Index& sbwt_query, const uint8_t* query,
                                   size_t len, double max_error_prob) {
    // variant_calling.rs:249-294
    ensure(sbwt_ref.k == sbwt_query.k, "variant_calling.rs:258");
    size_t k = (size_t)sbwt_ref.k;
    size_t d = random_match_threshold(k, sbwt_ref.n_kmers, 4, max_error_prob);
    std::vector<Variant> calls;
    std::vector<MsEntry> ms_vs_ref = matching_statistics(sbwt_ref, query, len);
    for (size_t i = 1; i < len; ++i) {
        if (ms_vs_ref[i].d < ms_vs_ref[i - 1].d && ms_vs_ref[i - 1].d >= d && ms_vs_ref[i].d < d) {
            for (size_t j = i + 1; j < std::min(i + k + 1, len); ++j) {
                if (ms_vs_ref[j].d >= d && ms_vs_ref[j].r - ms_vs_ref[j].l == 1) {
                    size_t ref_colex = ms_vs_ref[j].l;
                    std::vector<uint8_t> query_kmer = get_kmer_ending_at(query, j, k);
                    std::vector<uint8_t> ref_kmer = access_kmer(sbwt_ref, ref_colex);
                    std::vector<MsEntry> ms_q_vs_ref = matching_statistics(sbwt_ref, query_kmer.data(), k);
                    std::vector<MsEntry> ms_vs_query = matching_statistics(sbwt_query, ref_kmer.data(), k);
                    std::vector<uint8_t> qc, rc;
                    if (resolve_variant(query_kmer, ref_kmer, ms_vs_query, ms_q_vs_ref, d, &qc, &rc))
                        calls.push_back(Variant{i, qc, rc});
                    break;
                }
            }
        }
    }
    return calls;
}

// ---------------------------------------------------------------------------
// gap_filling.rs
// ---------------------------------------------------------------------------
static size_t count_right_overlaps(const std::vector<uint8_t>& kmer, const uint8_t* ref_seq, size_t ref_len,
                                   size_t ref_match_end) {
    // gap_filling.rs:20-43 (release-profile wrapping arithmetic, bounds-checked indexing)
    ensure(!kmer.empty(), "gap_filling.rs:25");
    ensure(ref_len > 0, "gap_filling.rs:26");
    ensure(ref_len >= ref_match_end, "gap_filling.rs:27");
    size_t kmer_pos = kmer.size() - 1;
    size_t ref_pos = ref_match_end - 1;  // wraps when ref_match_end == 0
    size_t matches = 0;
    while (kmer_pos > 0) {
        ensure(ref_pos < ref_len, "gap_filling.rs:33 index out of bounds");
        if (ref_seq[ref_pos] == kmer[kmer_pos])
            matches += 1;
        else
            break;
        kmer_pos -= 1;
        ref_pos -= 1;
    }
    return matches;
}

static size_t count_left_overlaps(const std::vector<uint8_t>& kmer, const uint8_t* ref_seq, size_t ref_len,
                                  size_t ref_match_start) {
    // gap_filling.rs:45-67
    ensure(!kmer.empty(), "gap_filling.rs:50");
    ensure(ref_len > 0, "gap_filling.rs:51");
    ensure(ref_len > ref_match_start, "gap_filling.rs:52");
    size_t kmer_pos = 0, ref_pos = ref_match_start, matches = 0;
    while (kmer_pos < kmer.size()) {
        ensure(ref_pos < ref_len, "gap_filling.rs:58 index out of bounds");
        if (ref_seq[ref_pos] == kmer[kmer_pos])
            matches += 1;
        else
            break;
        kmer_pos += 1;
        ref_pos += 1;
    }
    return matches;
}

std::pair<size_t, std::vector<uint8_t>> nearest_unique_context(const std::vector<MsEntry>& ms,
                                                                       const Index& sbwt, size_t range_start,
                                                                       size_t range_end) {
    // gap_filling.rs:127-151
    ensure(sbwt.k > 0, "gap_filling.rs:133");
    ensure(!ms.empty(), "gap_filling.rs:134");
    ensure(range_end >= range_start, "gap_filling.rs:135");
    ensure(range_end < ms.size(), "gap_filling.rs:136");
    std::vector<uint8_t> kmer;
    size_t kmer_idx = range_end;
    while (kmer_idx >= range_start) {
        ensure(kmer_idx < ms.size(), "gap_filling.rs:142 index out of bounds");
        if (ms[kmer_idx].r - ms[kmer_idx].l == 1) {
            kmer = access_kmer(sbwt, ms[kmer_idx].l);
            break;
        }
        kmer_idx -= 1;  // wraps at 0 -> caught by the bounds check above
    }
    return {kmer_idx, kmer};
}

std::vector<uint8_t> left_extend_kmer(const std::vector<uint8_t>& kmer_start, const Index& sbwt,
                                             size_t max_extension_len) {
    // gap_filling.rs:205-232
    ensure(!kmer_start.empty(), "gap_filling.rs:210");
    size_t left_extension_len = 0;
    std::vector<uint8_t> kmer = kmer_start;
    while (left_extension_len < max_extension_len) {
        std::vector<std::pair<std::vector<uint8_t>, std::pair<size_t, size_t>>> new_kmers;
        for (int c = 0; c < 4; ++c) {
            std::vector<uint8_t> nk;
            nk.push_back((uint8_t)ALPHABET[c]);
            nk.insert(nk.end(), kmer.begin(), kmer.begin() + (kmer.size() - (left_extension_len + 1)));
            size_t l, r;
            if (search(sbwt, nk.data(), nk.size(), &l, &r)) new_kmers.push_back({nk, {l, r}});
        }
        if (new_kmers.size() == 1 && new_kmers[0].second.second - new_kmers[0].second.first == 1) {
            kmer.insert(kmer.begin(), new_kmers[0].first[0]);
        } else {
            break;
        }
        left_extension_len += 1;
    }
    return kmer;
}

std::vector<uint8_t> left_extend_over_gap(const std::vector<MsEntry>& ms, const uint8_t* ref_seq,
                                                 size_t ref_len, const Index& sbwt, size_t left_overlap_req,
                                                 size_t right_overlap_req, size_t gap_start, size_t gap_end,
                                                 size_t search_radius) {
    // gap_filling.rs:295-361
    size_t k = (size_t)sbwt.k;
    ensure(k > 0, "gap_filling.rs:305");
    ensure(ms.size() == ref_len, "gap_filling.rs:306");
    ensure(left_overlap_req <= gap_start, "gap_filling.rs:307");
    ensure(gap_end <= ref_len && right_overlap_req <= ref_len - gap_end, "gap_filling.rs:308");
    ensure(gap_end > gap_start, "gap_filling.rs:309");
    ensure(gap_end < ms.size(), "gap_filling.rs:310");

    size_t search_start = std::min(gap_end + search_radius, ref_len - 1);
    size_t search_end = gap_end + right_overlap_req;
    auto slice = [&](const std::vector<uint8_t>& v, size_t a, size_t b) {
        ensure(a <= b && b <= v.size(), "gap_filling.rs:336/350 slice out of range");
        return std::vector<uint8_t>(v.begin() + a, v.begin() + b);
    };

    std::vector<uint8_t> kmer;
    size_t kmer_idx = search_start;
    while (kmer_idx >= search_end) {
        auto ctx = nearest_unique_context(ms, sbwt, search_end, kmer_idx);
        kmer_idx = ctx.first;
        kmer = ctx.second;
        if (!kmer.empty()) {
            size_t right_matches_want = search_start - (gap_end - 1) - (search_start - kmer_idx);
            size_t right_matches_got = count_right_overlaps(kmer, ref_seq, ref_len, gap_end + right_matches_want);
            size_t ref_start_pos = gap_start > left_overlap_req ? gap_start - left_overlap_req : 0;
            size_t left_matches_got = count_left_overlaps(kmer, ref_seq, ref_len, ref_start_pos);
            bool should_extend = kmer.size() < left_overlap_req + (gap_end - gap_start) + right_matches_got;
            if (right_matches_got >= std::min(right_matches_want, k) && left_matches_got >= left_overlap_req) {
                size_t start = left_matches_got - left_overlap_req;
                ensure(right_matches_got >= right_overlap_req, "gap_filling.rs:335 usize underflow");
                ensure(kmer.size() >= right_matches_got - right_overlap_req, "gap_filling.rs:335 usize underflow");
                size_t end = kmer.size() - (right_matches_got - right_overlap_req);
                kmer = slice(kmer, start, end);
                break;
            } else if (should_extend && right_matches_got >= std::min(right_matches_want, k) &&
                       left_matches_got < left_overlap_req) {
                size_t left_extend_length = left_overlap_req + (gap_end - gap_start) + right_matches_got - k;
                kmer = left_extend_kmer(kmer, sbwt, left_extend_length);
                size_t lm = count_left_overlaps(kmer, ref_seq, ref_len, ref_start_pos);
                if (lm >= left_overlap_req) {
                    size_t start = lm - left_overlap_req;
                    ensure(right_matches_got >= right_overlap_req, "gap_filling.rs:349 usize underflow");
                    ensure(kmer.size() >= right_matches_got - right_overlap_req, "gap_filling.rs:349 usize underflow");
                    size_t end = kmer.size() - (right_matches_got - right_overlap_req);
                    kmer = slice(kmer, start, end);
                    break;
                }
            }
            kmer.clear();
        }
        ensure(kmer_idx >= 1, "gap_filling.rs:357 usize underflow");
        kmer_idx -= 1;
    }
    return kmer;
}

std::vector<char> fill_gaps(const std::vector<char>& translation, const std::vector<MsEntry>& noisy_ms,
                            const uint8_t* ref_seq, size_t len, const Index& query_sbwt, size_t threshold,
                            double max_err_prob) {
    // gap_filling.rs:444-526
    size_t n_elements = translation.size();
    ensure(!translation.empty(), "gap_filling.rs:453");
    ensure(translation.size() == noisy_ms.size(), "gap_filling.rs:454");
    size_t k = (size_t)query_sbwt.k;
    ensure(k > 0, "gap_filling.rs:457");
    std::vector<char> refined = translation;
    ensure(refined.size() >= threshold, "gap_filling.rs:467 usize underflow");

    size_t i = threshold + 1;
    while (i < refined.size() - threshold) {
        if (refined[i - 1] == '-' || refined[i - 1] == 'X') {
            size_t start_index = i - 1;
            while (i < n_elements && refined[i] == '-') i += 1;
            size_t end_index = std::min(i, refined.size() - threshold);

            bool overlap_without_extend = end_index - start_index + 2 * threshold <= k;
            size_t search_radius = k - (threshold * (size_t)overlap_without_extend);
            std::vector<uint8_t> kmer = left_extend_over_gap(noisy_ms, ref_seq, len, query_sbwt, threshold, threshold,
                                                             start_index, end_index, search_radius);

            bool kmer_found = !kmer.empty() && std::find(kmer.begin(), kmer.end(), (uint8_t)'$') == kmer.end();
            bool no_indels = kmer.size() == threshold + (end_index - start_index) + threshold;

            size_t a = std::min(threshold, kmer.size());
            size_t b = std::min(threshold + end_index - start_index, kmer.size());
            std::vector<bool> matching_bases;
            for (size_t t = a, p = start_index; t < b && p < end_index; ++t, ++p)
                matching_bases.push_back(kmer[t] == ref_seq[p]);

            size_t total_overlaps = 0;
            for (bool x : matching_bases) total_overlaps += x;
            double log_probs = 0.0;
            {
                size_t consecutive_overlaps = 0;
                for (size_t w = 0; w + 1 < matching_bases.size(); ++w) {
                    if (matching_bases[w] && matching_bases[w + 1]) {
                        consecutive_overlaps += 1;
                        log_probs += 0.0;
                    } else {
                        double lp = consecutive_overlaps > 0 ? log_rm_max_cdf(consecutive_overlaps + 1, 4, 1) : 0.0;
                        consecutive_overlaps = 0;
                        log_probs += lp;
                    }
                }
            }
            bool fill_overlaps = log_probs > std::log1p(-max_err_prob);
            bool fill_flanked = !matching_bases.empty() && !matching_bases[0] &&
                                !matching_bases[matching_bases.size() - 1] &&
                                total_overlaps + 2 == end_index - start_index;
            bool pass_checks = kmer_found && no_indels && (overlap_without_extend || fill_overlaps || fill_flanked);
            if (pass_checks) {
                for (size_t p = start_index, t = threshold; p < end_index; ++p, ++t)
                    refined[p] = (kmer[t] == ref_seq[p]) ? 'M' : (char)kmer[t];
            }
        }
        i += 1;
    }
    return refined;
}

// ---------------------------------------------------------------------------
// lib.rs
// ---------------------------------------------------------------------------
std::vector<char> matches(const Index& ix, const uint8_t* q, size_t len, double max_error_prob) {
    // lib.rs:612-628
    size_t k = (size_t)ix.k;
    size_t threshold = random_match_threshold(k, ix.n_kmers, 4, max_error_prob);
    std::vector<MsEntry> ms = query_sbwt(ix, q, len);
    std::vector<size_t> noisy(ms.size());
    for (size_t i = 0; i < ms.size(); ++i) noisy[i] = ms[i].d;
    std::vector<int64_t> derand = derandomize_ms_vec(noisy, k, threshold);
    return translate_ms_vec(derand, k, threshold);
}

std::vector<RLE> find(const Index& ix, const uint8_t* q, size_t len, double max_error_prob, size_t max_gap_len) {
    // lib.rs:808-821
    std::vector<char> aln = matches(ix, q, len, max_error_prob);
    if (max_gap_len > 0) return run_lengths_gapped(aln, max_gap_len);
    return run_lengths(aln);
}

std::vector<Variant> call(const Index& sbwt_query, const uint8_t* ref_seq, size_t len, double max_error_prob,
                          int build_k, bool build_add_revcomp) {
    // lib.rs:547-573: builds the SBWT of ref_seq, then call_variants with the
    // ASSEMBLY index in the `sbwt_ref` slot (argument order at lib.rs:561-567).
    std::vector<std::vector<uint8_t>> v(1, std::vector<uint8_t>(ref_seq, ref_seq + len));
    Index sbwt_ref = build_index(v, build_k, build_add_revcomp);
    ensure(sbwt_ref.k == sbwt_query.k, "lib.rs:559");
    return call_variants(sbwt_query, sbwt_ref, ref_seq, len, max_error_prob);
}

std::vector<uint8_t> map(const Index& query_sbwt_ix, const uint8_t* ref_seq, size_t len, const MapOpts& opts) {
    // lib.rs:720-761
    size_t k = (size_t)query_sbwt_ix.k;
    if (opts.call_variants) ensure((int)k == opts.build_k, "lib.rs:729");
    size_t threshold = random_match_threshold(k, query_sbwt_ix.n_kmers, 4, opts.max_error_prob);
    std::vector<MsEntry> noisy_ms = query_sbwt(query_sbwt_ix, ref_seq, len);
    std::vector<size_t> noisy(noisy_ms.size());
    for (size_t i = 0; i < noisy.size(); ++i) noisy[i] = noisy_ms[i].d;
    std::vector<int64_t> derand = derandomize_ms_vec(noisy, k, threshold);
    std::vector<char> translation = translate_ms_vec(derand, k, threshold);
    std::vector<char> refined = opts.fill_gaps ? fill_gaps(translation, noisy_ms, ref_seq, len, query_sbwt_ix,
                                                           threshold, opts.max_error_prob)
                                               : translation;
    std::vector<char> with_variants;
    if (opts.call_variants) {
        std::vector<Variant> variants =
            call(query_sbwt_ix, ref_seq, len, opts.max_error_prob, opts.build_k, opts.build_add_revcomp);
        with_variants = add_variants(refined, variants);
    } else {
        with_variants = refined;
    }
    if (opts.format) return relative_to_ref(ref_seq, len, with_variants);
    return std::vector<uint8_t>(with_variants.begin(), with_variants.end());
}

}  // namespace kbo_oracle
