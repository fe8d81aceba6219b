// ===========================================================================
// kbo_b200/csrc/refine_host.cpp -- see refine_host.hpp.
// Each routine states the reference lines whose behaviour it reproduces
// (paths relative to the reference checkout, tmaklin/kbo 0.5.1).
// ===========================================================================
#include "refine_host.hpp"

#include <algorithm>
#include <cmath>
#include <thread>
#include <utility>

namespace kbo_b200 {
namespace {

[[noreturn]] void panic(const char* what) { throw RefinePanic{what}; }
inline void require(bool ok, const char* what) {
    if (!ok) panic(what);
}

typedef std::vector<uint8_t> Bytes;

// rightmost i in [0, k-2] with d[i] >= thr and d[i] > d[i+1]     (variant_calling.rs:74-83)
int rightmost_significant_peak(const uint8_t* d, uint32_t k, uint64_t thr) {
    for (int i = (int)k - 2; i >= 0; --i)
        if (d[i] >= thr && d[i] > d[i + 1]) return i;
    return -1;
}

uint32_t common_suffix(const uint8_t* a, const uint8_t* b, uint32_t k) {  // variant_calling.rs:61-72
    uint32_t n = 0;
    while (n < k && a[k - 1 - n] == b[k - 1 - n]) ++n;
    return n;
}

// ---- gap filling helpers ------------------------------------------------------
// number of matching characters walking left from kmer's last base / ref[ref_match_end-1];
// the k-mer's first base is never compared                       (gap_filling.rs:20-43)
uint64_t right_overlaps(const Bytes& kmer, const uint8_t* ref, uint64_t ref_len, uint64_t ref_match_end) {
    require(!kmer.empty() && ref_len > 0 && ref_len >= ref_match_end, "gap_filling.rs:25-27");
    uint64_t n = 0;
    for (uint64_t kp = kmer.size() - 1, rp = ref_match_end - 1; kp > 0; --kp, --rp) {
        require(rp < ref_len, "gap_filling.rs:33 index out of bounds");
        if (ref[rp] != kmer[kp]) break;
        ++n;
    }
    return n;
}

// number of matching characters walking right from kmer[0] / ref[ref_match_start]  (gap_filling.rs:45-67)
uint64_t left_overlaps(const Bytes& kmer, const uint8_t* ref, uint64_t ref_len, uint64_t ref_match_start) {
    require(!kmer.empty() && ref_len > 0 && ref_len > ref_match_start, "gap_filling.rs:50-52");
    uint64_t n = 0;
    for (uint64_t kp = 0, rp = ref_match_start; kp < kmer.size(); ++kp, ++rp) {
        require(rp < ref_len, "gap_filling.rs:58 index out of bounds");
        if (ref[rp] != kmer[kp]) break;
        ++n;
    }
    return n;
}

// walk from range_end down to range_start; first position with a singleton interval gives the k-mer
// (gap_filling.rs:127-151).  Returns the stopping index (range_start - 1 when nothing was found).
uint64_t unique_context(const MsArrays& ms, const HostIndex& ix, uint64_t range_start, uint64_t range_end,
                        Bytes* kmer) {
    require(ix.k > 0 && ms.n > 0 && range_end >= range_start && range_end < ms.n, "gap_filling.rs:133-136");
    kmer->clear();
    uint64_t idx = range_end;
    while (idx >= range_start) {
        require(idx < ms.n, "gap_filling.rs:142 index out of bounds");
        if (ms.r[idx] - ms.l[idx] == 1) {
            kmer->resize(ix.k);
            ix.access_kmer(ms.l[idx], kmer->data());
            break;
        }
        --idx;  // wraps at 0; caught by the bound check above
    }
    return idx;
}

// prepend bases while exactly one base extends the k-mer's first k-1 characters to a unique node
// (gap_filling.rs:205-232).  The reference searches the four k-length patterns c + S (S = those k-1
// characters); a k-length pattern matches at most one node, and the nodes ending with S are exactly the
// interval of S, so ONE search of S answers all four: the candidates are the full (non-dummy) nodes of that
// interval and their first characters.
Bytes extend_left(const Bytes& start, const HostIndex& ix, uint64_t max_extension) {
    require(!start.empty(), "gap_filling.rs:210");
    Bytes kmer = start;
    Bytes node(ix.k);
    for (uint64_t ext = 0; ext < max_extension; ++ext) {
        const uint64_t slen = kmer.size() - ext - 1;  // = k - 1 when `start` is a k-mer
        int hits = 0;
        uint8_t hit_base = 0;
        if (slen + 1 == ix.k) {
            uint64_t l = 0, r = 0;
            if (ix.search(kmer.data(), slen, &l, &r)) {
                for (uint64_t v = l; v < r; ++v) {
                    ix.access_kmer(v, node.data());
                    if (node[0] != '$') { ++hits; hit_base = node[0]; }
                }
            }
        } else {  // general pattern length: the literal four searches
            static const uint8_t letters[4] = {'A', 'C', 'G', 'T'};
            uint64_t hit_width = 0;
            Bytes probe(slen + 1);
            std::copy(kmer.begin(), kmer.begin() + slen, probe.begin() + 1);
            for (uint8_t c : letters) {
                probe[0] = c;
                uint64_t l = 0, r = 0;
                if (ix.search(probe.data(), probe.size(), &l, &r)) {
                    ++hits;
                    hit_base = c;
                    hit_width = r - l;
                }
            }
            if (hit_width != 1) hits = 0;
        }
        if (hits != 1) break;
        kmer.insert(kmer.begin(), hit_base);
    }
    return kmer;
}

// gap_filling.rs:295-361
Bytes bridge_gap(const MsArrays& ms, const uint8_t* ref, uint64_t ref_len, const HostIndex& ix, uint64_t left_req,
                 uint64_t right_req, uint64_t gap_start, uint64_t gap_end, uint64_t search_radius) {
    const uint64_t k = ix.k;
    require(k > 0 && ms.n == ref_len && left_req <= gap_start, "gap_filling.rs:305-307");
    require(gap_end <= ref_len && right_req <= ref_len - gap_end, "gap_filling.rs:308");
    require(gap_end > gap_start && gap_end < ms.n, "gap_filling.rs:309-310");
    const uint64_t search_start = std::min(gap_end + search_radius, ref_len - 1);
    const uint64_t search_end = gap_end + right_req;
    const uint64_t gap_len = gap_end - gap_start;
    const uint64_t ref_start = gap_start > left_req ? gap_start - left_req : 0;
    auto trim = [&](const Bytes& v, uint64_t a, uint64_t right_got) {
        require(right_got >= right_req && v.size() >= right_got - right_req, "gap_filling.rs:335 usize underflow");
        const uint64_t b = v.size() - (right_got - right_req);
        require(a <= b, "gap_filling.rs:336 slice out of range");
        return Bytes(v.begin() + a, v.begin() + b);
    };
    Bytes kmer;
    uint64_t idx = search_start;
    while (idx >= search_end) {
        idx = unique_context(ms, ix, search_end, idx, &kmer);
        if (!kmer.empty()) {
            const uint64_t right_want = idx + 1 - gap_end;  // = search_start-(gap_end-1)-(search_start-idx)
            const uint64_t right_got = right_overlaps(kmer, ref, ref_len, gap_end + right_want);
            const uint64_t left_got = left_overlaps(kmer, ref, ref_len, ref_start);
            const bool right_ok = right_got >= std::min(right_want, k);
            if (right_ok && left_got >= left_req) return trim(kmer, left_got - left_req, right_got);
            if (right_ok && left_got < left_req && kmer.size() < left_req + gap_len + right_got) {
                Bytes longer = extend_left(kmer, ix, left_req + gap_len + right_got - k);
                const uint64_t lg = left_overlaps(longer, ref, ref_len, ref_start);
                if (lg >= left_req) return trim(longer, lg - left_req, right_got);
            }
            kmer.clear();
        }
        require(idx >= 1, "gap_filling.rs:357 usize underflow");
        --idx;
    }
    return kmer;
}

}  // namespace

// ---------------------------------------------------------------------------
// variant_calling.rs
// ---------------------------------------------------------------------------
bool resolve_variant(const uint8_t* query_kmer, const uint8_t* ref_kmer, const uint8_t* ms_vs_query_d,
                     const uint8_t* ms_vs_ref_d, uint32_t k, uint64_t thr, Bytes* query_chars, Bytes* ref_chars) {
    const uint32_t suffix = common_suffix(query_kmer, ref_kmer, k);
    require(suffix > 0, "variant_calling.rs:153 assert!(common_suffix_len > 0)");
    const int qpeak = rightmost_significant_peak(ms_vs_ref_d, k, thr);
    const int rpeak = rightmost_significant_peak(ms_vs_query_d, k, thr);
    if (qpeak < 0 || rpeak < 0) return false;
    const int64_t suffix_start = (int64_t)k - suffix;
    const int64_t qgap = suffix_start - qpeak - 1, rgap = suffix_start - rpeak - 1;  // negative = overlap
    auto piece = [&](const uint8_t* s, int64_t a, int64_t b, Bytes* out) {
        require(a <= b && b <= (int64_t)k, "variant_calling.rs:139-201 slice out of range");
        out->assign(s + a, s + b);
    };
    query_chars->clear();
    ref_chars->clear();
    if (qgap > 0 && rgap > 0) {
        piece(query_kmer, qpeak + 1, suffix_start, query_chars);
        piece(ref_kmer, rpeak + 1, suffix_start, ref_chars);
        return true;
    }
    const int64_t qov = -qgap, rov = -rgap;
    if (qov == rov) return false;
    const int64_t vlen = qov > rov ? qov - rov : rov - qov;
    if (qov > rov) piece(ref_kmer, rpeak + 1, rpeak + 1 + vlen, ref_chars);      // deletion in the query
    else piece(query_kmer, qpeak + 1, qpeak + 1 + vlen, query_chars);            // insertion in the query
    return true;
}

std::vector<VariantCandidate64> find_variant_candidates(const MsArrays& ms, uint64_t len, uint32_t k, uint64_t thr) {
    // candidates = significant drops followed (within k) by a significant unique match
    std::vector<VariantCandidate64> cands;
    for (uint64_t i = 1; i < len; ++i) {
        if (!(ms.d[i] < ms.d[i - 1] && ms.d[i - 1] >= thr && ms.d[i] < thr)) continue;
        const uint64_t stop = std::min<uint64_t>(i + k + 1, len);
        for (uint64_t j = i + 1; j < stop; ++j) {
            if (ms.d[j] >= thr && ms.r[j] - ms.l[j] == 1) {
                cands.push_back(VariantCandidate64{i, j, ms.l[j]});
                break;
            }
        }
    }
    return cands;
}

std::vector<VariantRec> call_variants_from(const HostIndex& sbwt_ref, const std::vector<VariantCandidate64>& cands,
                                           const uint8_t* query, uint64_t len, uint64_t thr, const KmerMsFn& kmer_ms,
                                           const AccessKmersFn* access) {
    (void)len;
    const uint32_t k = sbwt_ref.k;
    std::vector<VariantRec> calls;
    if (cands.empty()) return calls;
    // the k-mer ending at j in the query ('$'-padded at the start, variant_calling.rs:46-59) and the index k-mer of
    // the unique node; their MS against the two indexes in ONE batch each
    const uint64_t nc = cands.size();
    Bytes qk(nc * k), rk(nc * k), ms_q_vs_ref(nc * k), ms_r_vs_query(nc * k);
    for (uint64_t c = 0; c < nc; ++c) {
        const uint64_t j = cands[c].j;
        uint8_t* dst = qk.data() + c * k;
        if (j + 1 >= k) {
            std::copy(query + j + 1 - k, query + j + 1, dst);
        } else {
            const uint64_t dollars = k - (j + 1);
            std::fill(dst, dst + dollars, (uint8_t)'$');
            std::copy(query, query + j + 1, dst + dollars);
        }
        if (!access) sbwt_ref.access_kmer(cands[c].node, rk.data() + c * k);
    }
    if (access) (*access)(cands, k, rk.data());
    kmer_ms(0, qk.data(), nc, k, ms_q_vs_ref.data());
    kmer_ms(1, rk.data(), nc, k, ms_r_vs_query.data());
    for (uint64_t c = 0; c < nc; ++c) {
        VariantRec v;
        v.query_pos = cands[c].i;
        if (resolve_variant(qk.data() + c * k, rk.data() + c * k, ms_r_vs_query.data() + c * k,
                            ms_q_vs_ref.data() + c * k, k, thr, &v.query_chars, &v.ref_chars))
            calls.push_back(v);
    }
    return calls;
}

std::vector<VariantRec> call_variants(const HostIndex& sbwt_ref, const MsArrays& ms, const uint8_t* query,
                                      uint64_t len, uint64_t thr, const KmerMsFn& kmer_ms) {
    return call_variants_from(sbwt_ref, find_variant_candidates(ms, len, sbwt_ref.k, thr), query, len, thr, kmer_ms);
}

// ---------------------------------------------------------------------------
// translate.rs:350-386
// ---------------------------------------------------------------------------
void add_variants(Bytes* translation, const std::vector<VariantRec>& variants) {
    Bytes& t = *translation;
    auto at = [&](uint64_t i) -> uint8_t& {
        require(i < t.size(), "translate.rs:350-386 index out of bounds");
        return t[i];
    };
    for (const VariantRec& v : variants) {
        const uint64_t ql = v.query_chars.size(), rl = v.ref_chars.size();
        if (ql == rl) {
            for (uint64_t i = 0; i < rl; ++i) at(v.query_pos + i) = v.ref_chars[i];
        } else if (ql == 0) {
            require(v.query_pos >= 1, "translate.rs:366 usize underflow");
            at(v.query_pos - 1) = 'I';
            at(v.query_pos) = 'I';
        } else if (rl == 0) {
            for (uint64_t i = 0; i < ql; ++i) at(v.query_pos + i) = 'D';
        } else {
            const bool same = std::all_of(v.ref_chars.begin(), v.ref_chars.end(),
                                          [&](uint8_t c) { return c == v.ref_chars[0]; });
            const uint8_t fill = same ? v.ref_chars[0] : (uint8_t)'N';
            for (uint64_t i = 0; i < ql; ++i) at(v.query_pos + i) = fill;
        }
    }
}

// ---------------------------------------------------------------------------
// gap_filling.rs:444-526
// ---------------------------------------------------------------------------
double gap_run_log_term(uint64_t m) {  // gap_filling.rs:489-501: ln_1p(-(exp(ln 1 - ln 4)).powi(m)), m = run + 1 + 1
    return std::log1p(-__builtin_powi(std::exp(std::log(1.0) - std::log(4.0)), (int)m));
}

// The gaps are found on the translation as it comes in: filling one only rewrites positions inside it, which the
// scan never looks at again, so the list of gaps does not depend on the fills and every gap can be bridged
// independently (num_threads > 1: contiguous ranges of the list on host threads; a reference panic is reported for the
// first gap in order, as the sequential loop would).
void fill_gaps(Bytes* translation, const MsArrays& noisy_ms, const uint8_t* ref_seq, uint64_t len,
               const HostIndex& ix, uint64_t thr, double max_err_prob, uint32_t num_threads) {
    Bytes& a = *translation;
    const uint64_t n = a.size();
    require(n > 0 && n == noisy_ms.n, "gap_filling.rs:453-454");
    const uint64_t k = ix.k;
    require(k > 0, "gap_filling.rs:457");
    require(n >= thr, "gap_filling.rs:467 usize underflow");
    const double log_bound = std::log1p(-max_err_prob);
    std::vector<std::pair<uint64_t, uint64_t>> gaps;  // [start, end)
    for (uint64_t i = thr + 1; i < n - thr; ++i) {
        if (a[i - 1] != '-' && a[i - 1] != 'X') continue;
        const uint64_t gs = i - 1;
        while (i < n && a[i] == '-') ++i;
        gaps.emplace_back(gs, std::min(i, n - thr));
    }
    auto bridge = [&](uint64_t gs, uint64_t ge) {
        const uint64_t glen = ge - gs;
        const bool fits_in_kmer = glen + 2 * thr <= k;
        const Bytes kmer = bridge_gap(noisy_ms, ref_seq, len, ix, thr, thr, gs, ge, k - (fits_in_kmer ? thr : 0));
        const bool found = !kmer.empty() && std::find(kmer.begin(), kmer.end(), (uint8_t)'$') == kmer.end();
        const bool no_indels = kmer.size() == thr + glen + thr;
        // agreement of the bridging bases with the reference inside the gap
        std::vector<char> same;
        for (uint64_t t = std::min<uint64_t>(thr, kmer.size()), p = gs;
             t < std::min<uint64_t>(thr + glen, kmer.size()) && p < ge; ++t, ++p)
            same.push_back(kmer[t] == ref_seq[p]);
        uint64_t agree = 0;
        for (char s : same) agree += s;
        double log_probs = 0.0;  // gap_filling.rs:489-501: runs of consecutive agreements, scored when they break
        uint64_t run = 0;
        for (size_t w = 0; w + 1 < same.size(); ++w) {
            if (same[w] && same[w + 1]) {
                ++run;
            } else {
                if (run > 0) log_probs += 1.0 * gap_run_log_term(run + 2);
                run = 0;
            }
        }
        const bool by_overlap = log_probs > log_bound;
        const bool flanked = !same.empty() && !same.front() && !same.back() && agree + 2 == glen;
        if (found && no_indels && (fits_in_kmer || by_overlap || flanked)) {
            for (uint64_t p = gs, t = thr; p < ge; ++p, ++t) a[p] = (kmer[t] == ref_seq[p]) ? (uint8_t)'M' : kmer[t];
        }
    };
    const size_t nt = std::max<size_t>(1, std::min<size_t>(num_threads, gaps.size() / 64 + 1));
    if (nt == 1) {
        for (const auto& g : gaps) bridge(g.first, g.second);
        return;
    }
    std::vector<std::string> panic_what(nt);
    std::vector<char> panicked(nt, 0);
    std::vector<std::thread> workers;
    for (size_t t = 0; t < nt; ++t) {
        workers.emplace_back([&, t]() {
            const size_t lo = gaps.size() * t / nt, hi = gaps.size() * (t + 1) / nt;
            try {
                for (size_t g = lo; g < hi; ++g) bridge(gaps[g].first, gaps[g].second);
            } catch (const RefinePanic& e) {
                panic_what[t] = e.what;  // ranges are in gap order: the lowest t holds the first panic
                panicked[t] = 1;
            }
        });
    }
    for (auto& w : workers) w.join();
    for (size_t t = 0; t < nt; ++t)
        if (panicked[t]) throw RefinePanic{panic_what[t]};
}

}  // namespace kbo_b200
