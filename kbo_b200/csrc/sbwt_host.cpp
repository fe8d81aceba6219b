// ===========================================================================
// kbo_b200/csrc/sbwt_host.cpp -- host-side SBWT + LCS construction and lookups.
// See sbwt_host.hpp.  Semantics follow SURVEY.md section 8c (sbwt 0.3.4 is not
// vendored under the reference; its behaviour is pinned by the reference's
// tests, which tests/ replays through the C ABI).
//
// Construction outline (k-mers packed 2 bits per base, LAST base in the most
// significant bits so that integer order == colexicographic order):
//   1. slide over every maximal ACGT run, emit packed k-mers (+ reverse
//      complements), sort (optionally multi-threaded) and deduplicate -> R;
//   2. k-mers whose (k-1)-prefix is no other k-mer's (k-1)-suffix get the chain
//      of '$'-padded prefixes ("dummy" nodes); with the all-'$' root this gives
//      the padded set P, kept as (key, len) pairs ordered by (key, len);
//   3. LCS[i] from the common leading bits of neighbouring keys;
//   4. outgoing labels by ONE merge pass: within the block of nodes ending in
//      c, nodes are ordered by their first k-1 characters, i.e. in the same
//      order as the source groups that reach them, so four cursors suffice.
// ===========================================================================
#include "sbwt_host.hpp"

#include <algorithm>
#include <thread>

namespace kbo_b200 {
namespace {

typedef unsigned __int128 u128;

template <typename K> struct KeyTraits;
template <> struct KeyTraits<uint64_t> {
    static constexpr int BITS = 64;
    static int clz(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
};
template <> struct KeyTraits<u128> {
    static constexpr int BITS = 128;
    static int clz(u128 x) {
        uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
        return hi ? __builtin_clzll(hi) : (lo ? 64 + __builtin_clzll(lo) : 128);
    }
};

inline int base_code(uint8_t ch) {
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

template <typename K>
void parallel_sort(std::vector<K>& v, uint32_t threads) {
    if (threads <= 1 || v.size() < (1u << 16)) {
        std::sort(v.begin(), v.end());
        return;
    }
    const size_t parts = threads;
    std::vector<size_t> cut(parts + 1);
    for (size_t i = 0; i <= parts; ++i) cut[i] = v.size() * i / parts;
    std::vector<std::thread> pool;
    for (size_t i = 0; i < parts; ++i)
        pool.emplace_back([&v, &cut, i]() { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1]); });
    for (auto& t : pool) t.join();
    for (size_t width = 1; width < parts; width *= 2) {
        std::vector<std::thread> mpool;
        for (size_t i = 0; i + width < parts; i += 2 * width) {
            size_t a = cut[i], b = cut[i + width], c = cut[std::min(i + 2 * width, parts)];
            mpool.emplace_back([&v, a, b, c]() { std::inplace_merge(v.begin() + a, v.begin() + b, v.begin() + c); });
        }
        for (auto& t : mpool) t.join();
    }
}

template <typename K>
struct PNode {
    K key;
    uint8_t len;  // number of real (non-'$') characters
};

template <typename K>
std::string build_typed(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k,
                        bool add_revcomp, uint32_t num_threads, bool keep_nodes, HostIndex* out) {
    constexpr int BITS = KeyTraits<K>::BITS;
    const int top = BITS - 2;                 // bit offset of the last character
    const int first_shift = BITS - 2 * (int)k;  // bit offset of the first character
    const K ones = ~(K)0;
    const K kmask = (2 * k >= (uint32_t)BITS) ? ones : (K)(ones << (BITS - 2 * k));
    const K sufmask = (k == 1) ? (K)0 : (K)(ones << (BITS - 2 * (k - 1)));

    // ---- 1. k-mers ---------------------------------------------------------
    std::vector<K> R;
    {
        uint64_t total = 0;
        for (uint64_t s = 0; s < n_seqs; ++s) total += lens[s];
        R.reserve((size_t)(add_revcomp ? 2 * total : total));
    }
    for (uint64_t s = 0; s < n_seqs; ++s) {
        K fwd = 0, rev = 0;
        uint32_t run = 0;
        const uint8_t* p = seqs[s];
        for (uint64_t i = 0; i < lens[s]; ++i) {
            int c = base_code(p[i]);
            if (c < 0) {
                run = 0;
                continue;
            }
            fwd = (K)((fwd >> 2) | ((K)c << top));
            rev = (K)((rev << 2) | ((K)(3 - c) << first_shift));
            if (run < k) ++run;
            if (run == k) {
                R.push_back(fwd & kmask);
                if (add_revcomp) R.push_back(rev & kmask);
            }
        }
    }
    parallel_sort(R, num_threads);
    R.erase(std::unique(R.begin(), R.end()), R.end());
    out->n_kmers = R.size();

    // ---- 2. dummy nodes ------------------------------------------------------
    std::vector<PNode<K>> dummy;
    dummy.push_back(PNode<K>{0, 0});
    if (k >= 2) {
        for (size_t i = 0; i < R.size(); ++i) {
            const K want = (K)(R[i] << 2);  // first k-1 characters, aligned like a (k-1)-suffix
            auto it = std::lower_bound(R.begin(), R.end(), want);
            const bool has_pred = it != R.end() && ((*it & sufmask) == want);
            if (has_pred) continue;
            for (uint32_t j = 1; j < k; ++j) dummy.push_back(PNode<K>{(K)(R[i] << (2 * (k - j))), (uint8_t)j});
        }
    }
    auto less = [](const PNode<K>& a, const PNode<K>& b) { return a.key != b.key ? a.key < b.key : a.len < b.len; };
    std::sort(dummy.begin(), dummy.end(), less);
    dummy.erase(std::unique(dummy.begin(), dummy.end(),
                            [](const PNode<K>& a, const PNode<K>& b) { return a.key == b.key && a.len == b.len; }),
                dummy.end());

    const uint64_t n = R.size() + dummy.size();
    if (n >= (1ull << 32) - 64) return "index too large: n_sets must be < 2^32";
    std::vector<PNode<K>> P;
    P.resize((size_t)n);
    {
        size_t a = 0, b = 0, o = 0;
        while (a < R.size() && b < dummy.size()) {
            PNode<K> ra{R[a], (uint8_t)k};
            if (less(ra, dummy[b])) { P[o++] = ra; ++a; } else { P[o++] = dummy[b++]; }
        }
        while (a < R.size()) P[o++] = PNode<K>{R[a++], (uint8_t)k};
        while (b < dummy.size()) P[o++] = dummy[b++];
    }
    std::vector<K>().swap(R);
    out->k = k;
    out->n_sets = n;

    // ---- 3. LCS ------------------------------------------------------------------
    out->lcs.assign((size_t)n, 0);
    for (uint64_t i = 1; i < n; ++i) {
        int same = KeyTraits<K>::clz((K)(P[i - 1].key ^ P[i].key)) >> 1;
        int lim = std::min<int>(P[i - 1].len, P[i].len);
        out->lcs[i] = (uint8_t)std::min(same, lim);
    }

    // ---- 4. labels by merging -------------------------------------------------------
    const size_t nwords = (size_t)(n + 63) / 64 + 1;
    for (int c = 0; c < 4; ++c) out->rows[c].assign(nwords, 0);
    // block of nodes whose last character is c: [sec[c], sec[c+1])
    uint64_t sec[5];
    sec[0] = 1;  // node 0 is the root '$'^k; every other node ends in a letter
    for (int c = 1; c < 4; ++c) {
        PNode<K> probe{(K)((K)c << top), 0};
        sec[c] = (uint64_t)(std::lower_bound(P.begin() + 1, P.end(), probe, less) - P.begin());
    }
    sec[4] = n;
    uint64_t cur[4] = {sec[0], sec[1], sec[2], sec[3]};
    for (uint64_t i = 0; i < n; ++i) {
        const bool group_first = (i == 0) || (out->lcs[i] + 1u < k);
        if (!group_first) continue;
        const K ukey = P[i].key & sufmask;                       // last k-1 characters of the source
        const uint32_t ulen = std::min<uint32_t>(P[i].len, k - 1);
        for (int c = 0; c < 4; ++c) {
            if (cur[c] >= sec[c + 1]) continue;
            const PNode<K>& t = P[cur[c]];
            // first k-1 characters of the target, aligned like a (k-1)-suffix
            if ((K)(t.key << 2) == ukey && (uint32_t)t.len - 1u == ulen) {
                out->rows[c][i >> 6] |= 1ull << (i & 63);
                ++cur[c];
            }
        }
    }
    for (int c = 0; c < 4; ++c)
        if (cur[c] != sec[c + 1]) return "internal error: label merge did not consume every node";
    if (keep_nodes) {
        out->node_hi.resize((size_t)n);
        out->node_len.resize((size_t)n);
        if (BITS == 128) out->node_lo.resize((size_t)n);
        for (uint64_t i = 0; i < n; ++i) {
            if (BITS == 128) {
                out->node_hi[i] = (uint64_t)((u128)P[i].key >> 64);
                out->node_lo[i] = (uint64_t)P[i].key;
            } else {
                out->node_hi[i] = (uint64_t)P[i].key;
            }
            out->node_len[i] = P[i].len;
        }
    }
    out->finalize();
    return std::string();
}

}  // namespace

void HostIndex::finalize() {
    const size_t nwords = rows[0].size();
    uint64_t acc = 1;
    for (int c = 0; c < 4; ++c) {
        row_cum[c].assign(nwords + 1, 0);
        uint32_t s = 0;
        for (size_t w = 0; w < nwords; ++w) {
            row_cum[c][w] = s;
            s += (uint32_t)__builtin_popcountll(rows[c][w]);
        }
        row_cum[c][nwords] = s;
        C[c] = acc;
        acc += s;
    }
}

uint64_t HostIndex::rank(int c, uint64_t p) const {
    const uint64_t w = p >> 6;
    uint64_t res = row_cum[c][w];
    if (p & 63) res += (uint64_t)__builtin_popcountll(rows[c][w] & (~0ull >> (64 - (p & 63))));
    return res;
}

uint64_t HostIndex::select(int c, uint64_t j) const {
    // largest word w with row_cum[w] <= j
    const std::vector<uint32_t>& cum = row_cum[c];
    size_t lo = 0, hi = cum.size() - 1;
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (cum[mid] <= j) lo = mid; else hi = mid;
    }
    uint64_t word = rows[c][lo];
    uint64_t need = j - cum[lo];
    for (uint64_t t = 0; t < need; ++t) word &= word - 1;
    return (uint64_t)lo * 64 + (uint64_t)__builtin_ctzll(word);
}

bool HostIndex::extend_right(uint64_t& l, uint64_t& r, uint8_t ch) const {
    int c = base_code(ch);
    if (c < 0) return false;
    uint64_t nl = C[c] + rank(c, l), nr = C[c] + rank(c, r);
    if (nl >= nr) return false;
    l = nl;
    r = nr;
    return true;
}

bool HostIndex::search(const uint8_t* pat, uint64_t len, uint64_t* l, uint64_t* r) const {
    uint64_t a = 0, b = n_sets;
    for (uint64_t i = 0; i < len; ++i)
        if (!extend_right(a, b, pat[i])) return false;
    *l = a;
    *r = b;
    return true;
}

void HostIndex::access_kmer(uint64_t colex, uint8_t* out_k) const {
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    if (!node_len.empty()) {  // decode the stored node
        const uint64_t hi = node_hi[colex], lo = node_lo.empty() ? 0 : node_lo[colex];
        const uint32_t len = node_len[colex];
        for (uint32_t t = 0; t < k; ++t) {
            uint8_t ch = '$';
            if (t < len) {
                const uint32_t code = t < 32 ? (uint32_t)(hi >> (62 - 2 * t)) & 3u : (uint32_t)(lo >> (62 - 2 * (t - 32))) & 3u;
                ch = (uint8_t)letters[code];
            }
            out_k[k - 1 - t] = ch;
        }
        return;
    }
    // no stored nodes: walk the incoming edges back k times
    uint64_t node = colex;
    for (uint32_t t = 0; t < k; ++t) {
        uint8_t ch = '$';
        if (node != 0) {
            int c = 3;
            while (c > 0 && C[c] > node) --c;
            ch = (uint8_t)letters[c];
            node = select(c, node - C[c]);  // the group-first source of this node's incoming edge
        }
        out_k[k - 1 - t] = ch;
    }
}

std::string build_host_index(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k,
                             bool add_revcomp, uint32_t num_threads, HostIndex* out, bool keep_nodes) {
    if (k <= 32) return build_typed<uint64_t>(seqs, lens, n_seqs, k, add_revcomp, num_threads, keep_nodes, out);
    return build_typed<u128>(seqs, lens, n_seqs, k, add_revcomp, num_threads, keep_nodes, out);
}

}  // namespace kbo_b200
