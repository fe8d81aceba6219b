// ===========================================================================
// kbo_b200/csrc/refine.cuh -- the refinement steps of kbo::map / kbo::call on
// the device:
//
//   gap_list_kernel / fill_gaps_kernel   gap_filling::fill_gaps (reference
//       src/gap_filling.rs:444-526) with bridge_gap (:295-361), unique_context
//       (:127-151), extend_left (:205-232), left / right_overlaps (:20-67);
//   access_kmers_kernel                  SbwtIndex::access_kmer for the nodes of
//       the variant candidates (src/variant_calling.rs:276).
//
// Both need the k-mers of the nodes ("select support", BuildOpts.build_select):
// the GPU builder keeps its colex-sorted node keys on the device (NodeKeysView).
// SbwtIndex::search (gap_filling.rs:217) is a walk over the rank words that K1
// uses, started from the prefix-state table where there is one.
//
// fill_gaps looks for its gaps on the translation as it comes in, and filling a
// gap only rewrites positions inside it, so the list of gaps does not depend on
// the fills and every gap is bridged independently: one THREAD per gap (the work
// per gap is a short, branchy sequence of dependent lookups; there are tens of
// thousands of gaps per assembly).  The host version (refine_host.cpp) remains
// for indexes without node keys; both are compared with the oracle by the tests.
// ===========================================================================
#pragma once
#include "kernels.cuh"

namespace kbo_b200 {

// colex-sorted nodes: 2 bits per base, LAST base in the most significant bits (index_build.cuh); words == 2:
// little-endian (lo, hi) pairs of an unsigned __int128
struct NodeKeysView {
    const uint64_t* keys = nullptr;
    const uint8_t* len = nullptr;  // number of non-'$' bases of the node
    uint32_t words = 0;            // 0: the index carries no node keys
};

// a node's k-mer as a 2-bit window: string index j (0 = first character) at bits 2j of (lo, hi); the first
// k - len characters are '$' (zero bits in the window)
struct KmerWin {
    uint64_t lo, hi;
    uint32_t len;
};

__device__ __forceinline__ KmerWin load_kmer(const NodeKeysView& nk, uint32_t k, uint64_t node) {
    KmerWin w;
    w.len = nk.len[node];
    if (nk.words == 1) {
        w.lo = nk.keys[node] >> (64 - 2 * k);  // 1 <= k <= 32
        w.hi = 0;
    } else {
        const uint64_t lo = nk.keys[2 * node], hi = nk.keys[2 * node + 1];
        const uint32_t sh = 128 - 2 * k;  // 32 < k <= 64: 0 <= sh < 64
        w.lo = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        w.hi = hi >> sh;
    }
    return w;
}
__device__ __forceinline__ uint32_t kmer_code(const KmerWin& w, uint32_t j) {
    return (uint32_t)(j < 32 ? w.lo >> (2 * j) : w.hi >> (2 * (j - 32))) & 3u;
}
__device__ __forceinline__ uint8_t code_char(uint32_t c) { return (uint8_t)((0x54474341u >> (8 * c)) & 0xffu); }  // "ACGT"
__device__ __forceinline__ uint8_t kmer_char(const KmerWin& w, uint32_t k, uint32_t j) {
    return j < k - w.len ? (uint8_t)'$' : code_char(kmer_code(w, j));
}

// SbwtIndex::access_kmer for a list of nodes: k ASCII bytes each ('$' padded at the start)
__global__ void access_kmers_kernel(NodeKeysView nk, uint32_t k, const uint32_t* __restrict__ nodes, uint64_t n_nodes,
                                    uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const KmerWin w = load_kmer(nk, k, nodes[i]);
    for (uint32_t j = 0; j < k; ++j) out[i * k + j] = kmer_char(w, k, j);
}

// one extend_right through the rank words; false (state untouched) when the interval becomes empty
__device__ __forceinline__ bool rank_step(const IndexView& ix, uint32_t c, uint32_t& l, uint32_t& r) {
    const uint32_t rowoff = c * ix.rank_stride;
    const uint64_t wl = __ldg(ix.rank + (rowoff + (l >> 5))), wr = __ldg(ix.rank + (rowoff + (r >> 5)));
    const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
    const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
    if (nl >= nr) return false;
    l = nl;
    r = nr;
    return true;
}
// SbwtIndex::search of the first m characters of a window without '$': extend_right folded over them from [0, n).
// The state after pref_len characters is the table entry when all of them extended (its depth is pref_len).
__device__ __forceinline__ bool search_window(const IndexView& ix, uint64_t lo, uint64_t hi, uint32_t m, uint32_t& l,
                                              uint32_t& r) {
    l = 0;
    r = ix.n;
    uint32_t j = 0;
    const uint32_t P = ix.pref ? ix.pref_len : 0u;
    if (P && m >= P) {
        uint32_t d = 0;
        if (pref_decode(__ldg(ix.pref + ((uint32_t)lo & ((1u << (2 * P)) - 1u))), ix.n, l, r, d)) {
            if (d != P) return false;  // some base of the first P did not extend
            j = P;
        } else {
            l = 0;
            r = ix.n;
        }
    }
    for (; j < m; ++j) {
        const uint32_t c = (uint32_t)(j < 32 ? lo >> (2 * j) : hi >> (2 * (j - 32))) & 3u;
        if (!rank_step(ix, c, l, r)) return false;
    }
    return true;
}

// ---- fill_gaps -------------------------------------------------------------------------------------------------
// The scan of gap_filling.rs:458-470 visits i = thr+1 .. n-thr-1 and opens a gap at i-1 when that character is '-'
// or 'X'; the run of '-' that follows is skipped.  Hence position p in [thr, n-thr-2] starts a gap iff it holds 'X',
// or it holds '-' and is not inside the run of an earlier start (p == thr, or the character before it is neither '-'
// nor 'X').  The gap ends at the first character after p that is not '-' (at most n - thr).
__global__ void gap_list_kernel(const uint8_t* __restrict__ aln, uint64_t n, uint32_t thr, uint2* __restrict__ gaps,
                                uint32_t cap, unsigned int* __restrict__ n_gaps) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + thr;
    if (p + 1 + thr >= n) return;
    const uint8_t c = aln[p];
    if (c != '-' && c != 'X') return;
    if (c == '-' && p != thr) {
        const uint8_t prev = aln[p - 1];
        if (prev == '-' || prev == 'X') return;
    }
    uint64_t e = p + 1;
    while (e < n && aln[e] == '-') ++e;
    if (e > n - thr) e = n - thr;
    const unsigned int i = atomicAdd(n_gaps, 1u);
    if (i < cap) gaps[i] = make_uint2((uint32_t)p, (uint32_t)e);
}

// where the reference would panic (reported for the first gap in sequence order, like the sequential loop)
enum RefinePanicCode {
    RP_NONE = 0,
    RP_BRIDGE_ARGS,      // gap_filling.rs:305-310
    RP_RIGHT_ARGS,       // :25-27
    RP_RIGHT_OOB,        // :33
    RP_LEFT_ARGS,        // :50-52
    RP_LEFT_OOB,         // :58
    RP_TRIM_UNDERFLOW,   // :335
    RP_TRIM_RANGE,       // :336
    RP_IDX_UNDERFLOW,    // :357
    RP_CODES
};

struct FillGapsParams {
    IndexView ix;
    NodeKeysView nk;
    const uint32_t* l;  // index::query_sbwt intervals of ref_seq against the index (K1 with INTERVALS)
    const uint32_t* r;
    const uint8_t* ref;  // ASCII
    uint8_t* aln;        // translation, edited in place
    uint64_t n;
    uint32_t thr;
    const double* run_terms;  // run_terms[m] = log1p(-(1/4)^m), computed on the host exactly as gap_filling.rs:489-501 does
    uint32_t n_terms;         // beyond the table the power underflows to 0 and the term is -0.0
    double log_bound;         // ln(1 - max_err_prob)
    const uint2* gaps;
    uint32_t n_gaps;
    uint8_t* arena;  // extension characters of extend_left: (thr + gap length) bytes per gap that needs them
    unsigned long long* arena_used;
    unsigned long long* panic;  // min over panicking gaps of (gap start << 8 | code); ~0 = none
};

// the bridging string: `n_ext` prepended bases (codes, ext[0] = first character) followed by the node's k-mer
struct Bridge {
    KmerWin node;
    const uint8_t* ext;
    uint64_t n_ext;
};
__device__ __forceinline__ uint8_t bridge_char(const Bridge& b, uint32_t k, uint64_t idx) {
    return idx < b.n_ext ? code_char(b.ext[idx]) : kmer_char(b.node, k, (uint32_t)(idx - b.n_ext));
}

// gap_filling.rs:45-67 on the bridging string (size = n_ext + k)
__device__ __forceinline__ int left_overlaps_dev(const Bridge& b, uint32_t k, const uint8_t* __restrict__ ref, uint64_t ref_len,
                                                 uint64_t ref_match_start, uint64_t* out) {
    const uint64_t size = b.n_ext + k;
    if (!(size > 0 && ref_len > 0 && ref_len > ref_match_start)) return RP_LEFT_ARGS;
    uint64_t n = 0;
    for (uint64_t kp = 0, rp = ref_match_start; kp < size; ++kp, ++rp) {
        if (rp >= ref_len) return RP_LEFT_OOB;
        if (ref[rp] != bridge_char(b, k, kp)) break;
        ++n;
    }
    *out = n;
    return RP_NONE;
}
// gap_filling.rs:20-43 on a node's k-mer (the first base is never compared)
__device__ __forceinline__ int right_overlaps_dev(const KmerWin& w, uint32_t k, const uint8_t* __restrict__ ref, uint64_t ref_len,
                                                  uint64_t ref_match_end, uint64_t* out) {
    if (!(k > 0 && ref_len > 0 && ref_len >= ref_match_end)) return RP_RIGHT_ARGS;
    uint64_t n = 0;
    for (uint64_t kp = k - 1, rp = ref_match_end - 1; kp > 0; --kp, --rp) {
        if (rp >= ref_len) return RP_RIGHT_OOB;  // (also the wrap below zero)
        if (ref[rp] != kmer_char(w, k, (uint32_t)kp)) break;
        ++n;
    }
    *out = n;
    return RP_NONE;
}

// gap_filling.rs:205-232: prepend bases while exactly one base extends the first k-1 characters to a unique node.
// One search of those k-1 characters answers the reference's four k-length searches (refine_host.cpp extend_left):
// the candidates are the full nodes of its interval.  The codes are stored from the END of `store` (capacity
// max_extension) backwards, so that the result is contiguous; returns their number.
__device__ __forceinline__ uint64_t extend_left_dev(const IndexView& ix, const NodeKeysView& nk, const KmerWin& start,
                                                    uint64_t max_extension, uint8_t* store) {
    const uint32_t k = ix.k;
    if (start.len != k) return 0;  // a '$' among the first k-1 characters: nothing matches
    uint64_t lo = start.lo, hi = start.hi, n_ext = 0;
    while (n_ext < max_extension) {
        uint32_t l, r;
        if (!search_window(ix, lo, hi, k - 1, l, r)) break;
        uint32_t hits = 0, hit_code = 0;
        for (uint32_t v = l; v < r; ++v) {
            if (nk.len[v] != k) continue;
            ++hits;
            hit_code = kmer_code(load_kmer(nk, k, v), 0);
        }
        if (hits != 1) break;
        hi = (hi << 2) | (lo >> 62);
        lo = (lo << 2) | hit_code;
        ++n_ext;
        store[max_extension - n_ext] = (uint8_t)hit_code;
    }
    return n_ext;
}

// gap_filling.rs:295-361.  On success *b / [*va, *vb) describe the returned (trimmed) string; *vb == *va: nothing found.
__device__ __forceinline__ int bridge_gap_dev(const FillGapsParams& p, uint64_t gap_start, uint64_t gap_end,
                                              uint64_t search_radius, Bridge* b, uint64_t* va, uint64_t* vb) {
    const uint64_t k = p.ix.k, ref_len = p.n, left_req = p.thr, right_req = p.thr;
    *va = *vb = 0;
    if (!(k > 0 && left_req <= gap_start)) return RP_BRIDGE_ARGS;
    if (!(gap_end <= ref_len && right_req <= ref_len - gap_end)) return RP_BRIDGE_ARGS;
    if (!(gap_end > gap_start && gap_end < ref_len)) return RP_BRIDGE_ARGS;
    const uint64_t search_start = gap_end + search_radius < ref_len - 1 ? gap_end + search_radius : ref_len - 1;
    const uint64_t search_end = gap_end + right_req;
    const uint64_t gap_len = gap_end - gap_start;
    const uint64_t ref_start = gap_start > left_req ? gap_start - left_req : 0;
    uint8_t* store = nullptr;  // extension characters of this gap (claimed at the first need)
    uint64_t idx = search_start;
    while (idx >= search_end) {
        // unique_context (gap_filling.rs:127-151): the nearest position at or below idx with a single-node interval
        bool have = false;
        while (idx >= search_end) {
            if (p.r[idx] - p.l[idx] == 1u) { have = true; break; }
            --idx;
        }
        if (have) {
            b->node = load_kmer(p.nk, (uint32_t)k, p.l[idx]);
            b->n_ext = 0;
            b->ext = nullptr;
            const uint64_t right_want = idx + 1 - gap_end;
            uint64_t right_got = 0, left_got = 0;
            int rc = right_overlaps_dev(b->node, (uint32_t)k, p.ref, ref_len, gap_end + right_want, &right_got);
            if (rc) return rc;
            rc = left_overlaps_dev(*b, (uint32_t)k, p.ref, ref_len, ref_start, &left_got);
            if (rc) return rc;
            const bool right_ok = right_got >= (right_want < k ? right_want : k);
            uint64_t size = k, a = 0;
            bool accept = false;
            if (right_ok && left_got >= left_req) {
                accept = true;
                a = left_got - left_req;
            } else if (right_ok && left_got < left_req && k < left_req + gap_len + right_got) {
                const uint64_t max_ext = left_req + gap_len + right_got - k;  // < left_req + gap_len
                if (!store) store = p.arena + atomicAdd(p.arena_used, (unsigned long long)(left_req + gap_len));
                b->n_ext = extend_left_dev(p.ix, p.nk, b->node, max_ext, store);
                b->ext = store + (max_ext - b->n_ext);
                uint64_t lg = 0;
                rc = left_overlaps_dev(*b, (uint32_t)k, p.ref, ref_len, ref_start, &lg);
                if (rc) return rc;
                if (lg >= left_req) {
                    accept = true;
                    a = lg - left_req;
                    size = k + b->n_ext;
                }
            }
            if (accept) {  // trim (gap_filling.rs:335-336)
                if (!(right_got >= right_req && size >= right_got - right_req)) return RP_TRIM_UNDERFLOW;
                const uint64_t e = size - (right_got - right_req);
                if (!(a <= e)) return RP_TRIM_RANGE;
                *va = a;
                *vb = e;
                return RP_NONE;
            }
        }
        if (idx < 1) return RP_IDX_UNDERFLOW;
        --idx;
    }
    return RP_NONE;
}

// one thread per gap (gap_filling.rs:472-523)
__global__ void fill_gaps_kernel(FillGapsParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_gaps; g += stride) {
        const uint64_t gs = p.gaps[g].x, ge = p.gaps[g].y;
        const uint64_t glen = ge - gs, thr = p.thr, k = p.ix.k;
        const bool fits_in_kmer = glen + 2 * thr <= k;
        Bridge b;
        uint64_t va = 0, vb = 0;
        const int rc = bridge_gap_dev(p, gs, ge, k - (fits_in_kmer ? thr : 0), &b, &va, &vb);
        if (rc) {
            atomicMin(p.panic, (unsigned long long)((gs << 8) | (uint64_t)rc));
            continue;
        }
        const uint64_t size = vb - va;
        if (size == 0 || size != thr + glen + thr) continue;  // nothing found, or an indel (gap_filling.rs:476-478)
        // a '$' inside the returned string?  (the node's first k - len characters, at string index n_ext + j)
        const uint64_t d0 = b.n_ext > va ? b.n_ext : va;
        const uint64_t d1 = b.n_ext + (k - b.node.len) < vb ? b.n_ext + (k - b.node.len) : vb;
        if (d0 < d1) continue;
        // agreement of the bridging bases with the reference inside the gap (gap_filling.rs:480-501)
        uint64_t agree = 0, run = 0;
        double log_probs = 0.0;
        bool first_same = false, prev_same = false;
        for (uint64_t t = 0; t < glen; ++t) {
            const bool same = bridge_char(b, (uint32_t)k, va + thr + t) == p.ref[gs + t];
            agree += same;
            if (t == 0) {
                first_same = same;
            } else if (prev_same && same) {
                ++run;
            } else {
                if (run > 0) {
                    const uint64_t m = run + 2;
                    log_probs += 1.0 * (m < p.n_terms ? p.run_terms[m] : -0.0);
                }
                run = 0;
            }
            prev_same = same;
        }
        const bool by_overlap = log_probs > p.log_bound;
        const bool flanked = glen > 0 && !first_same && !prev_same && agree + 2 == glen;
        if (!(fits_in_kmer || by_overlap || flanked)) continue;
        for (uint64_t t = 0; t < glen; ++t) {
            const uint8_t c = bridge_char(b, (uint32_t)k, va + thr + t);
            p.aln[gs + t] = c == p.ref[gs + t] ? (uint8_t)'M' : c;
        }
    }
}

}  // namespace kbo_b200
