// ===========================================================================
// kbo_b200/csrc/kernels.cuh -- hand-written sm_100a kernels of the kbo hot path
//
//   K0 pack_queries_kernel      ASCII CSR batch -> 2-bit packed "padded space"
//   K1 ms_kernel                k-bounded matching statistics (index::query_sbwt,
//                               reference src/index.rs:243-256 -> sbwt StreamingIndex)
//   K2 derand_translate_kernel  derandomize_ms_vec + translate_ms_vec fused
//                               (src/derandomize.rs:269-288, src/translate.rs:263-293)
//   K3 translate_i64_kernel     translate_ms_vec on an arbitrary i64 vector
//   G* derandomize_general_*    derandomize_ms_vec on an arbitrary MS vector
//
// "Padded space": all queries of a batch concatenated with ONE separator
// position after each query.  A separator is a non-ACGT symbol, which resets
// the MS state to (0,[0,n)) exactly like starting a new query does, so K1 can
// cut the padded sequence into uniform chunks without knowing query borders.
//
// The file also compiles under tests/emu/host_emu.hpp (KBO_HOST_EMU) so that
// the kernel logic can be checked against the oracle on a CPU-only box; that
// path is test infrastructure and is never part of the shipped library.
// ===========================================================================
#pragma once
#include <stdint.h>

#ifdef KBO_HOST_EMU
#include "host_emu.hpp"
#else
#include <cuda_runtime.h>
#endif

namespace kbo_b200 {

// ---------------------------------------------------------------------------
// Device-resident index (DESIGN.md "Index layout").
//   rank : 4 rows (A,C,G,T) of `rank_stride` 64-bit words.  Word b of row c
//          covers subset-matrix positions [32b, 32b+32):
//            low  32 bits = the 32 bits of row c,
//            high 32 bits = C[c] + number of set bits of row c before 32b,
//          so extend_right needs ONE 8-byte load per interval end and a popc.
//          Four consecutive words (128 positions) share a 32-byte L2 sector.
//   lcs  : one byte per node, zero padded past n (sentinel for the right scan).
//   links: one 32-bit word per node (and one for n): bits 0-6 LCS[q], bit 7 LINK_SLOW, bits 8-19 q - PSV(q), bits 20-31 NSV(q) - q,
//          PSV / NSV = nearest position to the left / right whose LCS is smaller; 4095 = farther than that (or none).
//          contract_left to the first depth that changes the interval is then two loads and a few additions.
//   rank2: see IndexView::rank2 (DESIGN.md "two bases per probe").
//   pref : for k >= PREF_MIN_K, the MS state after any string of pref_len bases fed to the empty state (index: first
//          base in the low bits; 8 bytes per entry, pref_encode).  A chunk's warm-up starts from its entry instead of
//          stepping through its first pref_len bases.
// ---------------------------------------------------------------------------
struct IndexView {
    const uint64_t* rank;
    const uint64_t* rank2;  // optional (== rank + 4 * rank_stride when present: the rows follow rank's in one allocation):
                            // 16 rows (first base | second base << 2) in the same word format, for TWO bases per
                            // probe: word b of row (a, c) = (C[c] + rank_c(C[a]) + #{i < 32b : node i has label a and its
                            // a-successor has label c}) << 32 | those 32 bits, so that
                            // extend_right(extend_right([l, r), a), c) = [rank2(l), rank2(r)) with one load per end
    uint32_t rank_stride;  // words per row (< 2^28 for n_sets < 2^32)
    const uint8_t* lcs;
    const uint32_t* links;  // n + 1 entries: LCS | distance to the previous smaller LCS << 8 | to the next smaller << 20
    const uint64_t* pref;   // optional: MS state after pref_len bases fed to the empty state, 4^pref_len entries (pref_encode)
    uint32_t pref_len;
    uint32_t n;  // n_sets
    uint32_t k;
};

struct QueryView {
    const uint64_t* pack;  // 32 bases per word, 2 bits each (A0 C1 G2 T3), base i of the word at bits [2i,2i+2)
    const uint32_t* inv;   // bit i: base is not ACGT (includes separators and the tail past Lp)
    const uint32_t* sep;   // bit i: separator or past Lp
    const uint32_t* wq;    // number of separators before the first base of the word
    uint64_t Lp;           // padded length = sum(len) + n_queries
    uint64_t n_words;      // words filled by K0 (covers Lp rounded up to a K2 tile, plus slack)
};

// counters 0-5 cover all work incl. chunk warm-up; 6-9 only events of emitted positions (the algorithmic figure)
enum { CNT_ATTEMPTS = 0, CNT_SPLIT = 1, CNT_CONTRACT = 2, CNT_EXTRA_LCS = 3, CNT_PROCESSED = 4, CNT_EMITTED = 5,
       CNT_ATT_EMIT = 6, CNT_SPLIT_EMIT = 7, CNT_CON_EMIT = 8, CNT_EXTRA_EMIT = 9, CNT_N = 10 };

// ---------------------------------------------------------------------------
// K0: pack.  One thread per 32 padded positions (a warp covers 1024).
//   * the number of separators before the warp's first position is found by ONE search per warp, 32 probes per
//     step; a lane then walks forward from there to its own word (a few steps unless queries are tiny);
//   * the word's source bytes are contiguous in the CSR buffer even across query borders (a separator occupies a
//     padded position but no source byte), so 32 of them are read with aligned 8-byte loads, shifted together,
//     and converted four per register with byte-parallel arithmetic;
//   * each separator inside the word is then inserted by shifting the upper part of the masks up by one position.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t sep_pos(const uint64_t* offsets, uint64_t q) {
    // padded position of the separator that follows query q
    return offsets[q + 1] - offsets[0] + q;
}

// Four ASCII bytes -> 8 bits of 2-bit codes (base i at bits [2i,2i+2), 0 where the byte is not ACGT)
// and 4 bits of "not ACGT" flags.
__device__ __forceinline__ void swar_pack4(uint32_t v, uint32_t& code8, uint32_t& inv4) {
    const uint32_t x = (v >> 1) & 0x03030303u;           // A0 C1 T2 G3
    const uint32_t x1 = (x >> 1) & 0x01010101u;
    uint32_t code = x ^ x1;                              // A0 C1 G2 T3
    const uint32_t is_t = x1 & ~x;                       // 1 in the bytes with x == 2
    const uint32_t expect = is_t * 0x0fu + 0x41414141u;  // the byte without bits 1,2: 0x41 (A C G) or 0x50 (T)
    const uint32_t diff = (v & 0xf9f9f9f9u) ^ expect;    // non-zero byte <=> not ACGT
    const uint32_t nz = (((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu) | diff) & 0x80808080u;
    code &= ~((nz >> 6) | (nz >> 7));
    code8 = (code * 0x01041040u) >> 24;                  // bits 8i+{0,1} -> 24+2i+{0,1}, no two terms collide
    inv4 = (nz * 0x00204081u) >> 28;                     // bits 8i+7 -> 28+i
}

__global__ void pack_queries_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ offsets,
                                    uint64_t nq, QueryView qv, uint64_t* __restrict__ pack,
                                    uint32_t* __restrict__ inv, uint32_t* __restrict__ sep,
                                    uint32_t* __restrict__ wq) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x is a multiple of 32
    const int lane = threadIdx.x & 31;
    const uint64_t pp0 = w * 32;
    const uint64_t off0 = offsets[0];
    // ---- first query whose separator is at or after the warp's first position: 33-ary search, whole warp ----
    uint64_t lo = 0, hi = nq;  // the answer is in [lo, hi]
    {
        const uint64_t warp_pp0 = pp0 - 32ull * lane;
        while (lo < hi) {
            const uint64_t span = hi - lo;
            uint64_t t = lo + (span * (uint64_t)(lane + 1)) / 33;  // increasing in lane, < hi
            const bool before = sep_pos(offsets, t) < warp_pp0;  // monotone: true for a prefix of the lanes
            const uint32_t c = (uint32_t)__popc(__ballot_sync(0xffffffffu, before));
            const uint64_t t_last_true = __shfl_sync(0xffffffffu, t, c ? c - 1 : 0);
            const uint64_t t_first_false = __shfl_sync(0xffffffffu, t, c < 32 ? c : 31);
            if (c) lo = t_last_true + 1;
            if (c < 32) hi = t_first_false;
        }
    }
    if (w >= qv.n_words) return;
    // ---- this word's first query: walk forward, a bounded binary search if that takes long ----------------------
    uint64_t q = lo;
    for (int step = 0; step < 4 && q < nq && sep_pos(offsets, q) < pp0; ++step) ++q;
    if (q < nq && sep_pos(offsets, q) < pp0) {
        uint64_t a = q + 1, b = nq;
        while (a < b) {
            const uint64_t mid = (a + b) >> 1;
            if (sep_pos(offsets, mid) < pp0) a = mid + 1; else b = mid;
        }
        q = a;
    }
    wq[w] = (uint32_t)q;
    uint64_t pk = 0;
    uint32_t iv = 0, sp = 0;
    if (q >= nq) {  // past the last separator: all tail
        iv = sp = ~0u;
    } else {
        // ---- 32 source bytes from off0 + pp0 - q on (no byte outside the batch is touched) ----
        const uint64_t src = off0 + pp0 - q, src_end = offsets[nq];
        const uintptr_t addr = reinterpret_cast<uintptr_t>(ascii) + src;
        const uintptr_t beg_addr = reinterpret_cast<uintptr_t>(ascii) + off0;
        const uintptr_t end_addr = reinterpret_cast<uintptr_t>(ascii) + src_end;
        const uintptr_t a0 = addr & ~(uintptr_t)7;
        const uint32_t sh = 8u * (uint32_t)(addr & 7u);
        uint64_t wv[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const uintptr_t a = a0 + 8 * j;
            if (a >= beg_addr && a + 8 <= end_addr) {
                wv[j] = __ldg(reinterpret_cast<const uint64_t*>(a));
            } else {  // the aligned word sticks out of the batch (its first / last one): byte by byte
                uint64_t v = 0;
                for (int t = 0; t < 8; ++t)
                    if (a + t >= beg_addr && a + t < end_addr) v |= (uint64_t)__ldg(reinterpret_cast<const uint8_t*>(a + t)) << (8 * t);
                wv[j] = v;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t v = sh ? ((wv[j] >> sh) | (wv[j + 1] << (64 - sh))) : wv[j];
            uint32_t c0, c1, i0, i1;
            swar_pack4((uint32_t)v, c0, i0);
            swar_pack4((uint32_t)(v >> 32), c1, i1);
            pk |= (uint64_t)(c0 | (c1 << 8)) << (16 * j);
            iv |= (i0 | (i1 << 4)) << (8 * j);
        }
        // ---- separators inside the word: open one position at each ------------------------------------------
        uint64_t next_sep = sep_pos(offsets, q);
        while (next_sep < pp0 + 32) {
            const uint32_t b = (uint32_t)(next_sep - pp0);
            const uint64_t low2 = (1ull << (2 * b)) - 1ull;
            const uint32_t low1 = (1u << b) - 1u;
            pk = (pk & low2) | ((pk & ~low2) << 2);  // b <= 31: the shifts stay below 64
            pk &= ~(3ull << (2 * b));
            iv = (iv & low1) | ((iv & ~low1) << 1) | (1u << b);
            sp = (sp & low1) | ((sp & ~low1) << 1) | (1u << b);
            ++q;
            if (q >= nq) {  // everything after the last separator is tail
                const uint32_t tail = b == 31 ? 0u : (~0u << (b + 1));
                iv |= tail;
                sp |= tail;
                pk &= b == 31 ? ~0ull : ((1ull << (2 * (b + 1))) - 1ull);
                break;
            }
            next_sep = sep_pos(offsets, q);
        }
    }
    pack[w] = pk;
    inv[w] = iv;
    sep[w] = sp;
}

// ---------------------------------------------------------------------------
// K1: matching statistics.  One LANE per chunk of `chunk_len` padded positions
// (a warp therefore runs 32 independent dependent-load chains; see DESIGN.md
// for why this beats one warp per chunk).  Each loop iteration performs ONE
// extend attempt for the lane's current base: on success (or at d == 0) the
// lane emits and advances, on failure it contracts and retries, so lanes never
// wait for each other's contraction chains.
//
// Exactness: (d_i, I_i) depends only on the k-1 bases before i,
// so a chunk warms up from start-(k-1) with the empty state and emits from
// `start`.  Contraction jumps straight to t = max(LCS[l], LCS[r]): for targets
// in (t, d-1] contract_left returns the same interval, so the reference's
// retries there fail again by construction; the emitted (d, [l,r)) is identical.
// ---------------------------------------------------------------------------
struct MsParams {
    IndexView ix;
    QueryView q;
    uint32_t chunk_len;    // multiple of 32
    uint32_t flags;        // experiment switches (none read by K1 at present; bit1 is host side: K2 instead of K2b)
    uint64_t n_chunks;
    uint8_t* ms;         // padded space, 1 byte per position
    uint32_t* l_out;     // optional (INTERVALS)
    uint32_t* r_out;
    unsigned long long* counters;  // optional (COUNT)
};

__device__ __forceinline__ uint64_t lcs_lt_mask64(uint64_t w, uint64_t t_rep) {
    // 0x80 in every byte of w that is < t (bytes and t are < 128)
    const uint64_t H = 0x8080808080808080ull;
    return ~((w | H) - t_rep) & H;
}

enum { LINK_FAR = 4095, LINK_SCAN_WORDS = 512, LINK_SLOW = 0x80 };  // LCS values are < 128: bit 7 of the value byte is free

// largest q <= from with LCS[q] < t (t >= 1; LCS[0] = 0 ends every scan).  Gives up after max_words 8-byte words
// (returns 0xffffffff); max_words == 0: no limit.
__device__ __noinline__ uint32_t lcs_scan_left(const uint8_t* __restrict__ lcs, uint32_t from, uint32_t t, uint32_t max_words) {
    const uint64_t* __restrict__ L8 = reinterpret_cast<const uint64_t*>(lcs);
    const uint64_t T = (uint64_t)t * 0x0101010101010101ull;
    uint32_t b = from >> 3;
    uint64_t m = lcs_lt_mask64(L8[b], T);
    if ((from & 7) != 7) m &= (1ull << (8 * ((from & 7) + 1))) - 1ull;
    for (uint32_t words = 1; m == 0; ++words) {
        if (max_words && words >= max_words) return 0xffffffffu;
        --b;
        m = lcs_lt_mask64(L8[b], T);
    }
    return (b << 3) + ((63 - __clzll((long long)m)) >> 3);
}
// smallest q >= from with LCS[q] < t (the zero padding at n ends every scan)
__device__ __noinline__ uint32_t lcs_scan_right(const uint8_t* __restrict__ lcs, uint32_t from, uint32_t t, uint32_t max_words) {
    const uint64_t* __restrict__ L8 = reinterpret_cast<const uint64_t*>(lcs);
    const uint64_t T = (uint64_t)t * 0x0101010101010101ull;
    uint32_t b = from >> 3;
    uint64_t m = lcs_lt_mask64(L8[b], T) & (~0ull << (8 * (from & 7)));
    for (uint32_t words = 1; m == 0; ++words) {
        if (max_words && words >= max_words) return 0xffffffffu;
        ++b;
        m = lcs_lt_mask64(L8[b], T);
    }
    return (b << 3) + ((__ffsll((long long)m) - 1) >> 3);
}

// links[q] for q in [0, n]; thread per node.  LCS values are < 128 (k <= 128).
__global__ void lcs_links_kernel(const uint8_t* __restrict__ lcs, uint32_t n, uint32_t* __restrict__ links) {
    const uint64_t q64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q64 > n) return;
    const uint32_t q = (uint32_t)q64;
    const uint32_t v = q < n ? lcs[q] : 0u;
    uint32_t dl = LINK_FAR, dr = LINK_FAR;
    if (v > 0) {  // q >= 1 here because LCS[0] = 0
        const uint32_t a = lcs_scan_left(lcs, q - 1, v, LINK_SCAN_WORDS);
        if (a != 0xffffffffu && q - a < LINK_FAR) dl = q - a;
        const uint32_t b = lcs_scan_right(lcs, q + 1, v, LINK_SCAN_WORDS);
        if (b != 0xffffffffu && b - q < LINK_FAR) dr = b - q;
    }
    // LINK_SLOW: the contraction's common case needs v > 0 and both distances in reach; everything else (a handful of
    // nodes with LCS 0, the shallow nodes whose nearest smaller value is far away) is flagged once here instead of
    // being tested for on every contraction
    const uint32_t slow = (v == 0 || dl == LINK_FAR || dr == LINK_FAR) ? (uint32_t)LINK_SLOW : 0u;
    links[q] = v | slow | (dl << 8) | (dr << 20);
}

// ---- rank2: two bases per probe (IndexView::rank2) -------------------------------------------------------------
// Thread per 32-node word: for every node i of the word with label a, its a-successor s = C[a] + rank_a(i) is read
// off the rank word, and bit i of row (a | c << 2) is set iff s carries label c.  (No warp collectives, so the
// CPU emulation of tests/emu can run it thread by thread.)
__global__ void rank2_bits_kernel(IndexView ix, uint32_t* __restrict__ rows2) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = ix.rank_stride;
    if (w >= stride) return;
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const uint64_t wa = ix.rank[a * stride + w];
        uint32_t bits = (uint32_t)wa;
        while (bits) {
            const uint32_t j = (uint32_t)__ffs((int)bits) - 1u;
            bits &= bits - 1;
            if (w * 32 + j >= ix.n) break;
            const uint32_t s = (uint32_t)(wa >> 32) + __popc((uint32_t)wa & ((1u << j) - 1u));
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if ((ix.rank[c * stride + (s >> 5)] >> (s & 31)) & 1ull) m[a | (c << 2)] |= 1u << j;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) rows2[i * stride + w] = m[i];
}

// word b of row (a | c << 2) = (C[c] + rank_c(C[a]) + ones of the row before the word) << 32 | bits;
// prefix = exclusive sum of the popcounts of the 16 rows laid end to end
__global__ void compose_rank2_kernel(IndexView ix, const uint32_t* __restrict__ rows2, const uint32_t* __restrict__ prefix,
                                     uint64_t* __restrict__ rank2) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = ix.rank_stride;
    if (i >= 16 * stride) return;
    const uint32_t row = (uint32_t)(i / stride), a = row & 3u, c = row >> 2;
    const uint32_t Ca = (uint32_t)(ix.rank[a * stride] >> 32);           // C[a]
    const uint64_t wc = ix.rank[c * stride + (Ca >> 5)];
    const uint32_t base = (uint32_t)(wc >> 32) + __popc((uint32_t)wc & ((1u << (Ca & 31)) - 1u));  // C[c] + rank_c(C[a])
    rank2[i] = ((uint64_t)(base + prefix[i] - prefix[row * stride]) << 32) | rows2[i];
}

// contract_left to the largest depth that changes the interval, t = max(LCS[l], LCS[r]), given the link words of
// l and r.  Returns true when an end was farther than the links reach and had to be found by scanning.
// The rare cases (t == 0, an end beyond the reach of the links, the impossible t > d - 1) are kept out of line.
__device__ __noinline__ uint4 ms_contract_rare(const uint8_t* __restrict__ lcs, uint32_t n, uint32_t el, uint32_t er,
                                               uint32_t l, uint32_t r, uint32_t d) {  // returns (l, r, d, scanned)
    const uint32_t vl = el & 0x7fu, vr = er & 0x7fu;
    uint32_t t = vl > vr ? vl : vr;
    uint32_t scanned = 0;
    if (t == 0) {
        l = 0; r = n; d = 0;
    } else if (t > d - 1) {  // cannot happen for a maximal interval; keeps the literal bound
        t = d - 1;
        d = t;
        if (t == 0) { l = 0; r = n; }
        else { l = lcs_scan_left(lcs, l, t, 0); r = lcs_scan_right(lcs, r, t, 0); }
    } else {
        d = t;
        if (vl == t) {  // the left end moves to the previous position with a smaller LCS
            const uint32_t dl = (el >> 8) & 0xfffu;
            if (dl == LINK_FAR) { l = lcs_scan_left(lcs, l - 1, t, 0); scanned = 1; }
            else l -= dl;
        }
        if (vr == t) {
            const uint32_t dr = er >> 20;
            if (dr == LINK_FAR) { r = lcs_scan_right(lcs, r + 1, t, 0); scanned = 1; }
            else r += dr;
        }
    }
    return make_uint4(l, r, d, scanned);
}
__device__ __forceinline__ bool ms_contract(const IndexView& ix, uint32_t el, uint32_t er, uint32_t& l, uint32_t& r,
                                            uint32_t& d) {
    if ((el | er) & (uint32_t)LINK_SLOW) {  // an end with LCS 0 or with a distance beyond the links' reach
        const uint4 s = ms_contract_rare(ix.lcs, ix.n, el, er, l, r, d);
        l = s.x; r = s.y; d = s.z;
        return s.w != 0;
    }
    // Both values are > 0 and both distances are exact.  t <= d - 1 holds for every state the recurrence produces
    // ([l, r) is the whole colex range of its d-suffix, so LCS[l] < d and LCS[r] < d); the literal bound of
    // contract_left is kept in ms_contract_rare only.
    const uint32_t vl = el & 0xffu, vr = er & 0xffu;
    d = vl > vr ? vl : vr;
    if (vl >= vr) l -= (el >> 8) & 0xfffu;  // the left end moves to the previous position with a smaller LCS
    if (vr >= vl) r += er >> 20;
    return false;
}

// One base through the MS recurrence (extend; on failure at d > 0 contract and retry): what K1's loop does to a
// lane's state between two advances.  Used to tabulate the states after pref_len bases.
enum { PREF_LEN = 10, PREF_MIN_K = 16, PREF_MAX_LEN = 14, PREF_WIDTH_SAT = (1 << 27) - 1 };  // PREF_LEN: the default depth
__device__ __forceinline__ void ms_feed_base(const IndexView& ix, uint32_t c, uint32_t& l, uint32_t& r, uint32_t& d) {
    for (;;) {
        const uint32_t rowoff = c * ix.rank_stride;
        const uint64_t wl = ix.rank[rowoff + (l >> 5)], wr = ix.rank[rowoff + (r >> 5)];
        const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
        const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
        if (nl < nr) {
            l = nl; r = nr;
            d = d + 1 < ix.k ? d + 1 : ix.k;
            return;
        }
        if (d == 0) return;
        ms_contract(ix, ix.links[l], ix.links[r], l, r, d);
    }
}
// A table entry: l in the low word, (r - l) << 5 | d in the high word (d <= pref_len < 32).  d == 0 stands for the
// empty state (0, [0, n)); an interval of 2^27 - 1 or more nodes does not fit and reads back as "no entry" (only
// possible at depths below 3 for n < 2^32: the reader then steps or contracts as if there were no table).
__device__ __forceinline__ uint64_t pref_encode(uint32_t l, uint32_t r, uint32_t d) {
    uint32_t w = r - l;
    if (w > (uint32_t)PREF_WIDTH_SAT) w = (uint32_t)PREF_WIDTH_SAT;
    if (d == 0) { l = 0; w = 0; }
    return (uint64_t)l | ((uint64_t)((w << 5) | d) << 32);
}
__device__ __forceinline__ bool pref_decode(uint64_t e, uint32_t n, uint32_t& l, uint32_t& r, uint32_t& d) {
    const uint32_t hi = (uint32_t)(e >> 32), w = hi >> 5, dd = hi & 31u;
    if (dd == 0) { l = 0; r = n; d = 0; return true; }
    if (w == (uint32_t)PREF_WIDTH_SAT) return false;
    l = (uint32_t)e; r = l + w; d = dd;
    return true;
}
// level j (1 ..): cur[idx] for the 4^j strings of j bases, from prev (level j-1; unused for j == 1)
__global__ void prefix_table_level_kernel(IndexView ix, const uint64_t* __restrict__ prev, uint64_t* __restrict__ cur, uint32_t j) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1u << (2 * j))) return;
    const uint32_t c = idx >> (2 * (j - 1));               // the newest base sits in the highest two bits
    const uint32_t parent = idx & ((1u << (2 * (j - 1))) - 1u);
    uint32_t l = 0, r = ix.n, d = 0;
    if (j > 1 && !pref_decode(prev[parent], ix.n, l, r, d)) {  // the parent's interval did not fit: feed its bases again
        l = 0; r = ix.n; d = 0;
        for (uint32_t t = 0; t + 1 < j; ++t) ms_feed_base(ix, (parent >> (2 * t)) & 3u, l, r, d);
    }
    ms_feed_base(ix, c, l, r, d);
    cur[idx] = pref_encode(l, r, d);
}

// One extend attempt per loop iteration and lane.  A lane whose extension fails (at d > 0) contracts in the same
// iteration -- two loads of `links` and a few additions -- and retries the base in the next one.
// (Round 2 measured two more forms of this loop and dropped them, profiles/README.md "K1, last measurements":
// contractions gated to every second warp iteration -- 117 vs 107 us -- and whole chunks staged in shared memory with
// a block-wide copy-out instead of the in-loop flush -- 107.8 vs 107.3 us; commit c63fe50 has both.)
template <bool INTERVALS, bool COUNT>
__global__ void __launch_bounds__(256, 6) ms_kernel(MsParams p) {
    __shared__ __align__(16) uint8_t ms_stage[256 * 36];
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cnt_att = 0, cnt_split = 0, cnt_con = 0, cnt_extra = 0, cnt_proc = 0, cnt_emit = 0;
    unsigned long long cnt_att_e = 0, cnt_split_e = 0, cnt_con_e = 0, cnt_extra_e = 0;
    if (g < p.n_chunks) {
        const uint32_t n = p.ix.n, k = p.ix.k;
        const uint64_t start = g * p.chunk_len;
        const uint64_t remain = p.q.Lp - start;
        const uint32_t len = remain < p.chunk_len ? (uint32_t)remain : p.chunk_len;
        // Warm-up: the state at `start` depends on the k-1 bases before it, and only on those after the last
        // non-ACGT position among them (such a position resets the state).  The state after the first PREF_LEN of
        // the remaining bases comes from the table; the rest are stepped through.
        uint32_t warm = start >= (uint64_t)(k - 1) ? k - 1 : (uint32_t)start;
        for (uint64_t w = start >> 5; warm && w-- > ((start - warm) >> 5);) {  // chunk starts are multiples of 32
            uint32_t iv = __ldg(p.q.inv + w);
            if (w == ((start - warm) >> 5)) iv &= ~0u << ((start - warm) & 31);
            if (iv) {
                warm = (uint32_t)(start - (w * 32 + (31 - __clz((int)iv)) + 1));
                break;
            }
        }
        uint32_t l = 0, r = n, d = 0;
        const uint32_t P = p.ix.pref ? p.ix.pref_len : 0u;
        if (P && warm >= P) {
            const uint64_t first = start - warm;
            const uint32_t sh = 2 * (uint32_t)(first & 31);
            uint64_t bits = __ldg(p.q.pack + (first >> 5)) >> sh;
            if (sh > 64 - 2 * P) bits |= __ldg(p.q.pack + (first >> 5) + 1) << (64 - sh);
            if (pref_decode(__ldg(p.ix.pref + ((uint32_t)bits & ((1u << (2 * P)) - 1u))), n, l, r, d)) warm -= P;
        }
        const uint64_t pos0 = start - warm;
        const uint64_t wbase = pos0 >> 5;
        const uint64_t* __restrict__ qptr = p.q.pack + wbase;
        const uint32_t* __restrict__ iptr = p.q.inv + wbase;
        uint8_t* msw = p.ms + (wbase << 5);        // address of "bit position" 0 of this chunk
        uint32_t bp = (uint32_t)(pos0 & 31);       // position relative to the first loaded query word
        const uint32_t bp_emit = bp + warm;        // first emitted position
        const uint32_t bp_end = bp_emit + len;
        uint64_t qw = __ldg(qptr) >> (2 * bp);
        uint32_t iw = __ldg(iptr) >> bp;
        // emitted MS bytes are staged in shared memory (36-byte stride per lane: conflict-free word access) and
        // flushed as two 16-byte stores per 32 positions; chunk starts are multiples of 32, so flushes are aligned
        uint8_t* const stg = ms_stage + threadIdx.x * 36u;
        while (bp < bp_end) {
            bool advance = true;
            // (a non-ACGT position probes like any other and then resets the state: such positions are rare, and
            // a branch around the probe costs every iteration its test and the compiler's speculated reset)
            const bool inval = (iw & 1u) != 0;
            const uint32_t rowoff = ((uint32_t)qw & 3u) * p.ix.rank_stride;  // 32-bit word index
            const uint32_t bl = l >> 5, br = r >> 5;
            const uint64_t wl = __ldg(p.ix.rank + (rowoff + bl));
            const uint64_t wr = (br == bl) ? wl : __ldg(p.ix.rank + (rowoff + br));
            const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
            const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
            if (COUNT && !inval) {
                const bool sp = (bl >> 2) != (br >> 2);
                ++cnt_att; cnt_split += sp;
                if (bp >= bp_emit) { ++cnt_att_e; cnt_split_e += sp; }
            }
            if (nl < nr && !inval) {
                l = nl; r = nr;
                d = d + 1 < k ? d + 1 : k;
            } else if (d != 0 && !inval) {
                // contract_left to the largest depth that changes the interval: t = max(LCS[l], LCS[r])
                advance = false;
                const uint32_t el = __ldg(p.ix.links + l), er = __ldg(p.ix.links + r);
                const bool scanned = ms_contract(p.ix, el, er, l, r, d);
                if (COUNT) {
                    ++cnt_con; cnt_extra += scanned;
                    if (bp >= bp_emit) { ++cnt_con_e; cnt_extra_e += scanned; }
                }
            } else if (inval) {
                l = 0; r = n; d = 0;
            }
            if (advance) {
                if (COUNT) ++cnt_proc;
                // (warm-up positions are staged as well: their slots are rewritten before the first flush, because
                // bp_emit is a multiple of 32 and a flush needs bp > bp_emit)
                stg[bp & 31u] = (uint8_t)d;
                if (bp >= bp_emit) {
                    if (COUNT) ++cnt_emit;
                    if (INTERVALS) {
                        p.l_out[(wbase << 5) + bp] = l;
                        p.r_out[(wbase << 5) + bp] = r;
                    }
                }
                ++bp;
                qw >>= 2;
                iw >>= 1;
                if ((bp & 31) == 0 || bp == bp_end) {
                    if (bp > bp_emit) {  // flush the 32 (or last, partial) staged positions
                        const uint32_t* w = reinterpret_cast<const uint32_t*>(stg);
                        uint4* dst = reinterpret_cast<uint4*>(msw + ((bp - 1) & ~31u));
                        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                    if (bp < bp_end) {
                        qw = __ldg(qptr + (bp >> 5));
                        iw = __ldg(iptr + (bp >> 5));
                    }
                }
            }
        }
    }
    if (COUNT) {
        atomicAdd(p.counters + CNT_ATTEMPTS, cnt_att);
        atomicAdd(p.counters + CNT_SPLIT, cnt_split);
        atomicAdd(p.counters + CNT_CONTRACT, cnt_con);
        atomicAdd(p.counters + CNT_EXTRA_LCS, cnt_extra);
        atomicAdd(p.counters + CNT_PROCESSED, cnt_proc);
        atomicAdd(p.counters + CNT_EMITTED, cnt_emit);
        atomicAdd(p.counters + CNT_ATT_EMIT, cnt_att_e);
        atomicAdd(p.counters + CNT_SPLIT_EMIT, cnt_split_e);
        atomicAdd(p.counters + CNT_CON_EMIT, cnt_con_e);
        atomicAdd(p.counters + CNT_EXTRA_EMIT, cnt_extra_e);
    }
}

// K1p: K1 without intervals, probing TWO bases per rank word (IndexView::rank2) while the lane is in a matching
// stretch.  One code path per iteration: the lane picks the row set (rank2 / rank) and the advance (2 / 1) by select,
// not by branch, so pair lanes and single lanes of a warp execute the same instructions.  An empty pair says that one
// of the two extensions fails: the lane retries the first base alone (no contraction).  After any failed probe the
// lane probes single bases until two in a row have extended at the first try (noise stretches fail at every base).
template <bool COUNT>
__global__ void __launch_bounds__(256, 6) ms_pairs_kernel(MsParams p) {
    __shared__ __align__(16) uint8_t ms_stage[256 * 36];
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cnt_att = 0, cnt_split = 0, cnt_con = 0, cnt_extra = 0, cnt_proc = 0, cnt_emit = 0;
    unsigned long long cnt_att_e = 0, cnt_split_e = 0, cnt_con_e = 0, cnt_extra_e = 0;
    // state of the lane's chain (declared here: the main loop below is executed by the WHOLE warp, see there)
    const uint32_t n = p.ix.n, k = p.ix.k;
    uint32_t l = 0, r = n, d = 0, bp = 0, bp_emit = 0, bp_end = 0, iw = 0;
    uint64_t qw = 0, wbase = 0;
    const uint64_t* __restrict__ qptr = p.q.pack;
    const uint32_t* __restrict__ iptr = p.q.inv;
    uint8_t* msw = p.ms;
    if (g < p.n_chunks) {
        const uint64_t start = g * p.chunk_len;
        const uint64_t remain = p.q.Lp - start;
        const uint32_t len = remain < p.chunk_len ? (uint32_t)remain : p.chunk_len;
        uint32_t warm = start >= (uint64_t)(k - 1) ? k - 1 : (uint32_t)start;
        for (uint64_t w = start >> 5; warm && w-- > ((start - warm) >> 5);) {  // chunk starts are multiples of 32
            uint32_t iv = __ldg(p.q.inv + w);
            if (w == ((start - warm) >> 5)) iv &= ~0u << ((start - warm) & 31);
            if (iv) {
                warm = (uint32_t)(start - (w * 32 + (31 - __clz((int)iv)) + 1));
                break;
            }
        }
        const uint32_t P = p.ix.pref ? p.ix.pref_len : 0u;
        if (P && warm >= P) {
            const uint64_t first = start - warm;
            const uint32_t sh = 2 * (uint32_t)(first & 31);
            uint64_t bits = __ldg(p.q.pack + (first >> 5)) >> sh;
            if (sh > 64 - 2 * P) bits |= __ldg(p.q.pack + (first >> 5) + 1) << (64 - sh);
            if (pref_decode(__ldg(p.ix.pref + ((uint32_t)bits & ((1u << (2 * P)) - 1u))), n, l, r, d)) warm -= P;
        }
        const uint64_t pos0 = start - warm;
        wbase = pos0 >> 5;
        qptr = p.q.pack + wbase;
        iptr = p.q.inv + wbase;
        msw = p.ms + (wbase << 5);
        bp = (uint32_t)(pos0 & 31);
        bp_emit = bp + warm;
        bp_end = bp_emit + len;
        qw = __ldg(qptr) >> (2 * bp);
        iw = __ldg(iptr) >> bp;
    }
    {
        const uint32_t stride = p.ix.rank_stride;
        uint32_t cool = 0;  // successful advances left before the lane probes pairs again (set by every failed probe)
        uint8_t* const stg = ms_stage + threadIdx.x * 36u;
        while (bp < bp_end) {
          {
            uint32_t adv = 1, dA = 0, dB = 0;
            if (iw & 1u) {
                l = 0; r = n; d = 0;
                cool = 0;
            } else {
                // Integer arithmetic on `two` and ONE base pointer (rank2's 16 rows follow rank's 4 rows in the index
                // allocation), so pair lanes and single lanes run the same instructions: with a pointer select the
                // compiler emitted the probe twice, once per kind of lane, and the warp executed both.
                const uint32_t two = (uint32_t)(cool == 0) & (~(iw >> 1) & 1u) & (uint32_t)((bp & 31u) != 31u) &
                                     (uint32_t)(bp + 1 < bp_end);
                const uint32_t rowoff = (((uint32_t)qw & (3u + 12u * two)) + 4u * two) * stride;
                const uint32_t bl = l >> 5, br = r >> 5;
                const uint64_t wl = __ldg(p.ix.rank + (rowoff + bl));
                const uint64_t wr = (br == bl) ? wl : __ldg(p.ix.rank + (rowoff + br));
                const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
                const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
                if (COUNT) {
                    const bool sp = (bl >> 2) != (br >> 2);
                    ++cnt_att; cnt_split += sp;
                    if (bp + two >= bp_emit) { ++cnt_att_e; cnt_split_e += sp; }
                }
                if (nl < nr) {
                    l = nl; r = nr;
                    dA = d + 1 < k ? d + 1 : k;
                    dB = d + 2 < k ? d + 2 : k;
                    adv = 1u + two;
                    d = two ? dB : dA;
                    cool = cool ? cool - 1u : 0u;
                } else {
                    cool = 2;  // single probes until two bases in a row have extended at the first try
                    adv = (two | d) ? 0u : 1u;  // (d == 0 and a single base that matches nothing: emitted with d == 0)
                    if (!two && d != 0) {
                        // contract_left to the largest depth that changes the interval: t = max(LCS[l], LCS[r])
                        const uint32_t el = __ldg(p.ix.links + l), er = __ldg(p.ix.links + r);
                        const bool scanned = ms_contract(p.ix, el, er, l, r, d);
                        if (COUNT) {
                            ++cnt_con; cnt_extra += scanned;
                            if (bp >= bp_emit) { ++cnt_con_e; cnt_extra_e += scanned; }
                        }
                    }
                }
            }
            if (adv) {
                if (COUNT) { cnt_proc += adv; cnt_emit += (bp >= bp_emit) + (adv == 2 && bp + 1 >= bp_emit); }
                if (bp >= bp_emit) stg[bp & 31u] = (uint8_t)(adv == 2 ? dA : d);
                if (adv == 2 && bp + 1 >= bp_emit) stg[(bp + 1) & 31u] = (uint8_t)dB;
                bp += adv;
                qw >>= 2 * adv;
                iw >>= adv;
                if ((bp & 31) == 0 || bp == bp_end) {
                    if (bp > bp_emit) {  // flush the 32 (or last, partial) staged positions
                        const uint32_t* w = reinterpret_cast<const uint32_t*>(stg);
                        uint4* dst = reinterpret_cast<uint4*>(msw + ((bp - 1) & ~31u));
                        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                    if (bp < bp_end) {
                        qw = __ldg(qptr + (bp >> 5));
                        iw = __ldg(iptr + (bp >> 5));
                    }
                }
            }
          }
        }
    }
    if (COUNT) {
        atomicAdd(p.counters + CNT_ATTEMPTS, cnt_att);
        atomicAdd(p.counters + CNT_SPLIT, cnt_split);
        atomicAdd(p.counters + CNT_CONTRACT, cnt_con);
        atomicAdd(p.counters + CNT_EXTRA_LCS, cnt_extra);
        atomicAdd(p.counters + CNT_PROCESSED, cnt_proc);
        atomicAdd(p.counters + CNT_EMITTED, cnt_emit);
        atomicAdd(p.counters + CNT_ATT_EMIT, cnt_att_e);
        atomicAdd(p.counters + CNT_SPLIT_EMIT, cnt_split_e);
        atomicAdd(p.counters + CNT_CON_EMIT, cnt_con_e);
        atomicAdd(p.counters + CNT_EXTRA_EMIT, cnt_extra_e);
    }
}

// ---------------------------------------------------------------------------
// K2: derandomize + translate, fused, on the u8 MS vector K1 wrote.
//
// K1's output satisfies ms[i+1] <= ms[i] + 1 (a match can grow by at most one
// base).  Under that invariant the right-to-left recurrence of
// derandomize_ms_val (derandomize.rs:221-247) has the closed form
//     out[i] = ms[i] - eps[i]                       if ms[i] > threshold
//     out[i] = out[i+1] - 1                         otherwise
// with eps[i] in {0,1}:  eps = 0 if ms[i] == k, or ms[i+1] < ms[i], or i is the
// last position;  eps[i] = eps[i+1] if ms[i+1] == ms[i] + 1;  eps[i] = 1 - eps[i+1]
// if ms[i+1] == ms[i]   (DESIGN.md "Derandomize as two scans" gives the proof).
// translate_ms_vec only distinguishes out <= 0, == 1, < thr, > thr, so the
// kernel carries max(out, 0) in a byte; max(., 0) commutes with the recurrence.
//
// One warp per tile of 512 padded positions, 16 per lane.  The value entering
// the tile from the right (c[e]) is obtained by a warp-parallel look-ahead that
// ends at the first position whose eps is known unconditionally.
// ---------------------------------------------------------------------------
struct TrParams {
    const uint8_t* ms;  // padded space; readable up to n_words*32 + 16
    QueryView q;
    uint32_t k, thr;
    uint8_t* out;       // out[off0 + pp - (#separators before pp)]
    uint64_t off0;
    uint64_t n_tiles;   // tiles of the kernel that is launched (K2: 512 positions, K2b: 1024)
    uint32_t* out_gap;  // K2b<false>: one bit per padded position, '-'
    uint32_t* out_match;  //                                         'M' or 'R'
    uint32_t* out_r;      //                                         'R'
};

// K2b covers this parameter range; K2 everything else
__host__ __device__ inline bool k2b_supported(uint32_t k, uint32_t thr) { return thr >= 2 && thr < k && k <= 127; }

enum { K2_PER_LANE = 16, K2_TILE = 512, K2_WARPS = 4 };
enum { OP_KEEP = 0, OP_TOGGLE = 1, OP_SET0 = 2 };

__device__ __forceinline__ uint32_t sep_bit(const QueryView& q, int64_t pp) {
    if (pp < 0 || (uint64_t)pp >= q.n_words * 32) return 1u;
    return (__ldg(q.sep + (pp >> 5)) >> (pp & 31)) & 1u;
}

// parity op of position with value m0, right neighbour m1 (derandomize closed form)
__device__ __forceinline__ uint32_t parity_op(bool elig, uint32_t m0, uint32_t m1, uint32_t k) {
    if (!elig || m0 == k || m1 < m0) return OP_SET0;
    return (m1 == m0) ? OP_TOGGLE : OP_KEEP;
}

// composite parity transform encoded as (is_const << 1) | val ; apply `left` after `right`
__device__ __forceinline__ uint32_t par_compose(uint32_t left, uint32_t right) {
    if (left & 2u) return left;
    return (right & 2u) | ((right ^ left) & 1u);
}
__device__ __forceinline__ uint32_t par_apply(uint32_t f, uint32_t eps_in) {
    return (f & 2u) ? (f & 1u) : ((f ^ eps_in) & 1u);
}

// "nearest source" transform encoded as (has << 16) | x : x = value if has, else distance
__device__ __forceinline__ uint32_t src_compose(uint32_t left, uint32_t right) {
    if (left >> 16) return left;
    uint32_t a = left & 0xffffu, x = right & 0xffffu;
    if (right >> 16) return (1u << 16) | (x > a ? x - a : 0u);
    uint32_t s = a + x;
    return s > 0xffffu ? 0xffffu : s;
}
__device__ __forceinline__ uint32_t src_apply(uint32_t f, uint32_t c_in) {
    uint32_t x = f & 0xffffu;
    if (f >> 16) return x;
    return c_in > x ? c_in - x : 0u;
}

// MS bytes beyond the part of the vector that is in memory (the fused kernel keeps the MS of one tile plus a short
// look-ahead in shared memory).  When the look-ahead of derandomize runs past `end`, lane 0 continues the MS
// recurrence from the state at `end` and the warp reads the new values from `ring` (64 bytes, indexed by position
// modulo 64).  Rare: it takes a run of more than MS_LOOKAHEAD positions whose derandomized value stays undecided.
struct MsTail {
    uint64_t end;        // MS is in memory for positions < end
    uint32_t l, r, d;    // MS state after position end - 1
    uint8_t* ring;       // this warp's 64 bytes
    const IndexView* ix;
};
__device__ __forceinline__ void ms_feed_base(const IndexView& ix, uint32_t c, uint32_t& l, uint32_t& r, uint32_t& d);

// Clamped derandomized value of position e (first position right of a tile).  Warp-uniform.
// `ms` is indexed by padded position; `tail` (optional) says where it ends and how to continue it.
__device__ __forceinline__ uint32_t lookahead_c(const TrParams& p, const uint8_t* ms, uint64_t e, int lane,
                                                const MsTail* tail = nullptr) {
    if (sep_bit(p.q, (int64_t)e)) return 0u;  // tile ends exactly at a query end: nothing enters
    uint32_t dist = 0;       // N positions skipped before the first source
    bool in_run = false;     // source found, waiting for the first SET0
    uint32_t src_val = 0, src_dist = 0, parity = 0;
    uint64_t have = tail ? tail->end : ~0ull;  // MS known for positions < have (memory, then the ring)
    uint32_t tl = 0, tr = 0, td = 0;
    if (tail) { tl = tail->l; tr = tail->r; td = tail->d; }
    for (uint64_t base = e;; base += 32) {
        const uint64_t pp = base + lane;
        if (tail && base + 33 > have) {  // warp-uniform: this round reads positions up to base + 32
            if (lane == 0) {
#ifdef KBO_HOST_EMU
                ++emu_tail_extensions();  // (tests check that this path is reached)
#endif
                for (uint64_t q = have; q < base + 33; ++q) {
                    uint32_t v = 0;
                    if (q < p.q.Lp && !((__ldg(p.q.inv + (q >> 5)) >> (q & 31)) & 1u)) {
                        ms_feed_base(*tail->ix, (uint32_t)(__ldg(p.q.pack + (q >> 5)) >> (2 * (q & 31))) & 3u, tl, tr, td);
                        v = td;
                    } else {
                        tl = 0; tr = tail->ix->n; td = 0;
                    }
                    tail->ring[q & 63] = (uint8_t)v;
                }
            }
            have = base + 33;
            __syncwarp();
        }
        const bool s0 = sep_bit(p.q, (int64_t)pp), s1 = sep_bit(p.q, (int64_t)pp + 1);
        uint32_t m0 = 0, m1 = 0;  // (K1 never wrote past the batch)
        if (!s0) m0 = (tail && pp >= tail->end) ? tail->ring[pp & 63] : ms[pp];
        if (!s1) m1 = (tail && pp + 1 >= tail->end) ? tail->ring[(pp + 1) & 63] : ms[pp + 1];
        const bool last = !s0 && s1;
        const bool elig = !s0 && !last && (m0 > p.thr || m0 == p.k);
        const bool source = s0 || last || elig;
        const uint32_t op = parity_op(elig, m0, m1, p.k);
        const uint32_t srcmask = __ballot_sync(0xffffffffu, source);
        const uint32_t setmask = __ballot_sync(0xffffffffu, op == OP_SET0);
        const uint32_t togmask = __ballot_sync(0xffffffffu, op == OP_TOGGLE);
        uint32_t from = 0;  // first lane of this round that belongs to the run
        if (!in_run) {
            if (srcmask == 0) {
                dist += 32;
                if (dist >= p.k) return 0u;  // c[e] <= k - dist <= 0
                continue;
            }
            const int j = __ffs((int)srcmask) - 1;
            src_dist = dist + (uint32_t)j;
            const uint32_t mj = __shfl_sync(0xffffffffu, m0, j);
            const uint32_t kind = __shfl_sync(0xffffffffu, (uint32_t)(elig ? 2 : (last ? 1 : 0)), j);
            if (kind != 2) {  // anchor (or separator, which cannot come first): value known
                const uint32_t v = (kind == 1 && mj > p.thr) ? mj : 0u;
                return v > src_dist ? v - src_dist : 0u;
            }
            src_val = mj;
            in_run = true;
            from = (uint32_t)j;
        }
        const uint32_t sm = setmask & (~0u << from);
        if (sm) {
            const int z = __ffs((int)sm) - 1;
            const uint32_t between = (z == 0) ? 0u : (togmask & (~0u << from) & ((1u << z) - 1u));
            parity ^= __popc(between) & 1u;
            const uint32_t v = src_val - parity;
            return v > src_dist ? v - src_dist : 0u;
        }
        parity ^= __popc(togmask & (~0u << from)) & 1u;
    }
}

// the X / - / M rule and the R rules of translate_ms_vec in closed form (translate.rs:180-216,263-293)
__device__ __forceinline__ uint8_t translate_char(uint32_t prevc, uint32_t cur, uint32_t nextc, bool first0,
                                                  bool first1, bool last, uint32_t k, uint32_t thr) {
    const uint32_t prev = (first0 || first1) ? k : prevc;
    const uint32_t next = last ? cur : nextc;
    const bool trig = !last && cur > thr && next > 0 && next < thr;
    const bool trig_prev = !first0 && !first1 && !last && prevc > thr && cur > 0 && cur < thr;
    if (trig || trig_prev) return 'R';
    if (cur == 0) return (next == 1 && prev > 0) ? 'X' : '-';
    return 'M';
}

__global__ void __launch_bounds__(K2_WARPS * 32) derand_translate_kernel(TrParams p) {
    __shared__ uint8_t stage[K2_WARPS][K2_TILE + 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * K2_WARPS + warp;
    if (tile >= p.n_tiles) return;  // warp-uniform
    const uint64_t s = tile * K2_TILE;
    const uint64_t P = s + (uint64_t)K2_PER_LANE * lane;
    const uint32_t k = p.k, thr = p.thr;

    // ---- loads ------------------------------------------------------------
    // separator bits of positions P-2 .. P+16 -> sf bit (t+2) = sep(P+t)
    uint32_t sf;
    {
        const uint32_t sw = __ldg(p.q.sep + (P >> 5));
        sf = ((sw >> (P & 31)) & 0xffffu) << 2;
        sf |= sep_bit(p.q, (int64_t)P - 2) | (sep_bit(p.q, (int64_t)P - 1) << 1);
        sf |= sep_bit(p.q, (int64_t)P + 16) << 18;
    }
    uint32_t m[K2_PER_LANE + 1];
    {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (((sf >> 2) & 0xffffu) != 0xffffu) v = *reinterpret_cast<const uint4*>(p.ms + P);  // K1 never wrote past the batch
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < K2_PER_LANE; ++t) m[t] = (w[t >> 2] >> (8 * (t & 3))) & 0xffu;
        m[K2_PER_LANE] = ((sf >> 18) & 1u) ? 0u : p.ms[P + K2_PER_LANE];
    }
    const uint32_t c_e = lookahead_c(p, p.ms, s + K2_TILE, lane);
    const bool e_sep = sep_bit(p.q, (int64_t)(s + K2_TILE));
    const uint32_t m_e = e_sep ? 0u : p.ms[s + K2_TILE];
    const bool e_last = !e_sep && sep_bit(p.q, (int64_t)(s + K2_TILE) + 1);
    const uint32_t eps_e = (!e_sep && !e_last && (m_e > thr || m_e == k)) ? (m_e - c_e) & 1u : 0u;

    // ---- phase A: parity transform of the lane, suffix-composed over the warp --
    uint32_t ops = 0;   // 2 bits per position
    uint32_t kinds = 0; // 2 bits per position: 0 N, 1 last(anchor), 2 eligible, 3 separator
    uint32_t F = 0;     // identity
#pragma unroll
    for (int t = K2_PER_LANE - 1; t >= 0; --t) {
        const bool s0 = (sf >> (t + 2)) & 1u, s1 = (sf >> (t + 3)) & 1u;
        const bool last = !s0 && s1;
        const bool elig = !s0 && !last && (m[t] > thr || m[t] == k);
        const uint32_t op = parity_op(elig, m[t], m[t + 1], k);
        ops |= op << (2 * t);
        kinds |= (s0 ? 3u : (last ? 1u : (elig ? 2u : 0u))) << (2 * t);
        F = par_compose(op == OP_SET0 ? 2u : op, F);
    }
    uint32_t G = F;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, G, off);
        if (lane + off < 32) G = par_compose(G, o);
    }
    uint32_t Gn = __shfl_down_sync(0xffffffffu, G, 1);
    if (lane == 31) Gn = 0;
    const uint32_t eps_r = par_apply(Gn, eps_e);

    // ---- phase B: source values, then nearest-source transform over the warp ----
    uint32_t cv[K2_PER_LANE];
    uint32_t H = 0;  // no source yet, distance 0
    {
        uint32_t eps = eps_r;
#pragma unroll
        for (int t = K2_PER_LANE - 1; t >= 0; --t) {
            const uint32_t op = (ops >> (2 * t)) & 3u, kind = (kinds >> (2 * t)) & 3u;
            eps = (op == OP_SET0) ? 0u : (eps ^ op);
            uint32_t v = 0;
            if (kind == 2) v = m[t] - eps;
            else if (kind == 1) v = m[t] > thr ? m[t] : 0u;
            cv[t] = v;
            const uint32_t f = (kind == 0) ? 1u : ((1u << 16) | v);  // N: distance 1 ; source: value
            H = src_compose(f, H);
        }
    }
    uint32_t HH = H;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, HH, off);
        if (lane + off < 32) HH = src_compose(HH, o);
    }
    uint32_t Hn = __shfl_down_sync(0xffffffffu, HH, 1);
    if (lane == 31) Hn = 0;
    const uint32_t c_r = src_apply(Hn, c_e);  // clamped derandomized value of position P+16

    // ---- phase B3: final clamped values of the lane's positions ----------------
    uint32_t c[K2_PER_LANE + 1];
    c[K2_PER_LANE] = c_r;
#pragma unroll
    for (int t = K2_PER_LANE - 1; t >= 0; --t) {
        const uint32_t kind = (kinds >> (2 * t)) & 3u;
        c[t] = (kind == 0) ? (c[t + 1] > 0 ? c[t + 1] - 1 : 0u) : cv[t];
    }

    // ---- value of position P-1 (left neighbour lane, or computed for lane 0) ----
    uint32_t c_left = __shfl_up_sync(0xffffffffu, c[K2_PER_LANE - 1], 1);
    if (lane == 0) {
        c_left = 0;
        if (s > 0 && !((sf >> 1) & 1u) && !((sf >> 2) & 1u)) {  // P-1 and P are in the same query
            const uint32_t mp = p.ms[s - 1];
            if (mp > thr || mp == k) {
                const uint32_t kind0 = kinds & 3u;
                const uint32_t eps0 = (kind0 == 2) ? (m[0] - c[0]) & 1u : 0u;
                const uint32_t op = parity_op(true, mp, m[0], k);
                const uint32_t eps = (op == OP_SET0) ? 0u : (eps0 ^ op);
                c_left = mp - eps;
            } else {
                c_left = c[0] > 0 ? c[0] - 1 : 0u;
            }
        }
    }

    // ---- phase C: translate and stage ------------------------------------------
    // separators before P inside this tile -> local output index
    uint32_t nsep_lane = __popc((sf >> 2) & 0xffffu);
    uint32_t incl = nsep_lane;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    const uint32_t nsep_tile = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t li = (uint32_t)(K2_PER_LANE * lane) - (incl - nsep_lane);
#pragma unroll
    for (int t = 0; t < K2_PER_LANE; ++t) {
        const uint32_t kind = (kinds >> (2 * t)) & 3u;
        if (kind == 3) continue;
        const bool first0 = (sf >> (t + 1)) & 1u;                 // P+t-1 is a separator / before the batch
        const bool first1 = !first0 && ((sf >> t) & 1u);          // P+t-2 is
        const uint32_t prevc = (t == 0) ? c_left : c[t - 1];
        stage[warp][li++] = translate_char(prevc, c[t], c[t + 1], first0, first1, kind == 1, k, thr);
    }
    __syncwarp();

    // ---- coalesced copy of the tile's characters to the unpadded output -----------
    const uint32_t cnt = K2_TILE - nsep_tile;
    uint8_t* dst = p.out + p.off0 + s - __ldg(p.q.wq + (s >> 5));  // tile starts are word aligned
    const uint32_t head0 = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    const uint32_t head = head0 < cnt ? head0 : cnt;
    if ((uint32_t)lane < head) dst[lane] = stage[warp][lane];
    const uint32_t nwords = (cnt - head) >> 2;
    for (uint32_t w = lane; w < nwords; w += 32) {
        const uint8_t* sp = &stage[warp][head + 4 * w];
        const uint32_t v = (uint32_t)sp[0] | ((uint32_t)sp[1] << 8) | ((uint32_t)sp[2] << 16) | ((uint32_t)sp[3] << 24);
        *reinterpret_cast<uint32_t*>(dst + head + 4 * w) = v;
    }
    const uint32_t done = head + 4 * nwords;
    if (done + lane < cnt) dst[done + lane] = stage[warp][done + lane];
}

// ---------------------------------------------------------------------------
// K2b: the same function as K2, bit-parallel.  One LANE per 32 padded positions,
// one warp per tile of 1024.  Valid for 2 <= thr < k <= 127 (the host picks K2
// otherwise); within that range
//   * MS bytes are < 128, so four of them are compared per 32-bit operation and
//     the per-byte flags are gathered into per-position bit masks with one
//     multiplication;
//   * an eligible position has m > thr, hence c = m - eps >= thr >= 2: it is
//     never in the classes "c == 0", "c == 1", "0 < c < thr", and it leaves
//     "c > thr" only for m == thr + 1 with eps == 1;
//   * eps is a segmented suffix XOR over the TOGGLE mask (5 shift steps);
//   * the positions below the threshold form runs that end at a source (eligible
//     position, last position of a query, or the value entering the word from
//     the right); inside a run c = max(v - distance, 0), so each class is a bit
//     range computed from the run's end and v.
// translate_ms_vec's rules then are a dozen logic operations on the class masks.
// CHARS == true stores the characters (unpadded, like K2); CHARS == false stores
// the three masks K4b consumes ('-', match-like, 'R') in padded space.
// ---------------------------------------------------------------------------
enum { K2B_TILE = 1024, K2B_WARPS = 4 };

__device__ __forceinline__ uint32_t bits_le(int t) {  // bits 0..t
    return t < 0 ? 0u : (t >= 31 ? ~0u : ((2u << t) - 1u));
}
__device__ __forceinline__ uint32_t bits_ge(int t) {  // bits t..31
    return t <= 0 ? ~0u : (t >= 32 ? 0u : (~0u << t));
}
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t bits) {  // bits in [0,31]
    return bits ? ((lo >> bits) | (hi << (32 - bits))) : lo;
}

// copy n bytes from shared memory (any alignment) to global memory, one warp
__device__ __forceinline__ void warp_copy_out(const uint8_t* st, uint32_t src_off, uint8_t* dst, uint32_t n, int lane) {
    const uint32_t head0 = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const uint32_t head = head0 < n ? head0 : n;
    if ((uint32_t)lane < head) dst[lane] = st[src_off + lane];
    const uint32_t nwords = (n - head) >> 2;
    for (uint32_t w = lane; w < nwords; w += 32) {
        const uint32_t so = src_off + head + 4 * w;
        const uint32_t* a = reinterpret_cast<const uint32_t*>(st + (so & ~3u));
        *reinterpret_cast<uint32_t*>(dst + head + 4 * w) = funnel_r(a[0], a[1], 8 * (so & 3u));
    }
    const uint32_t done = head + 4 * nwords;
    if (done + lane < n) dst[done + lane] = st[src_off + done + lane];
}

// (4 bits b0, 4 bits b1) -> 4 characters; all threads of the block, followed by __syncthreads by the caller
__device__ __forceinline__ void k2b_fill_lut(uint32_t* lut) {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t v = 0;
        for (int j = 0; j < 4; ++j) {
            const uint32_t code = ((i >> j) & 1u) | (((i >> (4 + j)) & 1u) << 1);  // 0 M, 1 -, 2 X, 3 R
            const uint32_t ch = code == 0 ? 'M' : (code == 1 ? '-' : (code == 2 ? 'X' : 'R'));
            v |= ch << (8 * j);
        }
        lut[i] = v;
    }
}

// One warp, one group of up to 32 words (1024 padded positions) starting at position s (a multiple of 32): lanes
// 0..last hold one word each, lanes above `last` are inert (they behave like words full of separators and write
// nothing).  `ms` is indexed by padded position and must be readable for [s - 1, s + 32 (last + 1) + 4) -- global
// memory (K2b after K1) or the fused kernel's shared-memory tile; `tail` continues it for the look-ahead (MsTail).
template <bool CHARS>
__device__ __forceinline__ void k2b_group(const TrParams& p, const uint8_t* ms, const uint64_t s, const int last,
                                          const MsTail* tail, const uint32_t* lut, uint8_t* stage) {
    const int lane = threadIdx.x & 31;
    const uint64_t e = s + 32ull * (uint32_t)(last + 1);
    const uint64_t P = s + 32ull * lane;
    const uint32_t k = p.k, thr = p.thr;
    const uint32_t H = 0x80808080u;

    // ---- loads: 32 MS bytes + the first four of the next word; separator bits around the word -------------
    const uint32_t S = lane <= last ? __ldg(p.q.sep + (P >> 5)) : ~0u;
    uint32_t s_next = __shfl_down_sync(0xffffffffu, S, 1) & 1u;  // separator bit of P+32
    if (lane == last) s_next = sep_bit(p.q, (int64_t)e);
    uint32_t mw[9];
    {
        uint4 va = make_uint4(0u, 0u, 0u, 0u), vb = va;
        if (S != ~0u) {  // words past the end of the batch were never written by K1
            va = *reinterpret_cast<const uint4*>(ms + P);
            vb = *reinterpret_cast<const uint4*>(ms + P + 16);
        }
        mw[0] = va.x; mw[1] = va.y; mw[2] = va.z; mw[3] = va.w;
        mw[4] = vb.x; mw[5] = vb.y; mw[6] = vb.z; mw[7] = vb.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) mw[j] &= 0x7f7f7f7fu;  // bytes of separators may hold anything
        mw[8] = __shfl_down_sync(0xffffffffu, mw[0], 1);
        if (lane == last) mw[8] = s_next ? 0u : (*reinterpret_cast<const uint32_t*>(ms + P + 32) & 0x7f7f7f7fu);
    }
    uint32_t s_prev = __shfl_up_sync(0xffffffffu, S, 1) >> 30;   // bit0 = sep(P-2), bit1 = sep(P-1)
    if (lane == 0) s_prev = sep_bit(p.q, (int64_t)P - 2) | (sep_bit(p.q, (int64_t)P - 1) << 1);

    // ---- value entering the tile from the right ---------------------------------------------------------------
    const uint32_t c_e = lookahead_c(p, ms, e, lane, tail);
    const bool e_sep = sep_bit(p.q, (int64_t)e);
    const uint32_t m_e = e_sep ? 0u : (((tail && e >= tail->end) ? tail->ring[e & 63] : ms[e]) & 0x7fu);
    const bool e_last = !e_sep && sep_bit(p.q, (int64_t)e + 1);
    const uint32_t eps_e = (!e_sep && !e_last && (m_e > thr || m_e == k)) ? (m_e - c_e) & 1u : 0u;

    // ---- byte-parallel comparisons -> per-position masks ----------------------------------------------------
    uint32_t GT = 0, GT1 = 0, NEK = 0, GE = 0, B0 = 0;
    {
        const uint32_t thr1 = (thr + 1) * 0x01010101u, thr2 = (thr + 2) * 0x01010101u, krep = k * 0x01010101u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t w = mw[j];
            const uint32_t nx = (w >> 8) | (mw[j + 1] << 24);  // the right neighbours of the four positions
            const uint32_t gt = ((w | H) - thr1) & H;           // m >= thr+1
            const uint32_t gt1 = ((w | H) - thr2) & H;          // m >= thr+2
            const uint32_t nek = ((w ^ krep) + 0x7f7f7f7fu) & H;  // m != k
            const uint32_t d = (nx | H) - w;                    // byte = 128 + m1 - m0
            GT |= ((gt * 0x00204081u) >> 28) << (4 * j);
            GT1 |= ((gt1 * 0x00204081u) >> 28) << (4 * j);
            NEK |= ((nek * 0x00204081u) >> 28) << (4 * j);
            GE |= (((d & H) * 0x00204081u) >> 28) << (4 * j);   // m1 >= m0
            B0 |= (((d & 0x01010101u) * 0x10204080u) >> 28) << (4 * j);  // with m1 >= m0: m1 == m0 + 1
        }
    }
    const uint32_t L = ~S & ((S >> 1) | (s_next << 31));  // last position of its query
    const uint32_t E = ~S & ~L & GT;                       // eligible (m == k implies m > thr here)
    const uint32_t Z = ~E | ~NEK | ~GE;                    // parity op SET0
    const uint32_t T = E & ~Z & ~B0;                       // parity op TOGGLE (m1 == m0)

    // ---- eps: segmented suffix XOR of T inside the word, then across the warp -------------------------------
    uint32_t x = T, open = ~Z;  // open[i]: no SET0 in the window examined so far
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
        x ^= open & (x >> sft);
        open &= (open >> sft) | (~0u << (32 - sft));
    }
    // now x[i] = eps[i] if nothing entered from the right, open[i] = no SET0 in [i, 31]
    uint32_t G = (Z ? 2u : 0u) | (x & 1u);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, G, off);
        if (lane + off <= last) G = par_compose(G, o);
    }
    uint32_t Gn = __shfl_down_sync(0xffffffffu, G, 1);
    if (lane >= last) Gn = 0;
    const uint32_t eps = x ^ (open & (0u - par_apply(Gn, eps_e)));

    // ---- value of a source position of this word --------------------------------------------------------------
    auto source_value = [&](uint32_t i) -> uint32_t {
        if ((S >> i) & 1u) return 0u;
        const uint32_t m = ms[P + i] & 0x7fu;
        if ((L >> i) & 1u) return ((GT >> i) & 1u) ? m : 0u;
        return m - ((eps >> i) & 1u);
    };
    const uint32_t SRC = S | L | E;
    uint32_t Hs = 32u;  // no source: 32 positions of distance
    if (SRC) {
        const uint32_t f = (uint32_t)__ffs((int)SRC) - 1u;
        const uint32_t v = source_value(f);
        Hs = (1u << 16) | (v > f ? v - f : 0u);
    }
    uint32_t HH = Hs;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, HH, off);
        if (lane + off <= last) HH = src_compose(HH, o);
    }
    uint32_t Hn = __shfl_down_sync(0xffffffffu, HH, 1);
    if (lane >= last) Hn = 0;
    const uint32_t c_in = src_apply(Hn, c_e);  // clamped derandomized value of position P+32

    // ---- classes of the below-threshold runs ---------------------------------------------------------------------
    uint32_t Z0 = L & ~GT, O1 = 0, PL = 0, PG = (E & ~(GT & ~GT1 & eps)) | (L & GT);
    uint32_t c_first = 0;  // numeric value of position P (needed by lane 0 only)
    {
        uint32_t rem = ~SRC;
        while (rem) {
            const uint32_t a = (uint32_t)__ffs((int)rem) - 1u;
            const uint32_t inv_run = ~(rem >> a);
            const uint32_t len = inv_run ? (uint32_t)__ffs((int)inv_run) - 1u : 32u;
            const uint32_t end = a + len;  // the run's source (32 = next word)
            const uint32_t v = end == 32 ? c_in : source_value(end);
            const uint32_t run = bits_ge((int)a) & bits_le((int)end - 1);
            const int t0 = (int)end - (int)v;  // c == 0 at and below t0
            Z0 |= run & bits_le(t0);
            O1 |= run & bits_ge(t0 + 1) & bits_le(t0 + 1);
            PL |= run & bits_ge(t0 + 1) & bits_le(t0 + (int)thr - 1);
            PG |= run & bits_ge(t0 + (int)thr + 1);
            if (a == 0) c_first = v > end ? v - end : 0u;
            rem &= ~run;
        }
        if (SRC & 1u) c_first = source_value(0);
    }

    // ---- neighbours across the word / tile boundaries --------------------------------------------------------
    uint32_t nPL = __shfl_down_sync(0xffffffffu, PL, 1) & 1u, nO1 = __shfl_down_sync(0xffffffffu, O1, 1) & 1u;
    if (lane >= last) {
        nPL = (c_e > 0 && c_e < thr) ? 1u : 0u;
        nO1 = c_e == 1 ? 1u : 0u;
    }
    uint32_t pPG = __shfl_up_sync(0xffffffffu, PG, 1) >> 31, pZ0 = __shfl_up_sync(0xffffffffu, Z0, 1) >> 31;
    if (lane == 0) {
        uint32_t c_left = 0;
        if (s > 0 && !(s_prev & 2u) && !(S & 1u)) {  // P-1 and P are in the same query
            const uint32_t mp = ms[s - 1] & 0x7fu;
            if (mp > thr) {
                const uint32_t eps0 = (E & 1u) ? (eps & 1u) : 0u;
                const uint32_t op = parity_op(true, mp, mw[0] & 0xffu, k);
                c_left = mp - ((op == OP_SET0) ? 0u : (eps0 ^ op));
            } else {
                c_left = c_first > 0 ? c_first - 1 : 0u;
            }
        }
        pPG = c_left > thr ? 1u : 0u;
        pZ0 = c_left == 0 ? 1u : 0u;
    }
    const uint32_t next_PL = (PL >> 1) | (nPL << 31), next_O1 = (O1 >> 1) | (nO1 << 31);
    const uint32_t prev_PG = (PG << 1) | pPG, prev_Z0 = (Z0 << 1) | pZ0;
    const uint32_t F01 = (S << 1) | (S << 2) | (s_prev >> 1) | ((s_prev & 2u)) | ((s_prev & 1u));
    // (bit 0: sep(P-1) | sep(P-2); bit 1: sep(P) | sep(P-1))

    // ---- translate_ms_vec (translate.rs:180-216,263-293) on the class masks ------------------------------------
    const uint32_t V = ~S;
    const uint32_t R = V & ~L & ((PG & next_PL) | (~F01 & prev_PG & PL));
    const uint32_t X = V & ~R & Z0 & ~L & next_O1 & (F01 | ~prev_Z0);
    const uint32_t dash = V & Z0 & ~X;

    if (!CHARS) {
        if (lane <= last) {
            p.out_gap[P >> 5] = dash;
            p.out_match[P >> 5] = V & ~Z0;  // 'M' or 'R'
            p.out_r[P >> 5] = R;
        }
        return;
    }

    // ---- characters: padded layout in shared memory, then the runs between separators are copied out ---------
    {
        const uint32_t b0 = dash | R, b1 = X | R;
        uint32_t* st = reinterpret_cast<uint32_t*>(stage + 32 * lane);
        uint4 o;
        o.x = lut[(b0 & 15u) | ((b1 & 15u) << 4)];
        o.y = lut[((b0 >> 4) & 15u) | (((b1 >> 4) & 15u) << 4)];
        o.z = lut[((b0 >> 8) & 15u) | (((b1 >> 8) & 15u) << 4)];
        o.w = lut[((b0 >> 12) & 15u) | (((b1 >> 12) & 15u) << 4)];
        *reinterpret_cast<uint4*>(st) = o;
        o.x = lut[((b0 >> 16) & 15u) | (((b1 >> 16) & 15u) << 4)];
        o.y = lut[((b0 >> 20) & 15u) | (((b1 >> 20) & 15u) << 4)];
        o.z = lut[((b0 >> 24) & 15u) | (((b1 >> 24) & 15u) << 4)];
        o.w = lut[((b0 >> 28) & 15u) | (((b1 >> 28) & 15u) << 4)];
        *reinterpret_cast<uint4*>(st + 4) = o;
    }
    __syncwarp();
    uint8_t* dst = p.out + p.off0 + s - __ldg(p.q.wq + (s >> 5));  // tile starts are word aligned
    uint32_t seg = 0;  // start of the current run of non-separator positions (tile coordinates)
    uint32_t has = __ballot_sync(0xffffffffu, S != 0);
    while (has) {
        const int l = __ffs((int)has) - 1;
        has &= has - 1;
        uint32_t sw = __shfl_sync(0xffffffffu, S, l);
        while (sw) {
            const uint32_t sp = 32u * l + (uint32_t)__ffs((int)sw) - 1u;
            sw &= sw - 1;
            if (sp > seg) {
                warp_copy_out(stage, seg, dst, sp - seg, lane);
                dst += sp - seg;
            }
            seg = sp + 1;
        }
    }
    if (seg < K2B_TILE) warp_copy_out(stage, seg, dst, K2B_TILE - seg, lane);
}

template <bool CHARS>
__global__ void __launch_bounds__(K2B_WARPS * 32, 12) derand_translate_bits_kernel(TrParams p) {  // 40 registers, no spills: 12 blocks per SM
    __shared__ uint32_t lut[256];
    __shared__ __align__(16) uint8_t stage[K2B_WARPS][K2B_TILE + 16];
    if (CHARS) {
        k2b_fill_lut(lut);
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * K2B_WARPS + warp;
    if (tile >= p.n_tiles) return;  // warp-uniform
    k2b_group<CHARS>(p, p.ms, tile * K2B_TILE, 31, nullptr, lut, stage[warp]);
}

// ---------------------------------------------------------------------------
// K4: format::run_lengths_gapped (format.rs:143-193) on the PLAIN translation
// (alphabet M - X R only), data-parallel over the whole batch in padded space.
//
// Input: three bit masks per 32 positions ('-', match-like = M or R, 'R'), from
// K2b<false> directly or from chars_to_masks_kernel.  N = not '-' and not a
// separator.  The reference's state machine is equivalent to:
//   * two aligned characters with only '-' between them belong to the same
//     segment iff at most max_gap_len '-' separate them (D = max_gap_len + 1);
//     leading / trailing '-' runs never belong to a segment (format.rs:180-184);
//   * START[p] = N[p] and no N in [p-D, p-1] of the same query; END likewise;
//   * a segment [ps, pe]: matches / mismatches / jumps / gap opens are counts of
//     masks over [ps, pe], gap_bases = length - #N.  jumps counts 'R' preceded
//     by 'R'; gap opens counts '-' preceded by N (a '-' run strictly inside).
// With exclusive prefix counts per word every START / END test and every record is O(1).  The counts are kept in
// two levels (before the word inside its block of 256 words, before the block); each of the two counting kernels
// scans inside its blocks, and its last block to finish scans the block totals:
//   rle_word_counts -> rle_mark -> rle_query_offsets, rle_records.
// Records are ordered by position, hence by query; the slot of a segment is the
// number of STARTs before it.
// ---------------------------------------------------------------------------
struct RleRecord {
    uint64_t start, end, matches, mismatches, jumps, gap_bases, gap_opens;  // == kbo_rle
};

struct RleCounts {  // per word: counts before the word inside its block of RLE_BLOCK words; per block: before the block
    uint32_t n, m, j, go;
};

enum { RLE_BLOCK = 256 };  // words (= threads) per block of the two scanning kernels

struct RleParams {
    const uint32_t* gap;    // '-'                       (bit per padded position)
    const uint32_t* match;  // 'M' or 'R'
    const uint32_t* rr;     // 'R'
    const uint32_t* sep;    // QueryView::sep
    const uint32_t* wq;     // QueryView::wq
    uint64_t n_words;       // words covered (a multiple of 32, >= ceil(Lp / 32))
    uint64_t n_blocks;      // ceil(n_words / RLE_BLOCK)
    const uint64_t* offsets;
    uint64_t nq;
    uint32_t window;        // D = max_gap_len + 1
    uint32_t* jump;         // rle_word_counts out
    uint32_t* gopen;
    RleCounts* cnt;         // n_words entries (block-relative) ...
    RleCounts* cnt_blk;     // ... + n_blocks + 1 entries (before each block; last = grand total)
    uint32_t* start;        // rle_mark out
    uint32_t* end;
    uint64_t* cse;          // n_words entries: #START | #END << 32 before the word inside its block ...
    uint64_t* cse_blk;      // ... + n_blocks + 1 entries
    unsigned int* tickets;  // two zero-initialised counters (self-resetting) electing the last block of a launch
    uint64_t* rle_offsets;  // nq + 1 (device memory, or page-locked host memory mapped into the device)
    RleRecord* out;         // likewise
    uint64_t cap;           // records with a slot >= cap are dropped (the count is still exact)
    const uint64_t* base_in;  // optional: number of records of the sub-batches before this one (device counter);
    uint64_t* total_out;      // optional: receives base + the records of this sub-batch
    uint32_t write_first;     // != 0: rle_offsets[0] is written too (0 for the sub-batches after the first)
};

__device__ __forceinline__ uint32_t low_mask(uint32_t b) { return b ? (~0u >> (32 - b)) : 0u; }  // bits below b (b < 32)

// old K2 path / plain alignments: characters (unpadded) -> the three masks in padded space; thread per position
__global__ void chars_to_masks_kernel(const uint8_t* __restrict__ aln, uint64_t off0, const uint32_t* __restrict__ sep,
                                      const uint32_t* __restrict__ wq, uint64_t n_words, uint32_t* __restrict__ gap,
                                      uint32_t* __restrict__ match, uint32_t* __restrict__ rr) {
    const uint64_t pp = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // grid covers n_words * 32 exactly
    const uint64_t w = pp >> 5;
    const uint32_t b = (uint32_t)(pp & 31);
    uint8_t ch = 0;
    if (w < n_words) {
        const uint32_t sw = __ldg(sep + w);
        if (!((sw >> b) & 1u)) ch = aln[off0 + pp - __ldg(wq + w) - __popc(sw & low_mask(b))];
    }
    const uint32_t g = __ballot_sync(0xffffffffu, ch == '-');
    const uint32_t m = __ballot_sync(0xffffffffu, ch == 'M' || ch == 'R');
    const uint32_t r = __ballot_sync(0xffffffffu, ch == 'R');
    if (b == 0 && w < n_words) {
        gap[w] = g;
        match[w] = m;
        rr[w] = r;
    }
}

__device__ __forceinline__ uint32_t rle_nongap(const RleParams& p, uint64_t w) {
    return ~__ldg(p.sep + w) & ~__ldg(p.gap + w);
}

// Exclusive scan of N 32-bit counters over the RLE_BLOCK threads of a block (all threads must call).
// v: in = the thread's counts, out = counts of the threads before it; total = counts of the whole block.
template <int N>
__device__ __forceinline__ void rle_block_scan(uint32_t (&v)[N], uint32_t (&total)[N]) {
    __shared__ uint32_t warp_sum[RLE_BLOCK / 32][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) inc[i] = v[i];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc[i], off);
            if (lane >= off) inc[i] += o;
        }
    }
    __syncthreads();  // warp_sum may still be read by the previous use
    if (lane == 31)
        for (int i = 0; i < N; ++i) warp_sum[warp][i] = inc[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint32_t before = 0, all = 0;
        for (int w = 0; w < RLE_BLOCK / 32; ++w) {
            const uint32_t x = warp_sum[w][i];
            if (w < warp) before += x;
            all += x;
        }
        v[i] = before + inc[i] - v[i];
        total[i] = all;
    }
}

// One block turns the per-block totals (N counters of 32 bits each per entry) into "before the block", with the
// grand total at [n_blocks].  The totals were written by other blocks of the same launch: they are read at L2.
template <int N, typename T>
__device__ __forceinline__ void rle_scan_block_totals(T* blk, uint64_t n_blocks) {
    uint32_t run[N];
    for (int i = 0; i < N; ++i) run[i] = 0;
    for (uint64_t base = 0; base < n_blocks; base += RLE_BLOCK) {
        const uint64_t b = base + threadIdx.x;
        uint32_t v[N], tot[N];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(blk + (b < n_blocks ? b : 0));
        for (int i = 0; i < N; ++i) v[i] = b < n_blocks ? __ldcg(src + i) : 0u;
        rle_block_scan<N>(v, tot);
        if (b < n_blocks) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(blk + b);
            for (int i = 0; i < N; ++i) dst[i] = run[i] + v[i];
        }
        for (int i = 0; i < N; ++i) run[i] += tot[i];
    }
    if (threadIdx.x == 0) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(blk + n_blocks);
        for (int i = 0; i < N; ++i) dst[i] = run[i];
    }
}

// The block that finishes last scans the block totals.  The ticket counter resets itself for the next launch.
template <int N, typename T>
__device__ __forceinline__ void rle_finish_blocks(T* blk, uint64_t n_blocks, unsigned int* ticket) {
    // (only thread 0 wrote this block's totals -- just before the call -- so only it has to order that write before
    // its ticket; a fence in all 256 threads showed up as 3.5 warp-cycles of membar stall per issued instruction)
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    rle_scan_block_totals<N>(blk, n_blocks);
    if (threadIdx.x == 0) *ticket = 0;
}

// thread per word: jump / gap-open masks, the four counts before the word, block totals.
// MARKS (window == 1, i.e. max_gap_len == 0: a segment is a maximal run of aligned characters, so START / END need
// no prefix counts): also the START / END masks and their counts -- what rle_mark_kernel does in a second launch
// for the gapped form.
template <bool MARKS>
__global__ void __launch_bounds__(RLE_BLOCK, MARKS ? 6 : 5) rle_word_counts_kernel(RleParams p) {  // 40 / 48 registers, no spills
    const uint64_t w = (uint64_t)blockIdx.x * RLE_BLOCK + threadIdx.x;
    uint32_t c[MARKS ? 6 : 4], tot[MARKS ? 6 : 4];
#pragma unroll
    for (int i = 0; i < (MARKS ? 6 : 4); ++i) c[i] = 0;
    if (w < p.n_words) {
        const uint32_t N = rle_nongap(p, w), R = __ldg(p.rr + w), Gp = __ldg(p.gap + w);
        uint32_t n31 = 0, r31 = 0;
        if (w > 0) {
            n31 = rle_nongap(p, w - 1) >> 31;
            r31 = __ldg(p.rr + w - 1) >> 31;
        }
        const uint32_t J = R & ((R << 1) | r31);    // 'R' preceded by 'R'          (format.rs:175-177)
        const uint32_t GO = Gp & ((N << 1) | n31);  // '-' preceded by an aligned character
        p.jump[w] = J;
        p.gopen[w] = GO;
        c[0] = __popc(N);
        c[1] = __popc(__ldg(p.match + w));
        c[2] = __popc(J);
        c[3] = __popc(GO);
        if (MARKS) {
            const uint32_t n0 = w + 1 < p.n_words ? rle_nongap(p, w + 1) & 1u : 0u;
            const uint32_t st = N & ~((N << 1) | n31);  // first / last character of a run of aligned characters
            const uint32_t en = N & ~((N >> 1) | (n0 << 31));
            p.start[w] = st;
            p.end[w] = en;
            c[4] = __popc(st);
            c[5] = __popc(en);
        }
    }
    {   // a word adds at most 32 to a counter and a block has 256 words: two counters share one 32-bit scan value
        // (half as many shuffles, and the registers that kept this kernel at 4 blocks per SM)
        static_assert(RLE_BLOCK * 32 < 65536, "packed block scan: a block total must fit 16 bits");
        constexpr int NP = MARKS ? 3 : 2;
        uint32_t pk[NP], tp[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) pk[i] = c[2 * i] | (c[2 * i + 1] << 16);
        rle_block_scan<NP>(pk, tp);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            c[2 * i] = pk[i] & 0xffffu; c[2 * i + 1] = pk[i] >> 16;
            tot[2 * i] = tp[i] & 0xffffu; tot[2 * i + 1] = tp[i] >> 16;
        }
    }
    if (w < p.n_words) {
        RleCounts before = {c[0], c[1], c[2], c[3]};
        p.cnt[w] = before;
        if (MARKS) p.cse[w] = (uint64_t)c[4] | ((uint64_t)c[5] << 32);
    }
    if (threadIdx.x == 0) {
        RleCounts t = {tot[0], tot[1], tot[2], tot[3]};
        p.cnt_blk[blockIdx.x] = t;
        if (MARKS) p.cse_blk[blockIdx.x] = (uint64_t)tot[4] | ((uint64_t)tot[5] << 32);
    }
    if (MARKS) {  // one election, both arrays (thread 0 wrote the totals above and orders them before its ticket)
        __shared__ bool last;
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(p.tickets, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!last) return;
        __threadfence();
        rle_scan_block_totals<4>(p.cnt_blk, p.n_blocks);
        rle_scan_block_totals<2>(p.cse_blk, p.n_blocks);
        if (threadIdx.x == 0) *p.tickets = 0;
    } else {
        rle_finish_blocks<4>(p.cnt_blk, p.n_blocks, p.tickets);
    }
}

// number of aligned characters before padded position x
__device__ __forceinline__ uint32_t rle_pref_n(const RleParams& p, uint64_t x) {
    const uint64_t w = x >> 5;
    if (w >= p.n_words) return p.cnt_blk[p.n_blocks].n;
    uint32_t c = p.cnt_blk[w / RLE_BLOCK].n + p.cnt[w].n;
    const uint32_t b = (uint32_t)(x & 31);
    if (b) c += __popc(rle_nongap(p, w) & low_mask(b));
    return c;
}
__device__ __forceinline__ uint64_t rle_query_start(const RleParams& p, uint64_t q) {  // padded position of its first base
    return p.offsets[q] - p.offsets[0] + q;
}

// thread per word: START / END masks, their counts before the word, block totals
__global__ void __launch_bounds__(RLE_BLOCK) rle_mark_kernel(RleParams p) {
    const uint64_t w = (uint64_t)blockIdx.x * RLE_BLOCK + threadIdx.x;
    uint32_t st = 0, en = 0;
    if (w < p.n_words) {
        const uint32_t N = rle_nongap(p, w);
        const uint32_t n31 = w > 0 ? rle_nongap(p, w - 1) >> 31 : 0u;
        const uint32_t n0 = w + 1 < p.n_words ? rle_nongap(p, w + 1) & 1u : 0u;
        st = N & ~((N << 1) | n31);  // first character of a run of aligned characters
        en = N & ~((N >> 1) | (n0 << 31));
        if (p.window > 1 && (st | en)) {
            const uint32_t S = __ldg(p.sep + w);
            const uint64_t q0 = __ldg(p.wq + w);
            const uint64_t D = p.window;
            uint32_t keep = 0;
            for (uint32_t rem = st; rem; rem &= rem - 1) {
                const uint32_t b = (uint32_t)__ffs((int)rem) - 1u;
                const uint64_t pos = w * 32 + b;
                const uint64_t qs = rle_query_start(p, q0 + __popc(S & low_mask(b)));
                uint64_t lo = pos > D ? pos - D : 0;
                if (lo < qs) lo = qs;
                if (rle_pref_n(p, pos) == rle_pref_n(p, lo)) keep |= 1u << b;
            }
            st = keep;
            keep = 0;
            for (uint32_t rem = en; rem; rem &= rem - 1) {
                const uint32_t b = (uint32_t)__ffs((int)rem) - 1u;
                const uint64_t pos = w * 32 + b;
                const uint64_t qe = rle_query_start(p, q0 + __popc(S & low_mask(b)) + 1) - 1;  // its separator
                uint64_t hi = pos + D + 1;  // exclusive
                if (hi > qe) hi = qe;
                if (rle_pref_n(p, hi) == rle_pref_n(p, pos + 1)) keep |= 1u << b;
            }
            en = keep;
        }
        p.start[w] = st;
        p.end[w] = en;
    }
    uint32_t c[2] = {(uint32_t)__popc(st), (uint32_t)__popc(en)}, tot[2];
    rle_block_scan<2>(c, tot);
    if (w < p.n_words) p.cse[w] = (uint64_t)c[0] | ((uint64_t)c[1] << 32);
    if (threadIdx.x == 0) p.cse_blk[blockIdx.x] = (uint64_t)tot[0] | ((uint64_t)tot[1] << 32);
    rle_finish_blocks<2>(p.cse_blk, p.n_blocks, p.tickets + 1);
}

// STARTs (low half) / ENDs (high half) before word w
__device__ __forceinline__ uint64_t rle_cse(const RleParams& p, uint64_t w) {
    return p.cse_blk[w / RLE_BLOCK] + p.cse[w];
}

__device__ __forceinline__ RleCounts rle_pref_all(const RleParams& p, uint64_t x) {
    const uint64_t w = x >> 5;
    if (w >= p.n_words) return p.cnt_blk[p.n_blocks];
    const RleCounts blk = p.cnt_blk[w / RLE_BLOCK];
    RleCounts c = p.cnt[w];
    c.n += blk.n; c.m += blk.m; c.j += blk.j; c.go += blk.go;
    const uint32_t b = (uint32_t)(x & 31);
    if (b) {
        const uint32_t lm = low_mask(b);
        c.n += __popc(rle_nongap(p, w) & lm);
        c.m += __popc(__ldg(p.match + w) & lm);
        c.j += __popc(p.jump[w] & lm);
        c.go += __popc(p.gopen[w] & lm);
    }
    return c;
}

// Last K4 launch: thread t < n_words writes one record per END bit of word t; thread t <= nq writes the record
// offset of query t (records before the query = STARTs before its first base).  `base` = records of the sub-batches
// that precede this one in the same host call (0 when there are none).
__global__ void rle_finish_kernel(RleParams p) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t base = p.base_in ? *p.base_in : 0ull;
    if (t <= p.nq) {
        if (t == p.nq) {
            const uint64_t total = base + (uint32_t)p.cse_blk[p.n_blocks];
            p.rle_offsets[t] = total;
            if (p.total_out) *p.total_out = total;
        } else if (t > 0 || p.write_first) {
            const uint64_t x = rle_query_start(p, t);
            const uint32_t b = (uint32_t)(x & 31);
            p.rle_offsets[t] = base + (uint32_t)rle_cse(p, x >> 5) + (b ? __popc(p.start[x >> 5] & low_mask(b)) : 0);
        }
    }
    const uint64_t w = t;
    if (w >= p.n_words) return;
    uint32_t rem = p.end[w];
    if (!rem) return;
    const uint32_t S = __ldg(p.sep + w);
    const uint64_t q0 = __ldg(p.wq + w);
    uint32_t slot = (uint32_t)(rle_cse(p, w) >> 32);
    for (; rem; rem &= rem - 1, ++slot) {
        if (base + slot >= p.cap) break;
        const uint32_t b = (uint32_t)__ffs((int)rem) - 1u;
        const uint64_t pe = w * 32 + b;
        const uint64_t qs = rle_query_start(p, q0 + __popc(S & low_mask(b)));
        // the START with the same rank: last word at or after the query's first whose prefix count is <= slot
        uint64_t lo = qs >> 5, hi = w;
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) >> 1;
            if ((uint32_t)rle_cse(p, mid) <= slot) lo = mid; else hi = mid - 1;
        }
        uint32_t sb = p.start[lo];
        for (uint32_t skip = slot - (uint32_t)rle_cse(p, lo); skip; --skip) sb &= sb - 1;
        const uint64_t ps = lo * 32 + (uint32_t)__ffs((int)sb) - 1u;
        const RleCounts a = rle_pref_all(p, ps), z = rle_pref_all(p, pe + 1);
        const uint32_t n = z.n - a.n, m = z.m - a.m;
        RleRecord rec;
        rec.start = ps - qs;
        rec.end = pe + 1 - qs;
        rec.matches = m;
        rec.mismatches = n - m;
        rec.jumps = z.j - a.j;
        rec.gap_bases = (pe + 1 - ps) - n;
        rec.gap_opens = z.go - a.go;
        p.out[base + slot] = rec;
    }
}

// ---------------------------------------------------------------------------
// K3: translate_ms_vec (translate.rs:263-293) for one arbitrary i64 vector.
// ---------------------------------------------------------------------------
__global__ void translate_i64_kernel(const int64_t* __restrict__ d, uint64_t n, uint32_t k, uint32_t thr,
                                     uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t T = (int64_t)thr;
    const int64_t cur = d[i];
    const bool last = (i == n - 1);
    const int64_t next = last ? cur : d[i + 1];
    const int64_t prevv = i > 0 ? d[i - 1] : 0;
    const int64_t prev = i > 1 ? prevv : (int64_t)k;
    const bool trig = !last && cur > T && next > 0 && next < T;
    const bool trig_prev = i > 1 && !last && prevv > T && cur > 0 && cur < T;
    uint8_t ch;
    if (trig || trig_prev) ch = 'R';
    else if (cur <= 0) ch = (next == 1 && prev > 0) ? 'X' : '-';
    else ch = 'M';
    out[i] = ch;
}

// ---------------------------------------------------------------------------
// Refinement helpers on the device.
//
// variant_candidates_kernel: the candidate scan of call_variants (variant_calling.rs:268-272) on K1's (d, l, r) of
// ONE query: position i is a candidate when ms[i] < ms[i-1], ms[i-1] >= thr and ms[i] < thr, and some j in
// (i, min(i + k + 1, len)) has ms[j] >= thr with a one-node interval; the first such j is reported.  Thread per
// position; candidates are appended through an atomic counter (the host sorts the few of them by i).
// relative_to_ref_kernel: format::relative_to_ref (format.rs:266-287), byte per thread.
// ---------------------------------------------------------------------------
struct VariantCandidate {
    uint32_t i, j, node, pad;  // query position of the drop, position of the unique match, its colex rank
};

__global__ void variant_candidates_kernel(const uint8_t* __restrict__ d, const uint32_t* __restrict__ l,
                                          const uint32_t* __restrict__ r, uint64_t len, uint32_t k, uint32_t thr,
                                          VariantCandidate* __restrict__ out, uint32_t cap, unsigned int* __restrict__ count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= len) return;
    const uint32_t cur = d[i], prev = d[i - 1];
    if (!(cur < prev && prev >= thr && cur < thr)) return;
    const uint64_t stop = i + k + 1 < len ? i + k + 1 : len;
    for (uint64_t j = i + 1; j < stop; ++j) {
        if (d[j] >= thr && r[j] - l[j] == 1u) {
            const unsigned int slot = atomicAdd(count, 1u);
            if (slot < cap) out[slot] = VariantCandidate{(uint32_t)i, (uint32_t)j, l[j], 0u};
            return;
        }
    }
}

__global__ void relative_to_ref_kernel(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ aln, uint64_t n,
                                       uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t a = aln[i];
    uint8_t o = a;
    if (a == 'M' || a == 'R' || a == 'I') o = ref[i];
    else if (a == 'X' || a == 'D' || a == '-') o = '-';
    out[i] = o;
}

// ---------------------------------------------------------------------------
// G: derandomize_ms_vec (derandomize.rs:269-288) for one ARBITRARY MS vector
// (values <= k, no monotonicity assumed), exact i64 output.
//
// With b_i = ms_i - i for eligible i (ms_i > thr; the last position is always a
// "set" with value ms>thr?ms:0), w_i = out_i - i is a running maximum with
// hysteresis:  w_i = b_i if (w_{i+1} <= b_i - 2, or ms_i == k) else w_{i+1}.
// Let M_i = max{b_j : j >= i eligible}.  Then w_i = M_i - eps_i with eps in {0,1}:
//   Delta = b_i - M_{i+1}:  ms_i == k or Delta >= 2 -> eps_i = 0;
//   Delta == 1 -> eps_i = 1 - eps_{i+1};  Delta <= 0 -> eps_i = eps_{i+1}.
// So the recurrence is a suffix-max scan followed by a segmented parity scan
// (proof in DESIGN.md).  Five small passes over tiles of G_TILE elements:
//   g1 tile max -> g2 scan of tile maxima -> g3 tile parity transform ->
//   g4 scan of transforms -> g5 apply.
// ---------------------------------------------------------------------------
enum { G_TILE = 1024, G_THREADS = 256 };
#define KBO_NEG_INF (-(1ll << 62))

__device__ __forceinline__ int64_t g_b(const uint64_t* ms, uint64_t i, uint64_t n, uint32_t thr, uint32_t k) {
    // candidate value b_i (or -inf when position i cannot start a run)
    const uint64_t v = ms[i];
    if (i == n - 1) return (v > thr ? (int64_t)v : 0) - (int64_t)i;
    return (v > thr || v == k) ? (int64_t)v - (int64_t)i : KBO_NEG_INF;
}

__global__ void __launch_bounds__(G_THREADS) g1_tile_max_kernel(const uint64_t* __restrict__ ms, uint64_t n,
                                                                uint32_t k, uint32_t thr,
                                                                int64_t* __restrict__ tile_max) {
    __shared__ int64_t red[G_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * G_TILE;
    int64_t mx = KBO_NEG_INF;
    for (uint32_t j = threadIdx.x; j < G_TILE; j += G_THREADS) {
        const uint64_t i = base + j;
        if (i < n) {
            const int64_t b = g_b(ms, i, n, thr, k);
            mx = b > mx ? b : mx;
        }
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (uint32_t sft = G_THREADS / 2; sft > 0; sft >>= 1) {
        if (threadIdx.x < sft) red[threadIdx.x] = red[threadIdx.x] > red[threadIdx.x + sft] ? red[threadIdx.x] : red[threadIdx.x + sft];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_max[blockIdx.x] = red[0];
}

// exclusive suffix max over tiles: m_in[t] = max(tile_max[t+1..]); single thread (n_tiles is small)
__global__ void g2_scan_max_kernel(const int64_t* __restrict__ tile_max, uint64_t n_tiles, int64_t* __restrict__ m_in) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t acc = KBO_NEG_INF;
    for (uint64_t t = n_tiles; t-- > 0;) {
        m_in[t] = acc;
        acc = tile_max[t] > acc ? tile_max[t] : acc;
    }
}

// One thread block walks its tile right to left in 4 strips of 256 positions; inside a strip
// thread j owns position j.  M_{i+1} comes from a block suffix-max scan, the parity transform
// from a block suffix composition.  `apply` = false: write the tile's transform; true: write out.
__device__ __forceinline__ void g_block_suffix_max(int64_t* sh, int64_t& v) {
    // inclusive suffix max over the block (thread j sees max over threads >= j)
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t off = 1; off < G_THREADS; off <<= 1) {
        int64_t o = (threadIdx.x + off < G_THREADS) ? sh[threadIdx.x + off] : KBO_NEG_INF;
        __syncthreads();
        if (o > v) v = o;
        sh[threadIdx.x] = v;
        __syncthreads();
    }
}
__device__ __forceinline__ void g_block_suffix_par(uint32_t* sh, uint32_t& f) {
    sh[threadIdx.x] = f;
    __syncthreads();
    for (uint32_t off = 1; off < G_THREADS; off <<= 1) {
        uint32_t o = (threadIdx.x + off < G_THREADS) ? sh[threadIdx.x + off] : 0u;
        __syncthreads();
        f = par_compose(f, o);
        sh[threadIdx.x] = f;
        __syncthreads();
    }
}

template <bool APPLY>
__global__ void __launch_bounds__(G_THREADS) g35_tile_kernel(const uint64_t* __restrict__ ms, uint64_t n, uint32_t k,
                                                            uint32_t thr, const int64_t* __restrict__ m_in,
                                                            uint32_t* __restrict__ tile_par,
                                                            const uint32_t* __restrict__ eps_in,
                                                            int64_t* __restrict__ out) {
    __shared__ int64_t shm[G_THREADS];
    __shared__ uint32_t shp[G_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * G_TILE;
    int64_t M_right = m_in[blockIdx.x];             // max of everything right of the current strip
    uint32_t F_right = APPLY ? (2u | (eps_in[blockIdx.x] & 1u)) : 0u;  // transform of everything right of the strip
    for (int strip = G_TILE / G_THREADS - 1; strip >= 0; --strip) {
        const uint64_t i = base + (uint64_t)strip * G_THREADS + threadIdx.x;
        const bool valid = i < n;
        const int64_t b = valid ? g_b(ms, i, n, thr, k) : KBO_NEG_INF;
        int64_t Mi = b;  // becomes inclusive suffix max within the strip
        g_block_suffix_max(shm, Mi);
        // exclusive: max over positions > i
        int64_t Mex = (threadIdx.x + 1 < G_THREADS) ? shm[threadIdx.x + 1] : KBO_NEG_INF;
        if (M_right > Mex) Mex = M_right;
        const int64_t Minc = Mi > M_right ? Mi : M_right;
        uint32_t op = OP_KEEP;
        if (valid && b != KBO_NEG_INF) {
            if (i == n - 1 || ms[i] == k) op = OP_SET0;
            else {
                const int64_t delta = (Mex == KBO_NEG_INF) ? 2 : b - Mex;
                op = delta >= 2 ? OP_SET0 : (delta == 1 ? OP_TOGGLE : OP_KEEP);
            }
        }
        uint32_t f = (op == OP_SET0) ? 2u : op;
        g_block_suffix_par(shp, f);  // f = composite over positions >= i within the strip
        if (APPLY) {
            if (valid) {
                const uint32_t eps = par_apply(par_compose(f, F_right), 0u);
                out[i] = Minc + (int64_t)i - (int64_t)eps;
            }
        }
        // fold this strip into the running right-hand state (thread 0 holds the strip totals)
        const uint32_t strip_f = shp[0];
        const int64_t strip_m = shm[0];
        __syncthreads();
        F_right = par_compose(strip_f, F_right);
        if (strip_m > M_right) M_right = strip_m;
    }
    if (!APPLY && threadIdx.x == 0) tile_par[blockIdx.x] = F_right;
}

// eps entering each tile from the right: eps_in[t] = (T_{t+1} o T_{t+2} o ...)(0)
__global__ void g4_scan_par_kernel(const uint32_t* __restrict__ tile_par, uint64_t n_tiles, uint32_t* __restrict__ eps_in) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t acc = 0;  // identity
    for (uint64_t t = n_tiles; t-- > 0;) {
        eps_in[t] = par_apply(acc, 0u);
        acc = par_compose(tile_par[t], acc);
    }
}

}  // namespace kbo_b200
