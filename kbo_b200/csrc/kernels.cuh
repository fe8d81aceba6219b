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
// ---------------------------------------------------------------------------
struct IndexView {
    const uint64_t* rank;
    uint32_t rank_stride;  // words per row (< 2^28 for n_sets < 2^32)
    const uint8_t* lcs;
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
// K0: pack.  One thread per 32 padded positions.  The bases of a word are read
// eight at a time (aligned 8-byte loads, shifted together when the source is
// misaligned) and converted four per 32-bit register with byte-parallel
// arithmetic; a word that holds separators is cut into its runs of bases.
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

// n (1..8) bytes starting at src, little endian; touches only the aligned 8-byte words that hold them
__device__ __forceinline__ uint64_t load_bytes8(const uint8_t* src, uint32_t n) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)7;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 7u);
    uint64_t v = *reinterpret_cast<const uint64_t*>(a) >> (8 * mis);
    if (mis + n > 8) v |= *reinterpret_cast<const uint64_t*>(a + 8) << (64 - 8 * mis);
    return v;
}

__global__ void pack_queries_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ offsets,
                                    uint64_t nq, QueryView qv, uint64_t* __restrict__ pack,
                                    uint32_t* __restrict__ inv, uint32_t* __restrict__ sep,
                                    uint32_t* __restrict__ wq) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= qv.n_words) return;
    const uint64_t pp0 = w * 32;
    // q = first query whose separator is at or after pp0
    uint64_t lo = 0, hi = nq;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (sep_pos(offsets, mid) < pp0) lo = mid + 1; else hi = mid;
    }
    uint64_t q = lo;
    wq[w] = (uint32_t)q;
    const uint64_t off0 = offsets[0];
    uint64_t pk = 0;
    uint32_t iv = 0, sp = 0;
    uint32_t j = 0;
    while (j < 32) {
        if (q >= nq) {  // past the last separator: the tail is all separator
            iv |= ~0u << j;
            sp |= ~0u << j;
            break;
        }
        const uint64_t pp = pp0 + j;
        const uint64_t next_sep = sep_pos(offsets, q);
        if (pp == next_sep) {
            iv |= 1u << j;
            sp |= 1u << j;
            ++q;
            ++j;
            continue;
        }
        uint64_t n64 = next_sep - pp;
        uint32_t n = 32 - j;
        if (n64 < n) n = (uint32_t)n64;
        if (n > 8) n = 8;
        const uint64_t v = load_bytes8(ascii + off0 + pp - q, n);
        uint32_t c0, c1, i0, i1;
        swar_pack4((uint32_t)v, c0, i0);
        swar_pack4((uint32_t)(v >> 32), c1, i1);
        const uint32_t keep = (1u << n) - 1u;  // drop what was read past the run
        const uint32_t code16 = (c0 | (c1 << 8)) & ((1u << (2 * n)) - 1u);
        const uint32_t inv8 = (i0 | (i1 << 4)) & keep;
        pk |= (uint64_t)code16 << (2 * j);
        iv |= inv8 << j;
        j += n;
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
    uint32_t probe_iters;  // probe iterations per contraction phase (>= 1)
    uint32_t flags;        // experiment switches: bit0 = population count on the ALU pipe instead of POPC (XU pipe)
    uint64_t n_chunks;
    uint8_t* ms;         // padded space, 1 byte per position
    uint32_t* l_out;     // optional (INTERVALS)
    uint32_t* r_out;
    unsigned long long* counters;  // optional (COUNT)
};

__device__ __forceinline__ uint32_t popc_alu(uint32_t x) {
    x = x - ((x >> 1) & 0x55555555u);
    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
    x = (x + (x >> 4)) & 0x0f0f0f0fu;
    return (x * 0x01010101u) >> 24;
}

__device__ __forceinline__ uint64_t lcs_lt_mask64(uint64_t w, uint64_t t_rep) {
    // 0x80 in every byte of w that is < t (bytes and t are < 128)
    const uint64_t H = 0x8080808080808080ull;
    return ~((w | H) - t_rep) & H;
}

// Loop structure: `probe_iters` probe iterations (lanes whose extension failed sit out the rest of
// the group), then ONE contraction phase executed together by every lane that failed.  The divergent
// contraction code (~40 % of the static loop body, used by ~10 % of the lanes per iteration) is thus
// issued once per group instead of once per iteration.  All position arithmetic is 32-bit.
template <bool INTERVALS, bool COUNT>
__global__ void __launch_bounds__(256, 8) ms_kernel(MsParams p) {
    __shared__ __align__(16) uint8_t ms_stage[256 * 36];
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cnt_att = 0, cnt_split = 0, cnt_con = 0, cnt_extra = 0, cnt_proc = 0, cnt_emit = 0;
    unsigned long long cnt_att_e = 0, cnt_split_e = 0, cnt_con_e = 0, cnt_extra_e = 0;
    if (g < p.n_chunks) {
        const uint32_t n = p.ix.n, k = p.ix.k;
        const uint64_t start = g * p.chunk_len;
        const uint64_t remain = p.q.Lp - start;
        const uint32_t len = remain < p.chunk_len ? (uint32_t)remain : p.chunk_len;
        const uint32_t warm = start >= (uint64_t)(k - 1) ? k - 1 : (uint32_t)start;  // warm-up bases
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
        uint32_t l = 0, r = n, d = 0;
        bool failed = false;
        // emitted MS bytes are staged in shared memory (36-byte stride per lane: conflict-free word access) and
        // flushed as two 16-byte stores per 32 positions; chunk starts are multiples of 32, so flushes are aligned
        uint8_t* const stg = ms_stage + threadIdx.x * 36u;
        while (bp < bp_end) {
            // ---- probe phase ------------------------------------------------------------------
#pragma unroll 1
            for (uint32_t it = 0; it < p.probe_iters; ++it) {
                if (failed || bp >= bp_end) continue;
                bool advance;
                if (iw & 1u) {
                    l = 0; r = n; d = 0;
                    advance = true;
                } else {
                    const uint32_t rowoff = ((uint32_t)qw & 3u) * p.ix.rank_stride;  // 32-bit word index
                    const uint32_t bl = l >> 5, br = r >> 5;
                    const uint64_t wl = __ldg(p.ix.rank + (rowoff + bl));
                    const uint64_t wr = (br == bl) ? wl : __ldg(p.ix.rank + (rowoff + br));
                    const uint32_t ml = (uint32_t)wl & ((1u << (l & 31)) - 1u), mr = (uint32_t)wr & ((1u << (r & 31)) - 1u);
                    uint32_t nl, nr;
                    if (p.flags & 1u) {
                        nl = (uint32_t)(wl >> 32) + popc_alu(ml);
                        nr = (uint32_t)(wr >> 32) + popc_alu(mr);
                    } else {
                        nl = (uint32_t)(wl >> 32) + __popc(ml);
                        nr = (uint32_t)(wr >> 32) + __popc(mr);
                    }
                    if (COUNT) {
                        const bool sp = (bl >> 2) != (br >> 2);
                        ++cnt_att; cnt_split += sp;
                        if (bp >= bp_emit) { ++cnt_att_e; cnt_split_e += sp; }
                    }
                    if (nl < nr) {
                        l = nl; r = nr;
                        d = d + 1 < k ? d + 1 : k;
                        advance = true;
                    } else if (d == 0) {
                        advance = true;  // state stays (0,[0,n))
                    } else {
                        advance = false;
                        failed = true;
                    }
                }
                if (advance) {
                    if (COUNT) ++cnt_proc;
                    if (bp >= bp_emit) {
                        if (COUNT) ++cnt_emit;
                        stg[bp & 31u] = (uint8_t)d;
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
            // ---- contraction phase: contract_left to the largest target that changes the interval ----
            if (failed) {
                failed = false;
                // 16 LCS cells on each side are requested at once (two aligned 8-byte words per side), so the
                // scans below almost never need a further, dependent load
                const uint64_t* __restrict__ L8 = reinterpret_cast<const uint64_t*>(p.ix.lcs);
                const uint32_t wl_i = l >> 3, wr_i = r >> 3;
                const uint64_t Wl1 = L8[wl_i];
                const uint64_t Wl0 = L8[wl_i ? wl_i - 1 : 0];
                const uint64_t Wr0 = (wr_i == wl_i) ? Wl1 : L8[wr_i];
                const uint64_t Wr1 = L8[wr_i + 1];  // the zero padding past n makes this readable
                if (COUNT) {
                    ++cnt_con; cnt_extra += (wr_i != wl_i);
                    if (bp >= bp_emit) { ++cnt_con_e; cnt_extra_e += (wr_i != wl_i); }
                }
                const uint32_t vl = (uint32_t)(Wl1 >> (8 * (l & 7))) & 0xffu;
                const uint32_t vr = (uint32_t)(Wr0 >> (8 * (r & 7))) & 0xffu;  // LCS[n] reads the zero padding
                uint32_t t = vl > vr ? vl : vr;
                if (t > d - 1) t = d - 1;  // cannot happen for a maximal interval; keeps the literal bound
                if (t == 0) {
                    l = 0; r = n; d = 0;
                } else {
                    d = t;
                    const uint64_t T = (uint64_t)t * 0x0101010101010101ull;
                    // left: largest q <= l with LCS[q] < t (LCS[0] = 0 stops the scan)
                    uint64_t m = lcs_lt_mask64(Wl1, T);
                    if ((l & 7) != 7) m &= (1ull << (8 * ((l & 7) + 1))) - 1ull;
                    uint32_t b = wl_i;
                    if (m == 0) {
                        m = lcs_lt_mask64(Wl0, T);
                        --b;
                        while (m == 0) {
                            --b;
                            m = lcs_lt_mask64(L8[b], T);
                            if (COUNT) { ++cnt_extra; cnt_extra_e += (bp >= bp_emit); }
                        }
                    }
                    l = (b << 3) + ((63 - __clzll((long long)m)) >> 3);
                    // right: smallest q >= r with LCS[q] < t (the zero padding at n stops the scan)
                    m = lcs_lt_mask64(Wr0, T) & (~0ull << (8 * (r & 7)));
                    b = wr_i;
                    if (m == 0) {
                        m = lcs_lt_mask64(Wr1, T);
                        ++b;
                        while (m == 0) {
                            ++b;
                            m = lcs_lt_mask64(L8[b], T);
                            if (COUNT) { ++cnt_extra; cnt_extra_e += (bp >= bp_emit); }
                        }
                    }
                    r = (b << 3) + ((__ffsll((long long)m) - 1) >> 3);
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
    uint64_t n_tiles;
};

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

// Clamped derandomized value of position e (first position right of a tile).  Warp-uniform.
__device__ __forceinline__ uint32_t lookahead_c(const TrParams& p, uint64_t e, int lane) {
    if (sep_bit(p.q, (int64_t)e)) return 0u;  // tile ends exactly at a query end: nothing enters
    uint32_t dist = 0;       // N positions skipped before the first source
    bool in_run = false;     // source found, waiting for the first SET0
    uint32_t src_val = 0, src_dist = 0, parity = 0;
    for (uint64_t base = e;; base += 32) {
        const uint64_t pp = base + lane;
        const uint32_t m0 = p.ms[pp], m1 = p.ms[pp + 1];
        const bool s0 = sep_bit(p.q, (int64_t)pp), s1 = sep_bit(p.q, (int64_t)pp + 1);
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
    uint32_t m[K2_PER_LANE + 1];
    {
        const uint4 v = *reinterpret_cast<const uint4*>(p.ms + P);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < K2_PER_LANE; ++t) m[t] = (w[t >> 2] >> (8 * (t & 3))) & 0xffu;
        m[K2_PER_LANE] = p.ms[P + K2_PER_LANE];
    }
    // separator bits of positions P-2 .. P+16 -> sf bit (t+2) = sep(P+t)
    uint32_t sf;
    {
        const uint32_t sw = __ldg(p.q.sep + (P >> 5));
        sf = ((sw >> (P & 31)) & 0xffffu) << 2;
        sf |= sep_bit(p.q, (int64_t)P - 2) | (sep_bit(p.q, (int64_t)P - 1) << 1);
        sf |= sep_bit(p.q, (int64_t)P + 16) << 18;
    }
    const uint32_t c_e = lookahead_c(p, s + K2_TILE, lane);
    const uint32_t m_e = p.ms[s + K2_TILE];
    const bool e_sep = sep_bit(p.q, (int64_t)(s + K2_TILE));
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
// K4: format::run_lengths_gapped (format.rs:143-193) on the PLAIN translation
// K2 produces (alphabet M - X R only), one warp per query.  The warp walks the
// query 32 characters per round; the four ballots of a round are consumed by
// a warp-uniform state machine over the runs of gap / non-gap characters:
//   a gap run inside a segment is "pending" until the next aligned character;
//   it closes the segment as soon as it grows past max_gap_len, and a trailing
//   gap run is dropped (format.rs:180-184).  jumps counts 'R' preceded by 'R'.
// WRITE == false counts the segments of each query; after an exclusive scan of
// the counts (rle_scan_kernel) WRITE == true stores the records in query order.
// ---------------------------------------------------------------------------
struct RleRecord {
    uint64_t start, end, matches, mismatches, jumps, gap_bases, gap_opens;  // == kbo_rle
};

enum { RLE_STAGE = 8 };  // records per query kept in the staging buffer by the counting pass

struct RleState {
    uint64_t start, end;
    uint32_t nm, nx, nj, gb, go, pend, prev_r, n_seg;
    bool in_seg;
};

// Stores one finished record: staging slot (counting pass) or final slot (write pass).
template <bool WRITE>
__device__ __forceinline__ void rle_emit(const RleState& st, int lane, RleRecord* __restrict__ dst, uint64_t slot0,
                                         uint64_t cap) {
    const uint64_t slot = slot0 + st.n_seg;
    const bool ok = WRITE ? (slot < cap) : (st.n_seg < RLE_STAGE);
    if (lane == 0 && ok) {
        RleRecord rec = {st.start, st.end, st.nm, st.nx, st.nj, st.gb, st.go};
        dst[slot] = rec;
    }
}

// One round of 32 characters (lane i holds character base+i; `nv` of them are valid).
template <bool WRITE>
__device__ __forceinline__ void rle_round(RleState& st, uint8_t ch, uint64_t base, uint32_t nv, uint32_t max_gap_len,
                                          int lane, RleRecord* __restrict__ dst, uint64_t slot0, uint64_t cap) {
    const uint32_t valid = nv == 32 ? 0xffffffffu : ((1u << nv) - 1u);
    const uint32_t G = __ballot_sync(0xffffffffu, ch == '-') & valid;
    const uint32_t N = valid & ~G;
    const uint32_t Mm = __ballot_sync(0xffffffffu, ch == 'M' || ch == 'R' || ch == 'I') & N;
    const uint32_t Rm = __ballot_sync(0xffffffffu, ch == 'R') & N;
    const uint32_t J = Rm & ((Rm << 1) | st.prev_r);
    st.prev_r = Rm >> 31;
    if (G == 0 && nv == 32 && st.in_seg && st.pend == 0) {  // fast path: 32 aligned characters inside a segment
        st.nm += __popc(Mm);
        st.nx += __popc(~Mm);
        st.nj += __popc(J);
        st.end = base + 32;
        return;
    }
    uint32_t pos = 0;
    while (pos < nv) {
        if ((G >> pos) & 1u) {
            const uint32_t rest = ~G >> pos;  // first zero of G at or after pos
            uint32_t g = rest ? (uint32_t)(__ffs((int)rest) - 1) : 32u - pos;
            if (g > nv - pos) g = nv - pos;
            if (st.in_seg) {
                st.pend += g;
                if (st.pend > max_gap_len) {  // the gap outgrew max_gap_len: close without it
                    rle_emit<WRITE>(st, lane, dst, slot0, cap);
                    ++st.n_seg;
                    st.in_seg = false;
                    st.pend = 0;
                }
            }
            pos += g;
        } else {
            const uint32_t rest = ~N >> pos;
            uint32_t n = rest ? (uint32_t)(__ffs((int)rest) - 1) : 32u - pos;
            if (n > nv - pos) n = nv - pos;
            const uint32_t run = (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << pos;
            if (!st.in_seg) {
                st.in_seg = true;
                st.start = base + pos;
                st.nm = st.nx = st.nj = st.gb = st.go = 0;
            } else if (st.pend) {
                st.gb += st.pend;
                st.go += 1;
            }
            st.pend = 0;
            st.nm += __popc(Mm & run);
            st.nx += __popc(N & ~Mm & run);
            st.nj += __popc(J & run);
            st.end = base + pos + n;
            pos += n;
        }
    }
}

// WRITE == false: count the segments of each query and keep the first RLE_STAGE records in `stage`.
// WRITE == true : copy the staged records to their final slots; a query with more than RLE_STAGE
//                 segments is recomputed and written directly.
template <bool WRITE>
__global__ void __launch_bounds__(128) rle_kernel(const uint8_t* __restrict__ aln, const uint64_t* __restrict__ offsets,
                                                  uint64_t nq, uint32_t max_gap_len, uint32_t* __restrict__ counts,
                                                  RleRecord* __restrict__ stage,
                                                  const uint64_t* __restrict__ rle_offsets,
                                                  RleRecord* __restrict__ out, uint64_t cap) {
    const int lane = threadIdx.x & 31;
    const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;  // warp-uniform
    uint64_t slot0 = 0;
    RleRecord* dst = stage + q * RLE_STAGE;
    if (WRITE) {
        slot0 = rle_offsets[q];
        const uint32_t cnt = counts[q];
        if (cnt <= RLE_STAGE) {  // common case: move the staged records (7 words each)
            const uint64_t* src = reinterpret_cast<const uint64_t*>(stage + q * RLE_STAGE);
            uint64_t* d64 = reinterpret_cast<uint64_t*>(out + slot0);
            for (uint32_t w = lane; w < cnt * 7; w += 32)
                if (slot0 + w / 7 < cap) d64[w] = src[w];
            return;
        }
        dst = out;
    }
    const uint64_t a = offsets[q] - offsets[0];
    const uint64_t len = offsets[q + 1] - offsets[q];
    RleState st;
    st.start = st.end = 0;
    st.nm = st.nx = st.nj = st.gb = st.go = st.pend = st.prev_r = st.n_seg = 0;
    st.in_seg = false;
    for (uint64_t base = 0; base < len; base += 128) {
        // four rounds of loads in flight before the first ballot
        uint8_t ch[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t i = base + 32 * j + lane;
            ch[j] = i < len ? aln[a + i] : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t b = base + 32 * j;
            if (b < len) {
                const uint32_t nv = len - b < 32 ? (uint32_t)(len - b) : 32u;
                rle_round<WRITE>(st, ch[j], b, nv, max_gap_len, lane, dst, slot0, cap);
            }
        }
    }
    if (st.in_seg) {  // a trailing gap run is dropped
        rle_emit<WRITE>(st, lane, dst, slot0, cap);
        ++st.n_seg;
    }
    if (!WRITE && lane == 0) counts[q] = st.n_seg;
}

// exclusive scan of the per-query segment counts -> rle_offsets[0..nq]; one block
__global__ void __launch_bounds__(1024) rle_scan_kernel(const uint32_t* __restrict__ counts, uint64_t nq,
                                                        uint64_t* __restrict__ rle_offsets) {
    __shared__ uint64_t part[1024];
    const uint32_t t = threadIdx.x;
    const uint64_t per = (nq + 1023) / 1024;
    const uint64_t lo = (uint64_t)t * per, hi = lo + per < nq ? lo + per : nq;
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += counts[i];
    part[t] = s;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint64_t o = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += o;
        __syncthreads();
    }
    uint64_t run = t ? part[t - 1] : 0;
    for (uint64_t i = lo; i < hi; ++i) {
        rle_offsets[i] = run;
        run += counts[i];
    }
    if (t == 1023) rle_offsets[nq] = part[1023];
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
