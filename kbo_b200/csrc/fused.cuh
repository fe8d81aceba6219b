// ===========================================================================
// kbo_b200/csrc/fused.cuh -- K1 + K2b in ONE kernel: matching statistics,
// derandomize and translate of one tile of the batch per thread block, the MS
// vector living only in shared memory (north_star (c); reference chain
// src/lib.rs:624-627: query_sbwt -> derandomize_ms_vec -> translate_ms_vec).
//
// A block owns `tile_len` consecutive padded positions (a multiple of 32) and
//   0. stages the packed query words of its tile (+ the k-1 positions of
//      context before it and a short look-ahead after it) into shared memory
//      with two bulk (TMA) copies                              [cp.async.bulk];
//   A. "fast pass", one lane per chunk of the tile: the MS recurrence with TWO
//      bases per rank probe (IndexView::rank2) and NO contraction code.  When a
//      base does not extend, the lane records a task (position + the exact
//      state before it), restarts from the empty state after that base and goes
//      on.  What it writes after a restart is provisional: the longest match
//      that STARTS at or after the restart;
//   B. "repair pass", one lane per task: the exact recurrence (extend, contract
//      on the LCS links, retry) from the recorded state, overwriting the
//      provisional bytes, until its length equals the provisional one at a
//      position at or after the task's last restart -- from there on both
//      passes describe the same match, so everything the fast pass wrote after
//      it is already exact.  Noise stretches (a dozen bases after every
//      mismatch, three dependent loads per base) thus run densely packed, one
//      per lane, instead of stalling the 31 other chunks of a warp;
//   C. derandomize + translate (k2b_group of kernels.cuh) on the MS bytes in
//      shared memory; characters (kbo::matches) or the three masks K4 consumes
//      (kbo::find) go to global memory.  The MS vector never does.
//
// Exactness of A + B.  Let r be the position of a lane's latest restart.  Its
// state after position j >= r is "longest suffix of Q[r..j] in the index"; the
// exact state is the same with r replaced by the start of the lane's feed.
// (i)  If both lengths agree at some j >= r the two suffixes are the same
//      string, hence the same interval, and the recurrences coincide from j on.
// (ii) The state depends on the last k bases only, so they
//      agree at the latest at j = r + k - 1: a failure at f >= r + k happens in
//      an exact state and opens a NEW task; an earlier one cannot know and
//      extends the open task's window instead (the task must run to f at
//      least).  A task therefore always starts from an exact state, stops
//      before the next task starts, and tasks of one chunk never overlap.
// (iii) A non-ACGT base resets both recurrences: a sync point by (i).
// The first failure of a lane happens in the state the one-pass kernel would
// have (pref-table start + exact steps), warm-up included.
// ===========================================================================
#pragma once
#include "kernels.cuh"

namespace kbo_b200 {

enum { FUSED_THREADS = 128, FUSED_WARPS = 4, MS_LOOKAHEAD = 72, FUSED_STAGE_CHARS = K2B_TILE + 16 };
enum { FUSED_FLAG_NO_PAIRS = 4 };  // kbo_set_ms_flags bit 2: one base per probe in the fast pass (comparison runs)
// kbo_set_ms_flags bit 3 (together with bit 4): the ONE-PASS form -- pass A is K1's exact recurrence (one base per probe,
// contraction on the link array in the same iteration), there are no tasks and no pass B; staging and pass C as above.
enum { FUSED_FLAG_EXACT = 8 };

struct FusedParams {
    IndexView ix;
    QueryView q;
    TrParams tr;            // k, thr, outputs (tr.ms unused)
    uint32_t tile_len;      // padded positions per tile, multiple of 32
    uint32_t chunk;         // positions per lane: ceil((tile_len + MS_LOOKAHEAD) / FUSED_THREADS)
    uint32_t stage_words;   // capacity of the staged query window (words of 32 positions)
    uint32_t task_cap;      // capacity of the task list
    uint32_t flags;
    uint64_t mask_words;    // masks mode: words the three mask arrays hold (the last block zero-fills past the batch)
    unsigned long long* counters;  // optional (COUNT)
};

// shared-memory carve-up (all offsets multiples of 16)
struct FusedSmem {
    uint32_t off_pack, off_inv, off_ms, off_warm, off_tasks, off_misc, off_ring, off_lut, off_stage, total;
};
__host__ __device__ inline uint32_t fused_warm_row(uint32_t k) { return (k + 14u) & ~15u; }  // >= k - 1, multiple of 16
__host__ __device__ inline FusedSmem fused_smem_layout(uint32_t tile_len, uint32_t stage_words, uint32_t task_cap, uint32_t k,
                                                       bool chars) {
    FusedSmem s;
    uint32_t o = 0;
    s.off_pack = o;  o += ((stage_words * 8u) + 15u) & ~15u;
    s.off_inv = o;   o += ((stage_words * 4u) + 15u) & ~15u;
    s.off_ms = o;    o += (16u + tile_len + MS_LOOKAHEAD + 16u + 15u) & ~15u;
    s.off_warm = o;  o += task_cap ? FUSED_THREADS * fused_warm_row(k) : 0u;  // provisional MS of every lane's warm-up positions
    s.off_tasks = o; o += task_cap * 16u;
    s.off_misc = o;  o += 64u;   // mbarrier (8), task count, tail state
    s.off_ring = o;  o += FUSED_WARPS * 64u;
    s.off_lut = o;   o += chars ? 1024u : 0u;
    s.off_stage = o; o += chars ? FUSED_WARPS * (uint32_t)FUSED_STAGE_CHARS : 0u;
    s.total = o;
    return s;
}

// Tile geometry of one launch (host side).  Tiles are sized so that (a) a lane gets about `target` positions (64
// unless the caller says otherwise: the k-1 warm-up positions of every lane are overhead, but a 10^7-base batch has
// to be cut this fine to give every SM its lanes), and (b) the number of tiles is a multiple of the SM count, so that
// every SM gets the same number of blocks (round 1: 4.12 blocks per SM on average left the busiest SM with 5).
struct FusedGeom {
    uint32_t tile_len = 0, chunk = 0, stage_words = 0, task_cap = 0;
    uint64_t n_tiles = 0;
    FusedSmem smem;
};
inline bool fused_geometry(uint64_t Lp, uint32_t k, bool chars, int n_sms, uint32_t target, FusedGeom* out,
                           bool exact = false) {
    FusedGeom g;
    if (!target) {
        const uint64_t t = Lp / ((uint64_t)n_sms * 2048ull * 4ull);
        target = (uint32_t)(t < 64 ? 64 : (t > 256 ? 256 : t));
    }
    uint64_t tiles = (Lp + (uint64_t)FUSED_THREADS * target - 1) / ((uint64_t)FUSED_THREADS * target);
    if (tiles > (uint64_t)n_sms) tiles = ((tiles + n_sms / 2) / n_sms) * n_sms;  // nearest multiple of the SM count
    if (tiles == 0) tiles = 1;
    uint64_t tile_len = ((Lp + tiles - 1) / tiles + 31) & ~31ull;
    if (tile_len < 32) tile_len = 32;
    if (tile_len > 60000) return false;
    g.tile_len = (uint32_t)tile_len;
    g.n_tiles = (Lp + tile_len - 1) / tile_len;
    g.chunk = (g.tile_len + MS_LOOKAHEAD + FUSED_THREADS - 1) / FUSED_THREADS;
    g.stage_words = (g.tile_len + MS_LOOKAHEAD + 31) / 32 + ((k + 30) >> 5) + 8;
    g.task_cap = exact ? 0u : FUSED_THREADS * ((g.chunk + 2 * (k - 1)) / k + 2);  // a new task at most every k positions of a lane's feed
    g.smem = fused_smem_layout(g.tile_len, g.stage_words, g.task_cap, k, chars);
    if (g.smem.total > 200 * 1024) return false;
    *out = g;
    return true;
}

#ifndef KBO_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#define KBO_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#define KBO_GRID_CONSTANT __grid_constant__  // the kernel takes addresses of parameter fields (no local copy)
#else
#define KBO_DYN_SMEM(name) static __attribute__((aligned(16))) uint8_t name[232 * 1024]
#define KBO_GRID_CONSTANT
#endif

template <bool CHARS, bool COUNT, bool EXACT = false>
__global__ void __launch_bounds__(FUSED_THREADS, 8) ms_fused_kernel(const KBO_GRID_CONSTANT FusedParams p) {
    KBO_DYN_SMEM(smem);
    const FusedSmem lay = fused_smem_layout(p.tile_len, p.stage_words, p.task_cap, p.ix.k, CHARS);
    uint64_t* const pack_s = reinterpret_cast<uint64_t*>(smem + lay.off_pack);
    uint32_t* const inv_s = reinterpret_cast<uint32_t*>(smem + lay.off_inv);
    uint8_t* const ms_s = smem + lay.off_ms;
    uint8_t* const warm_s = smem + lay.off_warm;
    uint4* const tasks = reinterpret_cast<uint4*>(smem + lay.off_tasks);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + lay.off_misc);
    uint32_t* const misc = reinterpret_cast<uint32_t*>(smem + lay.off_misc) + 2;  // [0] task count, [1..3] tail l, r, d
    uint8_t* const ring = smem + lay.off_ring;
    uint32_t* const lut = reinterpret_cast<uint32_t*>(smem + lay.off_lut);
    uint8_t* const stage = smem + lay.off_stage;

    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t n = p.ix.n, k = p.ix.k;
    const uint64_t Lp = p.q.Lp;
    const uint64_t tile_start = (uint64_t)blockIdx.x * p.tile_len;
    unsigned long long cnt_att = 0, cnt_split = 0, cnt_con = 0, cnt_extra = 0, cnt_proc = 0, cnt_emit = 0;
    unsigned long long cnt_att_e = 0, cnt_split_e = 0, cnt_con_e = 0, cnt_extra_e = 0;
    if (tile_start >= Lp) return;  // (whole block)
    const uint64_t lp32 = (Lp + 31) & ~31ull;
    const uint64_t tile_end = tile_start + p.tile_len < lp32 ? tile_start + p.tile_len : lp32;  // emitted: [tile_start, tile_end)
    const uint64_t V = tile_start + p.tile_len + MS_LOOKAHEAD < Lp ? tile_start + p.tile_len + MS_LOOKAHEAD : Lp;  // MS: [tile_start, V)

    // ---- 0. stage the query words [ws, we) -------------------------------------------------------------------
    const uint32_t WB = (k + 30u) >> 5;  // words that hold the k-1 positions before the tile
    const uint64_t tw0 = tile_start >> 5;
    const uint64_t ws = (tw0 > WB ? tw0 - WB : 0ull) & ~3ull;
    uint64_t we = (((V + 31) >> 5) + 3) & ~3ull;
    if (we > p.q.n_words) we = p.q.n_words;  // (a multiple of 4 as well)
    const uint32_t n_stage = (uint32_t)(we - ws);
#ifndef KBO_HOST_EMU
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, n_stage * 12u);
        bulk_g2s(pack_s, p.q.pack + ws, n_stage * 8u, bar);
        bulk_g2s(inv_s, p.q.inv + ws, n_stage * 4u, bar);
        misc[0] = 0;
    }
    if (CHARS) k2b_fill_lut(lut);
    __syncthreads();
    mbar_wait(bar, 0);
#else
    if (tid == 0) {
        for (uint32_t i = 0; i < n_stage; ++i) { pack_s[i] = p.q.pack[ws + i]; inv_s[i] = p.q.inv[ws + i]; }
        misc[0] = 0;
    }
    (void)bar;
    if (CHARS) k2b_fill_lut(lut);
    __syncthreads();
#endif

    // positions below are RELATIVE to the first staged word: rel(pos) = pos - 32 ws
    const uint32_t rel_tile = (uint32_t)(tile_start - ws * 32);
    const uint32_t rel_end = (uint32_t)(tile_end - ws * 32);
    const uint32_t rel_V = (uint32_t)(V - ws * 32);
    uint8_t* const msb = ms_s + 16 - rel_tile;  // msb[rel] = MS byte of the position, rel >= rel_tile - 1
    const bool pairs = p.ix.rank2 != nullptr && !(p.flags & FUSED_FLAG_NO_PAIRS);
    const uint32_t stride = p.ix.rank_stride;

    // ---- A (one-pass form). the exact recurrence, as K1 -------------------------------------------------------------
    if (EXACT) {
        const uint32_t a = rel_tile + tid * p.chunk;
        if (a < rel_V) {
            const uint32_t b = a + p.chunk < rel_V ? a + p.chunk : rel_V;
            const uint64_t a64 = tile_start + (uint64_t)tid * p.chunk;
            uint32_t warm = a64 >= (uint64_t)(k - 1) ? k - 1 : (uint32_t)a64;
            if (warm) {  // cut after the last non-ACGT position of the window
                const uint32_t lo = a - warm;
                for (int32_t w = (int32_t)((a - 1) >> 5); w >= (int32_t)(lo >> 5); --w) {
                    uint32_t iv = inv_s[w];
                    if ((uint32_t)w == ((a - 1) >> 5) && (a & 31u)) iv &= (1u << (a & 31u)) - 1u;
                    if ((uint32_t)w == (lo >> 5)) iv &= ~0u << (lo & 31u);
                    if (iv) {
                        warm = a - ((uint32_t)w * 32u + (31u - (uint32_t)__clz((int)iv)) + 1u);
                        break;
                    }
                }
            }
            uint32_t l = 0, r = n, d = 0;
            const uint32_t P = p.ix.pref ? p.ix.pref_len : 0u;
            if (P && warm >= P) {
                const uint32_t first = a - warm;
                const uint32_t sh = 2u * (first & 31u);
                uint64_t bits = pack_s[first >> 5] >> sh;
                if (sh > 64 - 2 * P) bits |= pack_s[(first >> 5) + 1] << (64 - sh);
                if (pref_decode(__ldg(p.ix.pref + ((uint32_t)bits & ((1u << (2 * P)) - 1u))), n, l, r, d)) warm -= P;
            }
            uint32_t bp = a - warm;
            if (tid == 0) ms_s[15] = (uint8_t)d;  // MS of the position before the tile (overwritten when it is stepped through)
            uint64_t qw = pack_s[bp >> 5] >> (2u * (bp & 31u));
            uint32_t iw = inv_s[bp >> 5] >> (bp & 31u);
            while (bp < b) {
                bool advance = true;
                if (iw & 1u) {
                    l = 0; r = n; d = 0;
                } else {
                    const uint32_t rowoff = ((uint32_t)qw & 3u) * stride;
                    const uint32_t bl = l >> 5, br = r >> 5;
                    const uint64_t wl = __ldg(p.ix.rank + (rowoff + bl));
                    const uint64_t wr = (br == bl) ? wl : __ldg(p.ix.rank + (rowoff + br));
                    const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
                    const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
                    if (COUNT) {
                        const bool sp = (bl >> 2) != (br >> 2);
                        ++cnt_att; cnt_split += sp;
                        if (bp >= a && bp < rel_end) { ++cnt_att_e; cnt_split_e += sp; }
                    }
                    if (nl < nr) {
                        l = nl; r = nr;
                        d = d + 1 < k ? d + 1 : k;
                    } else if (d != 0) {
                        advance = false;
                        const uint32_t el = __ldg(p.ix.links + l), er = __ldg(p.ix.links + r);
                        const bool scanned = ms_contract(p.ix, el, er, l, r, d);
                        if (COUNT) {
                            ++cnt_con; cnt_extra += scanned;
                            if (bp >= a && bp < rel_end) { ++cnt_con_e; cnt_extra_e += scanned; }
                        }
                    }
                }
                if (advance) {
                    if (COUNT) { ++cnt_proc; cnt_emit += (bp >= a && bp < rel_end); }
                    if (bp >= a) msb[bp] = (uint8_t)d;
                    else if (tid == 0 && bp + 1 == a) ms_s[15] = (uint8_t)d;
                    ++bp;
                    qw >>= 2;
                    iw >>= 1;
                    if ((bp & 31u) == 0 && bp < b) {
                        qw = pack_s[bp >> 5];
                        iw = inv_s[bp >> 5];
                    }
                }
            }
            if (b == rel_V) { misc[1] = l; misc[2] = r; misc[3] = d; }  // exact state after the last MS position (pass C's tail)
        }
    }
    // ---- A. fast pass ---------------------------------------------------------------------------------------------
    if (!EXACT) {
        const uint32_t a = rel_tile + tid * p.chunk;
        if (a < rel_V) {
            const uint32_t b = a + p.chunk < rel_V ? a + p.chunk : rel_V;
            const uint64_t a64 = tile_start + (uint64_t)tid * p.chunk;
            // warm-up: the k-1 positions before the chunk, cut after the last non-ACGT position among them
            uint32_t warm = a64 >= (uint64_t)(k - 1) ? k - 1 : (uint32_t)a64;
            if (warm) {
                const uint32_t lo = a - warm;
                for (int32_t w = (int32_t)((a - 1) >> 5); w >= (int32_t)(lo >> 5); --w) {
                    uint32_t iv = inv_s[w];
                    if ((uint32_t)w == ((a - 1) >> 5) && (a & 31u)) iv &= (1u << (a & 31u)) - 1u;  // positions < a
                    if ((uint32_t)w == (lo >> 5)) iv &= ~0u << (lo & 31u);                         // positions >= lo
                    if (iv) {
                        warm = a - ((uint32_t)w * 32u + (31u - (uint32_t)__clz((int)iv)) + 1u);
                        break;
                    }
                }
            }
            uint32_t l = 0, r = n, d = 0;
            const uint32_t P = p.ix.pref ? p.ix.pref_len : 0u;
            if (P && warm >= P) {
                const uint32_t first = a - warm;
                const uint32_t sh = 2u * (first & 31u);
                uint64_t bits = pack_s[first >> 5] >> sh;
                if (sh > 64 - 2 * P) bits |= pack_s[(first >> 5) + 1] << (64 - sh);
                if (pref_decode(__ldg(p.ix.pref + ((uint32_t)bits & ((1u << (2 * P)) - 1u))), n, l, r, d)) warm -= P;
            }
            uint32_t bp = a - warm;
            // Bytes of positions >= a go to the tile's MS array; the lane's provisional bytes of its warm-up positions
            // go to its private row (index: distance below a), where the repair pass compares against them.
            // Lane 0 also produces the MS byte of the position before the tile (translate's left neighbour).
            uint8_t* const wrow = warm_s + tid * fused_warm_row(k);
            if (tid == 0) ms_s[15] = (uint8_t)d;  // overwritten below when that position is stepped through
            uint64_t qw = pack_s[bp >> 5] >> (2u * (bp & 31u));
            uint32_t iw = inv_s[bp >> 5] >> (bp & 31u);
            bool fast = true;         // probe two bases at once (cleared by an empty pair until a base extends again)
            bool known_fail = false;  // the current base is known not to extend (the pair was empty, its first base fine)
            uint32_t new_from = 0;    // a failure at or after this position opens a new task (else it extends the open one)
            uint32_t cur_task = 0, cur_start = 0;
            while (bp < b) {
                uint32_t adv = 1, dA = 0, dB = 0;
                if (iw & 1u) {  // not ACGT: both recurrences restart here
                    l = 0; r = n; d = 0;
                    fast = true; known_fail = false;
                    new_from = bp + 1;
                } else {
                    const bool two = pairs && fast && !known_fail && !(iw & 2u) && (bp & 31u) != 31u && bp + 1 < b;
                    bool ok = false;
                    if (!known_fail) {
                        const uint64_t* __restrict__ rows = two ? p.ix.rank2 : p.ix.rank;
                        const uint32_t rowoff = ((uint32_t)qw & (two ? 15u : 3u)) * stride;
                        const uint32_t bl = l >> 5, br = r >> 5;
                        const uint64_t wl = __ldg(rows + (rowoff + bl));
                        const uint64_t wr = (br == bl) ? wl : __ldg(rows + (rowoff + br));
                        const uint32_t nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
                        const uint32_t nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
                        if (COUNT) {
                            const bool sp = (bl >> 2) != (br >> 2);
                            ++cnt_att; cnt_split += sp;
                            if (bp + (two ? 1u : 0u) >= a && bp < rel_end) { ++cnt_att_e; cnt_split_e += sp; }
                        }
                        ok = nl < nr;
                        if (ok) { l = nl; r = nr; }
                    }
                    if (ok) {
                        dA = d + 1 < k ? d + 1 : k;
                        dB = d + 2 < k ? d + 2 : k;
                        adv = two ? 2u : 1u;
                        d = two ? dB : dA;
                        known_fail = !two && !fast;  // single step after an empty pair: the pair's second base fails
                        fast = true;
                    } else if (two) {
                        adv = 0;
                        fast = false;  // retry the first base alone
                    } else if (d != 0) {
                        // Failure in a state with d > 0: the exact continuation goes to the repair pass and the lane
                        // restarts from the empty state AFTER this base (provisional length 0 here).  After a
                        // substitution -- the common case -- the match that starts at the next base is the one the
                        // exact recurrence ends up with a dozen positions later, where the two then agree; a restart
                        // AT the failing base would die at about that depth and restart again behind the exact match
                        // (measured: 42 instead of ~13 positions per repair).
                        if (bp >= new_from) {
                            cur_task = atomicAdd(misc, 1u);
                            cur_start = bp;
                            if (cur_task < p.task_cap) tasks[cur_task] = make_uint4(bp, l, r, d | (tid << 8));
                        } else if (cur_task < p.task_cap) {
                            reinterpret_cast<uint16_t*>(&tasks[cur_task].w)[1] = (uint16_t)(bp - cur_start);
                        }
                        new_from = bp + 1 + k;
                        l = 0; r = n; d = 0;
                        dA = 0;
                        fast = true; known_fail = false;
                    } else {
                        dA = 0;  // nothing matches this base, not even alone
                        fast = true; known_fail = false;
                    }
                }
                if (adv) {
                    if (COUNT) { cnt_proc += adv; cnt_emit += (bp >= a && bp < rel_end) + (adv == 2 && bp + 1 >= a && bp + 1 < rel_end); }
                    {
                        const uint8_t v0 = (uint8_t)(adv == 2 ? dA : d);
                        if (bp >= a) msb[bp] = v0; else wrow[a - 1 - bp] = v0;
                        if (tid == 0 && bp + 1 == a) ms_s[15] = v0;
                    }
                    if (adv == 2) {
                        if (bp + 1 >= a) msb[bp + 1] = (uint8_t)dB; else wrow[a - 2 - bp] = (uint8_t)dB;
                        if (tid == 0 && bp + 2 == a) ms_s[15] = (uint8_t)dB;
                    }
                    bp += adv;
                    qw >>= 2 * adv;
                    iw >>= adv;
                    if ((bp & 31u) == 0 && bp < b) {
                        qw = pack_s[bp >> 5];
                        iw = inv_s[bp >> 5];
                    }
                }
            }
            if (b == rel_V) { misc[1] = l; misc[2] = r; misc[3] = d; }  // state after the last MS position (see pass B)
        }
    }
    __syncthreads();

    // ---- B. repair pass: one lane per task ------------------------------------------------------------------------
    // Most positions of a task are noise: the base fails to extend, the interval is contracted, the base extends.
    // The link words of the current interval are therefore loaded TOGETHER with the rank words of every probe, so a
    // failed probe costs no second round trip before the contraction (the extra sector is wasted when the probe
    // succeeds, which is the rarer case here).  The first probe of a task is known to fail (the fast pass saw it).
    if (!EXACT) {
        const uint32_t n_tasks = misc[0] < p.task_cap ? misc[0] : p.task_cap;  // (the capacity is a proven bound)
        for (uint32_t t = tid; t < n_tasks; t += FUSED_THREADS) {
            const uint4 T = tasks[t];
            uint32_t bp = T.x, l = T.y, r = T.z, d = T.w & 0xffu;
            const uint32_t owner = (T.w >> 8) & 0xffu, last_fail = T.x + (T.w >> 16);  // (restart = the position after it)
            const uint32_t a = rel_tile + owner * p.chunk;
            const uint32_t b = a + p.chunk < rel_V ? a + p.chunk : rel_V;
            uint8_t* const wrow = warm_s + owner * fused_warm_row(k);
            uint64_t qw = pack_s[bp >> 5] >> (2u * (bp & 31u));
            uint32_t iw = inv_s[bp >> 5] >> (bp & 31u);
            bool synced = false, known_fail = true;
            while (bp < b) {
                if (iw & 1u) {
                    l = 0; r = n; d = 0;
                } else {
                    const uint32_t rowoff = ((uint32_t)qw & 3u) * stride;
                    for (;;) {  // extend; on failure contract to the next depth that changes the interval and retry
                        const uint32_t bl = l >> 5, br = r >> 5;
                        uint32_t nl = 0, nr = 0;
                        const uint32_t el = __ldg(p.ix.links + l), er = __ldg(p.ix.links + r);
                        if (!known_fail) {
                            const uint64_t wl = __ldg(p.ix.rank + (rowoff + bl));
                            const uint64_t wr = (br == bl) ? wl : __ldg(p.ix.rank + (rowoff + br));
                            nl = (uint32_t)(wl >> 32) + __popc((uint32_t)wl & ((1u << (l & 31)) - 1u));
                            nr = (uint32_t)(wr >> 32) + __popc((uint32_t)wr & ((1u << (r & 31)) - 1u));
                            if (COUNT) {
                                const bool sp = (bl >> 2) != (br >> 2);
                                ++cnt_att; cnt_split += sp;
                                if (bp >= a && bp < rel_end) { ++cnt_att_e; cnt_split_e += sp; }
                            }
                        }
                        known_fail = false;
                        if (nl < nr) {
                            l = nl; r = nr;
                            d = d + 1 < k ? d + 1 : k;
                            break;
                        }
                        if (d == 0) break;
                        const bool scanned = ms_contract(p.ix, el, er, l, r, d);
                        if (COUNT) {
                            ++cnt_con; cnt_extra += scanned;
                            if (bp >= a && bp < rel_end) { ++cnt_con_e; cnt_extra_e += scanned; }
                        }
                    }
                }
                if (COUNT) ++cnt_proc;
                {
                    uint8_t* const slot = bp >= a ? msb + bp : wrow + (a - 1 - bp);
                    if (bp > last_fail && *slot == (uint8_t)d) { synced = true; break; }  // same match from here on
                    *slot = (uint8_t)d;
                    if (owner == 0 && bp + 1 == a) ms_s[15] = (uint8_t)d;
                }
                ++bp;
                qw >>= 2;
                iw >>= 1;
                if ((bp & 31u) == 0 && bp < b) {
                    qw = pack_s[bp >> 5];
                    iw = inv_s[bp >> 5];
                }
            }
            if (!synced && b == rel_V) { misc[1] = l; misc[2] = r; misc[3] = d; }  // exact state after the last position
        }
    }
    __syncthreads();

    // ---- C. derandomize + translate on the tile's MS bytes ---------------------------------------------------------
    {
        MsTail tail;
        tail.end = V;
        tail.l = misc[1]; tail.r = misc[2]; tail.d = misc[3];
        tail.ring = ring + warp * 64u;
        tail.ix = &p.ix;
        const uint8_t* ms_abs = msb - ws * 32;  // indexed by padded position
        const uint32_t tile_words = (uint32_t)((tile_end - tile_start) >> 5);
        const uint32_t n_groups = (tile_words + 31u) >> 5;
        for (uint32_t g = warp; g < n_groups; g += FUSED_WARPS) {
            const uint32_t left = tile_words - 32u * g;
            k2b_group<CHARS>(p.tr, ms_abs, tile_start + 1024ull * g, (int)(left < 32u ? left : 32u) - 1, &tail, lut,
                             stage + warp * (uint32_t)FUSED_STAGE_CHARS);
            if (CHARS) __syncwarp();
        }
        if (!CHARS && blockIdx.x == gridDim.x - 1) {  // K4 reads whole 1024-position tiles of masks
            for (uint64_t w = (tile_end >> 5) + tid; w < p.mask_words; w += FUSED_THREADS) {
                p.tr.out_gap[w] = 0u; p.tr.out_match[w] = 0u; p.tr.out_r[w] = 0u;
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

}  // namespace kbo_b200
