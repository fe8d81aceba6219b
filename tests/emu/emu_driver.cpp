// ===========================================================================
// tests/emu/emu_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs the kernels of kbo_b200/csrc/kernels.cuh on the CPU through
// tests/emu/host_emu.hpp, with the same launch sequence and buffer geometry as
// kbo_b200/csrc/capi.cu, so that tests can compare the kernel LOGIC with the
// oracle in a GPU-less container.  Also exposes the product's host-side index
// builder (plain C++) for CPU tests.  Built into tests/emu/libkbo_emu.so by
// tests/emu_lib.py; never loaded by the kbo_b200 package.
// ===========================================================================
#define KBO_HOST_EMU 1
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>

#include "../../kbo_b200/csrc/host_layout.hpp"
#include "../../kbo_b200/csrc/kernels.cuh"
#include "../../kbo_b200/csrc/fused.cuh"
#include "../../kbo_b200/csrc/refine.cuh"
#include "../../kbo_b200/csrc/refine_host.hpp"
#include "../../kbo_b200/csrc/sbwt_host.hpp"

using namespace kbo_b200;

static uint32_t g_emu_flags = 0;
extern "C" void emu_set_ms_flags(uint32_t v) { g_emu_flags = v; }
static int g_emu_prefix_table = 1;  // 0: indexes built afterwards get no prefix-state table
extern "C" void emu_set_prefix_table(int v) { g_emu_prefix_table = v; }
static uint32_t g_emu_prefix_len = PREF_LEN;  // depth of the table of indexes built afterwards
extern "C" void emu_set_prefix_len(uint32_t v) { g_emu_prefix_len = v ? v : PREF_LEN; }
static int g_emu_k2_mode = 0;  // 0: as the product dispatches, 1: always K2, 2: K2b (where supported)
extern "C" void emu_set_k2_mode(int v) { g_emu_k2_mode = v; }

// the product's K2 / K2b dispatch (capi.cu run_derand_translate)
static void emu_run_translate(TrParams tp, const Geometry& g) {
    const bool bits = g_emu_k2_mode != 1 && k2b_supported(tp.k, tp.thr);
    if (bits) {
        tp.n_tiles = g.n_tiles_b;
        emu_launch_par((unsigned)((g.n_tiles_b + K2B_WARPS - 1) / K2B_WARPS), K2B_WARPS * 32,
                       [&]() { derand_translate_bits_kernel<true>(tp); });
    } else {
        tp.n_tiles = g.n_tiles;
        emu_launch_par((unsigned)((g.n_tiles + K2_WARPS - 1) / K2_WARPS), K2_WARPS * 32,
                       [&]() { derand_translate_kernel(tp); });
    }
}

static int g_emu_rank2 = 1;  // 0: indexes built afterwards carry no rank2 rows
extern "C" void emu_set_rank2(int v) { g_emu_rank2 = v; }

struct EmuIndex {
    HostIndex host;
    DeviceLayout lay;
    std::vector<uint64_t> rank2;
    std::vector<uint32_t> links;
    std::vector<uint64_t> pref, pref_tmp;
    IndexView view;
};

static void finish(EmuIndex* e) {
    build_device_layout(e->host, &e->lay);
    e->view.rank = e->lay.rank.data();
    e->view.rank_stride = (uint32_t)e->lay.stride;
    e->view.lcs = e->lay.lcs.data();
    e->view.n = (uint32_t)e->host.n_sets;
    e->view.k = e->host.k;
    e->view.rank2 = nullptr;
    if (g_emu_rank2) {  // as capi.cu build_rank2 (the scan between the two kernels is plain C++ here)
        const uint64_t stride = e->lay.stride, words = 16 * stride;
        std::vector<uint32_t> rows2(words), prefix(words);
        emu_launch_seq((unsigned)((stride + 127) / 128), 128, [&]() { rank2_bits_kernel(e->view, rows2.data()); });
        uint32_t run = 0;
        for (uint64_t i = 0; i < words; ++i) { prefix[i] = run; run += (uint32_t)__builtin_popcount(rows2[i]); }
        // rank2's rows follow rank's four rows in ONE array (as in the device allocation; K1p relies on it)
        e->rank2.assign(4 * stride + words, 0);
        std::memcpy(e->rank2.data(), e->lay.rank.data(), 4 * stride * 8);
        emu_launch_seq((unsigned)((words + 255) / 256), 256,
                       [&]() { compose_rank2_kernel(e->view, rows2.data(), prefix.data(), e->rank2.data() + 4 * stride); });
        e->view.rank = e->rank2.data();
        e->view.rank2 = e->rank2.data() + 4 * stride;
    }
    const uint32_t n = e->view.n;
    e->links.assign((size_t)n + 1, 0);
    emu_launch_seq((unsigned)(((uint64_t)n + 1 + 255) / 256), 256,
                   [&]() { lcs_links_kernel(e->lay.lcs.data(), n, e->links.data()); });
    e->view.links = e->links.data();
    e->view.pref = nullptr;
    e->view.pref_len = 0;
    if (e->view.k >= PREF_MIN_K && g_emu_prefix_table) {  // as capi.cu build_pref_table
        const uint32_t P = std::min<uint32_t>(std::min<uint32_t>(g_emu_prefix_len, PREF_MAX_LEN), e->view.k - 1);
        const size_t last = (size_t)1 << (2 * P);
        e->pref_tmp.assign(std::max<size_t>(last / 4, 4), 0);
        e->pref.assign(last, 0);
        for (uint32_t j = 1; j <= P; ++j) {
            const uint32_t cnt = 1u << (2 * j);
            uint64_t* cur = ((P - j) & 1u) ? e->pref_tmp.data() : e->pref.data();
            const uint64_t* prev = ((P - j) & 1u) ? e->pref.data() : e->pref_tmp.data();
            emu_launch_seq((cnt + 255) / 256, 256, [&]() { prefix_table_level_kernel(e->view, prev, cur, j); });
        }
        e->view.pref = e->pref.data();
        e->view.pref_len = P;
    }
}

static int g_emu_fused = 0;         // 1: matches / find go through the fused K1 + K2b kernel (fused.cuh) where it applies
static uint32_t g_emu_fused_chunk = 0;  // target positions per lane (0 = automatic)
static int g_emu_sms = 4;           // "SM count" of the emulated device (tiles are rounded to a multiple of it)
static uint64_t g_emu_fused_launches = 0, g_emu_fused_tiles = 0;
extern "C" uint64_t emu_fused_launches() { return g_emu_fused_launches; }
extern "C" uint64_t emu_fused_tiles() { return g_emu_fused_tiles; }
extern "C" uint64_t emu_tail_extension_count() { return emu_tail_extensions(); }
extern "C" void emu_set_fused(int on, uint32_t chunk, int sms) { g_emu_fused = on; g_emu_fused_chunk = chunk; g_emu_sms = sms > 0 ? sms : 4; }

struct Staged {
    Geometry g;
    std::vector<uint64_t> pack;
    std::vector<uint32_t> inv, sep, wq;
    std::vector<uint8_t> ms;
    std::vector<uint32_t> l, r;
    QueryView qv;
};

static void stage_and_ms(EmuIndex* e, const uint8_t* concat, const uint64_t* offsets, uint64_t nq, uint32_t chunk_len,
                         bool intervals, unsigned long long* counters, Staged* s) {
    const uint64_t total = offsets[nq] - offsets[0];
    s->g = make_geometry(total, nq, chunk_len);
    const Geometry& g = s->g;
    s->pack.assign(g.n_words, 0xdeadbeefdeadbeefull);
    s->inv.assign(g.n_words, 0xdeadbeef);
    s->sep.assign(g.n_words, 0xdeadbeef);
    s->wq.assign(g.n_words, 0xdeadbeef);
    s->ms.assign(g.ms_bytes, 0xAB);  // garbage where K1 does not write, like uninitialised device memory
    if (intervals) {
        s->l.assign(g.n_words * 32, 0xdeadbeef);
        s->r.assign(g.n_words * 32, 0xdeadbeef);
    }
    QueryView& qv = s->qv;
    qv.pack = s->pack.data();
    qv.inv = s->inv.data();
    qv.sep = s->sep.data();
    qv.wq = s->wq.data();
    qv.Lp = g.Lp;
    qv.n_words = g.n_words;
    {
        const unsigned threads = 128, blocks = (unsigned)((g.n_words + threads - 1) / threads);
        emu_launch_par(blocks, threads, [&]() {
            pack_queries_kernel(concat, offsets, nq, qv, s->pack.data(), s->inv.data(), s->sep.data(), s->wq.data());
        });
    }
    MsParams mp;
    mp.ix = e->view;
    mp.q = qv;
    mp.chunk_len = g.chunk_len;
    mp.flags = g_emu_flags;
    mp.n_chunks = g.n_chunks;
    mp.ms = s->ms.data();
    mp.l_out = intervals ? s->l.data() : nullptr;
    mp.r_out = intervals ? s->r.data() : nullptr;
    mp.counters = counters;
    const unsigned threads = 256, blocks = (unsigned)((g.n_chunks + threads - 1) / threads);
    if (!intervals && (mp.flags & 32u) && mp.ix.rank2) {  // K1p uses warp collectives: one host thread per lane
        emu_launch_par(blocks, threads, [&]() {
            if (counters) ms_pairs_kernel<true>(mp); else ms_pairs_kernel<false>(mp);
        });
        return;
    }
    emu_launch_seq(blocks, threads, [&]() {
        if (intervals) {
            if (counters) ms_kernel<true, true>(mp); else ms_kernel<true, false>(mp);
        } else {
            if (counters) ms_kernel<false, true>(mp); else ms_kernel<false, false>(mp);
        }
    });
}

// K0 + fused kernel (as capi.cu run_pack + run_fused).  Returns false when the parameters are outside its range.
static bool emu_run_fused(EmuIndex* e, const uint8_t* concat, const uint64_t* offsets, uint64_t nq, uint32_t thr,
                          uint8_t* chars_out, uint64_t off0, uint32_t* gap, uint32_t* match, uint32_t* rr,
                          unsigned long long* counters, Staged* s) {
    if (!k2b_supported(e->host.k, thr) || g_emu_k2_mode == 1) return false;
    const uint64_t total = offsets[nq] - offsets[0];
    s->g = make_geometry(total, nq, 0);
    const Geometry& g = s->g;
    FusedGeom fg;
    const bool exact = (g_emu_flags & FUSED_FLAG_EXACT) != 0;
    if (!fused_geometry(g.Lp, e->host.k, chars_out != nullptr, g_emu_sms, g_emu_fused_chunk, &fg, exact)) return false;
    s->pack.assign(g.n_words, 0xdeadbeefdeadbeefull);
    s->inv.assign(g.n_words, 0xdeadbeef);
    s->sep.assign(g.n_words, 0xdeadbeef);
    s->wq.assign(g.n_words, 0xdeadbeef);
    QueryView& qv = s->qv;
    qv.pack = s->pack.data(); qv.inv = s->inv.data(); qv.sep = s->sep.data(); qv.wq = s->wq.data();
    qv.Lp = g.Lp; qv.n_words = g.n_words;
    emu_launch_par((unsigned)((g.n_words + 127) / 128), 128, [&]() {
        pack_queries_kernel(concat, offsets, nq, qv, s->pack.data(), s->inv.data(), s->sep.data(), s->wq.data());
    });
    FusedParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.ix = e->view; fp.q = qv;
    fp.tr.ms = nullptr; fp.tr.q = qv; fp.tr.k = e->host.k; fp.tr.thr = thr; fp.tr.out = chars_out; fp.tr.off0 = off0;
    fp.tr.out_gap = gap; fp.tr.out_match = match; fp.tr.out_r = rr;
    fp.tile_len = fg.tile_len; fp.chunk = fg.chunk; fp.stage_words = fg.stage_words; fp.task_cap = fg.task_cap;
    fp.flags = g_emu_flags; fp.mask_words = g.n_tiles_b * 32; fp.counters = counters;
    ++g_emu_fused_launches;
    g_emu_fused_tiles += fg.n_tiles;
    emu_launch_par((unsigned)fg.n_tiles, FUSED_THREADS, [&]() {
        if (exact) {
            if (chars_out) { if (counters) ms_fused_kernel<true, true, true>(fp); else ms_fused_kernel<true, false, true>(fp); }
            else { if (counters) ms_fused_kernel<false, true, true>(fp); else ms_fused_kernel<false, false, true>(fp); }
        } else {
            if (chars_out) { if (counters) ms_fused_kernel<true, true>(fp); else ms_fused_kernel<true, false>(fp); }
            else { if (counters) ms_fused_kernel<false, true>(fp); else ms_fused_kernel<false, false>(fp); }
        }
    });
    return true;
}

template <typename T>
static void unpad(const T* in, const QueryView& q, T* out) {
    for (uint64_t pp = 0; pp < q.Lp; ++pp) {
        const uint32_t sw = q.sep[pp >> 5];
        if ((sw >> (pp & 31)) & 1u) continue;
        const uint64_t nsep = q.wq[pp >> 5] + __builtin_popcount(sw & ((1u << (pp & 31)) - 1u));
        out[pp - nsep] = in[pp];
    }
}

extern "C" {

void* emu_host_build(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k, int revcomp,
                     uint32_t threads, char* err, uint64_t err_cap) {
    EmuIndex* e = new EmuIndex();
    std::string msg = build_host_index(seqs, lens, n_seqs, k, revcomp != 0, threads, &e->host);
    if (!msg.empty()) {
        std::snprintf(err, err_cap, "%s", msg.c_str());
        delete e;
        return nullptr;
    }
    finish(e);
    return e;
}

void* emu_from_parts(uint32_t k, uint64_t n_sets, uint64_t n_kmers, const uint64_t* const rows[4], const uint8_t* lcs) {
    EmuIndex* e = new EmuIndex();
    HostIndex& h = e->host;
    h.k = k;
    h.n_sets = n_sets;
    h.n_kmers = n_kmers;
    const size_t nw = (size_t)(n_sets + 63) / 64;
    for (int c = 0; c < 4; ++c) {
        h.rows[c].assign(nw + 1, 0);
        std::memcpy(h.rows[c].data(), rows[c], nw * 8);
        if (n_sets & 63) h.rows[c][nw - 1] &= ~0ull >> (64 - (n_sets & 63));
    }
    h.lcs.assign(lcs, lcs + n_sets);
    h.finalize();
    finish(e);
    return e;
}

void emu_free(void* h) { delete (EmuIndex*)h; }
uint64_t emu_n_sets(void* h) { return ((EmuIndex*)h)->host.n_sets; }
uint64_t emu_n_kmers(void* h) { return ((EmuIndex*)h)->host.n_kmers; }
uint32_t emu_k(void* h) { return ((EmuIndex*)h)->host.k; }
void emu_export(void* h, uint64_t* a, uint64_t* c, uint64_t* g, uint64_t* t, uint8_t* lcs, uint64_t* C4) {
    HostIndex& hi = ((EmuIndex*)h)->host;
    const size_t nw = (size_t)(hi.n_sets + 63) / 64;
    uint64_t* outs[4] = {a, c, g, t};
    for (int ch = 0; ch < 4; ++ch) std::memcpy(outs[ch], hi.rows[ch].data(), nw * 8);
    std::memcpy(lcs, hi.lcs.data(), (size_t)hi.n_sets);
    for (int ch = 0; ch < 4; ++ch) C4[ch] = hi.C[ch];
}
void emu_access_kmer(void* h, uint64_t colex, uint8_t* out) { ((EmuIndex*)h)->host.access_kmer(colex, out); }
int emu_search(void* h, const uint8_t* pat, uint64_t len, uint64_t* l, uint64_t* r) {
    return ((EmuIndex*)h)->host.search(pat, len, l, r) ? 1 : 0;
}

// K0 + K1 (+ unpad): index::query_sbwt for a CSR batch, compact outputs indexed like concat
void emu_query_sbwt_batch(void* h, const uint8_t* concat, const uint64_t* offsets, uint64_t nq, uint32_t chunk_len,
                          uint8_t* d_out, uint32_t* l_out, uint32_t* r_out, unsigned long long* counters10) {
    Staged s;
    if (counters10) std::memset(counters10, 0, CNT_N * sizeof(unsigned long long));
    stage_and_ms((EmuIndex*)h, concat + offsets[0], offsets, nq, chunk_len, l_out || r_out, counters10, &s);
    unpad<uint8_t>(s.ms.data(), s.qv, d_out + offsets[0]);
    if (l_out) unpad<uint32_t>(s.l.data(), s.qv, l_out + offsets[0]);
    if (r_out) unpad<uint32_t>(s.r.data(), s.qv, r_out + offsets[0]);
}

// K0 alone: the packed form of a CSR batch (n_words = make_geometry(total, nq).n_words entries per array)
uint64_t emu_pack(const uint8_t* ascii, const uint64_t* offsets, uint64_t nq, uint64_t* pack, uint32_t* inv,
                  uint32_t* sep, uint32_t* wq) {
    const Geometry g = make_geometry(offsets[nq] - offsets[0], nq, 0);
    if (!pack) return g.n_words;
    QueryView qv;
    qv.pack = pack; qv.inv = inv; qv.sep = sep; qv.wq = wq;
    qv.Lp = g.Lp; qv.n_words = g.n_words;
    emu_launch_par((unsigned)((g.n_words + 127) / 128), 128,
                   [&]() { pack_queries_kernel(ascii, offsets, nq, qv, pack, inv, sep, wq); });
    return g.n_words;
}

// K0 + K1 + K2: kbo::matches for a CSR batch
void emu_matches_batch(void* h, const uint8_t* concat, const uint64_t* offsets, uint64_t nq, uint32_t thr,
                       uint32_t chunk_len, uint8_t* chars_out) {
    EmuIndex* e = (EmuIndex*)h;
    Staged s;
    if (g_emu_fused && emu_run_fused(e, concat + offsets[0], offsets, nq, thr, chars_out, offsets[0], nullptr, nullptr,
                                     nullptr, nullptr, &s))
        return;
    stage_and_ms(e, concat + offsets[0], offsets, nq, chunk_len, false, nullptr, &s);
    TrParams tp;
    tp.ms = s.ms.data();
    tp.q = s.qv;
    tp.k = e->host.k;
    tp.thr = thr;
    tp.out = chars_out;
    tp.off0 = offsets[0];
    emu_run_translate(tp, s.g);
}

// K2 alone on a caller-supplied u8 MS vector of ONE query (no separators except the final one)
void emu_derand_translate_u8(const uint8_t* ms, uint64_t n, uint32_t k, uint32_t thr, uint8_t* chars_out) {
    const uint64_t offsets[2] = {0, n};
    Geometry g = make_geometry(n, 1, 0);
    std::vector<uint64_t> pack(g.n_words);
    std::vector<uint32_t> inv(g.n_words), sep(g.n_words), wq(g.n_words);
    std::vector<uint8_t> ascii(n, 'A'), msbuf(g.ms_bytes, 0xAB);
    std::memcpy(msbuf.data(), ms, n);
    QueryView qv;
    qv.pack = pack.data(); qv.inv = inv.data(); qv.sep = sep.data(); qv.wq = wq.data();
    qv.Lp = g.Lp; qv.n_words = g.n_words;
    emu_launch_par((unsigned)((g.n_words + 127) / 128), 128, [&]() {
        pack_queries_kernel(ascii.data(), offsets, 1, qv, pack.data(), inv.data(), sep.data(), wq.data());
    });
    TrParams tp;
    tp.ms = msbuf.data(); tp.q = qv; tp.k = k; tp.thr = thr; tp.out = chars_out; tp.off0 = 0;
    emu_run_translate(tp, g);
}

void emu_derandomize_general(const uint64_t* ms, uint64_t n, uint32_t k, uint32_t thr, int64_t* out) {
    const uint64_t n_tiles = (n + G_TILE - 1) / G_TILE;
    std::vector<int64_t> tmax(n_tiles), min_(n_tiles);
    std::vector<uint32_t> tpar(n_tiles), eps(n_tiles);
    emu_launch_par((unsigned)n_tiles, G_THREADS, [&]() { g1_tile_max_kernel(ms, n, k, thr, tmax.data()); });
    emu_launch_seq(1, 32, [&]() { g2_scan_max_kernel(tmax.data(), n_tiles, min_.data()); });
    emu_launch_par((unsigned)n_tiles, G_THREADS,
                   [&]() { g35_tile_kernel<false>(ms, n, k, thr, min_.data(), tpar.data(), nullptr, nullptr); });
    emu_launch_seq(1, 32, [&]() { g4_scan_par_kernel(tpar.data(), n_tiles, eps.data()); });
    emu_launch_par((unsigned)n_tiles, G_THREADS,
                   [&]() { g35_tile_kernel<true>(ms, n, k, thr, min_.data(), nullptr, eps.data(), out); });
}

// K4 on masks in padded space (the product's run_rle_offsets + run_rle_records)
static uint64_t emu_run_rle(const uint32_t* gap, const uint32_t* match, const uint32_t* rr, const QueryView& qv,
                            uint64_t n_words, const uint64_t* offsets, uint64_t nq, uint32_t max_gap_len,
                            uint64_t* out7, uint64_t cap, uint64_t* rle_offsets) {
    const uint64_t nb = (n_words + RLE_BLOCK - 1) / RLE_BLOCK;
    std::vector<uint32_t> jump(n_words), gopen(n_words), start(n_words), end(n_words);
    std::vector<RleCounts> cnt(n_words + nb + 1);
    std::vector<uint64_t> cse(n_words + nb + 1);
    unsigned int tickets[2] = {0, 0};
    RleParams p;
    p.gap = gap; p.match = match; p.rr = rr; p.sep = qv.sep; p.wq = qv.wq; p.n_words = n_words; p.n_blocks = nb;
    p.offsets = offsets; p.nq = nq; p.window = max_gap_len + 1;
    p.jump = jump.data(); p.gopen = gopen.data(); p.cnt = cnt.data(); p.cnt_blk = cnt.data() + n_words;
    p.start = start.data(); p.end = end.data(); p.cse = cse.data(); p.cse_blk = cse.data() + n_words;
    p.tickets = tickets; p.rle_offsets = rle_offsets; p.out = (RleRecord*)out7; p.cap = cap;
    p.base_in = nullptr; p.total_out = nullptr; p.write_first = 1;
    if (p.window == 1) {  // as capi.cu run_rle_counts
        emu_launch_par((unsigned)nb, RLE_BLOCK, [&]() { rle_word_counts_kernel<true>(p); });
    } else {
        emu_launch_par((unsigned)nb, RLE_BLOCK, [&]() { rle_word_counts_kernel<false>(p); });
        emu_launch_par((unsigned)nb, RLE_BLOCK, [&]() { rle_mark_kernel(p); });
    }
    const unsigned threads = 128;
    const uint64_t items = n_words > nq + 1 ? n_words : nq + 1;
    emu_launch_seq((unsigned)((items + threads - 1) / threads), threads, [&]() { rle_finish_kernel(p); });
    if (tickets[0] != 0 || tickets[1] != 0) std::abort();  // the last block must leave the counters at zero
    return rle_offsets[nq];
}

// K4 on plain translations (characters), CSR batch; returns the number of records (out has `cap` slots of 7 u64)
uint64_t emu_rle_batch(const uint8_t* aln, const uint64_t* offsets, uint64_t nq, uint32_t max_gap_len, uint64_t* out7,
                       uint64_t cap, uint64_t* rle_offsets) {
    const Geometry g = make_geometry(offsets[nq] - offsets[0], nq, 0);
    std::vector<uint64_t> pack(g.n_words);
    std::vector<uint32_t> inv(g.n_words), sep(g.n_words), wq(g.n_words);
    QueryView qv;
    qv.pack = pack.data(); qv.inv = inv.data(); qv.sep = sep.data(); qv.wq = wq.data();
    qv.Lp = g.Lp; qv.n_words = g.n_words;
    emu_launch_par((unsigned)((g.n_words + 127) / 128), 128,
                   [&]() { pack_queries_kernel(aln, offsets, nq, qv, pack.data(), inv.data(), sep.data(), wq.data()); });
    const uint64_t nw = g.n_tiles_b * 32;
    std::vector<uint32_t> gap(nw), match(nw), rr(nw);
    emu_launch_par((unsigned)(nw * 32 / 128), 128, [&]() {
        chars_to_masks_kernel(aln, offsets[0], sep.data(), wq.data(), nw, gap.data(), match.data(), rr.data());
    });
    return emu_run_rle(gap.data(), match.data(), rr.data(), qv, nw, offsets, nq, max_gap_len, out7, cap, rle_offsets);
}

// K0 + K1 + K2b<masks> (or K2 + chars_to_masks) + K4: kbo::find for a CSR batch
uint64_t emu_find_batch(void* h, const uint8_t* concat, const uint64_t* offsets, uint64_t nq, uint32_t thr,
                        uint32_t max_gap_len, uint64_t* out7, uint64_t cap, uint64_t* rle_offsets) {
    EmuIndex* e = (EmuIndex*)h;
    Staged s;
    if (g_emu_fused) {
        const Geometry g0 = make_geometry(offsets[nq] - offsets[0], nq, 0);
        const uint64_t nw0 = g0.n_tiles_b * 32;
        std::vector<uint32_t> gap(nw0, 0xdeadbeef), match(nw0, 0xdeadbeef), rr(nw0, 0xdeadbeef);
        if (emu_run_fused(e, concat + offsets[0], offsets, nq, thr, nullptr, 0, gap.data(), match.data(), rr.data(), nullptr, &s)) {
            std::vector<uint64_t> rel(nq + 1);
            for (uint64_t i = 0; i <= nq; ++i) rel[i] = offsets[i] - offsets[0];
            return emu_run_rle(gap.data(), match.data(), rr.data(), s.qv, nw0, rel.data(), nq, max_gap_len, out7, cap, rle_offsets);
        }
    }
    stage_and_ms(e, concat + offsets[0], offsets, nq, 0, false, nullptr, &s);
    const uint64_t nw = s.g.n_tiles_b * 32;
    std::vector<uint32_t> gap(nw), match(nw), rr(nw);
    TrParams tp;
    tp.ms = s.ms.data(); tp.q = s.qv; tp.k = e->host.k; tp.thr = thr; tp.off0 = 0;
    if (g_emu_k2_mode != 1 && k2b_supported(tp.k, tp.thr)) {
        tp.out = nullptr;
        tp.out_gap = gap.data(); tp.out_match = match.data(); tp.out_r = rr.data();
        tp.n_tiles = s.g.n_tiles_b;
        emu_launch_par((unsigned)((s.g.n_tiles_b + K2B_WARPS - 1) / K2B_WARPS), K2B_WARPS * 32,
                       [&]() { derand_translate_bits_kernel<false>(tp); });
    } else {
        std::vector<uint8_t> chars(s.g.total + 16);
        tp.out = chars.data();
        emu_run_translate(tp, s.g);
        emu_launch_par((unsigned)(nw * 32 / 128), 128, [&]() {
            chars_to_masks_kernel(chars.data(), 0, s.qv.sep, s.qv.wq, nw, gap.data(), match.data(), rr.data());
        });
    }
    std::vector<uint64_t> rel(nq + 1);
    for (uint64_t i = 0; i <= nq; ++i) rel[i] = offsets[i] - offsets[0];
    return emu_run_rle(gap.data(), match.data(), rr.data(), s.qv, nw, rel.data(), nq, max_gap_len, out7, cap, rle_offsets);
}

// ---- host refinement logic (refine_host.cpp) driven by emulated-kernel MS -------------------------
static void emu_single_ms(EmuIndex* e, const uint8_t* seq, uint64_t len, uint32_t thr, std::vector<uint8_t>* d,
                          std::vector<uint32_t>* l, std::vector<uint32_t>* r, std::vector<uint8_t>* chars) {
    const uint64_t offsets[2] = {0, len};
    Staged s;
    stage_and_ms(e, seq, offsets, 1, 0, true, nullptr, &s);
    d->assign(s.ms.begin(), s.ms.begin() + len);
    l->assign(s.l.begin(), s.l.begin() + len);
    r->assign(s.r.begin(), s.r.begin() + len);
    if (thr) {
        chars->assign(len + 16, 0);
        TrParams tp;
        tp.ms = s.ms.data(); tp.q = s.qv; tp.k = e->host.k; tp.thr = thr; tp.out = chars->data(); tp.off0 = 0;
        emu_run_translate(tp, s.g);
        chars->resize(len);
    }
}

// ---- refinement on the "device" (refine.cuh), as capi.cu device_fill_gaps / device_access_kmers drive it ----------
static int g_emu_device_refine = 0;
extern "C" void emu_set_device_refine(int v) { g_emu_device_refine = v; }
static uint64_t g_emu_device_gaps = 0;
extern "C" uint64_t emu_device_gap_count() { return g_emu_device_gaps; }

struct EmuNodeKeys {
    std::vector<uint64_t> keys;
    NodeKeysView view;
};
static bool emu_node_keys(const EmuIndex* e, EmuNodeKeys* nk) {
    const HostIndex& h = e->host;
    if (h.node_len.empty()) return false;
    const size_t n = (size_t)h.n_sets;
    if (h.node_lo.empty()) {
        nk->keys = h.node_hi;
        nk->view.words = 1;
    } else {
        nk->keys.resize(2 * n);
        for (size_t i = 0; i < n; ++i) { nk->keys[2 * i] = h.node_lo[i]; nk->keys[2 * i + 1] = h.node_hi[i]; }
        nk->view.words = 2;
    }
    nk->view.keys = nk->keys.data();
    nk->view.len = h.node_len.data();
    return true;
}

static void emu_device_fill_gaps(EmuIndex* e, const EmuNodeKeys& nk, std::vector<uint8_t>* chars, const MsArrays& ms,
                                 const uint8_t* ref_seq, uint64_t len, uint32_t thr, double p) {
    if (len == 0) throw RefinePanic{"gap_filling.rs:453-454"};
    if (len < thr) throw RefinePanic{"gap_filling.rs:467 usize underflow"};
    if (len <= 2ull * thr + 1) return;
    const uint64_t n_pos = len - 2ull * thr - 1;
    std::vector<uint2> gaps((size_t)len);
    unsigned int n_gaps = 0;
    emu_launch_seq((unsigned)((n_pos + 255) / 256), 256,
                   [&]() { gap_list_kernel(chars->data(), len, thr, gaps.data(), (uint32_t)gaps.size(), &n_gaps); });
    g_emu_device_gaps += n_gaps;
    if (!n_gaps) return;
    std::vector<double> terms(544);
    for (size_t m = 0; m < terms.size(); ++m) terms[m] = gap_run_log_term(m);
    std::vector<uint8_t> arena((size_t)(len + (uint64_t)n_gaps * thr + 16));
    unsigned long long used = 0, panic = ~0ull;
    FillGapsParams fp;
    fp.ix = e->view;
    fp.nk = nk.view;
    fp.l = ms.l; fp.r = ms.r;
    fp.ref = ref_seq;
    fp.aln = chars->data();
    fp.n = len;
    fp.thr = thr;
    fp.run_terms = terms.data();
    fp.n_terms = (uint32_t)terms.size();
    fp.log_bound = std::log1p(-p);
    fp.gaps = gaps.data();
    fp.n_gaps = n_gaps;
    fp.arena = arena.data();
    fp.arena_used = &used;
    fp.panic = &panic;
    emu_launch_seq(3, 64, [&]() { fill_gaps_kernel(fp); });  // (fewer threads than gaps: the grid-stride loop runs)
    if (used > arena.size()) throw RefinePanic{"emu: arena overflow"};
    if (panic != ~0ull) throw RefinePanic{"gap_filling.rs panic on the device"};
}

static std::vector<VariantRec> emu_call_impl(EmuIndex* q, const uint8_t* ref_seq, uint64_t len, uint64_t thr,
                                             uint32_t build_k, int revcomp, const MsArrays& ms) {
    EmuIndex refix;
    const uint8_t* seqs[1] = {ref_seq};
    const uint64_t lens[1] = {len};
    std::string msg = build_host_index(seqs, lens, 1, build_k, revcomp != 0, 1, &refix.host);
    if (!msg.empty()) throw RefinePanic{msg};
    finish(&refix);
    if (refix.host.k != q->host.k) throw RefinePanic{"lib.rs:559 k mismatch"};
    KmerMsFn fn = [&](int which, const uint8_t* kmers, uint64_t n_kmers, uint32_t k, uint8_t* d_out) {
        std::vector<uint64_t> off(n_kmers + 1);
        for (uint64_t i = 0; i <= n_kmers; ++i) off[i] = i * k;
        Staged s;
        stage_and_ms(which == 0 ? q : &refix, kmers, off.data(), n_kmers, 0, false, nullptr, &s);
        unpad<uint8_t>(s.ms.data(), s.qv, d_out);
    };
    EmuNodeKeys nk;
    if (g_emu_device_refine && emu_node_keys(q, &nk)) {
        AccessKmersFn access = [&](const std::vector<VariantCandidate64>& cs, uint32_t k, uint8_t* out) {
            std::vector<uint32_t> nodes(cs.size());
            for (size_t i = 0; i < cs.size(); ++i) nodes[i] = (uint32_t)cs[i].node;
            emu_launch_seq((unsigned)((cs.size() + 127) / 128), 128,
                           [&]() { access_kmers_kernel(nk.view, k, nodes.data(), nodes.size(), out); });
        };
        return call_variants_from(q->host, find_variant_candidates(ms, len, q->host.k, thr), ref_seq, len, thr, fn, &access);
    }
    return call_variants(q->host, ms, ref_seq, len, thr, fn);
}

static int64_t emu_pack_variants(const std::vector<VariantRec>& vs, uint64_t* pos, uint32_t* qlen, uint32_t* rlen,
                                 uint8_t* qchars, uint8_t* rchars) {
    size_t qo = 0, ro = 0;
    for (size_t i = 0; i < vs.size(); ++i) {
        pos[i] = vs[i].query_pos;
        qlen[i] = (uint32_t)vs[i].query_chars.size();
        rlen[i] = (uint32_t)vs[i].ref_chars.size();
        std::memcpy(qchars + qo, vs[i].query_chars.data(), qlen[i]);
        std::memcpy(rchars + ro, vs[i].ref_chars.data(), rlen[i]);
        qo += qlen[i];
        ro += rlen[i];
    }
    return (int64_t)vs.size();
}

// kbo::call with the product's host logic; returns the number of variants or -1 (panic)
int64_t emu_call(void* h_query, const uint8_t* ref_seq, uint64_t len, uint64_t thr, uint32_t build_k, int revcomp,
                 uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars, uint8_t* rchars) {
    EmuIndex* q = (EmuIndex*)h_query;
    try {
        std::vector<uint8_t> d, chars;
        std::vector<uint32_t> l, r;
        emu_single_ms(q, ref_seq, len, 0, &d, &l, &r, &chars);
        MsArrays ms;
        ms.d = d.data(); ms.l = l.data(); ms.r = r.data(); ms.n = len;
        return emu_pack_variants(emu_call_impl(q, ref_seq, len, thr, build_k, revcomp, ms), pos, qlen, rlen, qchars,
                                 rchars);
    } catch (const RefinePanic&) {
        return -1;
    }
}

// kbo::map with the product's host logic; returns 0 or -1 (panic)
int emu_map(void* h_query, const uint8_t* ref_seq, uint64_t len, uint32_t thr, uint32_t call_thr, double p, int do_fill,
            int do_call, int format, uint32_t build_k, int revcomp, uint8_t* out) {
    EmuIndex* q = (EmuIndex*)h_query;
    try {
        std::vector<uint8_t> d, chars;
        std::vector<uint32_t> l, r;
        emu_single_ms(q, ref_seq, len, thr, &d, &l, &r, &chars);
        MsArrays ms;
        ms.d = d.data(); ms.l = l.data(); ms.r = r.data(); ms.n = len;
        EmuNodeKeys nk;
        if (do_fill && g_emu_device_refine && emu_node_keys(q, &nk)) emu_device_fill_gaps(q, nk, &chars, ms, ref_seq, len, thr, p);
        else if (do_fill) fill_gaps(&chars, ms, ref_seq, len, q->host, thr, p, 4);  // the threaded branch when there are >= 64 gaps
        if (do_call) add_variants(&chars, emu_call_impl(q, ref_seq, len, call_thr ? call_thr : thr, build_k, revcomp, ms));
        for (uint64_t i = 0; i < len; ++i) {
            const uint8_t a = chars[i];
            if (!format) out[i] = a;
            else if (a == 'M' || a == 'R' || a == 'I') out[i] = ref_seq[i];
            else if (a == 'X' || a == 'D' || a == '-') out[i] = '-';
            else out[i] = a;
        }
        return 0;
    } catch (const RefinePanic&) {
        return -1;
    }
}

void emu_translate_i64(const int64_t* d, uint64_t n, uint32_t k, uint32_t thr, uint8_t* out) {
    emu_launch_seq((unsigned)((n + 255) / 256), 256, [&]() { translate_i64_kernel(d, n, k, thr, out); });
}

}  // extern "C"
