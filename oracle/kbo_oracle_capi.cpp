// ============================================================================
// oracle/kbo_oracle_capi.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface of the CPU oracle so that tests/ (ctypes) and bench.py's
// cpu_baseline / --impl reference legs can drive it.  Status convention:
// >= 0 ok (often a count), -1 = the reference would have panicked (message via
// kbo_oracle_last_error), -2 = caller buffer too small.
// ============================================================================
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "kbo_oracle.hpp"

using namespace kbo_oracle;

static thread_local std::string g_err;

#define ORACLE_TRY try {
#define ORACLE_CATCH                  \
    }                                 \
    catch (const Panic& p) {          \
        g_err = p.what;               \
        return -1;                    \
    }                                 \
    catch (const std::exception& e) { \
        g_err = e.what();             \
        return -1;                    \
    }

extern "C" {

const char* kbo_oracle_last_error() { return g_err.c_str(); }

void* kbo_oracle_build(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, int k, int add_revcomp) {
    try {
        std::vector<std::vector<uint8_t>> v;
        for (uint64_t i = 0; i < n_seqs; ++i) v.emplace_back(seqs[i], seqs[i] + lens[i]);
        return new Index(build_index(v, k, add_revcomp != 0));
    } catch (const Panic& p) {
        g_err = p.what;
        return nullptr;
    }
}
void kbo_oracle_free(void* h) { delete (Index*)h; }
int kbo_oracle_k(void* h) { return ((Index*)h)->k; }
uint64_t kbo_oracle_n_sets(void* h) { return ((Index*)h)->n_sets; }
uint64_t kbo_oracle_n_kmers(void* h) { return ((Index*)h)->n_kmers; }
void kbo_oracle_C(void* h, uint64_t* out4) {
    for (int c = 0; c < 4; ++c) out4[c] = ((Index*)h)->C[c];
}
// rows: 4 arrays of ceil(n_sets/64) u64 words, bit i of row c = node i carries label c
void kbo_oracle_rows(void* h, uint64_t* a, uint64_t* c, uint64_t* g, uint64_t* t) {
    Index* ix = (Index*)h;
    size_t nw = (ix->n_sets + 63) / 64;
    uint64_t* outs[4] = {a, c, g, t};
    for (int ch = 0; ch < 4; ++ch) std::memcpy(outs[ch], ix->bits[ch].data(), nw * 8);
}
void kbo_oracle_lcs(void* h, uint8_t* out) {
    Index* ix = (Index*)h;
    std::memcpy(out, ix->lcs.data(), ix->n_sets);
}
int kbo_oracle_access_kmer(void* h, uint64_t colex, uint8_t* out_k) {
    ORACLE_TRY
    auto s = access_kmer(*(Index*)h, colex);
    std::memcpy(out_k, s.data(), s.size());
    return 0;
    ORACLE_CATCH
}
int kbo_oracle_search(void* h, const uint8_t* pat, uint64_t len, uint64_t* l, uint64_t* r) {
    size_t a, b;
    if (!search(*(Index*)h, pat, len, &a, &b)) return 0;
    *l = a;
    *r = b;
    return 1;
}

// index.rs:243 query_sbwt
int kbo_oracle_query_sbwt(void* h, const uint8_t* q, uint64_t len, uint64_t* d, uint64_t* l, uint64_t* r) {
    ORACLE_TRY
    auto ms = query_sbwt(*(Index*)h, q, len);
    for (size_t i = 0; i < ms.size(); ++i) {
        d[i] = ms[i].d;
        if (l) l[i] = ms[i].l;
        if (r) r[i] = ms[i].r;
    }
    return 0;
    ORACLE_CATCH
}

int kbo_oracle_log_rm_max_cdf(uint64_t t, uint64_t s, uint64_t n, double* out) {
    ORACLE_TRY
    *out = log_rm_max_cdf(t, s, n);
    return 0;
    ORACLE_CATCH
}
int64_t kbo_oracle_random_match_threshold(uint64_t k, uint64_t n_kmers, uint64_t s, double p) {
    ORACLE_TRY
    return (int64_t)random_match_threshold(k, n_kmers, s, p);
    ORACLE_CATCH
}
int kbo_oracle_derandomize_ms_val(uint64_t cur, int64_t next, uint64_t thr, uint64_t k, int64_t* out) {
    ORACLE_TRY
    *out = derandomize_ms_val(cur, next, thr, k);
    return 0;
    ORACLE_CATCH
}
int kbo_oracle_derandomize_ms_vec(const uint64_t* ms, uint64_t n, uint64_t k, uint64_t thr, int64_t* out) {
    ORACLE_TRY
    std::vector<size_t> v(ms, ms + n);
    auto d = derandomize_ms_vec(v, k, thr);
    std::memcpy(out, d.data(), n * 8);
    return 0;
    ORACLE_CATCH
}
int kbo_oracle_translate_ms_val(int64_t cur, int64_t next, int64_t prev, uint64_t thr, char* out2) {
    ORACLE_TRY
    auto p = translate_ms_val(cur, next, prev, thr);
    out2[0] = p.first;
    out2[1] = p.second;
    return 0;
    ORACLE_CATCH
}
int kbo_oracle_translate_ms_vec(const int64_t* derand, uint64_t n, uint64_t k, uint64_t thr, char* out) {
    ORACLE_TRY
    std::vector<int64_t> v(derand, derand + n);
    auto t = translate_ms_vec(v, k, thr);
    std::memcpy(out, t.data(), n);
    return 0;
    ORACLE_CATCH
}
int kbo_oracle_matches(void* h, const uint8_t* q, uint64_t len, double p, char* out) {
    ORACLE_TRY
    auto t = matches(*(Index*)h, q, len, p);
    std::memcpy(out, t.data(), t.size());
    return 0;
    ORACLE_CATCH
}

static int64_t pack_rles(const std::vector<RLE>& rl, uint64_t* out, uint64_t cap) {
    if (rl.size() > cap) return -2;
    for (size_t i = 0; i < rl.size(); ++i) {
        uint64_t* o = out + 7 * i;
        o[0] = rl[i].start;
        o[1] = rl[i].end;
        o[2] = rl[i].matches;
        o[3] = rl[i].mismatches;
        o[4] = rl[i].jumps;
        o[5] = rl[i].gap_bases;
        o[6] = rl[i].gap_opens;
    }
    return (int64_t)rl.size();
}
// format.rs:143 run_lengths_gapped on a char alignment; out = 7 u64 per RLE
int64_t kbo_oracle_run_lengths_gapped(const char* aln, uint64_t n, uint64_t max_gap_len, uint64_t* out, uint64_t cap) {
    ORACLE_TRY
    std::vector<char> v(aln, aln + n);
    return pack_rles(run_lengths_gapped(v, max_gap_len), out, cap);
    ORACLE_CATCH
}
int64_t kbo_oracle_find(void* h, const uint8_t* q, uint64_t len, double p, uint64_t max_gap_len, uint64_t* out,
                        uint64_t cap) {
    ORACLE_TRY
    return pack_rles(find(*(Index*)h, q, len, p, max_gap_len), out, cap);
    ORACLE_CATCH
}
int kbo_oracle_relative_to_ref(const uint8_t* ref, uint64_t len, const char* aln, uint64_t alen, uint8_t* out) {
    ORACLE_TRY
    std::vector<char> v(aln, aln + alen);
    auto o = relative_to_ref(ref, len, v);
    std::memcpy(out, o.data(), o.size());
    return (int)o.size();
    ORACLE_CATCH
}

static int64_t pack_variants(const std::vector<Variant>& vs, uint64_t* pos, uint32_t* qlen, uint32_t* rlen,
                             uint8_t* qchars, uint8_t* rchars, uint64_t cap_var, uint64_t cap_chars) {
    if (vs.size() > cap_var) return -2;
    size_t qo = 0, ro = 0;
    for (size_t i = 0; i < vs.size(); ++i) {
        pos[i] = vs[i].query_pos;
        qlen[i] = (uint32_t)vs[i].query_chars.size();
        rlen[i] = (uint32_t)vs[i].ref_chars.size();
        if (qo + qlen[i] > cap_chars || ro + rlen[i] > cap_chars) return -2;
        std::memcpy(qchars + qo, vs[i].query_chars.data(), qlen[i]);
        std::memcpy(rchars + ro, vs[i].ref_chars.data(), rlen[i]);
        qo += qlen[i];
        ro += rlen[i];
    }
    return (int64_t)vs.size();
}
// lib.rs:547 call(sbwt_query, lcs_query, ref_seq, CallOpts{max_error_prob, sbwt_build_opts{k, add_revcomp}})
int64_t kbo_oracle_call(void* h_query, const uint8_t* ref_seq, uint64_t len, double p, int build_k, int build_revcomp,
                        uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars, uint8_t* rchars,
                        uint64_t cap_var, uint64_t cap_chars) {
    ORACLE_TRY
    auto vs = call(*(Index*)h_query, ref_seq, len, p, build_k, build_revcomp != 0);
    return pack_variants(vs, pos, qlen, rlen, qchars, rchars, cap_var, cap_chars);
    ORACLE_CATCH
}
// variant_calling.rs:249 call_variants(sbwt_ref, lcs_ref, sbwt_query, lcs_query, query, p)
int64_t kbo_oracle_call_variants(void* h_ref, void* h_query, const uint8_t* query, uint64_t len, double p,
                                 uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars, uint8_t* rchars,
                                 uint64_t cap_var, uint64_t cap_chars) {
    ORACLE_TRY
    auto vs = call_variants(*(Index*)h_ref, *(Index*)h_query, query, len, p);
    return pack_variants(vs, pos, qlen, rlen, qchars, rchars, cap_var, cap_chars);
    ORACLE_CATCH
}
// translate.rs:350 add_variants
int kbo_oracle_add_variants(const char* aln, uint64_t n, uint64_t n_var, const uint64_t* pos, const uint32_t* qlen,
                            const uint32_t* rlen, const uint8_t* qchars, const uint8_t* rchars, char* out) {
    ORACLE_TRY
    std::vector<char> t(aln, aln + n);
    std::vector<Variant> vs;
    size_t qo = 0, ro = 0;
    for (uint64_t i = 0; i < n_var; ++i) {
        Variant v;
        v.query_pos = pos[i];
        v.query_chars.assign(qchars + qo, qchars + qo + qlen[i]);
        v.ref_chars.assign(rchars + ro, rchars + ro + rlen[i]);
        qo += qlen[i];
        ro += rlen[i];
        vs.push_back(v);
    }
    auto r = add_variants(t, vs);
    std::memcpy(out, r.data(), n);
    return 0;
    ORACLE_CATCH
}
// gap_filling.rs:444 fill_gaps (MS recomputed from the index so the caller passes plain buffers)
int kbo_oracle_fill_gaps(void* h_query, const char* translation, const uint8_t* ref_seq, uint64_t len, uint64_t thr,
                         double p, char* out) {
    ORACLE_TRY
    Index& ix = *(Index*)h_query;
    auto ms = query_sbwt(ix, ref_seq, len);
    std::vector<char> t(translation, translation + len);
    auto r = fill_gaps(t, ms, ref_seq, len, ix, thr, p);
    std::memcpy(out, r.data(), len);
    return 0;
    ORACLE_CATCH
}
// gap_filling.rs:127 nearest_unique_context over MS of `ref_seq` against the index
int64_t kbo_oracle_nearest_unique_context(void* h, const uint8_t* ref_seq, uint64_t len, uint64_t range_start,
                                          uint64_t range_end, uint64_t* kmer_idx, uint8_t* kmer_out) {
    ORACLE_TRY
    Index& ix = *(Index*)h;
    auto ms = query_sbwt(ix, ref_seq, len);
    auto r = nearest_unique_context(ms, ix, range_start, range_end);
    *kmer_idx = r.first;
    std::memcpy(kmer_out, r.second.data(), r.second.size());
    return (int64_t)r.second.size();
    ORACLE_CATCH
}
// gap_filling.rs:205 left_extend_kmer
int64_t kbo_oracle_left_extend_kmer(void* h, const uint8_t* kmer, uint64_t len, uint64_t max_ext, uint8_t* out,
                                    uint64_t cap) {
    ORACLE_TRY
    std::vector<uint8_t> k0(kmer, kmer + len);
    auto r = left_extend_kmer(k0, *(Index*)h, max_ext);
    if (r.size() > cap) return -2;
    std::memcpy(out, r.data(), r.size());
    return (int64_t)r.size();
    ORACLE_CATCH
}
// gap_filling.rs:295 left_extend_over_gap over MS of `ref_seq` against the index
int64_t kbo_oracle_left_extend_over_gap(void* h, const uint8_t* ref_seq, uint64_t len, uint64_t left_req,
                                        uint64_t right_req, uint64_t gap_start, uint64_t gap_end, uint64_t radius,
                                        uint8_t* out, uint64_t cap) {
    ORACLE_TRY
    Index& ix = *(Index*)h;
    auto ms = query_sbwt(ix, ref_seq, len);
    auto r = left_extend_over_gap(ms, ref_seq, len, ix, left_req, right_req, gap_start, gap_end, radius);
    if (r.size() > cap) return -2;
    std::memcpy(out, r.data(), r.size());
    return (int64_t)r.size();
    ORACLE_CATCH
}
// lib.rs:720 map
int64_t kbo_oracle_map(void* h_query, const uint8_t* ref_seq, uint64_t len, double p, int do_fill_gaps,
                       int do_call_variants, int do_format, int build_k, int build_revcomp, uint8_t* out) {
    ORACLE_TRY
    MapOpts o;
    o.max_error_prob = p;
    o.fill_gaps = do_fill_gaps != 0;
    o.call_variants = do_call_variants != 0;
    o.format = do_format != 0;
    o.build_k = build_k;
    o.build_add_revcomp = build_revcomp != 0;
    auto r = map(*(Index*)h_query, ref_seq, len, o);
    std::memcpy(out, r.data(), r.size());
    return (int64_t)r.size();
    ORACLE_CATCH
}

// ---------------------------------------------------------------------------
// CPU baseline driver: kbo::matches over a batch of queries, one std::thread per
// requested core pulling queries from a shared counter (the analogue of
// kbo-cli's per-query threading; the library itself is single-threaded).
// Returns wall seconds; out may be null (results discarded after checksum).
// ---------------------------------------------------------------------------
double kbo_oracle_matches_batch(void* h, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries, double p,
                                char* out, int n_threads, uint64_t* checksum) {
    Index& ix = *(Index*)h;
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint64_t> next(0);
    std::atomic<uint64_t> sum(0);
    std::atomic<int> failed(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        uint64_t local = 0;
        for (;;) {
            uint64_t qi = next.fetch_add(1);
            if (qi >= n_queries) break;
            try {
                auto t = matches(ix, concat + offsets[qi], offsets[qi + 1] - offsets[qi], p);
                if (out) std::memcpy(out + offsets[qi], t.data(), t.size());
                for (size_t i = 0; i < t.size(); ++i) local = local * 1099511628211ULL + (uint8_t)t[i] + qi;
            } catch (...) {
                failed = 1;
            }
        }
        sum += local;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (checksum) *checksum = sum.load();
    if (failed) return -1.0;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Same driver for kbo::find (lib.rs:808-821): matches + run_lengths[_gapped] per query.
// n_rle_out (optional) receives the total number of RLE records.
double kbo_oracle_find_batch(void* h, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries, double p,
                             uint64_t max_gap_len, int n_threads, uint64_t* n_rle_out, uint64_t* checksum) {
    Index& ix = *(Index*)h;
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint64_t> next(0), sum(0), nrle(0);
    std::atomic<int> failed(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        uint64_t local = 0, cnt = 0;
        for (;;) {
            uint64_t qi = next.fetch_add(1);
            if (qi >= n_queries) break;
            try {
                auto rl = find(ix, concat + offsets[qi], offsets[qi + 1] - offsets[qi], p, max_gap_len);
                cnt += rl.size();
                for (const RLE& r : rl)
                    local = local * 1099511628211ULL + r.start + 3 * r.end + 5 * r.matches + 7 * r.mismatches +
                            11 * r.jumps + 13 * r.gap_bases + 17 * r.gap_opens + qi;
            } catch (...) {
                failed = 1;
            }
        }
        sum += local;
        nrle += cnt;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (checksum) *checksum = sum.load();
    if (n_rle_out) *n_rle_out = nrle.load();
    if (failed) return -1.0;
    return std::chrono::duration<double>(t1 - t0).count();
}

int kbo_oracle_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
