// ============================================================================
// oracle/kbo_oracle.hpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the kbo 0.5.1 hot path (k-bounded matching statistics over
// an SBWT index -> derandomize -> translate) and of the host logic either side
// of it, used ONLY as the checker in tests/, __graft_entry__.smoke() and as the
// `cpu_baseline` / `--impl reference` leg of bench.py.  Nothing under
// kbo_b200/ may include, link or call anything in this directory.
//
// Parity pinning: the in-tree parts (derandomize.rs, translate.rs, format.rs,
// variant_calling.rs, gap_filling.rs, lib.rs) are restated line by line and
// are pinned by the reference's own unit tests / doctests, transcribed in
// tests/test_oracle_golden.py.  The SBWT engine itself lives in the crate
// `sbwt = "0.3.4"` (reference Cargo.toml:18), which is NOT vendored under
// /root/reference and cannot be built here (no cargo/rustc, no network); it is
// restated from its published algorithm and pinned
// through the reference's call sites and golden vectors that run through it
// (index.rs:264-274, lib.rs:526-545,600-610,647-661,670-717,786-806,
// gap_filling.rs:535-922, variant_calling.rs:312-553, translate.rs:535-676).
// Unpinned edges (no reference test covers them): add_revcomp=true, query
// bytes other than upper-case ACGT, absolute colex interval values, the
// .sbwt/.lcs on-disk format.
// ============================================================================
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace kbo_oracle {

typedef unsigned __int128 u128;

// One node of the padded k-mer set P.  `key` holds the letters
// 2 bits each (A=0,C=1,G=2,T=3), LAST character of the k-mer in the two most
// significant bits (colex-major); `len` = number of non-'$' characters.
struct Node {
    u128 key;
    uint8_t len;
};

struct Index {
    int k = 0;
    size_t n_sets = 0;
    size_t n_kmers = 0;
    size_t C[4] = {0, 0, 0, 0};
    std::vector<uint64_t> bits[4];   // subset-matrix rows, n_sets bits each
    std::vector<uint64_t> cum[4];    // rank samples: set bits before each 512-bit block
    std::vector<uint8_t> lcs;        // LCS[i] = longest common suffix of P[i-1], P[i]; LCS[0]=0
    std::vector<Node> nodes;         // sorted P (stands in for select support / access_kmer)

    size_t rank(int c, size_t p) const;
    bool get_bit(int c, size_t i) const { return (bits[c][i >> 6] >> (i & 63)) & 1; }
};

struct MsEntry {
    size_t d;
    size_t l;
    size_t r;  // colex interval [l, r)
};

struct Variant {
    size_t query_pos;
    std::vector<uint8_t> query_chars;
    std::vector<uint8_t> ref_chars;
    bool operator==(const Variant& o) const {
        return query_pos == o.query_pos && query_chars == o.query_chars && ref_chars == o.ref_chars;
    }
};

struct RLE {
    size_t start = 0, end = 0, matches = 0, mismatches = 0, jumps = 0, gap_bases = 0, gap_opens = 0;
};

// Thrown where the reference would panic (assert!/index out of bounds/unwrap).
struct Panic {
    std::string what;
};

// --- index (restates sbwt 0.3.4 builder semantics) -----------------
Index build_index(const std::vector<std::vector<uint8_t>>& seqs, int k, bool add_revcomp);
std::vector<MsEntry> matching_statistics(const Index& ix, const uint8_t* q, size_t len);
std::vector<MsEntry> query_sbwt(const Index& ix, const uint8_t* q, size_t len);  // index.rs:243-256
bool search(const Index& ix, const uint8_t* pat, size_t len, size_t* l, size_t* r);
std::vector<uint8_t> access_kmer(const Index& ix, size_t colex);

// --- derandomize.rs -----------------------------------------------------------
double log_rm_max_cdf(size_t t, size_t alphabet_size, size_t n_kmers);
size_t random_match_threshold(size_t k, size_t n_kmers, size_t alphabet_size, double max_error_prob);
int64_t derandomize_ms_val(size_t curr_noisy_ms, int64_t next_derand_ms, size_t threshold, size_t k);
std::vector<int64_t> derandomize_ms_vec(const std::vector<size_t>& noisy_ms, size_t k, size_t threshold);

// --- translate.rs -------------------------------------------------------------
std::pair<char, char> translate_ms_val(int64_t ms_curr, int64_t ms_next, int64_t ms_prev, size_t threshold);
std::vector<char> translate_ms_vec(const std::vector<int64_t>& derand_ms, size_t k, size_t threshold);
std::vector<char> add_variants(const std::vector<char>& translation, const std::vector<Variant>& variants);

// --- format.rs ----------------------------------------------------------------
std::vector<RLE> run_lengths_gapped(const std::vector<char>& aln, size_t max_gap_len);
std::vector<RLE> run_lengths(const std::vector<char>& aln);
std::vector<uint8_t> relative_to_ref(const uint8_t* ref_seq, size_t len, const std::vector<char>& alignment);

// --- variant_calling.rs -------------------------------------------------------
bool resolve_variant(const std::vector<uint8_t>& query_kmer, const std::vector<uint8_t>& ref_kmer,
                     const std::vector<MsEntry>& ms_vs_query, const std::vector<MsEntry>& ms_vs_ref,
                     size_t significant_match_threshold, std::vector<uint8_t>* query_chars,
                     std::vector<uint8_t>* ref_chars);
std::vector<Variant> call_variants(const Index& sbwt_ref, const Index& sbwt_query, const uint8_t* query,
                                   size_t len, double max_error_prob);

// --- gap_filling.rs -----------------------------------------------------------
std::pair<size_t, std::vector<uint8_t>> nearest_unique_context(const std::vector<MsEntry>& ms, const Index& sbwt,
                                                                size_t range_start, size_t range_end);
std::vector<uint8_t> left_extend_kmer(const std::vector<uint8_t>& kmer_start, const Index& sbwt,
                                      size_t max_extension_len);
std::vector<uint8_t> left_extend_over_gap(const std::vector<MsEntry>& ms, const uint8_t* ref_seq, size_t ref_len,
                                          const Index& sbwt, size_t left_overlap_req, size_t right_overlap_req,
                                          size_t gap_start, size_t gap_end, size_t search_radius);
std::vector<char> fill_gaps(const std::vector<char>& translation, const std::vector<MsEntry>& noisy_ms,
                            const uint8_t* ref_seq, size_t len, const Index& query_sbwt, size_t threshold,
                            double max_err_prob);

// --- lib.rs -------------------------------------------------------------------
struct MapOpts {
    double max_error_prob = 0.0000001;
    bool fill_gaps = true;
    bool call_variants = true;
    bool format = true;
    int build_k = 31;
    bool build_add_revcomp = false;
};
std::vector<char> matches(const Index& ix, const uint8_t* q, size_t len, double max_error_prob);
std::vector<RLE> find(const Index& ix, const uint8_t* q, size_t len, double max_error_prob, size_t max_gap_len);
std::vector<Variant> call(const Index& sbwt_query, const uint8_t* ref_seq, size_t len, double max_error_prob,
                          int build_k, bool build_add_revcomp);
std::vector<uint8_t> map(const Index& query_sbwt, const uint8_t* ref_seq, size_t len, const MapOpts& opts);

}  // namespace kbo_oracle
