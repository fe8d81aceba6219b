// ===========================================================================
// kbo_b200/csrc/sbwt_host.hpp -- host side of index::build_sbwt_from_vecs
// (reference src/index.rs:56-99 -> sbwt::SbwtIndexBuilder) and of the small
// index lookups the reference performs on the host (SbwtIndex::search,
// access_kmer; gap_filling.rs:144,217, variant_calling.rs:276).
//
// Produces the plain SubsetMatrix form (4 bit rows + LCS bytes + C array) that
// index_upload re-lays out for the GPU.  Semantics: SURVEY.md section 8c.
// ===========================================================================
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kbo_b200 {

struct HostIndex {
    uint32_t k = 0;
    uint64_t n_sets = 0;
    uint64_t n_kmers = 0;
    uint64_t C[4] = {0, 0, 0, 0};
    std::vector<uint64_t> rows[4];      // ceil(n_sets/64)+1 words each
    std::vector<uint32_t> row_cum[4];   // set bits before each 64-bit word (for host rank/select)
    std::vector<uint8_t> lcs;           // n_sets bytes
    // optional "select support" (BuildOpts.build_select): the colex-sorted nodes themselves, letters packed
    // 2 bits each with the LAST base in the top bits of (hi, lo); node_len = number of non-'$' bases
    std::vector<uint64_t> node_hi, node_lo;
    std::vector<uint8_t> node_len;

    uint64_t rank(int c, uint64_t p) const;                  // set bits of row c in [0,p)
    uint64_t select(int c, uint64_t j) const;                // position of the j-th (0-based) set bit of row c
    bool extend_right(uint64_t& l, uint64_t& r, uint8_t ch) const;
    bool search(const uint8_t* pat, uint64_t len, uint64_t* l, uint64_t* r) const;
    void access_kmer(uint64_t colex, uint8_t* out_k) const;  // k bytes, '$' padded
    void finalize();                                          // row_cum + C from rows
};

// Returns empty string on success, else an error message.
// keep_nodes: also keep the sorted nodes for O(1) access_kmer (BuildOpts.build_select).
std::string build_host_index(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k,
                             bool add_revcomp, uint32_t num_threads, HostIndex* out, bool keep_nodes = true);

}  // namespace kbo_b200
