// ===========================================================================
// kbo_b200/csrc/refine_host.hpp -- host logic that consumes the GPU path's
// (d, l, r) output: variant calling (reference src/variant_calling.rs),
// gap filling (src/gap_filling.rs) and add_variants (src/translate.rs:350-386).
//
// These are the sequential, low-volume callers on either side of the hot path
// (SURVEY.md section 8f rows 2 and 4).  They never compute matching statistics
// themselves: full-length MS arrives as arrays from K1, and the 2 x k-length MS
// runs per variant candidate are handed back to the caller IN ONE BATCH through
// `KmerMsFn`, which the C ABI layer implements with a K0+K1 launch.
// ===========================================================================
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "sbwt_host.hpp"

namespace kbo_b200 {

struct MsArrays {  // index::query_sbwt output in device widths
    const uint8_t* d = nullptr;
    const uint32_t* l = nullptr;
    const uint32_t* r = nullptr;
    uint64_t n = 0;
};

struct VariantRec {  // variant_calling::Variant (variant_calling.rs:9-19)
    uint64_t query_pos = 0;
    std::vector<uint8_t> query_chars;
    std::vector<uint8_t> ref_chars;
};

// Raised where the reference would panic.
struct RefinePanic {
    std::string what;
};

// Computes MS lengths of `n_kmers` strings of k bytes each (concatenated in `kmers`) against one
// index; writes n_kmers * k bytes to d_out.  which == 0: the index called `sbwt_ref` in
// call_variants (variant_calling.rs:249), which == 1: `sbwt_query`.
typedef std::function<void(int which, const uint8_t* kmers, uint64_t n_kmers, uint32_t k, uint8_t* d_out)> KmerMsFn;

// variant_calling::resolve_variant (variant_calling.rs:139-201) on MS length vectors of k entries.
bool resolve_variant(const uint8_t* query_kmer, const uint8_t* ref_kmer, const uint8_t* ms_vs_query_d,
                     const uint8_t* ms_vs_ref_d, uint32_t k, uint64_t significant_match_threshold,
                     std::vector<uint8_t>* query_chars, std::vector<uint8_t>* ref_chars);

// A variant candidate (variant_calling.rs:268-272): a significant drop of the MS at query position i followed, within
// k positions, by a significant match at j whose interval is the single node `node`.
struct VariantCandidate64 {
    uint64_t i, j, node;
};
// the candidate scan on host arrays (kbo::map, which needs the arrays on the host anyway for fill_gaps)
std::vector<VariantCandidate64> find_variant_candidates(const MsArrays& ms_vs_ref, uint64_t len, uint32_t k, uint64_t threshold);
// SbwtIndex::access_kmer (variant_calling.rs:276) for the nodes of all candidates at once: k bytes each into `out`
// (the C ABI layer implements it with access_kmers_kernel when the index keeps its node keys on the device)
typedef std::function<void(const std::vector<VariantCandidate64>& cands, uint32_t k, uint8_t* out)> AccessKmersFn;
// call_variants given the candidates in increasing i (kbo::call finds them on the device: variant_candidates_kernel);
// access == nullptr: the k-mers come from sbwt_ref on the host, else sbwt_ref is only asked for k
std::vector<VariantRec> call_variants_from(const HostIndex& sbwt_ref, const std::vector<VariantCandidate64>& cands,
                                           const uint8_t* query, uint64_t len, uint64_t threshold, const KmerMsFn& kmer_ms,
                                           const AccessKmersFn* access = nullptr);

// variant_calling::call_variants (variant_calling.rs:249-294).  `ms_vs_ref` = MS of `query` against
// `sbwt_ref` (already computed on the GPU); `threshold` = random_match_threshold(k, sbwt_ref.n_kmers, 4, p).
std::vector<VariantRec> call_variants(const HostIndex& sbwt_ref, const MsArrays& ms_vs_ref, const uint8_t* query,
                                      uint64_t len, uint64_t threshold, const KmerMsFn& kmer_ms);

// translate::add_variants (translate.rs:350-386), in place on a byte alignment.
void add_variants(std::vector<uint8_t>* translation, const std::vector<VariantRec>& variants);

// ln(1 - (1/4)^m) exactly as gap_filling.rs:489-501 evaluates it for a run of m - 2 consecutive agreements; the device
// version of fill_gaps reads these values from a table made with this function
double gap_run_log_term(uint64_t m);

// gap_filling::fill_gaps (gap_filling.rs:444-526), in place on a byte alignment.
void fill_gaps(std::vector<uint8_t>* translation, const MsArrays& noisy_ms, const uint8_t* ref_seq, uint64_t len,
               const HostIndex& query_sbwt, uint64_t threshold, double max_err_prob, uint32_t num_threads = 1);

}  // namespace kbo_b200
