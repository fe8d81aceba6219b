/* ===========================================================================
 * kbo_b200.h -- C ABI of the B200-native (sm_100a) implementation of kbo's
 * k-bounded matching statistics hot path.
 *
 * This is the drop-in boundary: every entry point replaces one function of the
 * reference crate tmaklin/kbo 0.5.1 (paths below are relative to the reference
 * checkout) and is what a thin Rust `kbo-b200-sys` FFI crate would bind (see
 * INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *  - Every function returning `int` returns a kbo_status (0 = ok).  Where the
 *    reference PANICS (assert!/unwrap) the ABI returns a distinct non-zero code
 *    and never aborts or throws; kbo_last_error_message() gives the text and
 *    the reference file:line of the violated precondition.  The Rust shim turns
 *    non-zero into panic!().
 *  - Inputs are borrowed for the duration of the call only; outputs are caller
 *    allocated.  The library never frees caller memory.
 *  - Host-pointer entry points copy host->device and device->host inside the
 *    call.  `_device` entry points take CUDA device pointers (same device as
 *    the index) and a cudaStream_t (as void*; NULL = the legacy default
 *    stream) and are asynchronous only with respect to that stream.  The library
 *    never changes attributes of a caller's stream.
 *  - Widths: the device works in u8 (MS length, alignment characters) and u32
 *    (colex ranks, n_sets < 2^32).  Entry points without a `_compact` suffix
 *    widen to the reference's usize/i64 on the way out; alignment characters
 *    are 1 byte each (Rust `char` is 4; the shim widens).
 *  - All functions are thread-safe; an index handle is immutable after
 *    creation and may be shared by concurrent callers (reference functions take
 *    `&SbwtIndexVariant`, src/lib.rs:612-617).  Concurrent `_device` calls that
 *    pass the same stream are serialised on that stream's workspace.
 *  - Device-wide state the library changes: creating an index raises
 *    cudaLimitPersistingL2CacheSize (never lowered again) so that the library's
 *    own streams can mark the index as persisting in L2.
 *  - There is no CPU fallback: without a CUDA device every compute entry point
 *    fails with KBO_ERR_CUDA.
 * ======================================================================== */
#ifndef KBO_B200_H
#define KBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum kbo_status {
    KBO_OK = 0,
    KBO_ERR_EMPTY_INPUT = 1,      /* index.rs:60 assert!(!slices.is_empty()); index.rs:248 assert!(!query.is_empty()) */
    KBO_ERR_BAD_THRESHOLD = 2,    /* derandomize.rs:228,275; translate.rs:186,269: threshold > 1 */
    KBO_ERR_TOO_SHORT = 3,        /* derandomize.rs:276; translate.rs:270: len > 2 */
    KBO_ERR_BAD_K = 4,            /* derandomize.rs:227,274 k > 0; this build: k <= KBO_MAX_K */
    KBO_ERR_K_MISMATCH = 5,       /* lib.rs:559, lib.rs:729 */
    KBO_ERR_BAD_PROB = 6,         /* derandomize.rs:136-137: 0 < max_error_prob <= 1 */
    KBO_ERR_BAD_ARGUMENT = 7,     /* null pointer, n_kmers == 0 (derandomize.rs:96,134), alphabet == 0, MS value > k (derandomize.rs:229) */
    KBO_ERR_CUDA = 8,             /* no device / CUDA runtime error */
    KBO_ERR_OOM = 9,              /* host or device allocation failed */
    KBO_ERR_INDEX_TOO_LARGE = 10, /* n_sets >= 2^32 */
    KBO_ERR_BUFFER_TOO_SMALL = 11,/* caller capacity too small (count is still returned) */
    KBO_ERR_PANIC = 12,           /* the reference would panic on these inputs (index out of bounds etc.) */
    KBO_ERR_BATCH_TOO_LARGE = 13, /* a device-resident batch of >= 2^32 - 2^20 positions (host-buffer calls split internally) */
    KBO_ERR_IO = 14,              /* index.rs:137,148,202,207: the index file cannot be created / opened (the reference panics) */
    KBO_ERR_FORMAT = 15           /* index.rs:204,209 .unwrap(): not an index file this library can read */
} kbo_status;

#define KBO_MAX_K 64 /* packed k-mers are two 64-bit words; reference tests use k <= 63 */

/* src/lib.rs:259-313 BuildOpts (fields this implementation does not need are accepted and ignored) */
typedef struct kbo_build_opts {
    uint32_t k;              /* default 31 */
    int32_t add_revcomp;     /* default 0 */
    uint32_t num_threads;    /* default 1 (host-side sort threads of the k > 32 builder; nothing else) */
    uint32_t prefix_precalc; /* default 8; ignored (the device keeps its own table of the MS states after 10 bases) */
    int32_t build_select;    /* default 0; != 0 keeps the sorted nodes on the host for O(1) access_kmer (map/call);
                              * without it access_kmer walks the index (slower, same result) */
    uint32_t mem_gb;         /* ignored */
    int32_t dedup_batches;   /* ignored */
    const char* temp_dir;    /* ignored (always in memory) */
} kbo_build_opts;

/* src/format.rs:17-33 RLE */
typedef struct kbo_rle {
    uint64_t start, end, matches, mismatches, jumps, gap_bases, gap_opens;
} kbo_rle;

/* Opaque: (SbwtIndexVariant::SubsetMatrix, sbwt::LcsArray) resident on one GPU in the
 * interleaved rank/LCS layout (DESIGN.md "Index layout"). */
typedef struct kbo_index kbo_index;

/* ---- library ------------------------------------------------------------ */
const char* kbo_last_error_message(void);       /* thread-local */
int kbo_device_count(int* out);
void kbo_default_build_opts(kbo_build_opts* o); /* lib.rs:294-312 */
/* pinned host staging buffers for callers that want full-speed PCIe copies */
int kbo_alloc_pinned(size_t bytes, void** out);
int kbo_free_pinned(void* p);

/* ---- index: src/index.rs ------------------------------------------------ */
/* index::build_sbwt_from_vecs (index.rs:56-99) / kbo::build (lib.rs:501-506) */
int kbo_index_build(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, const kbo_build_opts* opts,
                    int device, kbo_index** out);
/* Upload an index built elsewhere (e.g. by the sbwt crate): 4 subset-matrix bit rows of
 * ceil(n_sets/64) little-endian u64 words (bit i of row c = node i has outgoing label c,
 * A,C,G,T order) and n_sets LCS bytes (LCS[0] == 0, every value < k).  1 <= k <= 127 here (LCS bytes are compared
 * seven bits wide on the device); KBO_MAX_K bounds kbo_index_build only.  The arrays are validated (LCS range, and
 * the rows must hold exactly n_sets - 1 set bits: every node but the root has one incoming edge);
 * KBO_ERR_BAD_ARGUMENT otherwise.  access_kmer on such an index walks the index (no stored nodes). */
int kbo_index_from_parts(uint32_t k, uint64_t n_sets, uint64_t n_kmers, const uint64_t* const rows[4],
                         const uint8_t* lcs, int device, kbo_index** out);
void kbo_index_free(kbo_index* ix);
uint32_t kbo_index_k(const kbo_index* ix);        /* SbwtIndex::k()       lib.rs:618 */
uint64_t kbo_index_n_kmers(const kbo_index* ix);  /* SbwtIndex::n_kmers() lib.rs:620 */
uint64_t kbo_index_n_sets(const kbo_index* ix);   /* SbwtIndex::n_sets()  */
int kbo_index_device(const kbo_index* ix);
uint64_t kbo_index_device_bytes(const kbo_index* ix);
/* Read the index back in the kbo_index_from_parts format (for verification / serialization). */
int kbo_index_export_parts(const kbo_index* ix, uint64_t* rows[4], uint8_t* lcs, uint64_t C_out[4]);
/* index::serialize_sbwt (index.rs:128-153): writes `<outfile_prefix>.sbwt` and `<outfile_prefix>.lcs`.
 * index::load_sbwt (index.rs:195-212): reads the pair back and uploads the index to `device`.
 * `.sbwt` starts with the reference's variant header (u64 LE 12, "SubsetMatrix": index.rs:139-140).  What follows it
 * in the reference is the serialisation of the un-vendored sbwt crate, pinned by nothing but a round trip
 * (index.rs:277-296); here the body is this library's own layout (magic "KBOB200", version, k, n_sets, n_kmers, the
 * four kbo_index_from_parts rows, FNV-1a checksum; `.lcs`: magic, version, k, n_sets, the LCS bytes, checksum; all
 * little endian).  Files round-trip bit-exactly through this pair; a file written by kbo-cli / the sbwt crate is
 * recognised by the missing magic and refused with KBO_ERR_FORMAT (never misread).  A loaded index behaves like one
 * from kbo_index_from_parts (access_kmer walks the index). */
int kbo_index_serialize(const kbo_index* ix, const char* outfile_prefix);
int kbo_index_load(const char* index_prefix, int device, kbo_index** out);
/* SbwtIndex::access_kmer (variant_calling.rs:276, gap_filling.rs:144): k bytes, '$' padded. */
int kbo_index_access_kmer(const kbo_index* ix, uint64_t colex, uint8_t* out_k);
/* SbwtIndex::search (gap_filling.rs:217): *found = 0 when the pattern does not occur. */
int kbo_index_search(const kbo_index* ix, const uint8_t* pattern, uint64_t len, int* found, uint64_t* l, uint64_t* r);

/* index::query_sbwt (index.rs:243-256) = StreamingIndex::matching_statistics: per query
 * position (d, l..r).  l_out/r_out may be NULL (lengths only, as lib.rs:624 uses it). */
int kbo_query_sbwt(const kbo_index* ix, const uint8_t* query, uint64_t len, uint64_t* d_out, uint64_t* l_out,
                   uint64_t* r_out);
/* Batched form over a CSR batch: query i = concat[offsets[i] .. offsets[i+1]); outputs indexed like concat. */
int kbo_query_sbwt_batch_compact(const kbo_index* ix, const uint8_t* concat, const uint64_t* offsets,
                                 uint64_t n_queries, uint8_t* d_out, uint32_t* l_out, uint32_t* r_out);

/* ---- derandomize: src/derandomize.rs ------------------------------------ */
int kbo_log_rm_max_cdf(uint64_t t, uint64_t alphabet_size, uint64_t n_kmers, double* out);      /* :91-100  (host, f64) */
int kbo_random_match_threshold(uint64_t k, uint64_t n_kmers, uint64_t alphabet_size, double max_error_prob,
                               uint64_t* out);                                                   /* :127-145 (host, f64) */
/* derandomize_ms_vec (:269-288) on the GPU for an arbitrary MS vector (values <= k). */
int kbo_derandomize_ms_vec(const uint64_t* noisy_ms, uint64_t n, uint64_t k, uint64_t threshold, int64_t* out,
                           int device);

/* ---- translate: src/translate.rs ---------------------------------------- */
/* translate_ms_vec (:263-293) on the GPU; one byte per alignment character ('M','-','X','R'). */
int kbo_translate_ms_vec(const int64_t* derand_ms, uint64_t n, uint64_t k, uint64_t threshold, uint8_t* chars_out,
                         int device);

/* ---- format: src/format.rs ---------------------------------------------- */
/* run_lengths_gapped (:143-193); max_gap_len == 0 gives run_lengths (:98-102).  Writes at most
 * `cap` entries, *n_out = number of RLEs found. */
int kbo_run_lengths_gapped(const uint8_t* aln, uint64_t n, uint64_t max_gap_len, kbo_rle* out, uint64_t cap,
                           uint64_t* n_out);
/* relative_to_ref (:266-287) */
int kbo_relative_to_ref(const uint8_t* ref_seq, const uint8_t* aln, uint64_t n, uint8_t* out);

/* ---- API: src/lib.rs ----------------------------------------------------- */
/* kbo::matches (lib.rs:612-628): threshold -> MS -> derandomize -> translate, fused on the device. */
int kbo_matches(const kbo_index* ix, const uint8_t* query, uint64_t len, double max_error_prob, uint8_t* chars_out);
/* Same for a CSR batch of queries in ONE launch sequence; chars_out indexed like concat. */
int kbo_matches_batch(const kbo_index* ix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                      double max_error_prob, uint8_t* chars_out);
/* Device-resident form: d_concat (total bytes), d_offsets (n_queries+1 u64) and d_chars_out are device
 * pointers on the index's device; host_offsets is the same offsets array on the host (needed for
 * precondition checks and sizing).  Asynchronous on `stream`. */
int kbo_matches_batch_device(const kbo_index* ix, const uint8_t* d_concat, const uint64_t* d_offsets,
                             const uint64_t* host_offsets, uint64_t n_queries, double max_error_prob,
                             uint8_t* d_chars_out, void* stream);
/* kbo::find (lib.rs:808-821) for a CSR batch: matches + run_lengths[_gapped].  RLEs of query i are
 * rle_out[rle_offsets[i] .. rle_offsets[i+1]) (rle_offsets has n_queries+1 entries).  `concat` and `rle_out` may be
 * pageable; page-locked buffers (kbo_alloc_pinned) are read / written in place by the copy engines. */
int kbo_find_batch(const kbo_index* ix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                   double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                   uint64_t* rle_offsets);
/* Asynchronous form of kbo_find_batch: the copy-in and all kernels of the batch are enqueued and the call returns;
 * kbo_job_wait blocks until the results are in rle_out / rle_offsets, returns the number of records in *n_rle (may be
 * NULL) and destroys the job.  One host thread can keep several batches in flight (submit, submit, wait, submit, ...).
 * All buffers must stay valid until the wait.  With page-locked `rle_out` and `rle_offsets` (kbo_alloc_pinned) the
 * device writes the results straight into them and the job needs a single synchronisation; `concat` and `offsets`
 * should be page-locked too, otherwise the copy-in is staged by the driver inside the submit call. */
typedef struct kbo_job kbo_job;
int kbo_find_batch_submit(const kbo_index* ix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                          double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                          uint64_t* rle_offsets, kbo_job** job);
int kbo_job_wait(kbo_job* job, uint64_t* n_rle);
/* Device-resident form of kbo_find_batch: d_concat, d_offsets, d_rle_out (rle_cap records) and d_rle_offsets
 * (n_queries+1 u64) are device pointers; asynchronous on `stream`.  If more than rle_cap records are found
 * only the first rle_cap are stored; d_rle_offsets[n_queries] always holds the true count. */
int kbo_find_batch_device(const kbo_index* ix, const uint8_t* d_concat, const uint64_t* d_offsets,
                          const uint64_t* host_offsets, uint64_t n_queries, double max_error_prob,
                          uint64_t max_gap_len, kbo_rle* d_rle_out, uint64_t rle_cap, uint64_t* d_rle_offsets,
                          void* stream);
/* ---- multi-GPU (SURVEY 8b "ctx_create(n_gpus) + batch calls that shard internally", 8e) ------------------------
 * One process drives several GPUs: a context owns one worker thread per device, an index set holds one replica of
 * an index per device.  The batch calls cut the CSR batch into contiguous query ranges of equal base counts (one per
 * device), run the hot path of every range on its device, and gather the results into the caller's buffers: alignment
 * characters at their final positions, RLE records in query order exactly as the single-device call returns them
 * (every device writes its records straight into the caller's buffer once the record counts of the devices before it
 * are known).  `devices` may repeat an ordinal (several workers on one GPU).  Pageable output buffers are page-locked
 * for the duration of the call. */
typedef struct kbo_ctx kbo_ctx;
typedef struct kbo_index_set kbo_index_set;
int kbo_ctx_create(int n_gpus /* <= 0: all */, const int* devices /* NULL: 0 .. n_gpus-1 */, kbo_ctx** out);
void kbo_ctx_free(kbo_ctx* ctx);
int kbo_ctx_n_gpus(const kbo_ctx* ctx);
/* kbo::build on every device of the context (index::build_sbwt_from_vecs, index.rs:56-99) */
int kbo_index_set_build(kbo_ctx* ctx, const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs,
                        const kbo_build_opts* opts, kbo_index_set** out);
void kbo_index_set_free(kbo_index_set* set);
const kbo_index* kbo_index_set_get(const kbo_index_set* set, int i);
/* kbo::matches / kbo::find over a CSR batch on all devices of the set; same arguments and results as
 * kbo_matches_batch / kbo_find_batch. */
int kbo_matches_batch_multi(const kbo_index_set* set, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                            double max_error_prob, uint8_t* chars_out);
int kbo_find_batch_multi(const kbo_index_set* set, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                         double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                         uint64_t* rle_offsets);

/* kbo::map without refinement (lib.rs:726-738,756-760 with fill_gaps = call_variants = false):
 * `format` != 0 applies relative_to_ref, else the raw translation characters are returned. */
int kbo_map_unrefined(const kbo_index* query_index, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
                      int format, uint8_t* out);

/* kbo::call (lib.rs:547-573): builds the SBWT of ref_seq with `sbwt_build_opts` (NULL = CallOpts default:
 * BuildOpts::default() with build_select), then variant_calling::call_variants (variant_calling.rs:249-294)
 * on GPU matching statistics.  Variant i = (pos[i], query_chars, ref_chars); the characters of all variants
 * are concatenated in qchars / rchars with lengths qlen[i] / rlen[i]. */
int kbo_call(const kbo_index* query_index, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
             const kbo_build_opts* sbwt_build_opts, uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars,
             uint8_t* rchars, uint64_t cap_variants, uint64_t cap_chars, uint64_t* n_variants);
/* kbo::map (lib.rs:720-761) with MapOpts {max_error_prob, fill_gaps, call_variants, format, sbwt_build_opts}. */
int kbo_map(const kbo_index* query_index, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
            int fill_gaps, int call_variants, int format, const kbo_build_opts* sbwt_build_opts, uint8_t* out);

/* kbo::call and kbo::map rebuild the SBWT of ref_seq on every call (lib.rs:553).  A caller that streams many assemblies
 * against ONE reference can build that index once -- kbo_index_build of ref_seq with the sbwt_build_opts it would have
 * passed, on the device of the query indexes -- and hand it in here: everything else, and every result, is as
 * kbo_call / kbo_map.  ref_index is only read. */
int kbo_call_with_ref(const kbo_index* query_index, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                      double max_error_prob, uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars,
                      uint8_t* rchars, uint64_t cap_variants, uint64_t cap_chars, uint64_t* n_variants);
int kbo_map_with_ref(const kbo_index* query_index, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                     double max_error_prob, int fill_gaps, int call_variants, int format, uint8_t* out);

/* ---- instrumentation ------------------------------------------------------ */
/* Event counters of the last MS launch sequence on this index when profiling counters are enabled
 * (kbo_set_profile_counters(1)): extend attempts, attempts whose two rank probes fell in different
 * 32-byte sectors, LCS contractions, extra LCS words touched, query bases incl. chunk warm-up,
 * query bases emitted.  Used for the roofline's algorithmic-bytes figure. */
typedef struct kbo_ms_counters {
    /* all work, including the k-1 warm-up bases of every chunk */
    uint64_t extend_attempts, extend_split_sector, contractions, contraction_extra_words, bases_processed,
        bases_emitted;
    /* only the events of emitted positions: the algorithmic figure the roofline uses */
    uint64_t emit_extend_attempts, emit_extend_split_sector, emit_contractions, emit_contraction_extra_words;
} kbo_ms_counters;
int kbo_set_profile_counters(int enabled);
int kbo_get_ms_counters(const kbo_index* ix, kbo_ms_counters* out);
/* Tuning knob: bases per MS chunk.  0 = automatic: from the batch size, and for the stream-ordered (_device) entry
 * points also from the number of distinct caller streams among the last 8 calls (overlapping calls get longer chunks:
 * less warm-up work per base; a lone call gets the chunk length that makes it finish soonest).  Results never depend on it. */
int kbo_set_chunk_len(uint32_t chunk_len);
/* Host threads that bridge gaps in kbo_map (gaps are independent given the incoming translation; results never depend
 * on it).  0 = hardware concurrency, at most 16.  (BuildOpts.num_threads keeps the reference's meaning: builder threads.) */
int kbo_set_refine_threads(uint32_t n);
/* kbo_map / kbo_call on an index that the GPU builder made with build_select keep the node k-mers on the device and
 * run gap_filling::fill_gaps (gap_filling.rs:444-526) and the access_kmer of call_variants (variant_calling.rs:276)
 * there; enabled == 0 sends both to the host versions (comparison runs).  Results never depend on it. */
int kbo_set_device_refine(int enabled);
/* The kbo_set_* knobs are process-wide defaults; this sets a knob for ONE index (value -1 = back to the default), so
 * that two indexes / callers in one process can differ. */
typedef enum kbo_tuning_key {
    KBO_TUNE_CHUNK_LEN = 0, KBO_TUNE_PIPELINE_PARTS = 1, KBO_TUNE_DEVICE_PARTS = 2, KBO_TUNE_MS_FLAGS = 3,
    KBO_TUNE_REFINE_THREADS = 4
} kbo_tuning_key;
int kbo_index_set_tuning(kbo_index* ix, int key, int64_t value);
/* Tuning knob: number of concurrent sub-batches inside the device-pointer batch calls (0 = automatic). */
int kbo_set_device_parts(uint32_t parts);
/* Tuning knob: number of sub-batches the host-buffer batch calls are pipelined over (0 = automatic). */
int kbo_set_pipeline_parts(uint32_t parts);
/* Index construction runs on the GPU for 2 <= k <= 64 (128-bit keys above 32); enabled != 0 forces the host builder
 * (for comparison). */
int kbo_set_host_builder(int enabled);
/* enabled == 0: indexes created afterwards carry no prefix-state table (K1 then warms every chunk up over k-1 bases;
 * comparison runs).  Results never depend on it. */
int kbo_set_prefix_table(int enabled);
/* Depth P of the prefix-state table of indexes created afterwards (the MS state after every string of P bases: a
 * chunk's warm-up starts from its entry instead of stepping through those bases; 8 bytes x 4^P).  0 = the default, 10;
 * at most 14, capped at k - 1.  Results never depend on it. */
int kbo_set_prefix_len(uint32_t len);
/* enabled == 0: indexes created afterwards carry no rank2 rows (K1 then probes one base at a time; comparison runs).
 * Results never depend on it. */
int kbo_set_rank2(int enabled);
/* enabled == 0: streams created afterwards do not mark the index arrays as persisting in L2 (comparison runs). */
int kbo_set_l2_persist(int enabled);
/* Experiment switches.  bit 1 (2): run K2 where the bit-parallel K2b would be picked; bit 4 (16): run matching
 * statistics + derandomize + translate as ONE fused kernel (fused.cuh: MS bytes only in shared memory, two bases per
 * rank probe, mismatch stretches in a second pass) instead of K1 followed by K2b; bit 2 (4): the fused kernel probes
 * one base at a time; bit 3 (8, with bit 4): the fused kernel in its one-pass form (K1's own recurrence, then
 * derandomize + translate on the shared-memory MS bytes); bit 5 (32): K1 with two bases per probe.  Results never
 * depend on them. */
int kbo_set_ms_flags(uint32_t flags);
/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
uint64_t kbo_kernel_launch_count(void);
/* Elapsed device time of the last host-pointer call's kernel section in ms (CUDA events). */
float kbo_last_kernel_ms(const kbo_index* ix);
/* Per-kernel timing of kbo_matches_batch_device calls: when enabled, CUDA events are recorded on the
 * caller's stream around K0 (pack), K1 (MS) and K2 (derandomize+translate) of every call (up to 512
 * calls between collections).  kbo_collect_kernel_times must be called after the stream has been
 * synchronised; it returns the SUMMED elapsed ms per kernel and the number of calls, and resets. */
int kbo_set_kernel_timing(int enabled);
int kbo_collect_kernel_times(const kbo_index* ix, void* stream, double sum_ms_out[3], uint64_t* n_calls);
/* Roofline denominators that MEASURED_PEAKS.json does not hold: throughput of 8-byte loads that each
 * touch a random 32-byte sector of a `buffer_bytes` buffer (small buffer = L2-resident, large = HBM),
 * issued like K1 issues them (one dependent chain per lane, every resident lane busy) when
 * `dependent` != 0, or as independent loads when 0.  Result: sectors per second. */
int kbo_measure_random_sector_rate(int device, uint64_t buffer_bytes, int dependent, double* sectors_per_s);

#ifdef __cplusplus
}
#endif
#endif /* KBO_B200_H */
