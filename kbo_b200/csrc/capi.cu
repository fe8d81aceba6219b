// ===========================================================================
// kbo_b200/csrc/capi.cu -- implementation of include/kbo_b200.h
//
// Host orchestration of the hot path: stage a CSR batch of queries, run
// K0 (pack) -> K1 (matching statistics) -> K2 (derandomize + translate) on one
// stream, copy the result back.  No CPU fallback exists: every compute entry
// point needs a CUDA device and fails with KBO_ERR_CUDA otherwise.
// ===========================================================================
#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kbo_b200.h"
#include "kernels.cuh"
#include "fused.cuh"
#include "index_build.cuh"
#include "refine.cuh"
#include "host_layout.hpp"
#include "refine_host.hpp"
#include "sbwt_host.hpp"

using namespace kbo_b200;

// ---------------------------------------------------------------------------
// errors, globals
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches(0);
static std::atomic<int> g_profile_counters(0);
static std::atomic<uint32_t> g_chunk_len(0);
static std::atomic<int> g_kernel_timing(0);
static std::atomic<int> g_rank2(1);  // 0: indexes built afterwards carry no two-bases-per-probe rows (comparison runs)
static std::atomic<int> g_prefix_table(1);  // 0: indexes built afterwards get no prefix-state table (comparison runs)
static std::atomic<int> g_prefix_len(0);    // depth of the prefix-state table of indexes built afterwards (0: PREF_LEN)
static std::atomic<int> g_l2_persist(1);  // 0: do not mark the index persisting in L2 (comparison runs)
static std::atomic<uint32_t> g_ms_flags(0);
static std::atomic<int> g_host_builder(0);
static std::atomic<uint32_t> g_parts(0);
static std::atomic<uint32_t> g_dev_parts(0);
static std::atomic<uint32_t> g_ms_block(128);

static std::atomic<int> g_device_refine(1);  // 0: fill_gaps / access_kmer of kbo::map and kbo::call on the host (comparison runs)
static std::atomic<uint32_t> g_refine_threads(0);  // host threads of fill_gaps (0 = hardware concurrency, at most 16)
struct kbo_index;
static uint32_t tuned_chunk_len(const kbo_index* ix);
static uint32_t tuned_ms_flags(const kbo_index* ix);

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? KBO_ERR_OOM : KBO_ERR_CUDA,                         \
                        std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);                 \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// ---------------------------------------------------------------------------
// device buffers / workspaces
// ---------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, cudaStream_t st) {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {  // page-locked host staging: copies to/from it are truly asynchronous
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Workspace {
    std::mutex mu;  // held for the whole enqueue section of the stream-ordered (_device) entry points: a workspace bound
                    // to a caller stream (NULL included) is shared by every host thread that passes that stream
    PinnedBuf h_rel, h_roff, h_rle;
    std::vector<Workspace*> subs;          // per-part workspaces with their own streams (device-pointer calls)
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_join;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    const void* l2_blob = nullptr;  // index allocation the stream's persisting-L2 window points at
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf ascii, offsets, pack, inv, sep, wq, ms, l, r, out, out2, out3, tmp64, counters, counters2;
    PinnedBuf h_count, h_d, h_l, h_r, h_chars;  // (h_d .. h_chars: kbo_map's (d, l, r) and characters on the host)
    DevBuf masks, rle_words, rle_cnt, rle_cse, rle_tickets;  // K2b<false> masks and the K4 arrays
    DevBuf gaps, arena, terms;  // fill_gaps on the device (refine.cuh)
    std::vector<cudaEvent_t> timing;  // 4 events per timed call (before K0, after K0, after K1, after K2)
    size_t timed_calls = 0;
    void destroy() {
        for (cudaEvent_t e : timing) cudaEventDestroy(e);
        h_rel.release(); h_roff.release(); h_rle.release(); h_count.release();
        h_d.release(); h_l.release(); h_r.release(); h_chars.release();
        for (Workspace* w : subs) { w->destroy(); delete w; }
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (cudaEvent_t e : ev_join) cudaEventDestroy(e);
        DevBuf* all[] = {&ascii, &offsets, &pack, &inv, &sep, &wq, &ms, &l, &r, &out, &out2, &out3,
                         &tmp64, &counters, &counters2, &masks, &rle_words, &rle_cnt, &rle_cse, &rle_tickets, &gaps, &arena,
                         &terms};
        for (DevBuf* b : all) b->release();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

// Tuning knobs of ONE index (kbo_index_set_tuning); -1 = follow the process-wide default (kbo_set_*).
struct IndexTuning {
    std::atomic<int64_t> chunk_len{-1}, pipeline_parts{-1}, device_parts{-1}, ms_flags{-1}, refine_threads{-1};
};

struct kbo_index {
    int device = 0;
    IndexTuning tune;
    std::vector<PinnedBuf> pinned_pool;  // host staging for find (guarded by mu)
    std::atomic<int> host_calls{0};      // host-buffer batch calls currently inside the library (any thread)
    std::atomic<bool> seen_concurrency{false};  // some host-buffer call found another caller inside (sticky)
    cudaStream_t recent_streams[8] = {};       // caller streams of the last 8 stream-ordered calls (guarded by mu)
    unsigned recent_pos = 0;
    HostIndex host;
    uint64_t* d_rank = nullptr;
    uint64_t* d_rank2 = nullptr;  // 16 rows for two bases per probe (inside the blob; null for device-only helper indexes)
    uint8_t* d_lcs = nullptr;
    uint32_t* d_links = nullptr;  // per node: LCS and the distances to the nearest smaller LCS on both sides
    uint8_t* d_blob = nullptr;    // the one allocation holding rank | links | lcs (one L2 access-policy window)
    uint64_t* d_pref = nullptr;   // MS states after view.pref_len bases (k >= PREF_MIN_K)
    // "select support" (BuildOpts.build_select) on the device: the colex-sorted node keys of the GPU builder
    // (refine.cuh NodeKeysView); null for indexes made from parts or by the host builder
    uint64_t* d_node_keys = nullptr;
    uint8_t* d_node_len = nullptr;
    uint32_t node_key_words = 0;
    // the plain SubsetMatrix form in `host` (rows, LCS, nodes) is read back from the device on first use
    // (ensure_host_mirror): kbo::map / kbo::call on a freshly built index never need it
    std::atomic<bool> host_ready{true};
    std::mutex host_mu;
    std::atomic<bool> rank2_ready{false};  // the rank2 rows are computed when a kernel that probes pairs first runs
    // arrays of an index made by the GPU builder come from the device's stream-ordered pool (kept warm): kbo::call /
    // kbo::map build and free an index per assembly, and cudaMalloc / cudaFree cost them 3-5 ms each time
    bool pool_alloc = false;
    uint64_t blob_bytes = 0;
    float l2_hit_ratio = 0.f;     // 0: no persisting-L2 window available
    uint64_t rank_stride = 0;
    uint64_t device_bytes = 0;
    IndexView view;
    std::mutex mu;
    std::unordered_map<cudaStream_t, Workspace*> by_stream;  // workspaces bound to caller streams
    kbo_ms_counters last_counters = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float last_kernel_ms = 0.f;
};

static uint32_t tuned_chunk_len(const kbo_index* ix) {
    const int64_t v = ix ? ix->tune.chunk_len.load() : -1;
    return v >= 0 ? (uint32_t)v : g_chunk_len.load();
}
static uint32_t tuned_ms_flags(const kbo_index* ix) {
    const int64_t v = ix ? ix->tune.ms_flags.load() : -1;
    return v >= 0 ? (uint32_t)v : g_ms_flags.load();
}
static uint32_t tuned_parts(const kbo_index* ix, bool device) {
    const int64_t v = ix ? (device ? ix->tune.device_parts.load() : ix->tune.pipeline_parts.load()) : -1;
    return v >= 0 ? (uint32_t)v : (device ? g_dev_parts.load() : g_parts.load());
}
static uint32_t tuned_refine_threads(const kbo_index* ix) {
    int64_t v = ix ? ix->tune.refine_threads.load() : -1;
    if (v < 0) v = g_refine_threads.load();
    if (v == 0) v = std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    return (uint32_t)v;
}

// The index arrays live in ONE allocation so that a single access-policy window covers them: every stream that
// runs K1 marks that range "persisting" in L2 (the streaming batch buffers of K0/K2/K4 would otherwise keep
// evicting index lines, and a warp of K1 waits for the slowest of its ~60 random loads per iteration).
// cudaGetDeviceProperties costs milliseconds per call; kbo::call / kbo::map create two indexes per assembly
struct L2Props {
    bool ok = false;
    size_t persisting_max = 0, window_max = 0;
};
static const L2Props& l2_props(int dev) {
    static std::mutex mu;
    static std::unordered_map<int, L2Props> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    L2Props p;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
        p.ok = true;
        p.persisting_max = (size_t)prop.persistingL2CacheMaxSize;
        p.window_max = (size_t)prop.accessPolicyMaxWindowSize;
    }
    cudaGetLastError();
    return cache.emplace(dev, p).first->second;
}

static int alloc_index_arrays(kbo_index* ix, uint64_t rank_words, uint64_t lcs_bytes, uint64_t n, bool with_rank2) {
    const uint64_t rank_bytes = rank_words * 8;
    const uint64_t rank2_bytes = with_rank2 && g_rank2.load() ? 4 * rank_bytes : 0;  // 16 rows instead of 4
    const uint64_t links_bytes = ((n + 1) * 4 + 255) & ~255ull;
    ix->blob_bytes = rank_bytes + rank2_bytes + links_bytes + lcs_bytes;
    if (ix->pool_alloc) CUDA_TRY(cudaMallocAsync((void**)&ix->d_blob, ix->blob_bytes, 0));
    else CUDA_TRY(cudaMalloc((void**)&ix->d_blob, ix->blob_bytes));
    ix->d_rank = reinterpret_cast<uint64_t*>(ix->d_blob);
    ix->d_rank2 = rank2_bytes ? reinterpret_cast<uint64_t*>(ix->d_blob + rank_bytes) : nullptr;
    ix->d_links = reinterpret_cast<uint32_t*>(ix->d_blob + rank_bytes + rank2_bytes);
    ix->d_lcs = ix->d_blob + rank_bytes + rank2_bytes + links_bytes;
    ix->device_bytes = ix->blob_bytes;
    int dev = 0;
    const L2Props* lp = cudaGetDevice(&dev) == cudaSuccess ? &l2_props(dev) : nullptr;
    if (lp && lp->ok && lp->persisting_max > 0 && lp->window_max > 0) {
        size_t want = std::min<size_t>((size_t)ix->blob_bytes, lp->persisting_max);
        size_t have = 0;
        cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
        if (have < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) want = have;
        else if (have > want) want = std::min<size_t>(have, (size_t)ix->blob_bytes);
        const size_t window = std::min<size_t>((size_t)ix->blob_bytes, lp->window_max);
        ix->l2_hit_ratio = window ? (float)std::min(1.0, (double)want / (double)window) : 0.f;
    }
    cudaGetLastError();
    return KBO_OK;
}
static void apply_l2_window(const kbo_index* ix, cudaStream_t st) {
    if (!ix->d_blob || ix->l2_hit_ratio <= 0.f || g_l2_persist.load() == 0) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    const L2Props& prop = l2_props(dev);
    if (!prop.ok) return;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = ix->d_blob;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)ix->blob_bytes, prop.window_max);
    attr.accessPolicyWindow.hitRatio = ix->l2_hit_ratio;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();  // best effort: the window is an optimisation
}

// Idle workspaces (own stream + scratch buffers, nothing index-specific) are pooled per DEVICE, not per index: kbo::call /
// kbo::map build a fresh index for every assembly, and with per-index pools every one of them paid for its ~20 device
// allocations and page-locked staging buffers again.  The persisting-L2 window of the stream is re-pointed when a
// workspace moves to another index.
struct DevicePool {
    std::mutex mu;
    std::vector<Workspace*> idle;
};
static DevicePool* device_pool(int device) {
    static std::mutex mu;
    static std::unordered_map<int, DevicePool*> pools;
    std::lock_guard<std::mutex> g(mu);
    DevicePool*& p = pools[device];
    if (!p) p = new DevicePool();
    return p;
}

static int acquire_ws(kbo_index* ix, Workspace** out) {
    DevicePool* pool = device_pool(ix->device);
    Workspace* ws = nullptr;
    {
        std::lock_guard<std::mutex> g(pool->mu);
        if (!pool->idle.empty()) {
            ws = pool->idle.back();
            pool->idle.pop_back();
        }
    }
    if (!ws) {
        ws = new Workspace();
        ws->own_stream = true;
        CUDA_TRY(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&ws->ev0));
        CUDA_TRY(cudaEventCreate(&ws->ev1));
    }
    if (ws->l2_blob != ix->d_blob) {
        apply_l2_window(ix, ws->stream);
        ws->l2_blob = ix->d_blob;
    }
    *out = ws;
    return KBO_OK;
}
static void release_ws(kbo_index* ix, Workspace* ws) {
    DevicePool* pool = device_pool(ix->device);
    std::lock_guard<std::mutex> g(pool->mu);
    pool->idle.push_back(ws);
}
// How many launches of K1 a stream-ordered call can expect to share the machine with: K1 is about 60 % of a step, so
// of the distinct caller streams among the last 8 calls roughly that share is inside K1 at any time.
static uint32_t expected_overlap(kbo_index* ix) {
    std::lock_guard<std::mutex> g(ix->mu);
    uint32_t distinct = 0;
    for (int i = 0; i < 8; ++i) {
        bool seen = false;
        for (int j = 0; j < i; ++j) seen |= ix->recent_streams[j] == ix->recent_streams[i];
        distinct += !seen;  // (the initial null entries count as one stream)
    }
    return std::max<uint32_t>(1, distinct * 6 / 10);
}

static int stream_ws(kbo_index* ix, cudaStream_t st, Workspace** out) {
    std::lock_guard<std::mutex> g(ix->mu);
    ix->recent_streams[ix->recent_pos++ & 7u] = st;
    auto it = ix->by_stream.find(st);
    if (it != ix->by_stream.end()) { *out = it->second; return KBO_OK; }
    Workspace* ws = new Workspace();
    ws->stream = st;
    ws->own_stream = false;  // caller-owned stream: its attributes (access-policy window) are left alone
    ix->by_stream[st] = ws;
    *out = ws;
    return KBO_OK;
}

// Does the calling thread already have a CUDA context bound?  (Driver entry point looked up at run time so that the
// library has no link-time dependency on libcuda and still loads on a box without a driver.)
static bool thread_has_context() {
    typedef int (*ctx_get_current_t)(void**);
    static const ctx_get_current_t fn = []() -> ctx_get_current_t {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<ctx_get_current_t>(p);
    }();
    if (!fn) return true;  // unknown: behave like a thread that has one
    void* ctx = nullptr;
    return fn(&ctx) == 0 && ctx != nullptr;
}

// Makes the index's device current for the duration of a call.  The previous device is restored only when the
// thread really had a context on it: a fresh host thread reports device 0 as "current" without having touched it,
// and cudaSetDevice(0) on the way out would create a primary context on GPU 0 from every worker thread of every
// rank (round-1 finding: GPU 0 busy 21 % at 8 ranks while the other GPUs stayed at 1 %).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    bool restore = false;
    explicit DeviceGuard(int dev) {
        const bool had_ctx = thread_has_context();
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev) {
            if (cudaSetDevice(dev) != cudaSuccess) ok = false;
            else restore = had_ctx;
        }
    }
    ~DeviceGuard() { if (restore) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------
// host-only pieces of the path (derandomize.rs:91-145: f64, once per index)
// ---------------------------------------------------------------------------
static double host_log_rm_max_cdf(uint64_t t, uint64_t alphabet_size, uint64_t n_kmers) {
    // n_kmers * ln_1p(-(exp(ln 1 - ln s))^(t+1))      derandomize.rs:99
    const double q = std::exp(std::log(1.0) - std::log((double)alphabet_size));
    return (double)n_kmers * std::log1p(-__builtin_powi(q, (int)t + 1));
}

static int host_threshold(uint64_t k, uint64_t n_kmers, uint64_t alphabet, double p, uint64_t* out) {
    if (k == 0) return fail(KBO_ERR_BAD_K, "k must be > 0 (derandomize.rs:133)");
    if (n_kmers == 0) return fail(KBO_ERR_BAD_ARGUMENT, "n_kmers must be > 0 (derandomize.rs:134)");
    if (alphabet == 0) return fail(KBO_ERR_BAD_ARGUMENT, "alphabet_size must be > 0 (derandomize.rs:135)");
    if (!(p <= 1.0) || !(p > 0.0)) return fail(KBO_ERR_BAD_PROB, "0 < max_error_prob <= 1 (derandomize.rs:136-137)");
    const double bound = std::log1p(-p);
    for (uint64_t i = 1; i < k; ++i) {
        if (host_log_rm_max_cdf(i, alphabet, n_kmers) > bound) { *out = i; return KBO_OK; }
    }
    *out = k;
    return KBO_OK;
}

// ---------------------------------------------------------------------------
// index upload: SubsetMatrix rows + LCS -> interleaved rank words + padded LCS
// ---------------------------------------------------------------------------
// links array (kernels.cuh IndexView::links) from the device LCS bytes; called by both builders
// rank2 rows from the rank words that are already on the device (both builders and kbo_index_from_parts end here)
static int build_rank2(kbo_index* ix) {
    ix->view.rank2 = nullptr;
    if (!ix->d_rank2) return KBO_OK;
    const uint64_t stride = ix->rank_stride, words = 16 * stride;
    uint32_t *rows2 = nullptr, *pc = nullptr, *prefix = nullptr;
    void* scan_tmp = nullptr;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&rows2, words * 4));
        CUDA_TRY(cudaMalloc((void**)&pc, words * 4));
        CUDA_TRY(cudaMalloc((void**)&prefix, words * 4));
        rank2_bits_kernel<<<(unsigned)((stride + 127) / 128), 128>>>(ix->view, rows2);
        LAUNCHED();
        row_popc_kernel<<<(unsigned)((words + 255) / 256), 256>>>(rows2, words, pc);
        LAUNCHED();
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, pc, prefix, (int64_t)words);
        CUDA_TRY(cudaMalloc(&scan_tmp, need ? need : 1));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, need, pc, prefix, (int64_t)words));
        LAUNCHED();
        compose_rank2_kernel<<<(unsigned)((words + 255) / 256), 256>>>(ix->view, rows2, prefix, ix->d_rank2);
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        return KBO_OK;
    };
    const int rc = body();
    cudaFree(rows2); cudaFree(pc); cudaFree(prefix); cudaFree(scan_tmp);
    if (rc == KBO_OK) ix->view.rank2 = ix->d_rank2;
    return rc;
}

// The states after 1, 2, ... P bases, level by level (prefix_table_level_kernel); level P stays in *out.
static int build_pref_table(const IndexView& view, uint32_t P, uint64_t** out, bool pool) {
    const size_t last = (size_t)1 << (2 * P);
    uint64_t *a = nullptr, *b = nullptr;  // level j lives in b when P - j is even (so level P does), else in a
    auto alloc = [&](uint64_t** p, size_t bytes) { return pool ? cudaMallocAsync((void**)p, bytes, 0) : cudaMalloc((void**)p, bytes); };
    auto release = [&](uint64_t* p) { if (pool) cudaFreeAsync(p, 0); else cudaFree(p); };
    CUDA_TRY(alloc(&a, std::max<size_t>(last / 4, 4) * 8));
    if (alloc(&b, last * 8) != cudaSuccess) {
        release(a);
        cudaGetLastError();
        return fail(KBO_ERR_OOM, "prefix-state table: out of device memory");
    }
    for (uint32_t j = 1; j <= P; ++j) {
        const uint32_t cnt = 1u << (2 * j);
        uint64_t* cur = ((P - j) & 1u) ? a : b;
        const uint64_t* prev = ((P - j) & 1u) ? b : a;
        prefix_table_level_kernel<<<(cnt + 255) / 256, 256>>>(view, prev, cur, j);
        LAUNCHED();
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    release(a);
    if (e != cudaSuccess) {
        release(b);
        return fail(KBO_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in the prefix-state table");
    }
    *out = b;
    return KBO_OK;
}

// The rows for two bases per probe are only read by the fused kernel and K1p (kbo_set_ms_flags bits 4 / 5): their
// space is part of the index allocation, their contents are made on first use.
static int ensure_rank2(kbo_index* ix) {
    if (!ix->d_rank2 || ix->rank2_ready.load(std::memory_order_acquire)) return KBO_OK;
    std::lock_guard<std::mutex> g(ix->host_mu);
    if (ix->rank2_ready.load(std::memory_order_acquire)) return KBO_OK;
    const int rc = build_rank2(ix);
    if (rc == KBO_OK) ix->rank2_ready.store(true, std::memory_order_release);
    return rc;
}

static int build_links(kbo_index* ix, uint64_t n, bool with_prefix_table = true) {
    ix->view.rank2 = nullptr;
    lcs_links_kernel<<<(unsigned)((n + 1 + 255) / 256), 256>>>(ix->d_lcs, (uint32_t)n, ix->d_links);
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    ix->view.links = ix->d_links;
    ix->view.pref = nullptr;
    ix->view.pref_len = 0;
    if (with_prefix_table && ix->view.k >= PREF_MIN_K && g_prefix_table.load()) {
        const int want = g_prefix_len.load();
        const uint32_t P = std::min<uint32_t>(want > 0 ? (uint32_t)want : (uint32_t)PREF_LEN,
                                              std::min<uint32_t>(PREF_MAX_LEN, ix->view.k - 1));
        const int rc = build_pref_table(ix->view, P, &ix->d_pref, ix->pool_alloc);
        if (rc) return rc;
        ix->view.pref = ix->d_pref;
        ix->view.pref_len = P;
        ix->device_bytes += ((uint64_t)8 << (2 * P));
    } else {
        // the construction ran on the legacy default stream; queries run on non-blocking streams of their own
        CUDA_TRY(cudaStreamSynchronize(0));
    }
    return KBO_OK;
}

// Reads the plain SubsetMatrix form (4 bit rows, LCS bytes, C, and the stored nodes) back from the device arrays of an
// index that the GPU builder made.  Host-side lookups (kbo_index_search / access_kmer / export_parts, the host
// versions of fill_gaps and call_variants) call this first; the device paths never do.
static int ensure_host_mirror(const kbo_index* cix) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (ix->host_ready.load(std::memory_order_acquire)) return KBO_OK;
    std::lock_guard<std::mutex> g(ix->host_mu);
    if (ix->host_ready.load(std::memory_order_acquire)) return KBO_OK;
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    HostIndex& h = ix->host;
    const uint64_t n = h.n_sets, stride = ix->rank_stride;
    std::vector<uint64_t> rank((size_t)(4 * stride));
    CUDA_TRY(cudaMemcpy(rank.data(), ix->d_rank, 4 * stride * 8, cudaMemcpyDeviceToHost));
    h.lcs.resize((size_t)n);
    CUDA_TRY(cudaMemcpy(h.lcs.data(), ix->d_lcs, n, cudaMemcpyDeviceToHost));
    const size_t nw64 = (size_t)(n + 63) / 64 + 1;
    for (int c = 0; c < 4; ++c) {
        h.rows[c].assign(nw64, 0);
        for (size_t w = 0; w < nw64; ++w) {  // the low half of a rank word holds the 32 row bits
            const uint64_t lo = 2 * w < stride ? (uint32_t)rank[(size_t)(c * stride + 2 * w)] : 0;
            const uint64_t hi = 2 * w + 1 < stride ? (uint32_t)rank[(size_t)(c * stride + 2 * w + 1)] : 0;
            h.rows[c][w] = lo | (hi << 32);
        }
    }
    h.finalize();
    if (ix->d_node_keys) {
        h.node_hi.resize((size_t)n);
        h.node_len.resize((size_t)n);
        if (ix->node_key_words == 1) {
            CUDA_TRY(cudaMemcpy(h.node_hi.data(), ix->d_node_keys, n * 8, cudaMemcpyDeviceToHost));
        } else {  // 128-bit keys: little-endian (lo, hi) pairs on the device
            std::vector<uint64_t> keys((size_t)(2 * n));
            CUDA_TRY(cudaMemcpy(keys.data(), ix->d_node_keys, n * 16, cudaMemcpyDeviceToHost));
            h.node_lo.resize((size_t)n);
            for (size_t i = 0; i < (size_t)n; ++i) {
                h.node_lo[i] = keys[2 * i];
                h.node_hi[i] = keys[2 * i + 1];
            }
        }
        CUDA_TRY(cudaMemcpy(h.node_len.data(), ix->d_node_len, n, cudaMemcpyDeviceToHost));
    }
    ix->host_ready.store(true, std::memory_order_release);
    return KBO_OK;
}

static int upload_index(kbo_index* ix) {
    const HostIndex& h = ix->host;
    const uint64_t n = h.n_sets;
    if (n >= (1ull << 32) - 64) return fail(KBO_ERR_INDEX_TOO_LARGE, "n_sets must be < 2^32");
    if (h.k == 0 || h.k > 127) return fail(KBO_ERR_BAD_K, "device LCS compare needs 1 <= k <= 127");
    DeviceLayout lay;
    build_device_layout(h, &lay);
    const std::vector<uint64_t>& rank = lay.rank;
    const std::vector<uint8_t>& lcs = lay.lcs;
    const uint64_t stride = lay.stride;
    { int rc = alloc_index_arrays(ix, rank.size(), lcs.size(), n, true); if (rc) return rc; }
    CUDA_TRY(cudaMemcpy(ix->d_rank, rank.data(), rank.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ix->d_lcs, lcs.data(), lcs.size(), cudaMemcpyHostToDevice));
    ix->rank_stride = stride;
    ix->view.rank = ix->d_rank;
    ix->view.rank_stride = (uint32_t)stride;
    ix->view.lcs = ix->d_lcs;
    ix->view.n = (uint32_t)n;
    ix->view.k = h.k;
    return build_links(ix, n);
}

// ---------------------------------------------------------------------------
// index construction on the device (index_build.cuh), 2 <= k <= 64
// ---------------------------------------------------------------------------
// Scratch of one index construction: stream-ordered allocations from the device's default memory pool
// (kept warm between builds), so that the ~25 temporaries cost microseconds instead of a cudaMalloc /
// cudaFree (device synchronisation) each.
struct TmpBufs {
    std::vector<void*> ptrs;
    ~TmpBufs() { for (void* p : ptrs) cudaFreeAsync(p, 0); }
    static void warm_pool(int device) {
        static std::mutex mu;
        static std::vector<int> done;
        std::lock_guard<std::mutex> g(mu);
        if (std::find(done.begin(), done.end(), device) != done.end()) return;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        done.push_back(device);
    }
    template <typename T> cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, (count ? count : 1) * sizeof(T), 0);
        if (e == cudaSuccess) { ptrs.push_back(p); *out = (T*)p; }
        return e;
    }
};

// device_only: the index will only serve K1 on short queries (the per-call index of `ref_seq` in kbo::call): no
// host copy of the SubsetMatrix form, no prefix-state table.
// KBO_BUILD_TIMING=1 in the environment prints where an index construction spends its time (stderr)
struct BuildTimer {
    const char* mode = std::getenv("KBO_BUILD_TIMING");  // "1": synchronise the device at every lap; "2": host clock only
    bool on = mode != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        if (mode[0] != '2') cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[kbo build] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// (KBO_BUILD_TIMING: wall time of a whole entry point, printed when it returns)
struct ScopeTimer {
    const char* what;
    bool on = std::getenv("KBO_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit ScopeTimer(const char* w) : what(w) {}
    ~ScopeTimer() {
        if (on) std::fprintf(stderr, "[kbo total] %-28s %8.2f ms\n", what,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

template <typename K>
static int build_index_gpu_typed(kbo_index* ix, const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k,
                                 bool revcomp, bool keep_nodes, bool device_only) {
    constexpr int BITS = KmerKey<K>::BITS;
    BuildTimer bt;
    uint64_t total = 0;
    std::vector<uint64_t> offsets(n_seqs + 1, 0);
    for (uint64_t i = 0; i < n_seqs; ++i) { total += lens[i]; offsets[i + 1] = total; }
    const Geometry g = kbo_b200::make_geometry(total, n_seqs, 64);
    TmpBufs::warm_pool(ix->device);
    TmpBufs tmp;
    uint8_t *d_ascii, *d_flags, *d_nopred, *d_Dlen = nullptr, *d_Plen;
    uint64_t *d_off, *d_pack, *d_count;
    K *d_keys, *d_keys_rc = nullptr, *d_sel, *d_sorted, *d_R, *d_src, *d_Dkey = nullptr, *d_Pkey;
    uint32_t *d_inv, *d_sep, *d_wq, *d_rows32, *d_pc, *d_prefix;
    CUDA_TRY(tmp.alloc(&d_ascii, total));
    CUDA_TRY(tmp.alloc(&d_off, n_seqs + 1));
    CUDA_TRY(tmp.alloc(&d_pack, g.n_words));
    CUDA_TRY(tmp.alloc(&d_inv, g.n_words));
    CUDA_TRY(tmp.alloc(&d_sep, g.n_words));
    CUDA_TRY(tmp.alloc(&d_wq, g.n_words));
    CUDA_TRY(tmp.alloc(&d_count, 2));
    {   // sequences may live anywhere on the host: stage them contiguously
        uint64_t at = 0;
        for (uint64_t i = 0; i < n_seqs; ++i) {
            if (lens[i]) CUDA_TRY(cudaMemcpy(d_ascii + at, seqs[i], lens[i], cudaMemcpyHostToDevice));
            at += lens[i];
        }
    }
    CUDA_TRY(cudaMemcpy(d_off, offsets.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice));
    QueryView qv;
    qv.pack = d_pack; qv.inv = d_inv; qv.sep = d_sep; qv.wq = d_wq; qv.Lp = g.Lp; qv.n_words = g.n_words;
    pack_queries_kernel<<<(unsigned)((g.n_words + 127) / 128), 128>>>(d_ascii, d_off, n_seqs, qv, d_pack, d_inv, d_sep, d_wq);
    LAUNCHED();
    const uint64_t Lp = g.Lp;
    const uint64_t n_cand = revcomp ? 2 * Lp : Lp;
    CUDA_TRY(tmp.alloc(&d_keys, Lp));
    if (revcomp) CUDA_TRY(tmp.alloc(&d_keys_rc, Lp));
    CUDA_TRY(tmp.alloc(&d_flags, Lp));
    CUDA_TRY(tmp.alloc(&d_sel, n_cand));
    CUDA_TRY(tmp.alloc(&d_sorted, n_cand));
    CUDA_TRY(tmp.alloc(&d_R, n_cand));
    kmer_keys_kernel<K><<<(unsigned)((Lp + 255) / 256), 256>>>(d_pack, d_inv, Lp, k, revcomp ? 1 : 0, d_keys, d_keys_rc, d_flags);
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    bt.lap("copy-in, pack, k-mer keys");
    // compaction of the valid k-mers (forward, then reverse complements)
    size_t tb = 0, need = 0;
    void* d_tmp = nullptr;
    cub::DeviceSelect::Flagged(nullptr, need, d_keys, d_flags, d_sel, d_count, (int64_t)Lp); tb = std::max(tb, need);
    cub::DeviceRadixSort::SortKeys(nullptr, need, d_sel, d_sorted, (int64_t)n_cand, BITS - 2 * (int)k, BITS); tb = std::max(tb, need);
    cub::DeviceSelect::Unique(nullptr, need, d_sorted, d_R, d_count, (int64_t)n_cand); tb = std::max(tb, need);
    CUDA_TRY(cudaMallocAsync(&d_tmp, tb ? tb : 1, 0));
    tmp.ptrs.push_back(d_tmp);
    uint64_t h_count = 0, n_valid = 0;
    need = tb;
    CUDA_TRY(cub::DeviceSelect::Flagged(d_tmp, need, d_keys, d_flags, d_sel, d_count, (int64_t)Lp));
    CUDA_TRY(cudaMemcpy(&h_count, d_count, 8, cudaMemcpyDeviceToHost));
    n_valid = h_count;
    if (revcomp) {
        need = tb;
        CUDA_TRY(cub::DeviceSelect::Flagged(d_tmp, need, d_keys_rc, d_flags, d_sel + n_valid, d_count, (int64_t)Lp));
        CUDA_TRY(cudaMemcpy(&h_count, d_count, 8, cudaMemcpyDeviceToHost));
        n_valid += h_count;
    }
    LAUNCHED(); LAUNCHED();
    uint64_t nR = 0;
    if (n_valid) {
        need = tb;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, need, d_sel, d_sorted, (int64_t)n_valid, BITS - 2 * (int)k, BITS));
        need = tb;
        CUDA_TRY(cub::DeviceSelect::Unique(d_tmp, need, d_sorted, d_R, d_count, (int64_t)n_valid));
        CUDA_TRY(cudaMemcpy(&nR, d_count, 8, cudaMemcpyDeviceToHost));
    }
    bt.lap("select, sort, unique");
    // dummy nodes: few, made on the host from the k-mers that have no predecessor
    std::vector<K> h_src;
    if (nR) {
        CUDA_TRY(tmp.alloc(&d_nopred, nR));
        CUDA_TRY(tmp.alloc(&d_src, nR));
        no_predecessor_kernel<K><<<(unsigned)((nR + 255) / 256), 256>>>(d_R, nR, k, d_nopred);
        LAUNCHED();
        need = tb;
        CUDA_TRY(cub::DeviceSelect::Flagged(d_tmp, need, d_R, d_nopred, d_src, d_count, (int64_t)nR));
        uint64_t n_src = 0;
        CUDA_TRY(cudaMemcpy(&n_src, d_count, 8, cudaMemcpyDeviceToHost));
        h_src.resize(n_src);
        if (n_src) CUDA_TRY(cudaMemcpy(h_src.data(), d_src, n_src * sizeof(K), cudaMemcpyDeviceToHost));
    }
    std::vector<std::pair<K, uint8_t>> dummies;
    dummies.push_back({(K)0, (uint8_t)0});
    for (K x : h_src)
        for (uint32_t j = 1; j < k; ++j) dummies.push_back({(K)(x << (2 * (k - j))), (uint8_t)j});
    std::sort(dummies.begin(), dummies.end());
    dummies.erase(std::unique(dummies.begin(), dummies.end()), dummies.end());
    const uint64_t nD = dummies.size();
    const uint64_t n = nR + nD;
    if (n >= (1ull << 32) - 64) return fail(KBO_ERR_INDEX_TOO_LARGE, "n_sets must be < 2^32");
    std::vector<K> h_Dkey(nD);
    std::vector<uint8_t> h_Dlen(nD);
    for (uint64_t i = 0; i < nD; ++i) { h_Dkey[i] = dummies[i].first; h_Dlen[i] = dummies[i].second; }
    CUDA_TRY(tmp.alloc(&d_Dkey, nD));
    CUDA_TRY(tmp.alloc(&d_Dlen, nD));
    CUDA_TRY(cudaMemcpy(d_Dkey, h_Dkey.data(), nD * sizeof(K), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_Dlen, h_Dlen.data(), nD, cudaMemcpyHostToDevice));
    ix->pool_alloc = true;
    if (keep_nodes && !device_only) {  // "select support": the sorted nodes stay on the device with the index
        CUDA_TRY(cudaMallocAsync((void**)&ix->d_node_keys, (n ? n : 1) * sizeof(K), 0));
        CUDA_TRY(cudaMallocAsync((void**)&ix->d_node_len, n ? n : 1, 0));
        ix->node_key_words = sizeof(K) / 8;
        d_Pkey = reinterpret_cast<K*>(ix->d_node_keys);
        d_Plen = ix->d_node_len;
    } else {
        CUDA_TRY(tmp.alloc(&d_Pkey, n));
        CUDA_TRY(tmp.alloc(&d_Plen, n));
    }
    merge_nodes_kernel<K><<<(unsigned)((n + 255) / 256), 256>>>(d_R, nR, d_Dkey, d_Dlen, nD, k, d_Pkey, d_Plen);
    LAUNCHED();
    bt.lap("dummies, merge");
    // final device arrays
    const uint64_t nblk = (n >> 5) + 2;
    const uint64_t stride = (nblk + 3) & ~3ull;
    const uint64_t lcs_bytes = ((n + 8) & ~7ull) + 16;
    { int rc = alloc_index_arrays(ix, 4 * stride, lcs_bytes, n, !device_only); if (rc) return rc; }
    bt.lap("index allocation");
    CUDA_TRY(cudaMemset(ix->d_lcs, 0, lcs_bytes));
    CUDA_TRY(tmp.alloc(&d_rows32, 4 * stride));
    CUDA_TRY(tmp.alloc(&d_pc, 4 * stride));
    CUDA_TRY(tmp.alloc(&d_prefix, 4 * stride));
    CUDA_TRY(cudaMemset(d_rows32, 0, 4 * stride * 4));
    lcs_kernel<K><<<(unsigned)((n + 255) / 256), 256>>>(d_Pkey, d_Plen, n, ix->d_lcs);
    LAUNCHED();
    labels_kernel<K><<<(unsigned)((n + 255) / 256), 256>>>(d_Pkey, d_Plen, n, k, d_rows32, stride);
    LAUNCHED();
    row_popc_kernel<<<(unsigned)((4 * stride + 255) / 256), 256>>>(d_rows32, 4 * stride, d_pc);
    LAUNCHED();
    size_t scan_need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_need, d_pc, d_prefix, (int64_t)(4 * stride));
    void* d_scan_tmp = d_tmp;
    if (scan_need > tb) {
        CUDA_TRY(cudaMallocAsync(&d_scan_tmp, scan_need, 0));
        tmp.ptrs.push_back(d_scan_tmp);
    }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(d_scan_tmp, scan_need, d_pc, d_prefix, (int64_t)(4 * stride)));
    LAUNCHED();
    compose_rank_kernel<<<(unsigned)((4 * stride + 255) / 256), 256>>>(d_rows32, d_prefix, 4 * stride, ix->d_rank);
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    // host copy of the plain SubsetMatrix form (search / access_kmer / export)
    HostIndex& h = ix->host;
    h.k = k;
    h.n_sets = n;
    h.n_kmers = nR;
    ix->rank_stride = stride;
    ix->view.rank = ix->d_rank;
    ix->view.rank_stride = (uint32_t)stride;
    ix->view.lcs = ix->d_lcs;
    ix->view.n = (uint32_t)n;
    ix->view.k = k;
    bt.lap("lcs, labels, rank words");
    if (device_only) { const int rc = build_links(ix, n, false); bt.lap("links (device-only index)"); return rc; }
    // the plain SubsetMatrix form (search / access_kmer / export on the host) is read back only if somebody asks
    ix->host_ready.store(false, std::memory_order_release);
    if (ix->d_node_keys) ix->device_bytes += n * (sizeof(K) + 1);
    const int rc_links = build_links(ix, n);
    bt.lap("rank2, links, prefix table");
    return rc_links;
}

static int build_index_gpu(kbo_index* ix, const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, uint32_t k,
                           bool revcomp, bool keep_nodes, bool device_only = false) {
    if (k <= 32) return build_index_gpu_typed<uint64_t>(ix, seqs, lens, n_seqs, k, revcomp, keep_nodes, device_only);
    return build_index_gpu_typed<u128>(ix, seqs, lens, n_seqs, k, revcomp, keep_nodes, device_only);
}

// ---------------------------------------------------------------------------
// batch geometry
// ---------------------------------------------------------------------------
static Geometry batch_geometry(const kbo_index* ix, uint64_t total, uint64_t nq, uint32_t overlap = 1) {
    return kbo_b200::make_geometry(total, nq, tuned_chunk_len(ix), overlap);
}

static int check_offsets(const uint64_t* offsets, uint64_t nq, uint64_t min_len, uint64_t* total) {
    if (!offsets) return fail(KBO_ERR_BAD_ARGUMENT, "offsets is null");
    if (nq == 0) return fail(KBO_ERR_EMPTY_INPUT, "no queries (index.rs:248 assert!(!query.is_empty()))");
    for (uint64_t i = 0; i < nq; ++i) {
        if (offsets[i + 1] < offsets[i]) return fail(KBO_ERR_BAD_ARGUMENT, "offsets must be non-decreasing");
        const uint64_t len = offsets[i + 1] - offsets[i];
        if (len == 0) return fail(KBO_ERR_EMPTY_INPUT, "empty query (index.rs:248 assert!(!query.is_empty()))");
        if (len < min_len) return fail(KBO_ERR_TOO_SHORT, "query shorter than 3 bases (derandomize.rs:276 len > 2)");
    }
    *total = offsets[nq] - offsets[0];
    return KBO_OK;
}

// ---------------------------------------------------------------------------
// kernel sequences (all on ws->stream)
// ---------------------------------------------------------------------------
static int run_pack(Workspace* ws, const uint8_t* d_ascii, const uint64_t* d_offsets, uint64_t nq, const Geometry& g,
                    QueryView* qv) {
    cudaStream_t st = ws->stream;
    CUDA_TRY(ws->pack.ensure(g.n_words * 8, st));
    CUDA_TRY(ws->inv.ensure(g.n_words * 4, st));
    CUDA_TRY(ws->sep.ensure(g.n_words * 4, st));
    CUDA_TRY(ws->wq.ensure(g.n_words * 4, st));
    qv->pack = ws->pack.as<uint64_t>();
    qv->inv = ws->inv.as<uint32_t>();
    qv->sep = ws->sep.as<uint32_t>();
    qv->wq = ws->wq.as<uint32_t>();
    qv->Lp = g.Lp;
    qv->n_words = g.n_words;
    const unsigned threads = 128;
    const unsigned blocks = (unsigned)((g.n_words + threads - 1) / threads);
    pack_queries_kernel<<<blocks, threads, 0, st>>>(d_ascii, d_offsets, nq, *qv, ws->pack.as<uint64_t>(),
                                                    ws->inv.as<uint32_t>(), ws->sep.as<uint32_t>(),
                                                    ws->wq.as<uint32_t>());
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return KBO_OK;
}

static int run_ms(kbo_index* ix, Workspace* ws, const QueryView& qv, const Geometry& g, bool intervals) {
    cudaStream_t st = ws->stream;
    CUDA_TRY(ws->ms.ensure(g.ms_bytes, st));
    if (intervals) {
        CUDA_TRY(ws->l.ensure(g.n_words * 32 * 4, st));
        CUDA_TRY(ws->r.ensure(g.n_words * 32 * 4, st));
    }
    const bool count = g_profile_counters.load() != 0;
    if (count) {
        CUDA_TRY(ws->counters.ensure(CNT_N * 8, st));
        CUDA_TRY(cudaMemsetAsync(ws->counters.p, 0, CNT_N * 8, st));
    }
    if (!intervals && (tuned_ms_flags(ix) & 32u)) { const int rc = ensure_rank2(ix); if (rc) return rc; }
    MsParams mp;
    mp.ix = ix->view;
    mp.q = qv;
    mp.chunk_len = g.chunk_len;
    mp.flags = tuned_ms_flags(ix);
    mp.n_chunks = g.n_chunks;
    mp.ms = ws->ms.as<uint8_t>();
    mp.l_out = intervals ? ws->l.as<uint32_t>() : nullptr;
    mp.r_out = intervals ? ws->r.as<uint32_t>() : nullptr;
    mp.counters = count ? ws->counters.as<unsigned long long>() : nullptr;
    const unsigned threads = g_ms_block.load();
    const unsigned blocks = (unsigned)((g.n_chunks + threads - 1) / threads);
    if (intervals) {
        if (count) ms_kernel<true, true><<<blocks, threads, 0, st>>>(mp);
        else ms_kernel<true, false><<<blocks, threads, 0, st>>>(mp);
    } else if ((mp.flags & 32u) && mp.ix.rank2) {  // bit 5: two bases per probe in K1 (ms_pairs_kernel)
        if (count) ms_pairs_kernel<true><<<blocks, threads, 0, st>>>(mp);
        else ms_pairs_kernel<false><<<blocks, threads, 0, st>>>(mp);
    } else {
        if (count) ms_kernel<false, true><<<blocks, threads, 0, st>>>(mp);
        else ms_kernel<false, false><<<blocks, threads, 0, st>>>(mp);
    }
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return KBO_OK;
}

// K2 / K2b.  want_masks == false: characters to d_out (unpadded, at off0).  want_masks == true: the three
// per-position masks K4 consumes, in ws->masks (gap | match | r, n_tiles_b * 32 words each).
static int run_derand_translate(kbo_index* ix, Workspace* ws, const QueryView& qv, const Geometry& g, uint32_t thr,
                                uint8_t* d_out, uint64_t off0, bool want_masks = false) {
    cudaStream_t st = ws->stream;
    TrParams tp;
    tp.ms = ws->ms.as<uint8_t>();
    tp.q = qv;
    tp.k = ix->host.k;
    tp.thr = thr;
    tp.out = d_out;
    tp.off0 = off0;
    tp.out_gap = tp.out_match = tp.out_r = nullptr;
    const uint64_t nw = g.n_tiles_b * 32;
    if (want_masks) {
        CUDA_TRY(ws->masks.ensure(nw * 3 * 4, st));
        tp.out_gap = ws->masks.as<uint32_t>();
        tp.out_match = tp.out_gap + nw;
        tp.out_r = tp.out_match + nw;
    }
    if (k2b_supported(tp.k, tp.thr) && !(tuned_ms_flags(ix) & 2u)) {  // flag bit1: force K2 (experiments / tests)
        tp.n_tiles = g.n_tiles_b;
        const unsigned blocks = (unsigned)((g.n_tiles_b + K2B_WARPS - 1) / K2B_WARPS);
        if (want_masks) derand_translate_bits_kernel<false><<<blocks, K2B_WARPS * 32, 0, st>>>(tp);
        else derand_translate_bits_kernel<true><<<blocks, K2B_WARPS * 32, 0, st>>>(tp);
        LAUNCHED();
    } else {
        if (want_masks) {  // K2 writes characters; a second kernel turns them into masks
            CUDA_TRY(ws->out.ensure(g.total + 16, st));
            tp.out = ws->out.as<uint8_t>();
            tp.off0 = 0;
        }
        tp.n_tiles = g.n_tiles;
        const unsigned blocks = (unsigned)((g.n_tiles + K2_WARPS - 1) / K2_WARPS);
        derand_translate_kernel<<<blocks, K2_WARPS * 32, 0, st>>>(tp);
        LAUNCHED();
        if (want_masks) {
            chars_to_masks_kernel<<<(unsigned)(nw * 32 / 128), 128, 0, st>>>(tp.out, 0, qv.sep, qv.wq, nw, tp.out_gap,
                                                                             tp.out_match, tp.out_r);
            LAUNCHED();
        }
    }
    CUDA_TRY(cudaGetLastError());
    return KBO_OK;
}

// K4 up to the START / END marks: one launch for max_gap_len == 0, two for the gapped form (the scans are inside).  Returns the parameter block for
// run_rle_finish BY VALUE (workspaces bound to caller streams can be shared between host threads).
// d_offsets are the batch's own CSR offsets (offsets[0] may be non-zero).
static int run_rle_counts(Workspace* ws, const QueryView& qv, const Geometry& g, const uint64_t* d_offsets, uint64_t nq,
                          uint32_t max_gap_len, RleParams* out_params) {
    cudaStream_t st = ws->stream;
    const uint64_t nw = g.n_tiles_b * 32;
    const uint64_t nb = (nw + RLE_BLOCK - 1) / RLE_BLOCK;
    CUDA_TRY(ws->rle_words.ensure(nw * 4 * 4, st));
    CUDA_TRY(ws->rle_cnt.ensure((nw + nb + 1) * sizeof(RleCounts), st));
    CUDA_TRY(ws->rle_cse.ensure((nw + nb + 1) * 8, st));
    if (!ws->rle_tickets.p) {
        CUDA_TRY(ws->rle_tickets.ensure(8, st));
        CUDA_TRY(cudaMemsetAsync(ws->rle_tickets.p, 0, 8, st));  // the kernels leave them at zero
    }
    RleParams p;
    std::memset(&p, 0, sizeof(p));
    p.gap = ws->masks.as<uint32_t>();
    p.match = p.gap + nw;
    p.rr = p.match + nw;
    p.sep = qv.sep;
    p.wq = qv.wq;
    p.n_words = nw;
    p.n_blocks = nb;
    p.offsets = d_offsets;
    p.nq = nq;
    p.window = max_gap_len + 1;
    p.jump = ws->rle_words.as<uint32_t>();
    p.gopen = p.jump + nw;
    p.start = p.gopen + nw;
    p.end = p.start + nw;
    p.cnt = ws->rle_cnt.as<RleCounts>();
    p.cnt_blk = p.cnt + nw;
    p.cse = ws->rle_cse.as<uint64_t>();
    p.cse_blk = p.cse + nw;
    p.tickets = ws->rle_tickets.as<unsigned int>();
    if (p.window == 1) {  // max_gap_len == 0: counts and START / END marks in one launch
        rle_word_counts_kernel<true><<<(unsigned)nb, RLE_BLOCK, 0, st>>>(p);
        LAUNCHED();
    } else {
        rle_word_counts_kernel<false><<<(unsigned)nb, RLE_BLOCK, 0, st>>>(p);
        LAUNCHED();
        rle_mark_kernel<<<(unsigned)nb, RLE_BLOCK, 0, st>>>(p);
        LAUNCHED();
    }
    CUDA_TRY(cudaGetLastError());
    *out_params = p;
    return KBO_OK;
}

// K4, last launch: per-query record offsets + the records.  `rle_offsets` (nq + 1 entries) and `out` (`cap` records)
// must be device-visible: device memory, or page-locked host memory mapped into the device (the kernel then writes
// the results straight into the caller's buffer and no device->host copy follows).  `base_in` / `total_out` chain
// the sub-batches of one host call (RleParams).
static int run_rle_finish(cudaStream_t st, RleParams p, uint64_t* rle_offsets, RleRecord* out, uint64_t cap,
                          const uint64_t* base_in = nullptr, uint64_t* total_out = nullptr, bool write_first = true) {
    p.rle_offsets = rle_offsets;
    p.out = out;
    p.cap = cap;
    p.base_in = base_in;
    p.total_out = total_out;
    p.write_first = write_first ? 1u : 0u;
    const unsigned threads = 128;
    const uint64_t items = std::max<uint64_t>(p.n_words, p.nq + 1);
    rle_finish_kernel<<<(unsigned)((items + threads - 1) / threads), threads, 0, st>>>(p);
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return KBO_OK;
}
static_assert(sizeof(RleRecord) == sizeof(kbo_rle), "device and ABI RLE records must agree");

// Copies the first min(*total, cap) records from device memory to the caller's page-locked buffer, 8 bytes per thread,
// consecutive threads consecutive words.  The records kernel itself scatters 8-byte stores (one thread per record);
// done straight into host memory those cost ~30 us per 25,000 records of end-to-end throughput (measured,
// profiles/README.md round 2), a dense copy does not.  The count is read on the device: no host round trip.
__global__ void relay_records_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst,
                                     const uint64_t* __restrict__ total, uint64_t base, uint64_t cap) {
    const uint64_t cnt = *total - base;  // records of this job (`base` = records of the devices before it, multi-GPU)
    const uint64_t n = (cnt < cap ? cnt : cap) * (sizeof(RleRecord) / 8);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

template <typename T>
__global__ void unpad_kernel(const T* __restrict__ in, QueryView q, T* __restrict__ out) {
    const uint64_t pp = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pp >= q.Lp) return;
    const uint32_t sw = __ldg(q.sep + (pp >> 5));
    if ((sw >> (pp & 31)) & 1u) return;
    const uint64_t nsep = __ldg(q.wq + (pp >> 5)) + __popc(sw & ((1u << (pp & 31)) - 1u));
    out[pp - nsep] = in[pp];
}

static int fetch_counters(kbo_index* ix, Workspace* ws) {
    if (!g_profile_counters.load() || !ws->counters.p) return KBO_OK;
    unsigned long long h[CNT_N];
    CUDA_TRY(cudaMemcpyAsync(h, ws->counters.p, sizeof(h), cudaMemcpyDeviceToHost, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    std::lock_guard<std::mutex> g(ix->mu);
    ix->last_counters.extend_attempts = h[CNT_ATTEMPTS];
    ix->last_counters.extend_split_sector = h[CNT_SPLIT];
    ix->last_counters.contractions = h[CNT_CONTRACT];
    ix->last_counters.contraction_extra_words = h[CNT_EXTRA_LCS];
    ix->last_counters.bases_processed = h[CNT_PROCESSED];
    ix->last_counters.bases_emitted = h[CNT_EMITTED];
    ix->last_counters.emit_extend_attempts = h[CNT_ATT_EMIT];
    ix->last_counters.emit_extend_split_sector = h[CNT_SPLIT_EMIT];
    ix->last_counters.emit_contractions = h[CNT_CON_EMIT];
    ix->last_counters.emit_contraction_extra_words = h[CNT_EXTRA_EMIT];
    return KBO_OK;
}

// ---------------------------------------------------------------------------
// fused K1 + K2b (fused.cuh): tile geometry and launch
// ---------------------------------------------------------------------------
static int device_sm_count(int device) {
    static std::mutex mu;
    static std::unordered_map<int, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);
    if (it != cache.end()) return it->second;
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    cache[device] = n;
    return n;
}

template <bool CHARS, bool COUNT, bool EXACT>
static cudaError_t launch_fused(const FusedParams& fp, const FusedGeom& fg, cudaStream_t st) {
    static std::mutex mu;
    static std::vector<int> configured;  // devices on which this instantiation may use > 48 KB of dynamic shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(mu);
        if (std::find(configured.begin(), configured.end(), dev) == configured.end()) {
            cudaError_t e = cudaFuncSetAttribute(ms_fused_kernel<CHARS, COUNT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            configured.push_back(dev);
        }
    }
    ms_fused_kernel<CHARS, COUNT, EXACT><<<(unsigned)fg.n_tiles, FUSED_THREADS, fg.smem.total, st>>>(fp);
    return cudaGetLastError();
}

// K1 + K2b in one launch.  Returns KBO_OK and *done = false when the parameters are outside the fused kernel's range
// (the caller then runs K1 and K2/K2b separately).
static int run_fused(kbo_index* ix, Workspace* ws, const QueryView& qv, const Geometry& g, uint32_t thr, uint8_t* d_out,
                     uint64_t off0, bool want_masks, bool* done) {
    *done = false;
    const uint32_t flags = tuned_ms_flags(ix);
    // bit 4 selects the fused kernel.  It is NOT the default: measured on B200 (profiles/README.md, round 2) its two-pass
    // form takes 225-300 us per 10^7-base batch where K1 + K2b take 112 + 18 us; its one-pass form (bit 3: K1's
    // recurrence in pass A) takes 119 us and wins for a lone stream (63 vs 60 G bases/s) but loses when independent
    // batches overlap on several streams (78 vs 97 G), which is how the library is fastest.
    if (!k2b_supported(ix->host.k, thr) || (flags & 2u) || !(flags & 16u)) return KBO_OK;
    FusedGeom fg;
    const bool exact = (flags & FUSED_FLAG_EXACT) != 0;  // bit 3: the one-pass form (K1's recurrence + K2b in one kernel)
    // (positions per lane stay at ~64 also when calls overlap: with K1's longer chunks for overlapping calls the one-pass
    // form fell from 78 to 64 G bases/s on six streams -- a block waits for the slowest of its 128 lanes before pass C)
    if (!fused_geometry(g.Lp, ix->host.k, !want_masks, device_sm_count(ix->device), tuned_chunk_len(ix), &fg, exact)) return KBO_OK;
    cudaStream_t st = ws->stream;
    if (!exact && !(flags & 4u)) { const int rc = ensure_rank2(ix); if (rc) return rc; }  // (bit 2: one base per probe)
    FusedParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.ix = ix->view;
    fp.q = qv;
    fp.tr.ms = nullptr;
    fp.tr.q = qv;
    fp.tr.k = ix->host.k;
    fp.tr.thr = thr;
    fp.tr.out = d_out;
    fp.tr.off0 = off0;
    fp.tile_len = fg.tile_len;
    fp.chunk = fg.chunk;
    fp.stage_words = fg.stage_words;
    fp.task_cap = fg.task_cap;
    fp.flags = flags;
    const uint64_t nw = g.n_tiles_b * 32;
    if (want_masks) {
        CUDA_TRY(ws->masks.ensure(nw * 3 * 4, st));
        fp.tr.out_gap = ws->masks.as<uint32_t>();
        fp.tr.out_match = fp.tr.out_gap + nw;
        fp.tr.out_r = fp.tr.out_match + nw;
        fp.mask_words = nw;
    }
    const bool count = g_profile_counters.load() != 0;
    if (count) {
        CUDA_TRY(ws->counters.ensure(CNT_N * 8, st));
        CUDA_TRY(cudaMemsetAsync(ws->counters.p, 0, CNT_N * 8, st));
        fp.counters = ws->counters.as<unsigned long long>();
    }
    cudaError_t e;
    if (exact) {
        if (want_masks) e = count ? launch_fused<false, true, true>(fp, fg, st) : launch_fused<false, false, true>(fp, fg, st);
        else e = count ? launch_fused<true, true, true>(fp, fg, st) : launch_fused<true, false, true>(fp, fg, st);
    } else {
        if (want_masks) e = count ? launch_fused<false, true, false>(fp, fg, st) : launch_fused<false, false, false>(fp, fg, st);
        else e = count ? launch_fused<true, true, false>(fp, fg, st) : launch_fused<true, false, false>(fp, fg, st);
    }
    LAUNCHED();
    CUDA_TRY(e);
    *done = true;
    return KBO_OK;
}

// matches for a batch whose inputs are already on the device (ws->stream)
static int matches_device(kbo_index* ix, Workspace* ws, const uint8_t* d_concat, const uint64_t* d_offsets,
                          uint64_t nq, const Geometry& g, uint32_t thr, uint8_t* d_out, uint64_t off0,
                          bool want_masks = false, QueryView* qv_out = nullptr) {
    QueryView qv;
    cudaEvent_t* ev = nullptr;
    if (g_kernel_timing.load() && ws->timed_calls < 512) {
        const size_t need = (ws->timed_calls + 1) * 4;
        while (ws->timing.size() < need) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            ws->timing.push_back(e);
        }
        ev = ws->timing.data() + ws->timed_calls * 4;
        ws->timed_calls++;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev[0], ws->stream));
    int rc = run_pack(ws, d_concat, d_offsets, nq, g, &qv);
    if (rc) return rc;
    if (ev) CUDA_TRY(cudaEventRecord(ev[1], ws->stream));
    bool fused = false;
    rc = run_fused(ix, ws, qv, g, thr, d_out, off0, want_masks, &fused);
    if (rc) return rc;
    if (fused) {  // one kernel: its time is reported as "ms", derandomize + translate as 0
        if (ev) CUDA_TRY(cudaEventRecord(ev[2], ws->stream));
        if (ev) CUDA_TRY(cudaEventRecord(ev[3], ws->stream));
    } else {
        rc = run_ms(ix, ws, qv, g, false);
        if (rc) return rc;
        if (ev) CUDA_TRY(cudaEventRecord(ev[2], ws->stream));
        rc = run_derand_translate(ix, ws, qv, g, thr, d_out, off0, want_masks);
        if (rc) return rc;
        if (ev) CUDA_TRY(cudaEventRecord(ev[3], ws->stream));
    }
    if (qv_out) *qv_out = qv;
    return KBO_OK;
}

// ---------------------------------------------------------------------------
// random-sector load micro-benchmark (roofline denominators for L2 / HBM)
// ---------------------------------------------------------------------------
template <bool DEPENDENT>
__global__ void __launch_bounds__(256) random_sector_kernel(const uint64_t* __restrict__ buf, uint32_t n_sectors,
                                                            uint32_t iters, unsigned long long* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = tid * 2654435761u + 12345u;
    uint64_t acc = 0;
    if (DEPENDENT) {
        for (uint32_t i = 0; i < iters; ++i) {
            x = x * 1664525u + 1013904223u;
            const uint32_t s = (uint32_t)(((uint64_t)(x ^ (uint32_t)acc) * n_sectors) >> 32);
            acc += __ldg(buf + (size_t)s * 4 + (x & 3));  // the next address depends on this value
        }
    } else {
        for (uint32_t i = 0; i < iters; i += 8) {
            uint64_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x = x * 1664525u + 1013904223u;
                const uint32_t s = (uint32_t)(((uint64_t)x * n_sectors) >> 32);
                v[j] = __ldg(buf + (size_t)s * 4 + (x & 3));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
    }
    if (acc == 0x123456789abcdefull) atomicAdd(sink, 1ull);
}

// ---------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------
extern "C" {

const char* kbo_last_error_message(void) { return g_err.c_str(); }

int kbo_device_count(int* out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return fail(KBO_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    *out = n;
    return KBO_OK;
}

void kbo_default_build_opts(kbo_build_opts* o) {
    if (!o) return;
    o->k = 31;
    o->add_revcomp = 0;
    o->num_threads = 1;
    o->prefix_precalc = 8;
    o->build_select = 0;
    o->mem_gb = 4;
    o->dedup_batches = 0;
    o->temp_dir = nullptr;
}

int kbo_alloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return KBO_OK;
}
int kbo_free_pinned(void* p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return KBO_OK;
}

static int finish_index(kbo_index* ix, int device, kbo_index** out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        delete ix;
        return fail(KBO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) { delete ix; return fail(KBO_ERR_BAD_ARGUMENT, "device ordinal out of range"); }
    ix->device = device;
    DeviceGuard dg(device);
    if (!dg.ok) { delete ix; return fail(KBO_ERR_CUDA, "cudaSetDevice failed"); }
    int rc = upload_index(ix);
    if (rc) { kbo_index_free(ix); return rc; }
    *out = ix;
    return KBO_OK;
}

int kbo_index_build(const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs, const kbo_build_opts* opts,
                    int device, kbo_index** out) {
    ScopeTimer scope_timer("kbo_index_build");
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
    if (!seqs || !lens || n_seqs == 0) return fail(KBO_ERR_EMPTY_INPUT, "no input sequences (index.rs:60)");
    kbo_build_opts o;
    if (opts) o = *opts; else kbo_default_build_opts(&o);
    if (o.k == 0 || o.k > KBO_MAX_K) return fail(KBO_ERR_BAD_K, "1 <= k <= 64 in this build");
    kbo_index* ix = new kbo_index();
    if (o.k >= 2 && o.k <= KBO_MAX_K && !g_host_builder.load()) {  // device construction (index_build.cuh)
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            delete ix;
            return fail(KBO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
        }
        if (device < 0 || device >= ndev) { delete ix; return fail(KBO_ERR_BAD_ARGUMENT, "device ordinal out of range"); }
        ix->device = device;
        DeviceGuard dg(device);
        if (!dg.ok) { delete ix; return fail(KBO_ERR_CUDA, "cudaSetDevice failed"); }
        int rc = build_index_gpu(ix, seqs, lens, n_seqs, o.k, o.add_revcomp != 0, o.build_select != 0);
        if (rc) { kbo_index_free(ix); return rc; }
        *out = ix;
        return KBO_OK;
    }
    std::string err = build_host_index(seqs, lens, n_seqs, o.k, o.add_revcomp != 0, o.num_threads ? o.num_threads : 1,
                                       &ix->host, o.build_select != 0);
    if (!err.empty()) { delete ix; return fail(KBO_ERR_INDEX_TOO_LARGE, err); }
    return finish_index(ix, device, out);
}

int kbo_index_from_parts(uint32_t k, uint64_t n_sets, uint64_t n_kmers, const uint64_t* const rows[4],
                         const uint8_t* lcs, int device, kbo_index** out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
    if (!rows || !lcs || !rows[0] || !rows[1] || !rows[2] || !rows[3]) return fail(KBO_ERR_BAD_ARGUMENT, "null part");
    if (k == 0 || k > 127) return fail(KBO_ERR_BAD_K, "1 <= k <= 127");
    if (n_sets == 0) return fail(KBO_ERR_EMPTY_INPUT, "n_sets == 0");
    kbo_index* ix = new kbo_index();
    HostIndex& h = ix->host;
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
    // K1 relies on these: LCS[0] == 0 stops every left scan, LCS bytes < k <= 127 keep the seven-bit compares
    // exact, and every node but the root has exactly one incoming edge (so ranks stay inside [0, n_sets]).
    bool lcs_ok = h.lcs[0] == 0;
    for (uint64_t i = 0; i < n_sets && lcs_ok; ++i) lcs_ok = h.lcs[i] < k;
    uint64_t bits = 0;
    for (int c = 0; c < 4; ++c)
        for (size_t w = 0; w < nw; ++w) bits += (uint64_t)__builtin_popcountll(h.rows[c][w]);
    if (!lcs_ok || bits != n_sets - 1) {
        delete ix;
        return fail(KBO_ERR_BAD_ARGUMENT, !lcs_ok ? "LCS array invalid: need LCS[0] == 0 and every value < k"
                                                  : "subset rows must hold exactly n_sets - 1 set bits");
    }
    h.finalize();
    return finish_index(ix, device, out);
}

void kbo_index_free(kbo_index* ix) {
    if (!ix) return;
    ScopeTimer scope_timer("kbo_index_free");
    {
        BuildTimer bt;
        DeviceGuard dg(ix->device);
        // (idle pooled workspaces belong to the device, not to this index: they stay for the next index)
        for (auto& kv : ix->by_stream) { kv.second->destroy(); delete kv.second; }
        for (PinnedBuf& pb : ix->pinned_pool) if (pb.p) cudaFreeHost(pb.p);
        // pool memory goes back stream-ordered unless stream-ordered (_device) calls on caller streams may still be
        // running (cudaFree waits for them; it accepts pool memory as well)
        const bool async_free = ix->pool_alloc && ix->by_stream.empty();
        void* arrays[] = {ix->d_blob, ix->d_pref, ix->d_node_keys, ix->d_node_len};
        for (void* p : arrays) {
            if (!p) continue;
            if (async_free) cudaFreeAsync(p, 0);
            else cudaFree(p);
        }
        bt.lap("index free");
    }
    delete ix;
}

uint32_t kbo_index_k(const kbo_index* ix) { return ix ? ix->host.k : 0; }
uint64_t kbo_index_n_kmers(const kbo_index* ix) { return ix ? ix->host.n_kmers : 0; }
uint64_t kbo_index_n_sets(const kbo_index* ix) { return ix ? ix->host.n_sets : 0; }
int kbo_index_device(const kbo_index* ix) { return ix ? ix->device : -1; }
uint64_t kbo_index_device_bytes(const kbo_index* ix) { return ix ? ix->device_bytes : 0; }

int kbo_index_export_parts(const kbo_index* ix, uint64_t* rows[4], uint8_t* lcs, uint64_t C_out[4]) {
    if (!ix) return fail(KBO_ERR_BAD_ARGUMENT, "index is null");
    { const int rc = ensure_host_mirror(ix); if (rc) return rc; }
    const size_t nw = (size_t)(ix->host.n_sets + 63) / 64;
    if (rows)
        for (int c = 0; c < 4; ++c)
            if (rows[c]) std::memcpy(rows[c], ix->host.rows[c].data(), nw * 8);
    if (lcs) std::memcpy(lcs, ix->host.lcs.data(), (size_t)ix->host.n_sets);
    if (C_out)
        for (int c = 0; c < 4; ++c) C_out[c] = ix->host.C[c];
    return KBO_OK;
}

// ---- index files: index::serialize_sbwt / index::load_sbwt (index.rs:128-212) --------------------------
// The reference writes `<prefix>.sbwt` = u64 LE 12 + "SubsetMatrix" (index.rs:139-140) followed by the sbwt crate's own
// serialisation of the index, and `<prefix>.lcs` = the crate's serialisation of the LCS array.  The crate (sbwt
// 0.3.4) is not part of the reference tree and the reference pins the files only by a round trip (index.rs:277-296),
// so the byte layout AFTER the variant header is this library's own (below) and a file written by kbo-cli is
// recognised and refused (KBO_ERR_FORMAT) instead of being misread.
//   .sbwt: u64 12, "SubsetMatrix", "KBOB200\0", u32 version (1), u32 k, u64 n_sets, u64 n_kmers,
//          4 x ceil(n_sets/64) u64 words (the kbo_index_from_parts rows, A C G T), u64 FNV-1a of everything before
//   .lcs:  "KBOB200\0", u32 version (1), u32 k, u64 n_sets, n_sets bytes, u64 FNV-1a of everything before
// all little endian.
namespace {
const char IO_VARIANT[] = "SubsetMatrix";
const char IO_MAGIC[8] = {'K', 'B', 'O', 'B', '2', '0', '0', '\0'};
const uint32_t IO_VERSION = 1;
struct IoFile {
    FILE* f = nullptr;
    uint64_t h = 1469598103934665603ull;  // FNV-1a 64
    ~IoFile() { if (f) std::fclose(f); }
    bool close() {  // a writer checks this: buffered data may only fail to reach the disk here
        const int rc = f ? std::fclose(f) : 0;
        f = nullptr;
        return rc == 0;
    }
    bool write(const void* p, size_t n) {
        const uint8_t* b = (const uint8_t*)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        return std::fwrite(p, 1, n, f) == n;
    }
    bool read(void* p, size_t n) {
        if (std::fread(p, 1, n, f) != n) return false;
        const uint8_t* b = (const uint8_t*)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        return true;
    }
    bool put32(uint32_t v) { return write(&v, 4); }  // the host is little endian (x86-64 / aarch64)
    bool put64(uint64_t v) { return write(&v, 8); }
    bool get(uint32_t& v) { return read(&v, 4); }
    bool get(uint64_t& v) { return read(&v, 8); }
};
}  // namespace

int kbo_index_serialize(const kbo_index* ix, const char* outfile_prefix) {
    if (!ix || !outfile_prefix) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    { const int rc = ensure_host_mirror(ix); if (rc) return rc; }
    const HostIndex& h = ix->host;
    const size_t nw = (size_t)(h.n_sets + 63) / 64;
    const std::string sbwt_path = std::string(outfile_prefix) + ".sbwt", lcs_path = std::string(outfile_prefix) + ".lcs";
    {
        IoFile o;
        o.f = std::fopen(sbwt_path.c_str(), "wb");
        if (!o.f) return fail(KBO_ERR_IO, "Expected write access to " + sbwt_path);  // index.rs:137
        bool ok = o.put64(sizeof(IO_VARIANT) - 1) && o.write(IO_VARIANT, sizeof(IO_VARIANT) - 1) &&
                  o.write(IO_MAGIC, 8) && o.put32(IO_VERSION) && o.put32(h.k) &&
                  o.put64(h.n_sets) && o.put64(h.n_kmers);
        for (int c = 0; c < 4 && ok; ++c) ok = o.write(h.rows[c].data(), nw * 8);
        const uint64_t sum = o.h;
        ok = ok && o.put64(sum) && o.close();
        if (!ok) return fail(KBO_ERR_IO, "write failed: " + sbwt_path);
    }
    {
        IoFile o;
        o.f = std::fopen(lcs_path.c_str(), "wb");
        if (!o.f) return fail(KBO_ERR_IO, "Expected write access to " + lcs_path);  // index.rs:148
        bool ok = o.write(IO_MAGIC, 8) && o.put32(IO_VERSION) && o.put32(h.k) && o.put64(h.n_sets) &&
                  o.write(h.lcs.data(), (size_t)h.n_sets);
        const uint64_t sum = o.h;
        ok = ok && o.put64(sum) && o.close();
        if (!ok) return fail(KBO_ERR_IO, "write failed: " + lcs_path);
    }
    return KBO_OK;
}

int kbo_index_load(const char* index_prefix, int device, kbo_index** out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
    if (!index_prefix) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    const std::string sbwt_path = std::string(index_prefix) + ".sbwt", lcs_path = std::string(index_prefix) + ".lcs";
    uint32_t k = 0;
    uint64_t n_sets = 0, n_kmers = 0;
    std::vector<uint64_t> rows[4];
    std::vector<uint8_t> lcs;
    {
        IoFile in;
        in.f = std::fopen(sbwt_path.c_str(), "rb");
        if (!in.f) return fail(KBO_ERR_IO, "Expected SBWT at " + sbwt_path);  // index.rs:202
        uint64_t name_len = 0;
        char name[sizeof(IO_VARIANT)] = {0}, magic[8] = {0};
        uint32_t version = 0;
        if (!in.get(name_len) || name_len != sizeof(IO_VARIANT) - 1 || !in.read(name, name_len) ||
            std::memcmp(name, IO_VARIANT, name_len) != 0)
            return fail(KBO_ERR_FORMAT, sbwt_path + ": not a SubsetMatrix index file (index.rs:139-140)");
        if (!in.read(magic, 8) || std::memcmp(magic, IO_MAGIC, 8) != 0)
            return fail(KBO_ERR_FORMAT, sbwt_path + ": a SubsetMatrix file whose body was not written by this library "
                                                    "(the sbwt crate's own layout is not readable here; rebuild the "
                                                    "index from the sequences or pass its arrays to kbo_index_from_parts)");
        if (!in.get(version) || version != IO_VERSION || !in.get(k) || !in.get(n_sets) || !in.get(n_kmers))
            return fail(KBO_ERR_FORMAT, sbwt_path + ": unsupported version or truncated header");
        if (k == 0 || k > 127 || n_sets == 0 || n_sets >= (1ull << 32))
            return fail(KBO_ERR_FORMAT, sbwt_path + ": k or n_sets out of range");
        const size_t nw = (size_t)(n_sets + 63) / 64;
        for (int c = 0; c < 4; ++c) {
            rows[c].resize(nw);
            if (!in.read(rows[c].data(), nw * 8)) return fail(KBO_ERR_FORMAT, sbwt_path + ": truncated");
        }
        const uint64_t sum = in.h;
        uint64_t stored = 0;
        if (!in.get(stored) || stored != sum) return fail(KBO_ERR_FORMAT, sbwt_path + ": checksum mismatch");
    }
    {
        IoFile in;
        in.f = std::fopen(lcs_path.c_str(), "rb");
        if (!in.f) return fail(KBO_ERR_IO, "Expected LCS array at " + lcs_path);  // index.rs:207
        char magic[8] = {0};
        uint32_t version = 0, k2 = 0;
        uint64_t n2 = 0;
        if (!in.read(magic, 8) || std::memcmp(magic, IO_MAGIC, 8) != 0)
            return fail(KBO_ERR_FORMAT, lcs_path + ": not an LCS file written by this library");
        if (!in.get(version) || version != IO_VERSION || !in.get(k2) || !in.get(n2))
            return fail(KBO_ERR_FORMAT, lcs_path + ": unsupported version or truncated header");
        if (k2 != k || n2 != n_sets) return fail(KBO_ERR_FORMAT, lcs_path + ": does not belong to " + sbwt_path);
        lcs.resize((size_t)n_sets);
        if (!in.read(lcs.data(), (size_t)n_sets)) return fail(KBO_ERR_FORMAT, lcs_path + ": truncated");
        const uint64_t sum = in.h;
        uint64_t stored = 0;
        if (!in.get(stored) || stored != sum) return fail(KBO_ERR_FORMAT, lcs_path + ": checksum mismatch");
    }
    const uint64_t* rp[4] = {rows[0].data(), rows[1].data(), rows[2].data(), rows[3].data()};
    const int rc = kbo_index_from_parts(k, n_sets, n_kmers, rp, lcs.data(), device, out);
    return rc == KBO_ERR_BAD_ARGUMENT ? fail(KBO_ERR_FORMAT, sbwt_path + ": inconsistent arrays (" + g_err + ")") : rc;
}

int kbo_index_access_kmer(const kbo_index* ix, uint64_t colex, uint8_t* out_k) {
    if (!ix || !out_k) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    if (colex >= ix->host.n_sets) return fail(KBO_ERR_PANIC, "access_kmer: colex rank out of range");
    { const int rc = ensure_host_mirror(ix); if (rc) return rc; }
    ix->host.access_kmer(colex, out_k);
    return KBO_OK;
}

int kbo_index_search(const kbo_index* ix, const uint8_t* pattern, uint64_t len, int* found, uint64_t* l, uint64_t* r) {
    if (!ix || !found || (!pattern && len)) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t a = 0, b = 0;
    { const int rc = ensure_host_mirror(ix); if (rc) return rc; }
    *found = ix->host.search(pattern, len, &a, &b) ? 1 : 0;
    if (*found) {
        if (l) *l = a;
        if (r) *r = b;
    }
    return KBO_OK;
}

// ---- derandomize.rs host functions --------------------------------------------
int kbo_log_rm_max_cdf(uint64_t t, uint64_t alphabet_size, uint64_t n_kmers, double* out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    if (n_kmers == 0) return fail(KBO_ERR_BAD_ARGUMENT, "n_kmers must be > 0 (derandomize.rs:96)");
    if (alphabet_size == 0) return fail(KBO_ERR_BAD_ARGUMENT, "alphabet_size must be > 0 (derandomize.rs:97)");
    *out = host_log_rm_max_cdf(t, alphabet_size, n_kmers);
    return KBO_OK;
}

int kbo_random_match_threshold(uint64_t k, uint64_t n_kmers, uint64_t alphabet_size, double max_error_prob,
                               uint64_t* out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    return host_threshold(k, n_kmers, alphabet_size, max_error_prob, out);
}

// ---- query_sbwt -------------------------------------------------------------------
int kbo_query_sbwt_batch_compact(const kbo_index* cix, const uint8_t* concat, const uint64_t* offsets,
                                 uint64_t n_queries, uint8_t* d_out, uint32_t* l_out, uint32_t* r_out) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix || !concat || !d_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    int rc = check_offsets(offsets, n_queries, 1, &total);
    if (rc) return rc;
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    Workspace* ws = nullptr;
    rc = acquire_ws(ix, &ws);
    if (rc) return rc;
    const Geometry g = batch_geometry(ix, total, n_queries);
    const bool intervals = l_out || r_out;
    cudaStream_t st = ws->stream;
    auto body = [&]() -> int {
        CUDA_TRY(ws->ascii.ensure(total, st));
        CUDA_TRY(ws->offsets.ensure((n_queries + 1) * 8, st));
        CUDA_TRY(ws->out.ensure(total, st));
        if (intervals) {
            CUDA_TRY(ws->out2.ensure(total * 4, st));
            CUDA_TRY(ws->out3.ensure(total * 4, st));
        }
        CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, concat + offsets[0], total, cudaMemcpyHostToDevice, st));
        std::vector<uint64_t> rel(n_queries + 1);
        for (uint64_t i = 0; i <= n_queries; ++i) rel[i] = offsets[i] - offsets[0];
        CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, rel.data(), (n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(ws->ev0, st));
        QueryView qv;
        int rc2 = run_pack(ws, ws->ascii.as<uint8_t>(), ws->offsets.as<uint64_t>(), n_queries, g, &qv);
        if (rc2) return rc2;
        rc2 = run_ms(ix, ws, qv, g, intervals);
        if (rc2) return rc2;
        const unsigned threads = 256;
        const unsigned blocks = (unsigned)((g.Lp + threads - 1) / threads);
        unpad_kernel<uint8_t><<<blocks, threads, 0, st>>>(ws->ms.as<uint8_t>(), qv, ws->out.as<uint8_t>());
        LAUNCHED();
        if (intervals) {
            unpad_kernel<uint32_t><<<blocks, threads, 0, st>>>(ws->l.as<uint32_t>(), qv, ws->out2.as<uint32_t>());
            LAUNCHED();
            unpad_kernel<uint32_t><<<blocks, threads, 0, st>>>(ws->r.as<uint32_t>(), qv, ws->out3.as<uint32_t>());
            LAUNCHED();
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ws->ev1, st));
        CUDA_TRY(cudaMemcpyAsync(d_out + offsets[0], ws->out.p, total, cudaMemcpyDeviceToHost, st));
        if (l_out) CUDA_TRY(cudaMemcpyAsync(l_out + offsets[0], ws->out2.p, total * 4, cudaMemcpyDeviceToHost, st));
        if (r_out) CUDA_TRY(cudaMemcpyAsync(r_out + offsets[0], ws->out3.p, total * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ws->ev0, ws->ev1);
        ix->last_kernel_ms = ms;
        return fetch_counters(ix, ws);
    };
    rc = body();
    release_ws(ix, ws);
    return rc;
}

int kbo_query_sbwt(const kbo_index* ix, const uint8_t* query, uint64_t len, uint64_t* d_out, uint64_t* l_out,
                   uint64_t* r_out) {
    if (!ix || !d_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    if (!query || len == 0) return fail(KBO_ERR_EMPTY_INPUT, "empty query (index.rs:248 assert!(!query.is_empty()))");
    const uint64_t offsets[2] = {0, len};
    std::vector<uint8_t> d(len);
    std::vector<uint32_t> l, r;
    const bool intervals = l_out || r_out;
    if (intervals) { l.resize(len); r.resize(len); }
    int rc = kbo_query_sbwt_batch_compact(ix, query, offsets, 1, d.data(), intervals ? l.data() : nullptr,
                                          intervals ? r.data() : nullptr);
    if (rc) return rc;
    for (uint64_t i = 0; i < len; ++i) d_out[i] = d[i];
    if (l_out) for (uint64_t i = 0; i < len; ++i) l_out[i] = l[i];
    if (r_out) for (uint64_t i = 0; i < len; ++i) r_out[i] = r[i];
    return KBO_OK;
}

// ---- matches / find / map ----------------------------------------------------------
static int matches_prologue(kbo_index* ix, const uint64_t* offsets, uint64_t nq, double p, uint64_t* total,
                            uint32_t* thr) {
    if (!ix) return fail(KBO_ERR_BAD_ARGUMENT, "index is null");
    uint64_t t = 0;
    int rc = host_threshold(ix->host.k, ix->host.n_kmers, 4, p, &t);  // lib.rs:620
    if (rc) return rc;
    // the reference checks in this order: non-empty query (index.rs:248), threshold > 1 (derandomize.rs:275),
    // len > 2 (derandomize.rs:276); one pass over the offsets when the threshold is fine
    rc = check_offsets(offsets, nq, t <= 1 ? 1 : 3, total);
    if (rc) return rc;
    if (t <= 1) return fail(KBO_ERR_BAD_THRESHOLD, "threshold must be > 1 (derandomize.rs:275)");
    *thr = (uint32_t)t;
    return KBO_OK;
}

static std::vector<uint64_t> split_queries(const uint64_t* offsets, uint64_t nq, uint64_t parts);

// K4 keeps its prefix counts (and K0 / K2b their word indices) in 32 bits: one launch sequence handles at most
// KBO_MAX_LAUNCH_POSITIONS padded positions.  Host-buffer calls split larger batches into sub-batches; the
// stream-ordered (_device) calls, which are one launch sequence, reject them.
static const uint64_t KBO_MAX_LAUNCH_POSITIONS = (1ull << 32) - (1ull << 20);
static uint64_t min_parts_for(uint64_t total, uint64_t nq) {
    return (total + nq) / (KBO_MAX_LAUNCH_POSITIONS / 2) + 1;  // (halved: split points fall on query borders)
}
static int check_launch_size(uint64_t total, uint64_t nq) {
    if (total + nq >= KBO_MAX_LAUNCH_POSITIONS)
        return fail(KBO_ERR_BATCH_TOO_LARGE, "a device-resident batch must stay below 2^32 - 2^20 padded positions "
                                             "(sum of lengths + number of queries); split it");
    return KBO_OK;
}

// Sizes every buffer of a workspace for a WHOLE host call (not for the sub-batch it will run), so that a
// workspace never has to be regrown -- cudaFree synchronises the device -- when calls alternate between one and
// several sub-batches.
static int reserve_ws(Workspace* ws, uint64_t total, uint64_t nq, bool for_find) {
    cudaStream_t st = ws->stream;
    const Geometry g = batch_geometry(nullptr, total, nq);
    const uint64_t nw = g.n_tiles_b * 32;
    CUDA_TRY(ws->ascii.ensure(total, st));
    CUDA_TRY(ws->offsets.ensure((nq + 1) * 8, st));
    CUDA_TRY(ws->pack.ensure(g.n_words * 8, st));
    CUDA_TRY(ws->inv.ensure(g.n_words * 4, st));
    CUDA_TRY(ws->sep.ensure(g.n_words * 4, st));
    CUDA_TRY(ws->wq.ensure(g.n_words * 4, st));
    CUDA_TRY(ws->ms.ensure(g.ms_bytes, st));
    if (for_find) {
        CUDA_TRY(ws->masks.ensure(nw * 3 * 4, st));
        CUDA_TRY(ws->rle_words.ensure(nw * 4 * 4, st));
        const uint64_t nb = (nw + RLE_BLOCK - 1) / RLE_BLOCK;
        CUDA_TRY(ws->rle_cnt.ensure((nw + nb + 1) * sizeof(RleCounts), st));
        CUDA_TRY(ws->rle_cse.ensure((nw + nb + 1) * 8, st));
        CUDA_TRY(ws->tmp64.ensure((nq + 1) * 8, st));
        CUDA_TRY(ws->h_rel.ensure((nq + 1) * 8));
        CUDA_TRY(ws->h_roff.ensure((nq + 1) * 8));
    } else {
        CUDA_TRY(ws->out.ensure(total + 16, st));
    }
    return KBO_OK;
}

// Number of sub-batches a host-buffer batch call is cut into.  A lone caller overlaps its own copy-in with its
// kernels by pipelining sub-batches; when other host threads are inside the library at the same time their calls
// already overlap each other, and further splitting only multiplies driver calls (measured: 3 threads 40 vs 28 G
// bases/s, 6 threads 46 vs 31 G bases/s with 1 vs 4 sub-batches).
struct HostCallScope {
    kbo_index* ix;
    bool concurrent;  // this index is (or has been) used from several host threads at once; sticky, so that the
                      // choice below does not flip back and forth
    explicit HostCallScope(kbo_index* i) : ix(i) {
        if (i->host_calls.fetch_add(1) > 0) i->seen_concurrency.store(true);
        concurrent = i->seen_concurrency.load();
    }
    ~HostCallScope() { ix->host_calls.fetch_sub(1); }
};
static uint64_t pick_parts(const HostCallScope& scope, uint64_t total) {
    if (g_profile_counters.load()) return 1;
    if (tuned_parts(scope.ix, false)) return tuned_parts(scope.ix, false);
    if (scope.concurrent) return 1;
    return std::min<uint64_t>(4, std::max<uint64_t>(1, total >> 21));
}



// kbo::matches for a host CSR batch; large batches are pipelined over sub-batches on separate streams
// (copy-in of part i+1 and copy-out of part i-1 overlap the kernels of part i).
int kbo_matches_batch(const kbo_index* cix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                      double max_error_prob, uint8_t* chars_out) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!concat || !chars_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(ix, offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    const HostCallScope scope(ix);
    const uint64_t want_parts = std::max<uint64_t>(pick_parts(scope, total), min_parts_for(total, n_queries));
    const std::vector<uint64_t> cut = split_queries(offsets, n_queries, want_parts);
    const size_t np = cut.size() - 1;
    std::vector<Workspace*> wss(np, nullptr);
    for (size_t s = 0; s < np && !rc; ++s) {
        rc = acquire_ws(ix, &wss[s]);
        if (rc) break;
        rc = reserve_ws(wss[s], total, n_queries, false);
        if (rc) break;
        Workspace* ws = wss[s];
        cudaStream_t st = ws->stream;
        const uint64_t q0 = cut[s], q1 = cut[s + 1], nq = q1 - q0;
        const uint64_t bytes = offsets[q1] - offsets[q0];
        const Geometry g = batch_geometry(ix, bytes, nq);
        auto body = [&]() -> int {
            CUDA_TRY(ws->ascii.ensure(bytes, st));
            CUDA_TRY(ws->offsets.ensure((nq + 1) * 8, st));
            CUDA_TRY(ws->out.ensure(bytes + 16, st));
            CUDA_TRY(ws->h_rel.ensure((nq + 1) * 8));
            uint64_t* rel = ws->h_rel.as<uint64_t>();
            for (uint64_t i = 0; i <= nq; ++i) rel[i] = offsets[q0 + i] - offsets[q0];
            CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, concat + offsets[q0], bytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, rel, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaEventRecord(ws->ev0, st));
            int rc2 = matches_device(ix, ws, ws->ascii.as<uint8_t>(), ws->offsets.as<uint64_t>(), nq, g, thr,
                                     ws->out.as<uint8_t>(), 0);
            if (rc2) return rc2;
            CUDA_TRY(cudaEventRecord(ws->ev1, st));
            CUDA_TRY(cudaMemcpyAsync(chars_out + offsets[q0], ws->out.p, bytes, cudaMemcpyDeviceToHost, st));
            return KBO_OK;
        };
        rc = body();
    }
    for (size_t s = 0; s < np; ++s) {
        if (!wss[s]) continue;
        cudaError_t e = cudaStreamSynchronize(wss[s]->stream);
        if (e != cudaSuccess && !rc) rc = fail(KBO_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    }
    if (!rc && np == 1) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, wss[0]->ev0, wss[0]->ev1);
        ix->last_kernel_ms = ms;
        rc = fetch_counters(ix, wss[0]);
    }
    for (Workspace* w : wss) if (w) release_ws(ix, w);
    return rc;
}

int kbo_matches(const kbo_index* ix, const uint8_t* query, uint64_t len, double max_error_prob, uint8_t* chars_out) {
    const uint64_t offsets[2] = {0, len};
    if (!query && len) return fail(KBO_ERR_BAD_ARGUMENT, "query is null");
    return kbo_matches_batch(ix, query ? query : (const uint8_t*)"", offsets, 1, max_error_prob, chars_out);
}

// Device-resident batch: the batch is cut into sub-batches that run K0 -> K1 -> K2 (-> K4 count) on their own
// streams, forked from and joined back into the caller's stream.  K1 alone does not fill the machine (it is
// bound by instruction issue at ~50 % occupancy), so the streaming kernels of one part overlap K1 of another.
// `counts` / `stage` (optional) receive the per-query segment counts and staged records for kbo::find.
static int matches_device_forked(kbo_index* ix, Workspace* ws, const uint8_t* d_concat, const uint64_t* d_offsets,
                                 const uint64_t* host_offsets, uint64_t nq, uint32_t thr, uint8_t* d_chars) {
    const uint64_t total = host_offsets[nq];
    uint64_t want = tuned_parts(ix, true);
    if (!want) want = std::min<uint64_t>(2, std::max<uint64_t>(1, total >> 22));
    if (g_profile_counters.load() || g_kernel_timing.load()) want = 1;  // instrumentation passes stay serial
    const std::vector<uint64_t> cut = split_queries(host_offsets, nq, want);
    const size_t np = cut.size() - 1;
    cudaStream_t user = ws->stream;
    if (np == 1) {
        const bool instrumented = g_profile_counters.load() || g_kernel_timing.load();
        const Geometry g = batch_geometry(ix, total, nq, instrumented ? 1 : expected_overlap(ix));
        return matches_device(ix, ws, d_concat, d_offsets, nq, g, thr, d_chars, 0);
    }
    while (ws->subs.size() < np) {
        Workspace* sub = new Workspace();
        sub->own_stream = true;
        cudaEvent_t e = nullptr;
        cudaError_t ce = cudaStreamCreateWithFlags(&sub->stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (ce != cudaSuccess) {  // keep subs and ev_join the same length
            if (sub->stream) cudaStreamDestroy(sub->stream);
            delete sub;
            return fail(KBO_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(ce));
        }
        apply_l2_window(ix, sub->stream);
        ws->subs.push_back(sub);
        ws->ev_join.push_back(e);
    }
    if (!ws->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ws->ev_fork, user));
    for (size_t s = 0; s < np; ++s) {
        Workspace* sub = ws->subs[s];
        const uint64_t q0 = cut[s], q1 = cut[s + 1], n = q1 - q0;
        const Geometry g = batch_geometry(ix, host_offsets[q1] - host_offsets[q0], n);
        CUDA_TRY(cudaStreamWaitEvent(sub->stream, ws->ev_fork, 0));
        int rc = matches_device(ix, sub, d_concat, d_offsets + q0, n, g, thr, d_chars, host_offsets[q0]);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(ws->ev_join[s], sub->stream));
        CUDA_TRY(cudaStreamWaitEvent(user, ws->ev_join[s], 0));
    }
    return KBO_OK;
}

int kbo_matches_batch_device(const kbo_index* cix, const uint8_t* d_concat, const uint64_t* d_offsets,
                             const uint64_t* host_offsets, uint64_t n_queries, double max_error_prob,
                             uint8_t* d_chars_out, void* stream) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!d_concat || !d_offsets || !d_chars_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(ix, host_offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    if (host_offsets[0] != 0) return fail(KBO_ERR_BAD_ARGUMENT, "device batches must have offsets[0] == 0");
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    rc = check_launch_size(total, n_queries);
    if (rc) return rc;
    Workspace* ws = nullptr;
    rc = stream_ws(ix, (cudaStream_t)stream, &ws);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ws->mu);
    rc = matches_device_forked(ix, ws, d_concat, d_offsets, host_offsets, n_queries, thr, d_chars_out);
    if (rc) return rc;
    if (g_profile_counters.load()) return fetch_counters(ix, ws);
    return KBO_OK;
}

// format.rs:143-193 restated as a host pass over one alignment
static int host_run_lengths(const uint8_t* a, uint64_t n, uint64_t max_gap_len, std::vector<kbo_rle>* out) {
    uint64_t i = 0;
    while (i < n) {
        if (a[i] == '-' || a[i] == ' ') { ++i; continue; }
        kbo_rle e = {i, 0, 0, 0, 0, 0, 0};
        uint64_t gap_run = 0;
        bool in_gap = false;
        while (i < n && a[i] != ' ') {
            const uint8_t ch = a[i];
            const bool true_gap = ch == '-';
            if (true_gap && !in_gap) { in_gap = true; ++e.gap_opens; gap_run = 0; }
            if (!true_gap) in_gap = false;
            const bool is_match = ch == 'M' || ch == 'R' || ch == 'I';
            const bool is_gap = true_gap || ch == 'D';
            e.matches += is_match;
            e.gap_bases += is_gap;
            e.mismatches += (!is_match && !is_gap);
            if (!is_gap) e.end = i + 1;
            if (ch == 'R') {
                if (i == 0) return fail(KBO_ERR_PANIC, "alignment starts with 'R' (format.rs:176 aln[i - 1])");
                e.jumps += (a[i - 1] == 'R');
            }
            gap_run += true_gap;
            ++i;
            if (gap_run > max_gap_len || (is_gap && i == n && e.gap_opens > 0)) {
                if (e.gap_opens == 0 || e.gap_bases < gap_run)
                    return fail(KBO_ERR_PANIC, "usize underflow (format.rs:181-182)");
                --e.gap_opens;
                e.gap_bases -= gap_run;
                break;
            }
        }
        out->push_back(e);
    }
    return KBO_OK;
}

int kbo_run_lengths_gapped(const uint8_t* aln, uint64_t n, uint64_t max_gap_len, kbo_rle* out, uint64_t cap,
                           uint64_t* n_out) {
    if ((!aln && n) || !n_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    std::vector<kbo_rle> v;
    int rc = host_run_lengths(aln, n, max_gap_len, &v);
    if (rc) return rc;
    *n_out = v.size();
    if (v.size() > cap) return fail(KBO_ERR_BUFFER_TOO_SMALL, "rle capacity too small");
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(kbo_rle));
    return KBO_OK;
}

int kbo_relative_to_ref(const uint8_t* ref_seq, const uint8_t* aln, uint64_t n, uint8_t* out) {
    if (n && (!ref_seq || !aln || !out)) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t i = 0;
#if defined(__SSE2__)
    {   // format.rs:270-286, sixteen characters at a time (the scalar loop below took 6 ms per 5 Mbp inside kbo_map)
        const __m128i cM = _mm_set1_epi8('M'), cR = _mm_set1_epi8('R'), cI = _mm_set1_epi8('I');
        const __m128i cX = _mm_set1_epi8('X'), cD = _mm_set1_epi8('D'), cG = _mm_set1_epi8('-');
        for (; i + 16 <= n; i += 16) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(aln + i));
            const __m128i r = _mm_loadu_si128(reinterpret_cast<const __m128i*>(ref_seq + i));
            const __m128i take_ref = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(a, cM), _mm_cmpeq_epi8(a, cR)), _mm_cmpeq_epi8(a, cI));
            const __m128i to_gap = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(a, cX), _mm_cmpeq_epi8(a, cD)), _mm_cmpeq_epi8(a, cG));
            const __m128i keep = _mm_andnot_si128(_mm_or_si128(take_ref, to_gap), a);
            const __m128i v = _mm_or_si128(_mm_or_si128(_mm_and_si128(take_ref, r), _mm_and_si128(to_gap, cG)), keep);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), v);
        }
    }
#endif
    for (; i < n; ++i) {  // format.rs:270-286, as byte masks
        const uint8_t a = aln[i];
        const uint8_t take_ref = (uint8_t) - (uint8_t)((a == 'M') | (a == 'R') | (a == 'I'));
        const uint8_t to_gap = (uint8_t) - (uint8_t)((a == 'X') | (a == 'D') | (a == '-'));
        out[i] = (uint8_t)((ref_seq[i] & take_ref) | ((uint8_t)'-' & to_gap) | (a & (uint8_t) ~(take_ref | to_gap)));
    }
    return KBO_OK;
}

// Splits a CSR batch into up to `parts` contiguous query ranges of similar base counts.
static std::vector<uint64_t> split_queries(const uint64_t* offsets, uint64_t nq, uint64_t parts) {
    std::vector<uint64_t> cut(1, 0);
    const uint64_t total = offsets[nq] - offsets[0];
    for (uint64_t s = 1; s < parts; ++s) {
        const uint64_t target = offsets[0] + total * s / parts;
        uint64_t q = (uint64_t)(std::lower_bound(offsets, offsets + nq + 1, target) - offsets);
        if (q < cut.back()) q = cut.back();
        if (q > nq) q = nq;
        if (q != cut.back()) cut.push_back(q);
    }
    if (cut.back() != nq) cut.push_back(nq);
    return cut;
}

// ---- kbo::find for host CSR batches: asynchronous jobs ---------------------------------------------------------
// A job is one kbo::find batch in flight: copy-in, K0, K1, K2b (masks), K4 enqueued on the streams of its
// workspaces, nothing synchronised until kbo_job_wait.  When the caller's output buffers are device-visible
// (page-locked host memory from kbo_alloc_pinned / cudaHostAlloc / cudaHostRegister, or device memory) the last K4
// kernel writes the records and the per-query offsets straight into them: no device->host copy, no record count on
// the host in the middle of the call, ONE synchronisation per job.  Pageable output buffers take a staged path
// (records in device memory, copied out when the count is known).
struct kbo_job {
    kbo_index* ix = nullptr;
    std::vector<Workspace*> wss;
    std::vector<RleParams> rle;      // per sub-batch (staged path: to re-run the last kernel with a larger buffer)
    std::vector<uint64_t> cut;       // sub-batch query ranges
    uint64_t nq = 0;
    kbo_rle* rle_out = nullptr;
    uint64_t rle_cap = 0;
    uint64_t* rle_offsets = nullptr;
    bool direct = false;
    bool deferred = false;           // multi-GPU: the last K4 kernel waits for the record base of this device's slice
    bool relay = false;              // direct path, page-locked records buffer: records go through device memory
    uint64_t relay_cap = 0;          //   (relay_records_kernel); capacity of that device buffer
    uint64_t staged_cap = 0;         // staged path: records the device buffer of wss[0] holds
    int rc = KBO_OK;
    std::string err;
};

// device-visible address of `p` if the device can write to it (page-locked / registered host, device, managed), else null
static void* device_visible(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)
        return attr.devicePointer;
    return nullptr;
}

static void job_release(kbo_job* job) {
    for (Workspace* w : job->wss) if (w) release_ws(job->ix, w);
    job->wss.clear();
}

static int find_submit(kbo_index* ix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                       double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                       uint64_t* rle_offsets, uint64_t want_parts, kbo_job** out, bool defer_finish = false) {
    *out = nullptr;
    if (!rle_offsets || !concat) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(ix, offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    const uint32_t gap = (uint32_t)std::min<uint64_t>(max_gap_len, 0x7fffffffull);
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    want_parts = std::max<uint64_t>(want_parts, min_parts_for(total, n_queries));
    kbo_job* job = new kbo_job();
    job->ix = ix;
    job->nq = n_queries;
    job->rle_out = rle_out;
    job->rle_cap = rle_cap;
    job->rle_offsets = rle_offsets;
    job->cut = split_queries(offsets, n_queries, want_parts);
    const size_t np = job->cut.size() - 1;
    job->wss.assign(np, nullptr);
    job->rle.resize(np);
    RleRecord* dv_out = reinterpret_cast<RleRecord*>(device_visible(rle_out));
    uint64_t* dv_off = reinterpret_cast<uint64_t*>(device_visible(rle_offsets));
    job->direct = dv_off && (dv_out || rle_cap == 0);
    job->deferred = defer_finish;
    if (defer_finish && !job->direct) { delete job; return fail(KBO_ERR_BAD_ARGUMENT, "deferred jobs need device-visible outputs"); }
    auto body = [&]() -> int {
        for (size_t s = 0; s < np; ++s) {
            int r = acquire_ws(ix, &job->wss[s]);
            if (r) return r;
            r = reserve_ws(job->wss[s], total, n_queries, true);
            if (r) return r;
        }
        Workspace* ws0 = job->wss[0];
        // running record totals of the sub-batches (device) and the final count (page-locked host), in ws0
        CUDA_TRY(ws0->counters2.ensure((np + 1) * 8, ws0->stream));
        CUDA_TRY(ws0->h_count.ensure(8 * (np + 2)));  // [0] final count; [1] deferred base; [2..] deferred part totals
        uint64_t* d_totals = ws0->counters2.as<uint64_t>();
        CUDA_TRY(cudaMemsetAsync(d_totals, 0, 8, ws0->stream));
        uint64_t* st_off = nullptr;     // staged path: device copies of the outputs
        RleRecord* st_out = nullptr;
        RleRecord* rec_out = dv_out;    // where the records kernel writes
        uint64_t rec_cap = rle_cap;
        if (job->direct && !defer_finish && dv_out && rle_cap) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, rle_out) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
                job->relay = true;
                job->relay_cap = std::min<uint64_t>(rle_cap, 4 * n_queries + total / 64 + 1024);
                CUDA_TRY(ws0->out2.ensure(job->relay_cap * sizeof(RleRecord), ws0->stream));
                rec_out = ws0->out2.as<RleRecord>();
                rec_cap = job->relay_cap;
            }
            cudaGetLastError();
        }
        if (!job->direct) {
            job->staged_cap = std::min<uint64_t>(rle_cap, 4 * n_queries + total / 64 + 1024);
            CUDA_TRY(ws0->out2.ensure(std::max<uint64_t>(job->staged_cap, 1) * sizeof(RleRecord), ws0->stream));
            CUDA_TRY(ws0->tmp64.ensure((n_queries + 1) * 8, ws0->stream));
            CUDA_TRY(ws0->h_roff.ensure((n_queries + 1) * 8));
            st_off = ws0->tmp64.as<uint64_t>();
            st_out = ws0->out2.as<RleRecord>();
        }
        for (size_t s = 0; s < np; ++s) {
            Workspace* ws = job->wss[s];
            cudaStream_t st = ws->stream;
            const uint64_t q0 = job->cut[s], q1 = job->cut[s + 1], nq = q1 - q0;
            const uint64_t bytes = offsets[q1] - offsets[q0];
            const Geometry g = batch_geometry(ix, bytes, nq);
            // the kernels subtract offsets[0] themselves, so the caller's offsets are copied as they are
            CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, concat + offsets[q0], bytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, offsets + q0, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
            if (s == 0) CUDA_TRY(cudaEventRecord(ws->ev0, st));
            QueryView qv;
            // (K0 addresses the text as base + offsets[i]: bias the base instead of rebasing the offsets on the host)
            int r = matches_device(ix, ws, ws->ascii.as<uint8_t>() - offsets[q0], ws->offsets.as<uint64_t>(), nq, g, thr,
                                   nullptr, 0, true, &qv);
            if (r) return r;
            r = run_rle_counts(ws, qv, g, ws->offsets.as<uint64_t>(), nq, gap, &job->rle[s]);
            if (r) return r;
            if (defer_finish) {  // only this part's record total goes to the host for now
                CUDA_TRY(cudaMemcpyAsync(ws0->h_count.as<uint64_t>() + 2 + s, job->rle[s].cse_blk + job->rle[s].n_blocks, 8,
                                         cudaMemcpyDeviceToHost, st));
                continue;
            }
            if (s > 0) CUDA_TRY(cudaStreamWaitEvent(st, job->wss[s - 1]->ev1, 0));  // its total is this part's base
            else if (np > 1) { /* d_totals[0] was zeroed on this very stream */ }
            r = run_rle_finish(st, job->rle[s], (job->direct ? dv_off : st_off) + q0, job->direct ? rec_out : st_out,
                               job->direct ? rec_cap : job->staged_cap, d_totals + s, d_totals + s + 1, s == 0);
            if (r) return r;
            if (s + 1 == np && job->relay) {  // (the parts before this one have finished: their events were waited for)
                relay_records_kernel<<<64, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(rec_out),
                                                         reinterpret_cast<uint64_t*>(dv_out), d_totals + np, 0, job->relay_cap);
                LAUNCHED();
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaEventRecord(ws->ev1, st));
            if (s + 1 == np) {
                CUDA_TRY(cudaMemcpyAsync(ws0->h_count.p, d_totals + np, 8, cudaMemcpyDeviceToHost, st));
                if (!job->direct)
                    CUDA_TRY(cudaMemcpyAsync(ws0->h_roff.p, st_off, (n_queries + 1) * 8, cudaMemcpyDeviceToHost, st));
            }
        }
        return KBO_OK;
    };
    rc = body();
    if (rc) {
        for (Workspace* w : job->wss) if (w) cudaStreamSynchronize(w->stream);
        job_release(job);
        delete job;
        return rc;
    }
    *out = job;
    return KBO_OK;
}

static int find_wait(kbo_job* job, uint64_t* n_rle_out) {
    kbo_index* ix = job->ix;
    DeviceGuard dg(ix->device);
    int rc = KBO_OK;
    const size_t np = job->wss.size();
    for (size_t s = 0; s < np; ++s) {
        cudaError_t e = cudaStreamSynchronize(job->wss[s]->stream);
        if (e != cudaSuccess && !rc) rc = fail(KBO_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    }
    uint64_t count = 0;
    if (!rc) {
        Workspace* ws0 = job->wss[0];
        count = *ws0->h_count.as<uint64_t>();
        if (!job->direct) {  // staged path: pageable caller buffers
            std::memcpy(job->rle_offsets, ws0->h_roff.p, (job->nq + 1) * 8);
            if (count <= job->rle_cap && count) {
                auto body = [&]() -> int {
                    if (count > job->staged_cap) {  // the estimate was too small: re-run the last kernels into a larger buffer
                        CUDA_TRY(ws0->out2.ensure(count * sizeof(RleRecord), ws0->stream));
                        for (size_t s = 0; s < np; ++s) {
                            int r = run_rle_finish(ws0->stream, job->rle[s], ws0->tmp64.as<uint64_t>() + job->cut[s],
                                                   ws0->out2.as<RleRecord>(), count, ws0->counters2.as<uint64_t>() + s,
                                                   nullptr, s == 0);
                            if (r) return r;
                        }
                    }
                    CUDA_TRY(cudaMemcpyAsync(job->rle_out, ws0->out2.p, count * sizeof(RleRecord), cudaMemcpyDeviceToHost,
                                             ws0->stream));
                    CUDA_TRY(cudaStreamSynchronize(ws0->stream));
                    return KBO_OK;
                };
                rc = body();
            }
        }
        if (!rc && job->relay && count > job->relay_cap && count <= job->rle_cap) {
            // more records than the device-side estimate: write them again, straight into the caller's buffer
            auto body = [&]() -> int {
                RleRecord* dv_out = reinterpret_cast<RleRecord*>(device_visible(job->rle_out));
                uint64_t* dv_off = reinterpret_cast<uint64_t*>(device_visible(job->rle_offsets));
                for (size_t s = 0; s < np; ++s) {
                    int r = run_rle_finish(ws0->stream, job->rle[s], dv_off + job->cut[s], dv_out, job->rle_cap,
                                           ws0->counters2.as<uint64_t>() + s, nullptr, s == 0);
                    if (r) return r;
                }
                CUDA_TRY(cudaStreamSynchronize(ws0->stream));
                return KBO_OK;
            };
            rc = body();
        }
        if (!rc && count > job->rle_cap)  // rle_offsets[nq] holds the true count on both paths
            rc = fail(KBO_ERR_BUFFER_TOO_SMALL, "rle capacity too small");
    }
    if (n_rle_out) *n_rle_out = count;
    if (!rc && np == 1) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, job->wss[0]->ev0, job->wss[0]->ev1);
        ix->last_kernel_ms = ms;
        rc = fetch_counters(ix, job->wss[0]);
    }
    job_release(job);
    delete job;
    return rc;
}

// ---- deferred finish (multi-GPU gather): records of this job are placed after `base` records of other devices -----
static int find_deferred_count(kbo_job* job, uint64_t* count) {
    DeviceGuard dg(job->ix->device);
    uint64_t c = 0;
    for (size_t s = 0; s < job->wss.size(); ++s) {
        CUDA_TRY(cudaStreamSynchronize(job->wss[s]->stream));
        c += (uint32_t)job->wss[0]->h_count.as<uint64_t>()[2 + s];  // low half: STARTs = records of the part
    }
    *count = c;
    return KBO_OK;
}
static int find_deferred_finish(kbo_job* job, uint64_t base, bool first_slice) {
    DeviceGuard dg(job->ix->device);
    const size_t np = job->wss.size();
    Workspace* ws0 = job->wss[0];
    uint64_t* d_totals = ws0->counters2.as<uint64_t>();
    RleRecord* dv_out = reinterpret_cast<RleRecord*>(device_visible(job->rle_out));
    uint64_t* dv_off = reinterpret_cast<uint64_t*>(device_visible(job->rle_offsets));
    ws0->h_count.as<uint64_t>()[1] = base;
    CUDA_TRY(cudaMemcpyAsync(d_totals, ws0->h_count.as<uint64_t>() + 1, 8, cudaMemcpyHostToDevice, ws0->stream));
    // the record count of this device is known (find_deferred_count): records go through device memory and are copied
    // densely into the caller's buffer at their final place (see relay_records_kernel)
    uint64_t mine = 0;
    for (size_t s = 0; s < np; ++s) mine += (uint32_t)ws0->h_count.as<uint64_t>()[2 + s];
    const uint64_t room = base < job->rle_cap ? job->rle_cap - base : 0;
    const uint64_t keep = std::min<uint64_t>(mine, room);
    RleRecord* rec_out = dv_out;
    uint64_t rec_cap = job->rle_cap;
    if (keep && dv_out) {
        CUDA_TRY(ws0->out2.ensure(keep * sizeof(RleRecord), ws0->stream));
        rec_out = ws0->out2.as<RleRecord>() - base;  // the kernel indexes records by their global slot
        rec_cap = base + keep;
    }
    for (size_t s = 0; s < np; ++s) {
        cudaStream_t st = job->wss[s]->stream;
        if (s > 0) CUDA_TRY(cudaStreamWaitEvent(st, job->wss[s - 1]->ev1, 0));
        int r = run_rle_finish(st, job->rle[s], dv_off + job->cut[s], rec_out, rec_cap, d_totals + s, d_totals + s + 1,
                               first_slice && s == 0);
        if (r) return r;
        if (s + 1 == np && keep && dv_out) {
            relay_records_kernel<<<64, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(ws0->out2.p),
                                                     reinterpret_cast<uint64_t*>(dv_out + base), d_totals + np, base, keep);
            LAUNCHED();
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(job->wss[s]->ev1, st));
        if (s + 1 == np) CUDA_TRY(cudaMemcpyAsync(ws0->h_count.p, d_totals + np, 8, cudaMemcpyDeviceToHost, st));
    }
    return KBO_OK;
}

int kbo_find_batch_submit(const kbo_index* cix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                          double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                          uint64_t* rle_offsets, kbo_job** job) {
    if (!job) return fail(KBO_ERR_BAD_ARGUMENT, "job is null");
    return find_submit(const_cast<kbo_index*>(cix), concat, offsets, n_queries, max_error_prob, max_gap_len, rle_out,
                       rle_cap, rle_offsets, 1, job);
}

int kbo_job_wait(kbo_job* job, uint64_t* n_rle) {
    if (!job) return fail(KBO_ERR_BAD_ARGUMENT, "job is null");
    return find_wait(job, n_rle);
}

// kbo::find for a host CSR batch, synchronous.  A lone caller's large batch is cut into sub-batches that run on their
// own streams, so the host->device copy of sub-batch i+1 overlaps the kernels of sub-batch i.
int kbo_find_batch(const kbo_index* cix, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                   double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                   uint64_t* rle_offsets) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix) return fail(KBO_ERR_BAD_ARGUMENT, "index is null");
    uint64_t want_parts = 1;
    kbo_job* job = nullptr;
    int rc;
    {
        const HostCallScope scope(ix);
        uint64_t total = 0;
        if (offsets && n_queries) total = offsets[n_queries] - offsets[0];
        want_parts = pick_parts(scope, total);
        rc = find_submit(ix, concat, offsets, n_queries, max_error_prob, max_gap_len, rle_out, rle_cap, rle_offsets,
                         want_parts, &job);
        if (rc) return rc;
        rc = find_wait(job, nullptr);
    }
    return rc;
}

// ---- multi-GPU: one process, one worker thread and one index replica per device ----------------
// The path shards by independent units: the query batch is cut into contiguous ranges of equal base counts, one per
// device; every device runs the hot path on its range; the results are gathered into the caller's buffers -- alignment
// characters at their final positions straight away, RLE records after the record counts of all devices are known
// (one extra synchronisation per call), written by every device directly into the caller's page-locked buffer.
// No collective is needed: the gather is the devices' own writes over PCIe.
struct kbo_ctx {
    std::vector<int> devices;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::function<int()>> tasks;  // one slot per worker
    std::vector<int> results;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::mutex call_mu;  // one multi-device call at a time per context

    void worker(size_t w) {
        cudaSetDevice(devices[w]);
        uint64_t seen = 0;
        for (;;) {
            std::function<int()> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                fn = tasks[w];
            }
            const int rc = fn ? fn() : KBO_OK;
            {
                std::lock_guard<std::mutex> lk(mu);
                results[w] = rc;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs fn(w) on every worker, returns the first non-zero status
    int run_all(const std::function<int(size_t)>& fn, std::string* err) {
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t w = 0; w < workers.size(); ++w)
                tasks[w] = [&fn, w, err]() {
                    const int rc = fn(w);
                    if (rc && err) { static std::mutex emu; std::lock_guard<std::mutex> g(emu); if (err->empty()) *err = g_err; }
                    return rc;
                };
            pending = (int)workers.size();
            ++generation;
        }
        cv_work.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (int rc : results) if (rc) return rc;
        return KBO_OK;
    }
};

struct kbo_index_set {
    kbo_ctx* ctx = nullptr;
    std::vector<kbo_index*> replicas;
};

int kbo_ctx_create(int n_gpus, const int* devices, kbo_ctx** out) {
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(KBO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n_gpus <= 0) n_gpus = ndev;
    kbo_ctx* ctx = new kbo_ctx();
    for (int i = 0; i < n_gpus; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= ndev) { delete ctx; return fail(KBO_ERR_BAD_ARGUMENT, "device ordinal out of range"); }
        ctx->devices.push_back(d);
    }
    ctx->tasks.resize(ctx->devices.size());
    ctx->results.assign(ctx->devices.size(), 0);
    for (size_t w = 0; w < ctx->devices.size(); ++w) ctx->workers.emplace_back([ctx, w] { ctx->worker(w); });
    *out = ctx;
    return KBO_OK;
}

void kbo_ctx_free(kbo_ctx* ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->stop = true;
    }
    ctx->cv_work.notify_all();
    for (std::thread& t : ctx->workers) t.join();
    delete ctx;
}

int kbo_ctx_n_gpus(const kbo_ctx* ctx) { return ctx ? (int)ctx->devices.size() : 0; }

int kbo_index_set_build(kbo_ctx* ctx, const uint8_t* const* seqs, const uint64_t* lens, uint64_t n_seqs,
                        const kbo_build_opts* opts, kbo_index_set** out) {
    if (!ctx || !out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    *out = nullptr;
    kbo_index_set* set = new kbo_index_set();
    set->ctx = ctx;
    set->replicas.assign(ctx->devices.size(), nullptr);
    std::string err;
    std::lock_guard<std::mutex> call(ctx->call_mu);
    const int rc = ctx->run_all([&](size_t w) {  // construction is deterministic: every device builds its own replica
        return kbo_index_build(seqs, lens, n_seqs, opts, ctx->devices[w], &set->replicas[w]);
    }, &err);
    if (rc) {
        for (kbo_index* ix : set->replicas) kbo_index_free(ix);
        delete set;
        return fail(rc, err);
    }
    *out = set;
    return KBO_OK;
}

void kbo_index_set_free(kbo_index_set* set) {
    if (!set) return;
    for (kbo_index* ix : set->replicas) kbo_index_free(ix);
    delete set;
}

const kbo_index* kbo_index_set_get(const kbo_index_set* set, int i) {
    return (set && i >= 0 && (size_t)i < set->replicas.size()) ? set->replicas[(size_t)i] : nullptr;
}

int kbo_matches_batch_multi(const kbo_index_set* set, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                            double max_error_prob, uint8_t* chars_out) {
    if (!set || !concat || !chars_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(set->replicas[0], offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    kbo_ctx* ctx = set->ctx;
    const std::vector<uint64_t> cut = split_queries(offsets, n_queries, ctx->devices.size());
    std::string err;
    std::lock_guard<std::mutex> call(ctx->call_mu);
    rc = ctx->run_all([&](size_t w) {  // outputs are indexed like concat: every device writes its own range in place
        if (w + 1 >= cut.size()) return (int)KBO_OK;
        return kbo_matches_batch(set->replicas[w], concat, offsets + cut[w], cut[w + 1] - cut[w], max_error_prob, chars_out);
    }, &err);
    return rc ? fail(rc, err) : KBO_OK;
}

int kbo_find_batch_multi(const kbo_index_set* set, const uint8_t* concat, const uint64_t* offsets, uint64_t n_queries,
                         double max_error_prob, uint64_t max_gap_len, kbo_rle* rle_out, uint64_t rle_cap,
                         uint64_t* rle_offsets) {
    if (!set || !concat || !rle_offsets) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(set->replicas[0], offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    kbo_ctx* ctx = set->ctx;
    // the devices write into the caller's buffers: page-lock them for the call if they are pageable
    bool reg_out = false, reg_off = false;
    if (rle_out && rle_cap && !device_visible(rle_out))
        reg_out = cudaHostRegister(rle_out, rle_cap * sizeof(kbo_rle), cudaHostRegisterPortable) == cudaSuccess;
    if (!device_visible(rle_offsets))
        reg_off = cudaHostRegister(rle_offsets, (n_queries + 1) * 8, cudaHostRegisterPortable) == cudaSuccess;
    cudaGetLastError();
    const std::vector<uint64_t> cut = split_queries(offsets, n_queries, ctx->devices.size());
    const size_t ns = cut.size() - 1;
    std::vector<kbo_job*> jobs(ns, nullptr);
    std::vector<uint64_t> counts(ns, 0), bases(ns + 1, 0);
    std::string err;
    {
        std::lock_guard<std::mutex> call(ctx->call_mu);
        rc = ctx->run_all([&](size_t w) {  // phase 1: everything up to the record counts
            if (w >= ns) return (int)KBO_OK;
            int r = find_submit(set->replicas[w], concat, offsets + cut[w], cut[w + 1] - cut[w], max_error_prob, max_gap_len,
                                rle_out, rle_cap, rle_offsets + cut[w], 1, &jobs[w], true);
            if (r) return r;
            return find_deferred_count(jobs[w], &counts[w]);
        }, &err);
        for (size_t w = 0; w < ns; ++w) bases[w + 1] = bases[w] + counts[w];
        const int rc1 = rc;
        const int rc2 = ctx->run_all([&](size_t w) {  // phase 2: records and offsets at their final places
            if (w >= ns || !jobs[w]) return (int)KBO_OK;
            int r = rc1 ? rc1 : find_deferred_finish(jobs[w], bases[w], w == 0);
            uint64_t n = 0;
            const int rw = find_wait(jobs[w], &n);  // (always: it releases the job)
            jobs[w] = nullptr;
            if (r) return r;
            return rw == KBO_ERR_BUFFER_TOO_SMALL ? (int)KBO_OK : rw;  // capacity is judged on the grand total below
        }, &err);
        if (!rc) rc = rc2;
    }
    if (reg_out) cudaHostUnregister(rle_out);
    if (reg_off) cudaHostUnregister(rle_offsets);
    if (rc) return fail(rc, err);
    if (bases[ns] > rle_cap) { rle_offsets[n_queries] = bases[ns]; return fail(KBO_ERR_BUFFER_TOO_SMALL, "rle capacity too small"); }
    return KBO_OK;
}

int kbo_find_batch_device(const kbo_index* cix, const uint8_t* d_concat, const uint64_t* d_offsets,
                          const uint64_t* host_offsets, uint64_t n_queries, double max_error_prob,
                          uint64_t max_gap_len, kbo_rle* d_rle_out, uint64_t rle_cap, uint64_t* d_rle_offsets,
                          void* stream) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!d_concat || !d_offsets || !d_rle_out || !d_rle_offsets) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(ix, host_offsets, n_queries, max_error_prob, &total, &thr);
    if (rc) return rc;
    if (host_offsets[0] != 0) return fail(KBO_ERR_BAD_ARGUMENT, "device batches must have offsets[0] == 0");
    const uint32_t gap = (uint32_t)std::min<uint64_t>(max_gap_len, 0x7fffffffull);
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    rc = check_launch_size(total, n_queries);
    if (rc) return rc;
    Workspace* ws = nullptr;
    rc = stream_ws(ix, (cudaStream_t)stream, &ws);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ws->mu);
    const bool instrumented = g_profile_counters.load() || g_kernel_timing.load();
    const Geometry g = batch_geometry(ix, total, n_queries, instrumented ? 1 : expected_overlap(ix));
    QueryView qv;
    rc = matches_device(ix, ws, d_concat, d_offsets, n_queries, g, thr, nullptr, 0, true, &qv);
    if (rc) return rc;
    RleParams rp;
    rc = run_rle_counts(ws, qv, g, d_offsets, n_queries, gap, &rp);
    if (rc) return rc;
    rc = run_rle_finish(ws->stream, rp, d_rle_offsets, reinterpret_cast<RleRecord*>(d_rle_out), rle_cap);
    if (rc) return rc;
    if (g_profile_counters.load()) return fetch_counters(ix, ws);
    return KBO_OK;
}

// kbo::map without refinement (lib.rs:726-738, 756-760): K0 -> K1 -> K2b on the device, and for `format` also
// format::relative_to_ref (format.rs:266-287) before the result leaves the device (relative_to_ref_kernel: the
// reference bases are already there, staged for K0).
int kbo_map_unrefined(const kbo_index* cix, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
                      int format, uint8_t* out) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!out) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    if (!ref_seq && len) return fail(KBO_ERR_BAD_ARGUMENT, "ref_seq is null");
    const uint64_t offsets[2] = {0, len};
    uint64_t total = 0;
    uint32_t thr = 0;
    int rc = matches_prologue(ix, offsets, 1, max_error_prob, &total, &thr);
    if (rc) return rc;
    if (!format || total + 1 >= KBO_MAX_LAUNCH_POSITIONS / 2) {  // (very long sequences: the pipelined host path)
        rc = kbo_matches(cix, ref_seq, len, max_error_prob, out);
        if (rc || !format) return rc;
        return kbo_relative_to_ref(ref_seq, out, len, out);
    }
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    Workspace* ws = nullptr;
    rc = acquire_ws(ix, &ws);
    if (rc) return rc;
    cudaStream_t st = ws->stream;
    const Geometry g = batch_geometry(ix, len, 1);
    auto body = [&]() -> int {
        CUDA_TRY(ws->ascii.ensure(len, st));
        CUDA_TRY(ws->offsets.ensure(16, st));
        CUDA_TRY(ws->out.ensure(len + 16, st));
        CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, ref_seq, len, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, offsets, 16, cudaMemcpyHostToDevice, st));
        int rc2 = matches_device(ix, ws, ws->ascii.as<uint8_t>(), ws->offsets.as<uint64_t>(), 1, g, thr, ws->out.as<uint8_t>(), 0);
        if (rc2) return rc2;
        relative_to_ref_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(ws->ascii.as<uint8_t>(), ws->out.as<uint8_t>(),
                                                                              len, ws->out.as<uint8_t>());
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out, ws->out.p, len, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return KBO_OK;
    };
    rc = body();
    release_ws(ix, ws);
    return rc;
}

// ---- call / map with refinement (lib.rs:547-573, 720-761) ---------------------------------------
// index::query_sbwt output of ONE query on the host, in page-locked buffers of the workspace that computed it (copied
// at PCIe rate; into pageable vectors the 9 bytes per base took 15 of kbo_map's 25 ms).  The workspace stays
// acquired until release().  `cands` (optional): the candidate scan of call_variants, done on the device.
struct HostMs {
    kbo_index* ix = nullptr;
    Workspace* ws = nullptr;
    const uint8_t* d = nullptr;
    const uint8_t* chars = nullptr;
    const uint32_t* l = nullptr;
    const uint32_t* r = nullptr;
    void release() { if (ws) release_ws(ix, ws); ws = nullptr; }
    ~HostMs() { release(); }
};

static int scan_candidates(kbo_index* ix, Workspace* ws, uint64_t len, uint32_t thr, std::vector<VariantCandidate64>* out);

// One query through K0 -> K1 (with intervals) [-> K2 when thr != 0] [-> candidate scan]; results copied to the host.
static int device_fill_gaps(kbo_index* ix, Workspace* ws, uint64_t len, uint32_t thr, double max_err_prob);

// fill_on_device (with the error probability in fill_p): fill_gaps runs on the device on the characters K2 wrote, and
// (d, l, r) stay there -- only the refined characters (and the candidates) come back.
static int run_single_full(kbo_index* ix, const uint8_t* seq, uint64_t len, uint32_t thr, HostMs* out,
                           uint32_t cand_thr = 0, std::vector<VariantCandidate64>* cands = nullptr,
                           bool fill_on_device = false, double fill_p = 0.0) {
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    Workspace* ws = nullptr;
    int rc = acquire_ws(ix, &ws);
    if (rc) return rc;
    out->ix = ix;
    out->ws = ws;
    const Geometry g = batch_geometry(ix, len, 1);
    cudaStream_t st = ws->stream;
    const uint64_t offsets[2] = {0, len};
    auto body = [&]() -> int {
        CUDA_TRY(ws->ascii.ensure(len, st));
        CUDA_TRY(ws->offsets.ensure(16, st));
        CUDA_TRY(ws->out.ensure(len + 16, st));
        if (!fill_on_device) {
            CUDA_TRY(ws->h_d.ensure(len));
            CUDA_TRY(ws->h_l.ensure(len * 4));
            CUDA_TRY(ws->h_r.ensure(len * 4));
        }
        if (thr) CUDA_TRY(ws->h_chars.ensure(len));
        CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, seq, len, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, offsets, 16, cudaMemcpyHostToDevice, st));
        QueryView qv;
        int rc2 = run_pack(ws, ws->ascii.as<uint8_t>(), ws->offsets.as<uint64_t>(), 1, g, &qv);
        if (rc2) return rc2;
        rc2 = run_ms(ix, ws, qv, g, true);
        if (rc2) return rc2;
        if (thr) {
            rc2 = run_derand_translate(ix, ws, qv, g, thr, ws->out.as<uint8_t>(), 0);
            if (rc2) return rc2;
            if (!fill_on_device) CUDA_TRY(cudaMemcpyAsync(ws->h_chars.p, ws->out.p, len, cudaMemcpyDeviceToHost, st));
        }
        // a single query has its only separator at position len: padded == unpadded below len
        if (!fill_on_device) {
            CUDA_TRY(cudaMemcpyAsync(ws->h_d.p, ws->ms.p, len, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(ws->h_l.p, ws->l.p, len * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(ws->h_r.p, ws->r.p, len * 4, cudaMemcpyDeviceToHost, st));
        }
        if (cands) {
            rc2 = scan_candidates(ix, ws, len, cand_thr, cands);  // (synchronises the stream)
            if (rc2) return rc2;
        }
        if (fill_on_device) {
            rc2 = device_fill_gaps(ix, ws, len, thr, fill_p);
            if (rc2) return rc2;
            CUDA_TRY(cudaMemcpyAsync(ws->h_chars.p, ws->out.p, len, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        return KBO_OK;
    };
    rc = body();
    out->d = ws->h_d.as<uint8_t>();
    out->l = ws->h_l.as<uint32_t>();
    out->r = ws->h_r.as<uint32_t>();
    out->chars = ws->h_chars.as<uint8_t>();
    if (rc) out->release();
    return rc;
}

// The candidate scan of call_variants (variant_calling.rs:268-272) on the (d, l, r) that K1 has just left in the
// workspace (one query: padded == unpadded below len); candidates come back sorted by position.  Synchronises.
static int scan_candidates(kbo_index* ix, Workspace* ws, uint64_t len, uint32_t thr, std::vector<VariantCandidate64>* out) {
    cudaStream_t st = ws->stream;
    CUDA_TRY(ws->counters2.ensure(8, st));
    CUDA_TRY(ws->h_count.ensure(16));
    uint32_t cap = (uint32_t)std::min<uint64_t>(len / 16 + 4096, 1u << 26);
    for (int attempt = 0; attempt < 2; ++attempt) {
        CUDA_TRY(ws->out2.ensure((uint64_t)cap * sizeof(VariantCandidate), st));
        CUDA_TRY(cudaMemsetAsync(ws->counters2.p, 0, 4, st));
        variant_candidates_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(
            ws->ms.as<uint8_t>(), ws->l.as<uint32_t>(), ws->r.as<uint32_t>(), len, ix->host.k, thr,
            ws->out2.as<VariantCandidate>(), cap, ws->counters2.as<unsigned int>());
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(ws->h_count.p, ws->counters2.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const uint32_t n = *ws->h_count.as<uint32_t>();
        if (n <= cap) {
            std::vector<VariantCandidate> tmp(n);
            if (n) CUDA_TRY(cudaMemcpy(tmp.data(), ws->out2.p, (size_t)n * sizeof(VariantCandidate), cudaMemcpyDeviceToHost));
            std::sort(tmp.begin(), tmp.end(), [](const VariantCandidate& a, const VariantCandidate& b) { return a.i < b.i; });
            out->clear();
            out->reserve(n);
            for (const VariantCandidate& c : tmp) out->push_back(VariantCandidate64{c.i, c.j, c.node});
            return KBO_OK;
        }
        cap = n;  // more candidates than the first guess: once more with room for all of them
    }
    return fail(KBO_ERR_CUDA, "candidate scan did not converge");
}

// gap_filling::fill_gaps (gap_filling.rs:444-526) on the device, in place on the characters in ws->out, from the
// intervals K1 left in ws->l / ws->r and the sequence in ws->ascii (refine.cuh).  Synchronises the stream twice
// (number of gaps, panic word).
static const char* refine_panic_text(unsigned code) {
    switch (code) {
        case RP_BRIDGE_ARGS: return "gap_filling.rs:305-310";
        case RP_RIGHT_ARGS: return "gap_filling.rs:25-27";
        case RP_RIGHT_OOB: return "gap_filling.rs:33 index out of bounds";
        case RP_LEFT_ARGS: return "gap_filling.rs:50-52";
        case RP_LEFT_OOB: return "gap_filling.rs:58 index out of bounds";
        case RP_TRIM_UNDERFLOW: return "gap_filling.rs:335 usize underflow";
        case RP_TRIM_RANGE: return "gap_filling.rs:336 slice out of range";
        case RP_IDX_UNDERFLOW: return "gap_filling.rs:357 usize underflow";
        default: return "gap_filling.rs: panic";
    }
}
static const std::vector<double>& gap_run_terms() {
    static const std::vector<double> terms = []() {
        std::vector<double> t(544);  // (1/4)^m underflows to zero from m = 538 on: the term is -0.0 beyond the table
        for (size_t m = 0; m < t.size(); ++m) t[m] = gap_run_log_term(m);
        return t;
    }();
    return terms;
}
static int device_fill_gaps(kbo_index* ix, Workspace* ws, uint64_t len, uint32_t thr, double max_err_prob) {
    cudaStream_t st = ws->stream;
    if (len == 0) return fail(KBO_ERR_PANIC, "gap_filling.rs:453-454");
    if (len < thr) return fail(KBO_ERR_PANIC, "gap_filling.rs:467 usize underflow");
    if (len <= 2ull * thr + 1) return KBO_OK;  // the scan of gap_filling.rs:458 visits no position
    const uint64_t n_pos = len - 2ull * thr - 1;
    CUDA_TRY(ws->counters2.ensure(32, st));
    CUDA_TRY(ws->h_count.ensure(32));
    uint32_t cap = (uint32_t)std::min<uint64_t>(len / 16 + 4096, 1u << 30), n_gaps = 0;
    for (int attempt = 0;; ++attempt) {
        if (attempt == 2) return fail(KBO_ERR_CUDA, "gap scan did not converge");
        CUDA_TRY(ws->gaps.ensure((uint64_t)cap * sizeof(uint2), st));
        CUDA_TRY(cudaMemsetAsync(ws->counters2.p, 0, 4, st));
        gap_list_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, st>>>(ws->out.as<uint8_t>(), len, thr, ws->gaps.as<uint2>(), cap,
                                                                        ws->counters2.as<unsigned int>());
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(ws->h_count.p, ws->counters2.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_gaps = *ws->h_count.as<uint32_t>();
        if (n_gaps <= cap) break;
        cap = n_gaps;
    }
    if (n_gaps == 0) return KBO_OK;
    const std::vector<double>& terms = gap_run_terms();
    CUDA_TRY(ws->terms.ensure(terms.size() * 8, st));
    CUDA_TRY(cudaMemcpyAsync(ws->terms.p, terms.data(), terms.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ws->arena.ensure(len + (uint64_t)n_gaps * thr + 16, st));  // gaps are disjoint: sum of (thr + length) fits
    unsigned long long init[2] = {0ull, ~0ull};  // arena bytes used, panic word
    CUDA_TRY(cudaMemcpyAsync(ws->counters2.as<uint8_t>() + 8, init, 16, cudaMemcpyHostToDevice, st));
    FillGapsParams fp;
    fp.ix = ix->view;
    fp.nk.keys = ix->d_node_keys;
    fp.nk.len = ix->d_node_len;
    fp.nk.words = ix->node_key_words;
    fp.l = ws->l.as<uint32_t>();
    fp.r = ws->r.as<uint32_t>();
    fp.ref = ws->ascii.as<uint8_t>();
    fp.aln = ws->out.as<uint8_t>();
    fp.n = len;
    fp.thr = thr;
    fp.run_terms = ws->terms.as<double>();
    fp.n_terms = (uint32_t)terms.size();
    fp.log_bound = std::log1p(-max_err_prob);
    fp.gaps = ws->gaps.as<uint2>();
    fp.n_gaps = n_gaps;
    fp.arena = ws->arena.as<uint8_t>();
    fp.arena_used = reinterpret_cast<unsigned long long*>(ws->counters2.as<uint8_t>() + 8);
    fp.panic = reinterpret_cast<unsigned long long*>(ws->counters2.as<uint8_t>() + 16);
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_gaps + 63) / 64, 148ull * 32);
    fill_gaps_kernel<<<blocks, 64, 0, st>>>(fp);
    LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ws->h_count.p, fp.panic, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned long long panic = *ws->h_count.as<unsigned long long>();
    if (panic != ~0ull) return fail(KBO_ERR_PANIC, refine_panic_text((unsigned)(panic & 0xffu)));
    return KBO_OK;
}

// SbwtIndex::access_kmer for the nodes of all variant candidates (variant_calling.rs:276) from the node keys on the device
static int device_access_kmers(kbo_index* ix, const std::vector<VariantCandidate64>& cands, uint32_t k, uint8_t* out) {
    if (cands.empty()) return KBO_OK;
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    Workspace* ws = nullptr;
    int rc = acquire_ws(ix, &ws);
    if (rc) return rc;
    cudaStream_t st = ws->stream;
    const uint64_t nc = cands.size();
    std::vector<uint32_t> nodes(nc);
    for (uint64_t c = 0; c < nc; ++c) nodes[c] = (uint32_t)cands[c].node;
    auto body = [&]() -> int {
        CUDA_TRY(ws->out2.ensure(nc * 4, st));
        CUDA_TRY(ws->out3.ensure(nc * k, st));
        CUDA_TRY(cudaMemcpyAsync(ws->out2.p, nodes.data(), nc * 4, cudaMemcpyHostToDevice, st));
        NodeKeysView nk;
        nk.keys = ix->d_node_keys;
        nk.len = ix->d_node_len;
        nk.words = ix->node_key_words;
        access_kmers_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(nk, k, ws->out2.as<uint32_t>(), nc, ws->out3.as<uint8_t>());
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out, ws->out3.p, nc * k, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return KBO_OK;
    };
    rc = body();
    release_ws(ix, ws);
    return rc;
}

// One query through K0 -> K1 (with intervals) -> the candidate scan on the device: only the candidates come back to
// the host, not 9 bytes per base of (d, l, r)  (kbo::call, variant_calling.rs:266-272).
static int run_single_candidates(kbo_index* ix, const uint8_t* seq, uint64_t len, uint32_t thr,
                                 std::vector<VariantCandidate64>* out) {
    DeviceGuard dg(ix->device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    if (len >= KBO_MAX_LAUNCH_POSITIONS) return fail(KBO_ERR_BATCH_TOO_LARGE, "sequence too long for one launch");
    Workspace* ws = nullptr;
    int rc = acquire_ws(ix, &ws);
    if (rc) return rc;
    const Geometry g = batch_geometry(ix, len, 1);
    cudaStream_t st = ws->stream;
    const uint64_t offsets[2] = {0, len};
    auto body = [&]() -> int {
        CUDA_TRY(ws->ascii.ensure(len, st));
        CUDA_TRY(ws->offsets.ensure(16, st));
        CUDA_TRY(cudaMemcpyAsync(ws->ascii.p, seq, len, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ws->offsets.p, offsets, 16, cudaMemcpyHostToDevice, st));
        QueryView qv;
        int rc2 = run_pack(ws, ws->ascii.as<uint8_t>(), ws->offsets.as<uint64_t>(), 1, g, &qv);
        if (rc2) return rc2;
        rc2 = run_ms(ix, ws, qv, g, true);
        if (rc2) return rc2;
        return scan_candidates(ix, ws, len, thr, out);
    };
    rc = body();
    release_ws(ix, ws);
    return rc;
}

// lib.rs:547-573 given the full-length MS of ref_seq against the assembly index (computed by the caller)
static int call_impl(kbo_index* query_index, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
                     const kbo_build_opts* opts, const std::vector<VariantCandidate64>& cands,
                     std::vector<VariantRec>* variants, kbo_index* given_ref_index = nullptr) {
    kbo_build_opts o;
    if (opts) o = *opts; else { kbo_default_build_opts(&o); o.build_select = 1; }
    uint64_t thr = 0;
    int rc = host_threshold(query_index->host.k, query_index->host.n_kmers, 4, max_error_prob, &thr);  // variant_calling.rs:260
    if (rc) return rc;
    BuildTimer bt;
    kbo_index* ref_index = given_ref_index;
    struct RefOwner {  // the per-call index of ref_seq is freed on every way out; a caller-provided one is left alone
        kbo_index** p;
        bool own;
        ~RefOwner() { if (own && *p) kbo_index_free(*p); }
    } ref_owner{&ref_index, given_ref_index == nullptr};
    const uint8_t* seqs[1] = {ref_seq};
    const uint64_t lens[1] = {len};
    if (given_ref_index) {
        if (given_ref_index->device != query_index->device)
            return fail(KBO_ERR_BAD_ARGUMENT, "ref_index lives on another device than the query index");
    } else if (o.k >= 2 && o.k <= KBO_MAX_K && !g_host_builder.load()) {  // lib.rs:553; only its device arrays are used below
        if (len == 0) return fail(KBO_ERR_EMPTY_INPUT, "no input sequences (index.rs:60)");
        DeviceGuard dg(query_index->device);
        if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
        ref_index = new kbo_index();
        ref_index->device = query_index->device;
        ref_index->host.k = o.k;
        rc = build_index_gpu(ref_index, seqs, lens, 1, o.k, o.add_revcomp != 0, false, true);
        if (rc) return rc;
    } else {
        rc = kbo_index_build(seqs, lens, 1, &o, query_index->device, &ref_index);
        if (rc) return rc;
    }
    bt.lap("call: index of ref_seq");
    if (ref_index->host.k != query_index->host.k) {  // lib.rs:559
        return fail(KBO_ERR_K_MISMATCH, "k of the reference index differs from k of the query index (lib.rs:559)");
    }
    int inner_rc = KBO_OK;
    KmerMsFn kmer_ms = [&](int which, const uint8_t* kmers, uint64_t n_kmers, uint32_t k, uint8_t* d_out) {
        std::vector<uint64_t> off(n_kmers + 1);
        for (uint64_t i = 0; i <= n_kmers; ++i) off[i] = i * k;
        int r2 = kbo_query_sbwt_batch_compact(which == 0 ? query_index : ref_index, kmers, off.data(), n_kmers, d_out,
                                              nullptr, nullptr);
        if (r2 && !inner_rc) inner_rc = r2;
        bt.lap(which == 0 ? "call: (k-mers of ref_seq, access_kmer) + MS vs the assembly index" : "call: MS of the index k-mers vs the index of ref_seq");
    };
    const bool dev_access = query_index->d_node_keys && g_device_refine.load();
    AccessKmersFn access = [&](const std::vector<VariantCandidate64>& cs, uint32_t k, uint8_t* out) {
        const int r2 = device_access_kmers(query_index, cs, k, out);
        if (r2 && !inner_rc) inner_rc = r2;
    };
    if (!dev_access) {
        rc = ensure_host_mirror(query_index);
        if (rc) return rc;
    }
    try {
        *variants = call_variants_from(query_index->host, cands, ref_seq, len, thr, kmer_ms, dev_access ? &access : nullptr);
    } catch (const RefinePanic& p) {
        return fail(KBO_ERR_PANIC, p.what);
    }
    bt.lap("call: resolve_variant over the candidates");
    return inner_rc;
}

static int call_entry(const kbo_index* cix, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                      double max_error_prob, const kbo_build_opts* sbwt_build_opts, uint64_t* pos, uint32_t* qlen,
                      uint32_t* rlen, uint8_t* qchars, uint8_t* rchars, uint64_t cap_variants, uint64_t cap_chars,
                      uint64_t* n_variants) {
    ScopeTimer scope_timer("kbo_call");
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix || !n_variants) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    if (!ref_seq || len == 0) return fail(KBO_ERR_EMPTY_INPUT, "empty reference sequence");
    uint64_t thr = 0;
    int rc = host_threshold(ix->host.k, ix->host.n_kmers, 4, max_error_prob, &thr);  // variant_calling.rs:260
    if (rc) return rc;
    std::vector<VariantCandidate64> cands;
    BuildTimer bt;
    rc = run_single_candidates(ix, ref_seq, len, (uint32_t)std::min<uint64_t>(thr, 255), &cands);  // variant_calling.rs:266-272
    if (rc) return rc;
    bt.lap("call: K0, K1 (d,l,r), candidate scan");
    std::vector<VariantRec> vars;
    rc = call_impl(ix, ref_seq, len, max_error_prob, sbwt_build_opts, cands, &vars, const_cast<kbo_index*>(ref_index));
    if (rc) return rc;
    *n_variants = vars.size();
    if (vars.size() > cap_variants) return fail(KBO_ERR_BUFFER_TOO_SMALL, "variant capacity too small");
    uint64_t qo = 0, ro = 0;
    for (size_t i = 0; i < vars.size(); ++i) {
        const VariantRec& v = vars[i];
        if (qo + v.query_chars.size() > cap_chars || ro + v.ref_chars.size() > cap_chars)
            return fail(KBO_ERR_BUFFER_TOO_SMALL, "variant character capacity too small");
        pos[i] = v.query_pos;
        qlen[i] = (uint32_t)v.query_chars.size();
        rlen[i] = (uint32_t)v.ref_chars.size();
        if (!v.query_chars.empty()) std::memcpy(qchars + qo, v.query_chars.data(), v.query_chars.size());
        if (!v.ref_chars.empty()) std::memcpy(rchars + ro, v.ref_chars.data(), v.ref_chars.size());
        qo += v.query_chars.size();
        ro += v.ref_chars.size();
    }
    return KBO_OK;
}

int kbo_call(const kbo_index* cix, const uint8_t* ref_seq, uint64_t len, double max_error_prob,
             const kbo_build_opts* sbwt_build_opts, uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars,
             uint8_t* rchars, uint64_t cap_variants, uint64_t cap_chars, uint64_t* n_variants) {
    return call_entry(cix, nullptr, ref_seq, len, max_error_prob, sbwt_build_opts, pos, qlen, rlen, qchars, rchars,
                      cap_variants, cap_chars, n_variants);
}
int kbo_call_with_ref(const kbo_index* cix, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                      double max_error_prob, uint64_t* pos, uint32_t* qlen, uint32_t* rlen, uint8_t* qchars,
                      uint8_t* rchars, uint64_t cap_variants, uint64_t cap_chars, uint64_t* n_variants) {
    if (!ref_index) return fail(KBO_ERR_BAD_ARGUMENT, "ref_index is null");
    return call_entry(cix, ref_index, ref_seq, len, max_error_prob, nullptr, pos, qlen, rlen, qchars, rchars, cap_variants,
                      cap_chars, n_variants);
}

static int map_entry(const kbo_index* cix, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                     double max_error_prob, int do_fill_gaps, int do_call_variants, int format,
                     const kbo_build_opts* sbwt_build_opts, uint8_t* out) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix || !out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    kbo_build_opts o;
    if (sbwt_build_opts) o = *sbwt_build_opts; else { kbo_default_build_opts(&o); o.build_select = 1; }
    if (ref_index) o.k = ref_index->host.k;  // (the options the caller built it with)
    if (do_call_variants && ix->host.k != o.k)  // lib.rs:729
        return fail(KBO_ERR_K_MISMATCH, "index k differs from sbwt_build_opts.k (lib.rs:729)");
    const uint64_t offsets[2] = {0, len};
    uint64_t total = 0;
    uint32_t thr = 0;
    if (!ref_seq && len) return fail(KBO_ERR_BAD_ARGUMENT, "ref_seq is null");
    int rc = matches_prologue(ix, offsets, 1, max_error_prob, &total, &thr);  // lib.rs:731-738 preconditions
    if (rc) return rc;
    if (!do_fill_gaps && !do_call_variants)  // nothing needs the intervals or the host: translate (and format) on the device
        return kbo_map_unrefined(cix, ref_seq, len, max_error_prob, format, out);
    BuildTimer bt;  // (KBO_BUILD_TIMING=1: where kbo_map spends its time)
    HostMs ms;
    std::vector<VariantCandidate64> cands;
    uint64_t call_thr = 0;
    if (do_call_variants) {
        rc = host_threshold(ix->host.k, ix->host.n_kmers, 4, max_error_prob, &call_thr);  // variant_calling.rs:260
        if (rc) return rc;
    }
    // fill_gaps (lib.rs:743-744) on the device when the index keeps its node keys there; else on host threads
    const bool dev_fill = do_fill_gaps && ix->d_node_keys && g_device_refine.load();
    rc = run_single_full(ix, ref_seq, len, thr, &ms, (uint32_t)std::min<uint64_t>(call_thr, 255),
                         do_call_variants ? &cands : nullptr, dev_fill, max_error_prob);
    if (rc) return rc;
    bt.lap(dev_fill ? "map: K0, K1 (d,l,r), K2b, candidate scan, fill_gaps (device) + copy-out"
                    : "map: K0, K1 (d,l,r), K2b, candidate scan + copy-out");
    std::vector<uint8_t> aln(ms.chars, ms.chars + len);
    bt.lap("map: copy of the characters");
    MsArrays view;
    view.d = ms.d;
    view.l = ms.l;
    view.r = ms.r;
    view.n = len;
    try {
        if (do_fill_gaps && !dev_fill) {  // gaps are independent: bridged on kbo_set_refine_threads host threads
            rc = ensure_host_mirror(ix);
            if (rc) return rc;
            fill_gaps(&aln, view, ref_seq, len, ix->host, thr, max_error_prob, tuned_refine_threads(ix));
            bt.lap("map: fill_gaps (host)");
        }
        if (do_call_variants) {                                                             // lib.rs:749-751
            std::vector<VariantRec> vars;
            ms.release();  // (the arrays are not needed any more; call_impl takes workspaces of its own)
            rc = call_impl(ix, ref_seq, len, max_error_prob, &o, cands, &vars, const_cast<kbo_index*>(ref_index));
            if (rc) return rc;
            bt.lap("map: call (ref index, k-mer MS, resolve)");
            add_variants(&aln, vars);
            bt.lap("map: add_variants");
        }
    } catch (const RefinePanic& p) {
        return fail(KBO_ERR_PANIC, p.what);
    }
    if (format) {  // lib.rs:756-760
        const int rc2 = kbo_relative_to_ref(ref_seq, aln.data(), len, out);
        bt.lap("map: relative_to_ref");
        return rc2;
    }
    std::memcpy(out, aln.data(), len);
    return KBO_OK;
}
int kbo_map(const kbo_index* cix, const uint8_t* ref_seq, uint64_t len, double max_error_prob, int do_fill_gaps,
            int do_call_variants, int format, const kbo_build_opts* sbwt_build_opts, uint8_t* out) {
    return map_entry(cix, nullptr, ref_seq, len, max_error_prob, do_fill_gaps, do_call_variants, format, sbwt_build_opts, out);
}
int kbo_map_with_ref(const kbo_index* cix, const kbo_index* ref_index, const uint8_t* ref_seq, uint64_t len,
                     double max_error_prob, int do_fill_gaps, int do_call_variants, int format, uint8_t* out) {
    if (!ref_index) return fail(KBO_ERR_BAD_ARGUMENT, "ref_index is null");
    return map_entry(cix, ref_index, ref_seq, len, max_error_prob, do_fill_gaps, do_call_variants, format, nullptr, out);
}

// ---- standalone derandomize / translate ------------------------------------------------
static int pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(KBO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(KBO_ERR_BAD_ARGUMENT, "device ordinal out of range");
    return KBO_OK;
}

int kbo_derandomize_ms_vec(const uint64_t* noisy_ms, uint64_t n, uint64_t k, uint64_t threshold, int64_t* out,
                           int device) {
    if (k == 0) return fail(KBO_ERR_BAD_K, "k must be > 0 (derandomize.rs:274)");
    if (threshold <= 1) return fail(KBO_ERR_BAD_THRESHOLD, "threshold must be > 1 (derandomize.rs:275)");
    if (n <= 2) return fail(KBO_ERR_TOO_SHORT, "len must be > 2 (derandomize.rs:276)");
    if (!noisy_ms || !out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    if (k > 0xffffffffull) return fail(KBO_ERR_BAD_K, "k too large");
    for (uint64_t i = 0; i < n; ++i)
        if (noisy_ms[i] > k) return fail(KBO_ERR_BAD_ARGUMENT, "MS value > k (derandomize.rs:229)");
    int rc = pick_device(device);
    if (rc) return rc;
    DeviceGuard dg(device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t n_tiles = (n + G_TILE - 1) / G_TILE;
    const uint32_t thr = (uint32_t)std::min<uint64_t>(threshold, 0xffffffffull);
    uint64_t* d_ms = nullptr;
    int64_t *d_out = nullptr, *d_tmax = nullptr, *d_min = nullptr;
    uint32_t *d_tpar = nullptr, *d_eps = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_ms); cudaFree(d_out); cudaFree(d_tmax); cudaFree(d_min); cudaFree(d_tpar); cudaFree(d_eps);
    };
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&d_ms, n * 8));
        CUDA_TRY(cudaMalloc((void**)&d_out, n * 8));
        CUDA_TRY(cudaMalloc((void**)&d_tmax, n_tiles * 8));
        CUDA_TRY(cudaMalloc((void**)&d_min, n_tiles * 8));
        CUDA_TRY(cudaMalloc((void**)&d_tpar, n_tiles * 4));
        CUDA_TRY(cudaMalloc((void**)&d_eps, n_tiles * 4));
        CUDA_TRY(cudaMemcpy(d_ms, noisy_ms, n * 8, cudaMemcpyHostToDevice));
        g1_tile_max_kernel<<<(unsigned)n_tiles, G_THREADS>>>(d_ms, n, (uint32_t)k, thr, d_tmax);
        LAUNCHED();
        g2_scan_max_kernel<<<1, 32>>>(d_tmax, n_tiles, d_min);
        LAUNCHED();
        g35_tile_kernel<false><<<(unsigned)n_tiles, G_THREADS>>>(d_ms, n, (uint32_t)k, thr, d_min, d_tpar, nullptr, nullptr);
        LAUNCHED();
        g4_scan_par_kernel<<<1, 32>>>(d_tpar, n_tiles, d_eps);
        LAUNCHED();
        g35_tile_kernel<true><<<(unsigned)n_tiles, G_THREADS>>>(d_ms, n, (uint32_t)k, thr, d_min, nullptr, d_eps, d_out);
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(out, d_out, n * 8, cudaMemcpyDeviceToHost));
        return KBO_OK;
    };
    rc = body();
    cleanup();
    return rc;
}

int kbo_translate_ms_vec(const int64_t* derand_ms, uint64_t n, uint64_t k, uint64_t threshold, uint8_t* chars_out,
                         int device) {
    if (k == 0) return fail(KBO_ERR_BAD_K, "k must be > 0 (translate.rs:268)");
    if (threshold <= 1) return fail(KBO_ERR_BAD_THRESHOLD, "threshold must be > 1 (translate.rs:269)");
    if (n <= 2) return fail(KBO_ERR_TOO_SHORT, "len must be > 2 (translate.rs:270)");
    if (!derand_ms || !chars_out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    int rc = pick_device(device);
    if (rc) return rc;
    DeviceGuard dg(device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    int64_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&d_in, n * 8));
        CUDA_TRY(cudaMalloc((void**)&d_out, n));
        CUDA_TRY(cudaMemcpy(d_in, derand_ms, n * 8, cudaMemcpyHostToDevice));
        const uint32_t kk = (uint32_t)std::min<uint64_t>(k, 0x7fffffffull);
        const uint32_t tt = (uint32_t)std::min<uint64_t>(threshold, 0x7fffffffull);
        translate_i64_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_in, n, kk, tt, d_out);
        LAUNCHED();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(chars_out, d_out, n, cudaMemcpyDeviceToHost));
        return KBO_OK;
    };
    rc = body();
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

// ---- instrumentation ------------------------------------------------------------------
int kbo_set_profile_counters(int enabled) { g_profile_counters = enabled ? 1 : 0; return KBO_OK; }
int kbo_get_ms_counters(const kbo_index* cix, kbo_ms_counters* out) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix || !out) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> g(ix->mu);
    *out = ix->last_counters;
    return KBO_OK;
}
int kbo_set_chunk_len(uint32_t chunk_len) { g_chunk_len = chunk_len; return KBO_OK; }
int kbo_set_refine_threads(uint32_t n) { g_refine_threads = n > 256 ? 256 : n; return KBO_OK; }
int kbo_set_device_refine(int enabled) { g_device_refine = enabled ? 1 : 0; return KBO_OK; }
int kbo_index_set_tuning(kbo_index* ix, int key, int64_t value) {
    if (!ix) return fail(KBO_ERR_BAD_ARGUMENT, "index is null");
    switch (key) {
        case KBO_TUNE_CHUNK_LEN: ix->tune.chunk_len = value; break;
        case KBO_TUNE_PIPELINE_PARTS: ix->tune.pipeline_parts = value > 64 ? 64 : value; break;
        case KBO_TUNE_DEVICE_PARTS: ix->tune.device_parts = value > 16 ? 16 : value; break;
        case KBO_TUNE_MS_FLAGS: ix->tune.ms_flags = value < 0 ? -1 : (value & 0xff); break;
        case KBO_TUNE_REFINE_THREADS: ix->tune.refine_threads = value > 256 ? 256 : value; break;
        default: return fail(KBO_ERR_BAD_ARGUMENT, "unknown tuning key");
    }
    return KBO_OK;
}
int kbo_set_device_parts(uint32_t parts) { g_dev_parts = parts > 16 ? 16 : parts; return KBO_OK; }
int kbo_set_pipeline_parts(uint32_t parts) { g_parts = parts > 64 ? 64 : parts; return KBO_OK; }
int kbo_set_host_builder(int enabled) { g_host_builder = enabled ? 1 : 0; return KBO_OK; }
int kbo_set_prefix_table(int enabled) { g_prefix_table = enabled ? 1 : 0; return KBO_OK; }
int kbo_set_prefix_len(uint32_t len) { g_prefix_len = (int)std::min<uint32_t>(len, PREF_MAX_LEN); return KBO_OK; }
int kbo_set_rank2(int enabled) { g_rank2 = enabled ? 1 : 0; return KBO_OK; }
int kbo_set_l2_persist(int enabled) { g_l2_persist = enabled ? 1 : 0; return KBO_OK; }
int kbo_set_ms_flags(uint32_t flags) {
    g_ms_flags = flags & 0xffu;  // bit 1: K2 instead of K2b; bit 2: fused kernel without rank2 pairs; bit 3: its one-pass form; bit 4: fused K1+K2b kernel; bit 5: K1p
    const uint32_t blk = (flags >> 8) & 0x3ffu;  // bits 8..17: K1 block size (experiment)
    if (blk == 128 || blk == 256) g_ms_block = blk;
    return KBO_OK;
}
uint64_t kbo_kernel_launch_count(void) { return g_launches.load(); }
float kbo_last_kernel_ms(const kbo_index* ix) { return ix ? ix->last_kernel_ms : 0.f; }

int kbo_set_kernel_timing(int enabled) { g_kernel_timing = enabled ? 1 : 0; return KBO_OK; }

int kbo_collect_kernel_times(const kbo_index* cix, void* stream, double sum_ms_out[3], uint64_t* n_calls) {
    kbo_index* ix = const_cast<kbo_index*>(cix);
    if (!ix || !sum_ms_out || !n_calls) return fail(KBO_ERR_BAD_ARGUMENT, "null argument");
    DeviceGuard dg(ix->device);
    Workspace* ws = nullptr;
    int rc = stream_ws(ix, (cudaStream_t)stream, &ws);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ws->mu);
    sum_ms_out[0] = sum_ms_out[1] = sum_ms_out[2] = 0.0;
    for (size_t c = 0; c < ws->timed_calls; ++c) {
        cudaEvent_t* ev = ws->timing.data() + c * 4;
        for (int j = 0; j < 3; ++j) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, ev[j], ev[j + 1]));
            sum_ms_out[j] += ms;
        }
    }
    *n_calls = ws->timed_calls;
    ws->timed_calls = 0;
    return KBO_OK;
}

int kbo_measure_random_sector_rate(int device, uint64_t buffer_bytes, int dependent, double* sectors_per_s) {
    if (!sectors_per_s) return fail(KBO_ERR_BAD_ARGUMENT, "out is null");
    int rc = pick_device(device);
    if (rc) return rc;
    DeviceGuard dg(device);
    if (!dg.ok) return fail(KBO_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t n_sectors64 = buffer_bytes / 32;
    if (n_sectors64 == 0 || n_sectors64 > 0xffffffffull) return fail(KBO_ERR_BAD_ARGUMENT, "buffer size out of range");
    const uint32_t n_sectors = (uint32_t)n_sectors64;
    uint64_t* buf = nullptr;
    unsigned long long* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&buf, (size_t)n_sectors * 32));
        CUDA_TRY(cudaMalloc((void**)&sink, 8));
        CUDA_TRY(cudaMemset(buf, 0x5a, (size_t)n_sectors * 32));
        CUDA_TRY(cudaMemset(sink, 0, 8));
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        // dependent > 1: dependent chains at that many warps per SM (blocks of four warps, rounded up): the latency of a
        // warp-wide load of 32 distinct sectors at a given occupancy = lanes in flight / the rate reported
        const unsigned blocks = dependent > 1 ? 148u * (((unsigned)dependent + 3u) / 4u) : 148u * 8u;  // (else 2048 lanes per SM)
        const unsigned threads = dependent > 1 ? 128u : 256u;
        const uint32_t iters = dependent > 1 ? 2048 : (dependent ? 256 : 1024);
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            CUDA_TRY(cudaEventRecord(e0));
            if (dependent) random_sector_kernel<true><<<blocks, threads>>>(buf, n_sectors, iters, sink);
            else random_sector_kernel<false><<<blocks, threads>>>(buf, n_sectors, iters, sink);
            LAUNCHED();
            CUDA_TRY(cudaEventRecord(e1));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaGetLastError());
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            const double rate = (double)blocks * threads * iters / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;  // first repetition warms the cache
        }
        *sectors_per_s = best;
        return KBO_OK;
    };
    rc = body();
    if (buf) cudaFree(buf);
    if (sink) cudaFree(sink);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

}  // extern "C"
