// ===========================================================================
// kbo_b200/csrc/index_build.cuh -- SBWT + LCS construction on the GPU
// (index::build_sbwt_from_vecs, reference src/index.rs:56-99 -> sbwt builder;
// semantics: SURVEY.md section 8c).  k <= 32: one 64-bit word per k-mer;
// 32 < k <= 64: `unsigned __int128` keys through the same kernels (round 1 sent
// those to the host builder in sbwt_host.cpp, which remains as a cross-check).
//
//   1. K0 packs the input sequences (2 bits/base + "not ACGT" mask; inputs are
//      separated exactly like queries, so no k-mer spans two inputs);
//   2. one thread per position emits the colex-packed k-mer ending there (and its
//      reverse complement), flagged valid iff the k positions are all ACGT;
//   3. compaction, radix sort and unique give the k-mer set R        [CUB];
//   4. k-mers whose (k-1)-prefix is nobody's (k-1)-suffix are found by binary
//      search; their few '$'-padded prefixes (dummy nodes) are made on the host;
//   5. R and the dummies are merged by rank (binary searches) into P;
//   6. LCS from neighbouring keys; one thread per node finds the source group
//      of its incoming edge by binary search and sets the label bit;
//   7. popcounts are prefix-summed over the four rows laid end to end, which
//      directly yields C[c] + rank_c(32b) for the interleaved rank words.
// Sorting / scanning / compaction use CUB (library code); nothing here is on the
// query hot path.
// ===========================================================================
#pragma once
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "sbwt_host.hpp"

namespace kbo_b200 {

// Packed k-mers: 2 bits per base, LAST base in the most significant bits, so that integer order == colexicographic
// order.  One 64-bit word for k <= 32, `unsigned __int128` for 32 < k <= 64 (device arithmetic, CUB radix sort,
// select and unique all take it natively).
typedef unsigned __int128 u128;
template <typename K> struct KmerKey;
template <> struct KmerKey<uint64_t> {
    enum { BITS = 64 };
    static __device__ __forceinline__ uint32_t clz(uint64_t x) { return x ? (uint32_t)__clzll((long long)x) : 64u; }
    static __device__ __forceinline__ uint64_t brev_pairs(uint64_t win) {  // reverse the 2-bit groups
        uint64_t rv = __brevll(win);
        return ((rv & 0x5555555555555555ull) << 1) | ((rv >> 1) & 0x5555555555555555ull);
    }
};
template <> struct KmerKey<u128> {
    enum { BITS = 128 };
    static __device__ __forceinline__ uint32_t clz(u128 x) {
        const uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
        return hi ? (uint32_t)__clzll((long long)hi) : (lo ? 64u + (uint32_t)__clzll((long long)lo) : 128u);
    }
    static __device__ __forceinline__ u128 brev_pairs(u128 win) {
        const uint64_t hi = KmerKey<uint64_t>::brev_pairs((uint64_t)(win >> 64));
        const uint64_t lo = KmerKey<uint64_t>::brev_pairs((uint64_t)win);
        return ((u128)lo << 64) | hi;
    }
};

// k-mer ending at padded position i: bases [i-k+1, i], first base in the lowest bits of the 2k-bit window,
// so (window << (64-2k)) compares colexicographically.
template <typename K>
__global__ void kmer_keys_kernel(const uint64_t* __restrict__ pack, const uint32_t* __restrict__ inv, uint64_t Lp,
                                 uint32_t k, int revcomp, K* __restrict__ keys, K* __restrict__ keys_rc,
                                 uint8_t* __restrict__ flags) {
    constexpr uint32_t BITS = KmerKey<K>::BITS;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lp) return;
    uint8_t ok = 0;
    K key = 0, key_rc = 0;
    if (i + 1 >= k) {
        const uint64_t j0 = i + 1 - k;
        // invalid bits of positions [j0, i]
        const uint64_t w0 = j0 >> 5;
        const uint32_t sh = (uint32_t)(j0 & 31);
        uint64_t bad = (uint64_t)inv[w0] >> sh;
        bad |= (uint64_t)inv[w0 + 1] << (32 - sh);
        if (sh) bad |= (uint64_t)inv[w0 + 2] << (64 - sh);
        const uint64_t kmask = k == 64 ? ~0ull : ((1ull << k) - 1ull);
        if ((bad & kmask) == 0) {
            ok = 1;
            // the 2k-bit window, first base in its lowest bits (a pack word holds 32 bases)
            K win = (K)(pack[w0] >> (2 * sh));
            if (BITS == 64) {
                if (sh) win |= (K)(pack[w0 + 1] << (64 - 2 * sh));
            } else {
                win |= (K)pack[w0 + 1] << (64 - 2 * sh);
                if (sh) win |= (K)pack[w0 + 2] << (128 - 2 * sh);
            }
            const K ones = ~(K)0;
            const K wmask = (2 * k >= BITS) ? ones : (K)(((K)1 << (2 * k)) - 1);
            win &= wmask;
            key = (K)(win << (BITS - 2 * k));
            if (revcomp) {
                // reverse the 2-bit groups of the window and complement them: the reversed window sits in the top 2k bits
                const K rv = (K)~KmerKey<K>::brev_pairs(win);
                key_rc = rv & (K)(ones << (BITS - 2 * k));
            }
        }
    }
    flags[i] = ok;
    keys[i] = key;
    if (revcomp) keys_rc[i] = key_rc;
}

// flag[a] = 1 iff no k-mer of R ends with the first k-1 characters of R[a]
template <typename K>
__global__ void no_predecessor_kernel(const K* __restrict__ R, uint64_t n, uint32_t k, uint8_t* __restrict__ flag) {
    constexpr uint32_t BITS = KmerKey<K>::BITS;
    const uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const K want = (K)(R[a] << 2);
    const K sufmask = (K)(~(K)0 << (BITS - 2 * (k - 1)));
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (R[mid] < want) lo = mid + 1; else hi = mid;
    }
    flag[a] = !(lo < n && (R[lo] & sufmask) == want);
}

// P = R merged with the dummies D (both sorted by (key, len)); node order = (key, len)
template <typename K>
__global__ void merge_nodes_kernel(const K* __restrict__ R, uint64_t nR, const K* __restrict__ Dkey,
                                   const uint8_t* __restrict__ Dlen, uint64_t nD, uint32_t k,
                                   K* __restrict__ Pkey, uint8_t* __restrict__ Plen) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nR) {
        const K key = R[t];
        uint64_t lo = 0, hi = nD;  // dummies with key' <= key all precede this k-mer (their len < k)
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (Dkey[mid] <= key) lo = mid + 1; else hi = mid;
        }
        Pkey[t + lo] = key;
        Plen[t + lo] = (uint8_t)k;
    } else if (t < nR + nD) {
        const uint64_t b = t - nR;
        const K key = Dkey[b];
        uint64_t lo = 0, hi = nR;  // k-mers with key < key' precede this dummy
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (R[mid] < key) lo = mid + 1; else hi = mid;
        }
        Pkey[b + lo] = key;
        Plen[b + lo] = Dlen[b];
    }
}

template <typename K>
__global__ void lcs_kernel(const K* __restrict__ Pkey, const uint8_t* __restrict__ Plen, uint64_t n,
                           uint8_t* __restrict__ lcs) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = 0;
    if (i > 0) {
        const K x = Pkey[i - 1] ^ Pkey[i];
        const uint32_t same = KmerKey<K>::clz(x) >> 1;
        const uint32_t lim = Plen[i - 1] < Plen[i] ? Plen[i - 1] : Plen[i];
        v = same < lim ? same : lim;
    }
    lcs[i] = (uint8_t)v;
}

// node t (not the root) receives its edge from the first node of the group that ends with t's first k-1
// characters; that node gets label (last character of t)
template <typename K>
__global__ void labels_kernel(const K* __restrict__ Pkey, const uint8_t* __restrict__ Plen, uint64_t n,
                              uint32_t k, uint32_t* __restrict__ rows32, uint64_t row_words32) {
    constexpr uint32_t BITS = KmerKey<K>::BITS;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0 || t >= n) return;
    const K key = Pkey[t];
    const uint32_t c = (uint32_t)(key >> (BITS - 2));
    const K ukey = (K)(key << 2);
    const uint32_t ulen = (uint32_t)Plen[t] - 1u;
    uint64_t lo = 0, hi = n;  // first node >= (ukey, ulen)
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const K mk = Pkey[mid];
        if (mk < ukey || (mk == ukey && Plen[mid] < ulen)) lo = mid + 1; else hi = mid;
    }
    atomicOr(rows32 + (uint64_t)c * row_words32 + (lo >> 5), 1u << (lo & 31));
}

__global__ void row_popc_kernel(const uint32_t* __restrict__ rows32, uint64_t total_words, uint32_t* __restrict__ pc) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_words) pc[i] = __popc(rows32[i]);
}

// rank word b of row c = (1 + ones before it in the four rows laid end to end) << 32 | bits
__global__ void compose_rank_kernel(const uint32_t* __restrict__ rows32, const uint32_t* __restrict__ prefix,
                                    uint64_t total_words, uint64_t* __restrict__ rank) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_words) rank[i] = ((uint64_t)(1u + prefix[i]) << 32) | rows32[i];
}

}  // namespace kbo_b200
