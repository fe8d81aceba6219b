// ===========================================================================
// kbo_b200/csrc/host_layout.hpp -- host-side helpers shared by the C ABI
// implementation and the CPU-only kernel-logic tests: the device index layout
// and the batch geometry.  (Plain C++, no CUDA.)
// ===========================================================================
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "sbwt_host.hpp"

namespace kbo_b200 {

// Interleaved rank words + padded LCS (kernels.cuh "IndexView").
struct DeviceLayout {
    std::vector<uint64_t> rank;  // 4 rows x stride words: (C[c] + ones before block) << 32 | 32 row bits
    uint64_t stride = 0;
    std::vector<uint8_t> lcs;    // n bytes + zero padding (>= 9 bytes, 8-byte multiple)
};

inline void build_device_layout(const HostIndex& h, DeviceLayout* out) {
    const uint64_t n = h.n_sets;
    const uint64_t nblk = (n >> 5) + 2;           // block n>>5 must exist for rank(n)
    const uint64_t stride = (nblk + 3) & ~3ull;   // every row starts on a 32-byte sector
    out->stride = stride;
    out->rank.assign((size_t)(4 * stride), 0);
    for (int c = 0; c < 4; ++c) {
        uint64_t run = h.C[c];
        const std::vector<uint64_t>& row = h.rows[c];
        for (uint64_t b = 0; b < nblk; ++b) {
            const uint64_t w = b >> 1;
            uint32_t bits = 0;
            if (w < row.size()) bits = (uint32_t)(row[w] >> ((b & 1) * 32));
            out->rank[(size_t)(c * stride + b)] = (run << 32) | bits;
            run += (uint64_t)__builtin_popcount(bits);
        }
    }
    const uint64_t lcs_bytes = ((n + 8) & ~7ull) + 16;
    out->lcs.assign((size_t)lcs_bytes, 0);
    std::memcpy(out->lcs.data(), h.lcs.data(), (size_t)n);
}

// Sizes of one batch in "padded space" (one separator after every query).
struct Geometry {
    uint64_t total = 0;    // sum of query lengths
    uint64_t Lp = 0;       // padded length = total + n_queries
    uint64_t n_tiles = 0;  // K2 tiles of 512 positions
    uint64_t n_tiles_b = 0;  // K2b tiles of 1024 positions
    uint64_t n_words = 0;  // 32-position words written by K0 (all tiles + 4 slack words)
    uint32_t chunk_len = 0;
    uint64_t n_chunks = 0;
    size_t ms_bytes = 0;   // bytes of the u8 MS array (readable slack past n_words*32)
};

// `overlap` = number of launches of K1 expected to share the machine (1 for a lone call).  Measured on B200
// (profiles/README.md): one launch over a 10^7-base batch is fastest at chunk_len 64 (148 SMs x 1024 lanes wanted:
// a wider chunk leaves too few warps, a narrower one doubles the warm-up work); when independent batches overlap on
// several streams the warps come from the other launches and the longer chunk's smaller warm-up share wins
// (6 streams: 95 G bases/s at 64, 109 G at 192, 80 G at 256).
inline uint32_t auto_chunk_len(uint64_t Lp, uint32_t overlap = 1) {
    uint64_t c = (Lp * (overlap ? overlap : 1) / (148ull * 1024ull)) & ~31ull;
    if (c < 64) c = 64;
    if (c > 512) c = 512;
    return (uint32_t)c;
}

inline Geometry make_geometry(uint64_t total, uint64_t nq, uint32_t forced_chunk_len, uint32_t overlap = 1) {
    Geometry g;
    g.total = total;
    g.Lp = total + nq;
    g.n_tiles = (g.Lp + 511) / 512;
    g.n_tiles_b = (g.Lp + 1023) / 1024;
    g.n_words = g.n_tiles_b * 32 + 4;
    g.chunk_len = forced_chunk_len ? ((forced_chunk_len + 31u) & ~31u) : auto_chunk_len(g.Lp, overlap);
    g.n_chunks = (g.Lp + g.chunk_len - 1) / g.chunk_len;
    g.ms_bytes = (size_t)(g.n_words * 32 + 64);
    return g;
}

}  // namespace kbo_b200
