// ===========================================================================
// tests/emu/host_emu.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal CUDA-on-CPU shim so that kbo_b200/csrc/kernels.cuh can be compiled
// with g++ (-DKBO_HOST_EMU) and its kernel LOGIC checked against the oracle in
// this GPU-less container before GPU minutes are spent.  It is never linked
// into libkbo_b200.so; the product has no CPU path.
//
// Model: one block at a time.  Kernels without warp collectives / barriers run
// their threads sequentially (emu_launch_seq); kernels with them run one host
// thread per CUDA thread, warps synchronised with std::barrier (emu_launch_par).
// ===========================================================================
#pragma once
#include <barrier>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
};
struct uint4 {
    uint32_t x, y, z, w;
};
#include <cstdint>


inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct uint2 {
    uint32_t x, y;
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
#define __align__(n) __attribute__((aligned(n)))

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct EmuWarp {
    std::barrier<> bar{32};
    uint64_t scratch[32];
};
struct EmuBlock {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<EmuWarp>> warps;
};
inline thread_local EmuWarp* emu_warp = nullptr;
inline thread_local EmuBlock* emu_block = nullptr;
inline thread_local int emu_lane = 0;

template <typename T>
inline T emu_exchange(T v, int src) {
    emu_warp->scratch[emu_lane] = (uint64_t)v;
    emu_warp->bar.arrive_and_wait();
    T r = (T)emu_warp->scratch[src & 31];
    emu_warp->bar.arrive_and_wait();
    return r;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int src = emu_lane + (int)d;
    return emu_exchange(v, src < 32 ? src : emu_lane);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int src = emu_lane - (int)d;
    return emu_exchange(v, src >= 0 ? src : emu_lane);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu_warp->scratch[emu_lane] = pred ? 1u : 0u;
    emu_warp->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(emu_warp->scratch[i] & 1u) << i;
    emu_warp->bar.arrive_and_wait();
    return r;
}
inline int __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp->bar.arrive_and_wait(); }
inline void __syncthreads() { emu_block->bar->arrive_and_wait(); }

inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31u));
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

inline unsigned long long& emu_tail_extensions() { static unsigned long long n = 0; return n; }

// ---- launchers --------------------------------------------------------------
template <typename F>
inline void emu_launch_seq(unsigned grid, unsigned block, F&& body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            blockIdx.x = b;
            threadIdx.x = t;
            body();
        }
}

template <typename F>
inline void emu_launch_par(unsigned grid, unsigned block, F&& body) {
    for (unsigned b = 0; b < grid; ++b) {
        EmuBlock blk;
        blk.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        for (unsigned w = 0; w < (block + 31) / 32; ++w) blk.warps.push_back(std::make_unique<EmuWarp>());
        std::vector<std::thread> ths;
        for (unsigned t = 0; t < block; ++t)
            ths.emplace_back([&, t]() {
                gridDim.x = grid;
                blockDim.x = block;
                blockIdx.x = b;
                threadIdx.x = t;
                emu_block = &blk;
                emu_warp = blk.warps[t / 32].get();
                emu_lane = (int)(t % 32);
                body();
                // a thread that leaves early must not block its warp/block mates
                emu_warp->bar.arrive_and_drop();
                blk.bar->arrive_and_drop();
            });
        for (auto& th : ths) th.join();
    }
}
