// cuda_emu.h -- just enough of the CUDA execution model to run a kernel's SOURCE on the CPU (TEST INFRASTRUCTURE ONLY).
//
// One thread block at a time, one OS thread per CUDA thread: threadIdx / blockIdx / blockDim / gridDim are thread-local,
// __syncthreads() is a block-wide barrier, __shfl_xor_sync() exchanges through a per-warp scratch pad (all 32 lanes of a
// warp must call it together, as on the device), `__shared__` variables become statics (blocks run one after the other).
// It checks indexing, tiling, barrier placement and the reduction logic of a kernel against the oracle on a machine without
// a GPU; it says nothing about performance, memory-model races or alignment faults.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct EmuBlock {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint64_t> warp_pad;                 // 32 slots per warp
};
inline EmuBlock *g_emu_block = nullptr;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

inline void __syncthreads() { g_emu_block->block_bar->arrive_and_wait(); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
    static_assert(sizeof(T) <= 8, "emulated shuffle moves at most 8 bytes");
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    g_emu_block->warp_pad[warp * 32 + lane] = bits;
    g_emu_block->warp_bar[warp]->arrive_and_wait();
    const uint64_t other = g_emu_block->warp_pad[warp * 32 + (lane ^ (unsigned)lane_mask)];
    g_emu_block->warp_bar[warp]->arrive_and_wait();
    T out;
    std::memcpy(&out, &other, sizeof(T));
    return out;
}

inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }

// run `body` as a grid of `grid` blocks of `block` threads (1-D), blocks one after the other
inline void emu_launch(unsigned grid, unsigned block, const std::function<void()> &body)
{
    for (unsigned b = 0; b < grid; b++) {
        EmuBlock blk;
        blk.block_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        const unsigned warps = (block + 31) / 32;
        for (unsigned w = 0; w < warps; w++) {
            const unsigned lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
            blk.warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)lanes));
        }
        blk.warp_pad.assign((size_t)warps * 32, 0);
        g_emu_block = &blk;
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; t++)
            threads.emplace_back([&, t] {
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
                body();
            });
        for (auto &th : threads) th.join();
        g_emu_block = nullptr;
    }
}
