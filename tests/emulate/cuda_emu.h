// cuda_emu.h -- just enough of the CUDA execution model to run the library's kernel SOURCES on the CPU.
// TEST INFRASTRUCTURE ONLY: nothing under tinyknn_b200/ includes or links this; tests/emulate/emu_build.py translates the
// .cu files (launch syntax, dynamic shared memory) and compiles them against this header into tests/emulate/_build/.
//
// Execution model: blocks run one after the other; the threads of a block are FIBERS (ucontext) of one OS thread that are
// switched only at synchronisation points:
//   * __syncthreads / __syncthreads_or park a fiber until every live thread of the block has arrived;
//   * warp collectives (__shfl*_sync, __ballot_sync, __any_sync, __all_sync, __syncwarp) park a fiber until every live
//     lane of its warp has arrived, then all of them see the same gathered values -- a lane that calls a collective the
//     others never reach is reported as a deadlock instead of hanging;
//   * `__shared__` variables are statics, dynamic shared memory is a per-block buffer with a canary behind it.
// "Device memory" is host memory: the C ABI's *_dev entry points take numpy buffers. This checks indexing, tiling, barrier
// placement, warp-collective protocols and the integer/byte arithmetic of the kernels against the oracle on a machine
// without a GPU. It says nothing about performance, memory-model races between warps, or alignment faults.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>

// ---- vector types ---------------------------------------------------------------------------------------------------
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct uint3 { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline double2 make_double2(double x, double y) { return {x, y}; }

extern uint3 threadIdx, blockIdx;            // set by the scheduler whenever a fiber is resumed
extern dim3 blockDim, gridDim;
constexpr int warpSize = 32;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- runtime of the emulator (cuda_emu.cpp) ---------------------------------------------------------------------------
namespace emu {
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
unsigned char *dyn_smem();                          // the block's dynamic shared memory (16-byte aligned)
int block_barrier(int pred);                        // __syncthreads; returns the OR of pred over the block
const uint64_t *warp_gather(uint64_t v);            // every live lane contributes v; returns the 32 values (exited lanes: 0)
unsigned lane_id();
unsigned live_lane_mask();
struct Stats { long long launches, blocks, fibers, switches, collectives, barriers, collectives_with_exited_lanes; };
Stats &stats();
}  // namespace emu
extern "C" {
const char *emu_last_error(void);                   // deadlock / canary reports of the last launch ("" if none)
void emu_stats(long long *out7);
void emu_reset_stats(void);
}

// ---- synchronisation and warp collectives ---------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(0); }
inline int __syncthreads_or(int pred) { return emu::block_barrier(pred); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_gather(0); }

namespace emu {
template <typename T> inline uint64_t bits_of(T v)
{
    static_assert(sizeof(T) <= 8, "emulated shuffle moves at most 8 bytes");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> inline T from_bits(uint64_t b)
{
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace emu

template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const unsigned lane = emu::lane_id();
    const uint64_t *g = emu::warp_gather(emu::bits_of(v));
    const unsigned base = lane & ~(unsigned)(width - 1);
    return emu::from_bits<T>(g[base + ((unsigned)src & (unsigned)(width - 1))]);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32)
{
    const unsigned lane = emu::lane_id();
    const uint64_t *g = emu::warp_gather(emu::bits_of(v));
    const unsigned other = lane ^ (unsigned)lane_mask;
    if ((other & ~(unsigned)(width - 1)) != (lane & ~(unsigned)(width - 1))) return v;
    return emu::from_bits<T>(g[other & 31]);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
    const unsigned lane = emu::lane_id();
    const uint64_t *g = emu::warp_gather(emu::bits_of(v));
    const unsigned in_seg = lane & (unsigned)(width - 1);
    return in_seg >= delta ? emu::from_bits<T>(g[lane - delta]) : v;
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
    const unsigned lane = emu::lane_id();
    const uint64_t *g = emu::warp_gather(emu::bits_of(v));
    const unsigned in_seg = lane & (unsigned)(width - 1);
    return in_seg + delta < (unsigned)width ? emu::from_bits<T>(g[lane + delta]) : v;
}
inline unsigned __ballot_sync(unsigned, int pred)
{
    const uint64_t *g = emu::warp_gather(pred ? 1 : 0);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(g[i] & 1) << i;
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred)
{
    const unsigned live = emu::live_lane_mask();
    return (__ballot_sync(mask, pred) & live) == live;
}
inline unsigned __activemask() { return emu::live_lane_mask(); }

// ---- atomics (one OS thread: plain read-modify-write) ---------------------------------------------------------------
template <typename T, typename U> inline T atomicAdd(T *p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> inline T atomicMax(T *p, U v) { const T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U> inline T atomicMin(T *p, U v) { const T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <typename T, typename U> inline T atomicExch(T *p, U v) { const T old = *p; *p = (T)v; return old; }
template <typename T, typename U> inline T atomicOr(T *p, U v) { const T old = *p; *p = (T)(old | (T)v); return old; }
template <typename T, typename U, typename V> inline T atomicCAS(T *p, U cmp, V v) { const T old = *p; if (old == (T)cmp) *p = (T)v; return old; }
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---- scalar helpers CUDA puts in the global namespace ---------------------------------------------------------------
template <typename A, typename B> inline std::common_type_t<A, B> min(A a, B b)
{
    using C = std::common_type_t<A, B>;
    return (C)b < (C)a ? (C)b : (C)a;
}
template <typename A, typename B> inline std::common_type_t<A, B> max(A a, B b)
{
    using C = std::common_type_t<A, B>;
    return (C)a < (C)b ? (C)b : (C)a;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}

inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }          // compile with -ffp-contract=off
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline long long __double2ll_rn(double v) { return std::llrint(v); }      // (default rounding mode: to nearest even)
inline int __float2int_rn(float v) { return (int)std::lrintf(v); }
inline float __int_as_float(int v) { return emu::from_bits<float>((uint64_t)(uint32_t)v); }
inline int __float_as_int(float v) { return (int)(uint32_t)emu::bits_of(v); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>((uint64_t)v); }
inline long long __double_as_longlong(double v) { return (long long)emu::bits_of(v); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <typename T> inline T __ldg(const T *p) { return *p; }

// ---- byte / halfword SIMD intrinsics (semantics of the CUDA math API) ----------------------------------------------
namespace emu {
template <typename F> inline uint32_t per_half(uint32_t a, uint32_t b, uint32_t c, F f)
{
    const uint32_t lo = (uint32_t)f(a & 0xffffu, b & 0xffffu, c & 0xffffu) & 0xffffu;
    const uint32_t hi = (uint32_t)f(a >> 16, b >> 16, c >> 16) & 0xffffu;
    return lo | (hi << 16);
}
template <typename F> inline uint32_t per_byte(uint32_t a, uint32_t b, F f)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= ((uint32_t)f((a >> (8 * i)) & 0xffu, (b >> (8 * i)) & 0xffu) & 0xffu) << (8 * i);
    return r;
}
inline int s16(uint32_t h) { return (int)(int16_t)(uint16_t)h; }
inline int s8(uint32_t b) { return (int)(int8_t)(uint8_t)b; }
// prmt.b32 (default mode): result byte i = byte (sel nibble i & 7) of {a: bytes 0-3, b: bytes 4-7}; nibble bit 3 set:
// the byte's sign bit replicated over all 8 bits
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (s >> (4 * i)) & 0xfu;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
        r |= byte << (8 * i);
    }
    return r;
}
}  // namespace emu

inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    // __byte_perm has no sign-replicate mode: selector nibbles use 3 bits
    return emu::prmt(a, b, s & 0x7777u);
}
inline unsigned __vadd2(unsigned a, unsigned b) { return emu::per_half(a, b, 0, [](uint32_t x, uint32_t y, uint32_t) { return x + y; }); }
inline unsigned __vsub2(unsigned a, unsigned b) { return emu::per_half(a, b, 0, [](uint32_t x, uint32_t y, uint32_t) { return x - y; }); }
inline unsigned __vadd4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return x + y; }); }
inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::min(std::min(emu::s16(x), emu::s16(y)), emu::s16(z)); });
}
inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::max(std::max(emu::s16(x), emu::s16(y)), emu::s16(z)); });
}
inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::min(std::min(x, y), z); });
}
inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::max(std::max(x, y), z); });
}
inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c)      // max(a + b, c), the add wraps at 16 bits
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::max(emu::s16((x + y) & 0xffffu), emu::s16(z)); });
}
inline unsigned __viaddmin_s16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::min(emu::s16((x + y) & 0xffffu), emu::s16(z)); });
}
inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::min((x + y) & 0xffffu, z); });
}
inline unsigned __viaddmax_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu::per_half(a, b, c, [](uint32_t x, uint32_t y, uint32_t z) { return std::max((x + y) & 0xffffu, z); });
}
inline int __viaddmin_s32(int a, int b, int c) { const int s = (int)((unsigned)a + (unsigned)b); return s < c ? s : c; }
inline int __viaddmax_s32(int a, int b, int c) { const int s = (int)((unsigned)a + (unsigned)b); return s > c ? s : c; }
inline int __vimin3_s32(int a, int b, int c) { return std::min(std::min(a, b), c); }
inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
inline unsigned __vmins4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return std::min(emu::s8(x), emu::s8(y)); }); }
inline unsigned __vmaxs4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return std::max(emu::s8(x), emu::s8(y)); }); }
inline unsigned __vminu4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return std::min(x, y); }); }
inline unsigned __vmaxu4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return std::max(x, y); }); }
inline unsigned __vcmplts4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return emu::s8(x) < emu::s8(y) ? 0xffu : 0u; }); }
inline unsigned __vcmpltu4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return x < y ? 0xffu : 0u; }); }
inline unsigned __vcmpgts4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return emu::s8(x) > emu::s8(y) ? 0xffu : 0u; }); }
inline unsigned __vcmpeq4(unsigned a, unsigned b) { return emu::per_byte(a, b, [](uint32_t x, uint32_t y) { return x == y ? 0xffu : 0u; }); }

// ---- the slice of the runtime API the library calls -----------------------------------------------------------------
typedef int cudaError_t;
typedef struct EmuStream_ *cudaStream_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorLaunchFailure = 719,
                      cudaErrorNotSupported = 801;
#define cudaStreamPerThread ((cudaStream_t)2)
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
constexpr unsigned cudaIpcMemLazyEnablePeerAccess = 1;
struct cudaIpcMemHandle_t { char reserved[64]; };

cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaMalloc(void **p, size_t bytes);
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t = nullptr);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaDeviceGetAttribute(int *value, cudaDeviceAttr attr, int dev);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
