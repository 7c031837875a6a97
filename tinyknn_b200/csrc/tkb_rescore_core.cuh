// tkb_rescore_core.cuh -- the exact rescoring distance of one raw row, shared by gather_dists_kernel
// (tkb_rescore.cu) and the fused query kernel (tkb_fused.cu), so that both give the same bits.
//
// Replaces the arithmetic of knn_brute1 (ref: tinyknn/utils.py:89-91): diff = Y - x ; einsum('ij,ij->i', diff, diff),
// in the dtype numpy would use (f32 rows: f32; f64 rows: f64). Summation order differs from numpy's (tolerance 1e-5
// relative, stated in the tests): f32 rows whose length is a multiple of 4 are read as one float4 per lane (lane i sums
// elements 4i..4i+3, then 4i+128.. in order, fma), other rows one element per lane (i, i+32, ...); then an xor-shuffle tree.
#pragma once
#include "tkb_common.cuh"

namespace tkb {

template <typename T>
__device__ __forceinline__ T warp_row_partial(const T *__restrict__ y, const float *__restrict__ x, int d, int lane)
{
    T acc = (T)0;
    for (int i = lane; i < d; i += 32) {
        const T df = y[i] - (T)x[i];
        acc = fma(df, df, acc);
    }
    return acc;
}

template <>
__device__ __forceinline__ float warp_row_partial<float>(const float *__restrict__ y, const float *__restrict__ x, int d, int lane)
{
    float acc = 0.0f;
    if ((d & 3) == 0 && ((((uintptr_t)y) | ((uintptr_t)x)) & 15) == 0) {
        for (int i = 4 * lane; i < d; i += 128) {
            const float4 yv = *reinterpret_cast<const float4 *>(y + i);
            const float4 xv = *reinterpret_cast<const float4 *>(x + i);
            float df = yv.x - xv.x; acc = fmaf(df, df, acc);
            df = yv.y - xv.y; acc = fmaf(df, df, acc);
            df = yv.z - xv.z; acc = fmaf(df, df, acc);
            df = yv.w - xv.w; acc = fmaf(df, df, acc);
        }
        return acc;
    }
    for (int i = lane; i < d; i += 32) {
        const float df = y[i] - x[i];
        acc = fmaf(df, df, acc);
    }
    return acc;
}

// U rows against the same query with all row reads of a step issued before the first is consumed (the gather is bound
// by memory-level parallelism). Per row the operations and their order are those of warp_row_partial. Null rows give 0.
template <typename T, int U>
__device__ __forceinline__ void warp_rows_partial(const T *const (&y)[U], const float *__restrict__ x, int d, int lane, T (&acc)[U])
{
#pragma unroll
    for (int u = 0; u < U; u++) acc[u] = (T)0;
    for (int i = lane; i < d; i += 32) {
        const T xv = (T)x[i];
        T yv[U];
#pragma unroll
        for (int u = 0; u < U; u++) yv[u] = y[u] ? y[u][i] : xv;
#pragma unroll
        for (int u = 0; u < U; u++) { const T df = yv[u] - xv; acc[u] = fma(df, df, acc[u]); }
    }
}

template <int U>
__device__ __forceinline__ void warp_rows_partial_f32(const float *const (&y)[U], const float *__restrict__ x, int d, int lane, float (&acc)[U])
{
    bool vec = (d & 3) == 0 && (((uintptr_t)x) & 15) == 0;
#pragma unroll
    for (int u = 0; u < U; u++) vec = vec && (((uintptr_t)y[u]) & 15) == 0;
    if (!vec) {                                                            // warp_row_partial<float> decides per row; keep its path per row
#pragma unroll
        for (int u = 0; u < U; u++) acc[u] = y[u] ? warp_row_partial<float>(y[u], x, d, lane) : 0.0f;
        return;
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc[u] = 0.0f;
    for (int i = 4 * lane; i < d; i += 128) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + i);
        float4 yv[U];
#pragma unroll
        for (int u = 0; u < U; u++) yv[u] = y[u] ? *reinterpret_cast<const float4 *>(y[u] + i) : xv;
#pragma unroll
        for (int u = 0; u < U; u++) {
            float df = yv[u].x - xv.x; acc[u] = fmaf(df, df, acc[u]);
            df = yv[u].y - xv.y; acc[u] = fmaf(df, df, acc[u]);
            df = yv[u].z - xv.z; acc[u] = fmaf(df, df, acc[u]);
            df = yv[u].w - xv.w; acc[u] = fmaf(df, df, acc[u]);
        }
    }
}

template <typename T, int U>
__device__ __forceinline__ void warp_rows_dispatch(const T *const (&y)[U], const float *__restrict__ x, int d, int lane, T (&acc)[U])
{
    warp_rows_partial<T, U>(y, x, d, lane, acc);
}
template <int U>
__device__ __forceinline__ void warp_rows_dispatch(const float *const (&y)[U], const float *__restrict__ x, int d, int lane, float (&acc)[U])
{
    warp_rows_partial_f32<U>(y, x, d, lane, acc);
}

template <typename T>
__device__ __forceinline__ T warp_tree_sum(T a)
{
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
    return a;
}

template <typename T>
__device__ __forceinline__ T warp_row_dist(const T *__restrict__ y, const float *__restrict__ x, int d, int lane)
{
    return warp_tree_sum<T>(warp_row_partial<T>(y, x, d, lane));
}

}  // namespace tkb
