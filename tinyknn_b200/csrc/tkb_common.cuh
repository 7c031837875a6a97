// tkb_common.cuh -- shared helpers for the tinyknn_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tinyknn_b200.h"

namespace tkb {

// ---- per-thread error message ---------------------------------------------------------------
char *err_buf();                       // defined in tkb_api.cu
int set_err(int code, const char *fmt, ...);
void count_launch();                   // bumps the process-wide kernel launch counter (tkb_launch_count)

#define TKB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::tkb::set_err(TKB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define TKB_REQUIRE(cond, msg)                                                              \
    do {                                                                                    \
        if (!(cond)) return ::tkb::set_err(TKB_ERR_INVALID, "invalid argument: %s", msg);  \
    } while (0)

#define TKB_LAUNCH_CHECK()                                                                  \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return ::tkb::set_err(TKB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",         \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);              \
        ::tkb::count_launch();                                                              \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;
constexpr int32_t PROBE_SKIP = INT32_MIN;     // probe slot that does not exist (R_c < n_probes)

// ---- the reference heap (ref: tinyknn/_fast_pq.pyx:274-307) ---------------------------------
// Array-layout binary max-heap over (vals, indices). `sift_from_root` overwrites the root with
// (label, v) and sifts down: a child moves up when STRICTLY greater than the value being placed;
// the left child is tried first and the right one only wins when strictly greater than the left.
__host__ __device__ inline void heap_sift_from_root(int64_t *idx, int32_t *val, int R,
                                                    int64_t label, int v)
{
    int j = 0;
    for (;;) {
        int nxt = j, nxt_val = v;
        int l = 2 * j + 1, r = 2 * j + 2;
        if (l < R) { int lv = val[l]; if (lv > nxt_val) { nxt = l; nxt_val = lv; } }
        if (r < R) { int rv = val[r]; if (rv > nxt_val) { nxt = r; nxt_val = rv; } }
        if (nxt == j) { val[j] = v; idx[j] = label; return; }
        val[j] = val[nxt]; idx[j] = idx[nxt];
        j = nxt;
    }
}

// scalar insert: linear dedupe over all R slots, then replace-root (used on the host side and as
// the single-thread device reference).
__host__ __device__ inline void heap_insert(int64_t *idx, int32_t *val, int R, int64_t label, int v)
{
    for (int j = 0; j < R; j++)
        if (idx[j] == label) return;
    heap_sift_from_root(idx, val, R, label, v);
}

// ref: tinyknn/_fast_pq.pyx:256-271
inline void heap_insert_is(int64_t *idx, int32_t *val, int R, int64_t label, int v)
{
    for (int j = 0; j < R; j++)
        if (idx[j] == label) return;
    int j = 0;
    while (j + 1 != R && val[j + 1] > v) {
        idx[j] = idx[j + 1]; val[j] = val[j + 1];
        j++;
    }
    idx[j] = label; val[j] = v;
}

// ---- saturating 8-bit fold steps --------------------------------------------------------------
template <bool SIGNED>
__device__ __forceinline__ int sat_add8(int acc, int t)
{
    if (SIGNED) return max(__viaddmin_s32(acc, t, 127), -128);   // _mm_adds_epi8
    return min(acc + t, 255);                                    // _mm_adds_epu8 (t >= 0)
}

__device__ __forceinline__ uint4 ldg_nc_u4(const uint4 *p)
{
#ifdef TKB_EMULATE                       // tests/emulate: the source compiled for the CPU, no PTX
    return *p;
#else
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#endif
}

// kernels' host-side launchers (one per .cu file), used by tkb_api.cu
int launch_estimate(const uint64_t *codes, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                    uint8_t *est, int64_t est_stride, int order, int signd, cudaStream_t st);
int launch_ivf_scan(const uint64_t *codes, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                    const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est,
                    int64_t slot_stride, const int64_t *seg_off, int64_t max_list_chunks, int order, int signd, cudaStream_t st);
int launch_ivf_plan(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                    int mode, int rank, int n_ranks, int q_per_rank, const int64_t *home_base, int64_t *seg_off,
                    int64_t *group_bytes, void *workspace, int64_t workspace_bytes, cudaStream_t st,
                    const int64_t *send_bases = nullptr, int64_t capacity = 0);
int launch_pull_minima(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                       const int64_t *seg_addr, const int64_t *seg_local, const int64_t *cm_table, uint8_t *cmin_local,
                       cudaStream_t st);
int launch_encode(const void *rows, int rows_dtype, int64_t n_rows, int d, const int64_t *row_index, int64_t n_out,
                  const float *centers, const float *cnorm, int Dp, int dpb, const double *R, int Dpad, uint64_t *codes,
                  cudaStream_t st);
int launch_assign(const void *rows, int dtype, int64_t n, int d, const void *centers, int C, const void *xnorm,
                  const void *cnorm, int k, int32_t *nearest, void *scratch, int64_t scratch_bytes, cudaStream_t st);
int kmeans_workspace_bytes(int64_t n, int d, int k, int64_t *bytes);
int launch_kmeans(const float *rows, int64_t n, int d, int k, float *centers, int max_iters, double absmax, int32_t *assign,
                  int *iters_done, void *workspace, int64_t workspace_bytes, cudaStream_t st);
int launch_kmeans_pq(const float *rows, int64_t n, int D, int dpb, float *centers, int iters, double absmax, void *workspace,
                     int64_t workspace_bytes, cudaStream_t st);
int launch_codes_to_native(const uint64_t *ref, int64_t n_chunks, int M, void *native, cudaStream_t st);
int launch_codes_from_native(const void *native, int64_t n_chunks, int M, uint64_t *ref, cudaStream_t st);
int launch_estimate_native(const void *native, int64_t n_chunks, int M, const uint8_t *tables, int Q, uint8_t *est,
                           int64_t est_stride, int order, int signd, void *workspace, int64_t workspace_bytes,
                           cudaStream_t st);
int launch_ivf_scan_native(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                           const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est,
                           int64_t slot_stride, const int64_t *seg_off, int64_t max_chunks_per_query, int order, int signd,
                           void *workspace, int64_t workspace_bytes, cudaStream_t st, uint8_t *cmin = nullptr,
                           const int64_t *cm_home = nullptr, int q_per_rank = 0, const uint8_t *skip_q = nullptr);
int tc_supported();
int tc_workspace_bytes(int Q, int P, int n_lists, int64_t *bytes);
int launch_ivf_scan_tc(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                       const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est, const int64_t *seg_off,
                       uint8_t *cmin, const int64_t *cm_home, int q_per_rank, int64_t max_chunks_per_query, void *workspace,
                       int64_t workspace_bytes, cudaStream_t st);
int launch_heap_fill(int64_t *heap_idx, int32_t *heap_val, int64_t count, int signd, cudaStream_t st);
int launch_replay(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n, int64_t *heap_idx,
                  int32_t *heap_val, int Q, int R, int signd, const int64_t *labels, cudaStream_t st);
int launch_ivf_replay(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                      const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                      int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                      cudaStream_t st);
int launch_replay_fresh(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n, int64_t *heap_idx,
                        int32_t *heap_val, int Q, int R, int signd, cudaStream_t st);
int launch_ivf_replay_fresh(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                            const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                            int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                            int unique_labels, int *fallback, cudaStream_t st, const uint8_t *cmin = nullptr,
                            const int64_t *cm_seg = nullptr);
int launch_lut_build(const float *queries, int Q, int d, int normalize, float *q_out,
                     const float *centers, int Dp, int dpb, const double *R, int Dpad,
                     double sqrt_n_blocks, double log_n_blocks, int signd, uint8_t *tables,
                     double *q_rot, double *shift, double *scale, cudaStream_t st);
int launch_gather_dists(const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                        const int64_t *idx, int Q, int R, void *dists, cudaStream_t st);
int launch_select_probes(const int64_t *heap_idx, const void *dists, int dtype, int Q, int R, int P,
                         int32_t *probes, cudaStream_t st);
int launch_select_topk(const int64_t *heap_idx, const void *dists, int dtype, int Q, int R, int k,
                       int64_t *out_ids, void *out_dists, int32_t *out_count, cudaStream_t st);
int launch_coarse_probes(const void *native_centers, int64_t n_chunks, int C, int M, const uint8_t *tables, int Q,
                         const float *centers, int d, const float *queries, int R, int P, int order, int32_t *probes,
                         int64_t *heap_idx, int32_t *heap_val, float *dists, cudaStream_t st);
int fused_workspace_bytes(int Q, int P, int R, int M, int order, int rows_dtype, int64_t max_list_chunks, int64_t *bytes);
int launch_ivf_query_fused(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                           const uint8_t *tables, const int32_t *probes, int Q, int P, const int64_t *ids,
                           const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                           int R, int k, int order, int64_t max_list_chunks,
                           int64_t *out_ids, void *out_dists, int32_t *out_count, int64_t *heap_idx, int32_t *heap_val,
                           void *workspace, int64_t workspace_bytes, cudaStream_t st);

}  // namespace tkb
