// tkb_api.cu -- extern "C" boundary of libtinyknn_b200.so (see include/tinyknn_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "tkb_common.cuh"

namespace tkb {

static thread_local char g_err[512] = "";

char *err_buf() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Grow-only device scratch for the host-buffer surface (one set per host thread and device).
struct Scratch {
    void *p = nullptr;
    size_t cap = 0;
    int dev = -1;
    int reserve(size_t bytes)
    {
        int cur = 0;
        TKB_CUDA(cudaGetDevice(&cur));
        if (p && cur == dev && bytes <= cap) return TKB_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4;
        TKB_CUDA(cudaMalloc(&p, want));
        cap = want; dev = cur;
        return TKB_OK;
    }
};

static thread_local Scratch s_codes, s_tables, s_est, s_heap, s_labels;

static int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_err(TKB_ERR_NO_DEVICE,
                       "no CUDA device available (%s): tinyknn_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return TKB_OK;
}

}  // namespace tkb

using namespace tkb;

extern "C" {

int tkb_version(void) { return 100; }

const char *tkb_last_error(void) { return tkb::err_buf(); }

long long tkb_launch_count(void) { return tkb::g_launches.load(std::memory_order_relaxed); }

int tkb_device_count(int *count)
{
    TKB_REQUIRE(count, "null pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return TKB_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer surface
// ---------------------------------------------------------------------------------------------

int tkb_estimate_pq_host(const uint64_t *data, int64_t n_chunks, int M, const uint64_t *tables,
                         uint64_t *out, int order, int signd)
{
    TKB_REQUIRE(n_chunks >= 0 && M > 0, "bad extent");
    if (n_chunks == 0) return TKB_OK;
    TKB_REQUIRE(data && tables && out, "null pointer");
    if (int rc = require_device()) return rc;
    const size_t code_bytes = (size_t)n_chunks * M * 8, tab_bytes = (size_t)M * 16, est_bytes = (size_t)n_chunks * 16;
    if (int rc = s_codes.reserve(code_bytes)) return rc;
    if (int rc = s_tables.reserve(tab_bytes)) return rc;
    if (int rc = s_est.reserve(est_bytes)) return rc;
    cudaStream_t st = cudaStreamPerThread;
    TKB_CUDA(cudaMemcpyAsync(s_codes.p, data, code_bytes, cudaMemcpyHostToDevice, st));
    TKB_CUDA(cudaMemcpyAsync(s_tables.p, tables, tab_bytes, cudaMemcpyHostToDevice, st));
    if (int rc = launch_estimate((const uint64_t *)s_codes.p, n_chunks, M, (const uint8_t *)s_tables.p, 1,
                                 (uint8_t *)s_est.p, 16 * n_chunks, order, signd, st)) return rc;
    TKB_CUDA(cudaMemcpyAsync(out, s_est.p, est_bytes, cudaMemcpyDeviceToHost, st));
    TKB_CUDA(cudaStreamSynchronize(st));
    return TKB_OK;
}

int tkb_query_pq_host(const uint64_t *data, int64_t n_chunks, int M, int n, const uint64_t *tables,
                      int64_t *indices, int32_t *vals, int R, int order, int signd,
                      const int64_t *labels)
{
    TKB_REQUIRE(n_chunks >= 0 && M > 0 && R >= 0 && n >= 0, "bad extent");
    if (n_chunks == 0 || R == 0) return TKB_OK;
    TKB_REQUIRE(data && tables && indices && vals, "null pointer");
    if (int rc = require_device()) return rc;
    const size_t code_bytes = (size_t)n_chunks * M * 8, tab_bytes = (size_t)M * 16, est_bytes = (size_t)n_chunks * 16;
    // labels are only dereferenced for positions < n (ref: _fast_pq.pyx:193-197)
    const int64_t n_lab = n < 16 * n_chunks ? n : 16 * n_chunks;
    const size_t heap_bytes = (size_t)R * 12;
    if (int rc = s_codes.reserve(code_bytes)) return rc;
    if (int rc = s_tables.reserve(tab_bytes)) return rc;
    if (int rc = s_est.reserve(est_bytes)) return rc;
    if (int rc = s_heap.reserve(heap_bytes + 16)) return rc;
    if (labels) if (int rc = s_labels.reserve((size_t)n_lab * 8 + 8)) return rc;
    cudaStream_t st = cudaStreamPerThread;
    int64_t *d_idx = (int64_t *)s_heap.p;
    int32_t *d_val = (int32_t *)((char *)s_heap.p + (size_t)R * 8);
    TKB_CUDA(cudaMemcpyAsync(s_codes.p, data, code_bytes, cudaMemcpyHostToDevice, st));
    TKB_CUDA(cudaMemcpyAsync(s_tables.p, tables, tab_bytes, cudaMemcpyHostToDevice, st));
    TKB_CUDA(cudaMemcpyAsync(d_idx, indices, (size_t)R * 8, cudaMemcpyHostToDevice, st));
    TKB_CUDA(cudaMemcpyAsync(d_val, vals, (size_t)R * 4, cudaMemcpyHostToDevice, st));
    if (labels && n_lab > 0)
        TKB_CUDA(cudaMemcpyAsync(s_labels.p, labels, (size_t)n_lab * 8, cudaMemcpyHostToDevice, st));
    if (int rc = launch_estimate((const uint64_t *)s_codes.p, n_chunks, M, (const uint8_t *)s_tables.p, 1,
                                 (uint8_t *)s_est.p, 16 * n_chunks, order, signd, st)) return rc;
    if (int rc = launch_replay((const uint8_t *)s_est.p, 16 * n_chunks, n_chunks, n, d_idx, d_val, 1, R, signd,
                               labels ? (const int64_t *)s_labels.p : nullptr, st)) return rc;
    TKB_CUDA(cudaMemcpyAsync(indices, d_idx, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
    TKB_CUDA(cudaMemcpyAsync(vals, d_val, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
    TKB_CUDA(cudaStreamSynchronize(st));
    return TKB_OK;
}

int tkb_init_heap(int64_t *indices, int32_t *vals, int R, int signd)
{
    TKB_REQUIRE(R >= 0 && (R == 0 || (indices && vals)), "bad heap");
    for (int i = 0; i < R; i++) { indices[i] = -1; vals[i] = signd ? 127 : 255; }
    return TKB_OK;
}

int tkb_insert(int64_t *indices, int32_t *vals, int R, int64_t label, int v)
{
    TKB_REQUIRE(R > 0 && indices && vals, "bad heap");
    heap_insert(indices, vals, R, label, v);
    return TKB_OK;
}

int tkb_insert_is(int64_t *indices, int32_t *vals, int R, int64_t label, int v)
{
    TKB_REQUIRE(R > 0 && indices && vals, "bad heap");
    heap_insert_is(indices, vals, R, label, v);
    return TKB_OK;
}

// ---------------------------------------------------------------------------------------------
// device surface
// ---------------------------------------------------------------------------------------------

int tkb_lut_build_dev(const float *queries, int Q, int d, int normalize, float *q_out,
                      const float *centers, int Dp, int dpb, const double *R, int Dpad,
                      double sqrt_n_blocks, double log_n_blocks, int signd,
                      uint8_t *tables, double *q_rot, double *shift, double *scale, void *stream)
{
    return launch_lut_build(queries, Q, d, normalize, q_out, centers, Dp, dpb, R, Dpad, sqrt_n_blocks,
                            log_n_blocks, signd, tables, q_rot, shift, scale, (cudaStream_t)stream);
}

int tkb_estimate_dev(const uint64_t *codes, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                     uint8_t *est, int64_t est_stride, int order, int signd, void *stream)
{
    return launch_estimate(codes, n_chunks, M, tables, Q, est, est_stride, order, signd, (cudaStream_t)stream);
}

int tkb_ivf_scan_dev(const uint64_t *codes, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                     const uint8_t *tables, const int32_t *probes, int Q, int P,
                     uint8_t *est, int64_t slot_stride, const int64_t *seg_off, int64_t max_list_chunks,
                     int order, int signd, void *stream)
{
    return launch_ivf_scan(codes, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, est, slot_stride, seg_off,
                           max_list_chunks, order, signd, (cudaStream_t)stream);
}

int tkb_ivf_plan_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                     int n_lists, int mode, int rank, int n_ranks, int q_per_rank,
                     int64_t *seg_off, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream)
{
    TKB_REQUIRE(mode != TKB_PLAN_PUSH, "use tkb_ivf_plan_push_dev for TKB_PLAN_PUSH");
    return launch_ivf_plan(probes, Q, P, list_size, list_owner, n_lists, mode, rank, n_ranks, q_per_rank, nullptr, seg_off,
                           group_bytes, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_ivf_plan_push_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                          int n_lists, int rank, int n_ranks, int q_per_rank, const int64_t *home_base,
                          int64_t *seg_addr, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream)
{
    return launch_ivf_plan(probes, Q, P, list_size, list_owner, n_lists, TKB_PLAN_PUSH, rank, n_ranks, q_per_rank, home_base,
                           seg_addr, group_bytes, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_ivf_plan_pull_owner_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                                int n_lists, int rank, int n_ranks, int q_per_rank, int64_t capacity,
                                int64_t *seg_off, int64_t *group_bytes, void *workspace, int64_t workspace_bytes, void *stream)
{
    TKB_REQUIRE(capacity > 0, "capacity must be positive");
    return launch_ivf_plan(probes, Q, P, list_size, list_owner, n_lists, TKB_PLAN_SEND, rank, n_ranks, q_per_rank, nullptr, seg_off,
                           group_bytes, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, capacity);
}

int tkb_ivf_plan_pull_home_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner,
                               int n_lists, int rank, int n_ranks, int q_per_rank, const int64_t *owner_base,
                               const int64_t *owner_groups, int64_t *seg_addr, int64_t *group_bytes, void *workspace,
                               int64_t workspace_bytes, void *stream)
{
    return launch_ivf_plan(probes, Q, P, list_size, list_owner, n_lists, 3 /* pull, home side */, rank, n_ranks, q_per_rank,
                           owner_base, seg_addr, group_bytes, workspace, workspace_bytes, (cudaStream_t)stream, owner_groups, 0);
}

int tkb_ivf_pull_minima_dev(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                            const int64_t *seg_addr, const int64_t *seg_local, const int64_t *cm_table, uint8_t *cmin_local,
                            void *stream)
{
    return launch_pull_minima(probes, Q, P, list_size, list_owner, n_lists, seg_addr, seg_local, cm_table, cmin_local,
                              (cudaStream_t)stream);
}

int tkb_ivf_replay_fresh_pull_dev(const int64_t *seg_addr, const int64_t *cm_seg_off, const uint8_t *cmin,
                                  const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, const int64_t *ids,
                                  const int32_t *probes, int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                                  int unique_labels, int32_t *fallback, void *stream)
{
    TKB_REQUIRE(seg_addr, "null pointer");
    TKB_REQUIRE(!cmin || cm_seg_off, "the chunk minima need the layout they are addressed by (cm_seg_off)");
    return launch_ivf_replay_fresh(nullptr, 0, seg_addr, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx,
                                   heap_val, R, signd, unique_labels, fallback, (cudaStream_t)stream, cmin, cm_seg_off);
}

// ---------------------------------------------------------------------------------------------
// peer memory (one process per GPU; CUDA IPC handles travel through the caller's own channel)
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == TKB_PEER_HANDLE_BYTES, "handle size");

int tkb_peer_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle)
{
    TKB_REQUIRE(bytes > 0 && dev_ptr && handle, "bad argument");
    if (int rc = require_device()) return rc;
    void *p = nullptr;
    TKB_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return set_err(TKB_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return TKB_OK;
}

int tkb_peer_open(const unsigned char *handle, void **dev_ptr)
{
    TKB_REQUIRE(handle && dev_ptr, "null pointer");
    if (int rc = require_device()) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    TKB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return TKB_OK;
}

int tkb_peer_close(void *dev_ptr)
{
    if (!dev_ptr) return TKB_OK;
    TKB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return TKB_OK;
}

int tkb_peer_free(void *dev_ptr)
{
    if (!dev_ptr) return TKB_OK;
    TKB_CUDA(cudaFree(dev_ptr));
    return TKB_OK;
}

int tkb_kmeans_workspace(int64_t n, int d, int k, int64_t *bytes) { return kmeans_workspace_bytes(n, d, k, bytes); }

int tkb_kmeans_dev(const float *rows, int64_t n, int d, int k, float *centers, int max_iters, double absmax, int32_t *assign,
                   int *iters_done, void *workspace, int64_t workspace_bytes, void *stream)
{
    if (int rc = require_device()) return rc;
    return launch_kmeans(rows, n, d, k, centers, max_iters, absmax, assign, iters_done, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_kmeans_pq_dev(const float *rows, int64_t n, int D, int dpb, float *centers, int iters, double absmax, void *workspace,
                      int64_t workspace_bytes, void *stream)
{
    if (int rc = require_device()) return rc;
    return launch_kmeans_pq(rows, n, D, dpb, centers, iters, absmax, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_encode_dev(const void *rows, int rows_dtype, int64_t n_rows, int d, const int64_t *row_index, int64_t n_out,
                   const float *centers, const float *cnorm, int Dp, int dpb, const double *R, int Dpad,
                   uint64_t *codes, void *stream)
{
    return launch_encode(rows, rows_dtype, n_rows, d, row_index, n_out, centers, cnorm, Dp, dpb, R, Dpad, codes,
                         (cudaStream_t)stream);
}

int tkb_assign_dev(const void *rows, int dtype, int64_t n, int d, const void *centers, int C, const void *xnorm,
                   const void *cnorm, int k, int32_t *nearest, void *scratch, int64_t scratch_bytes, void *stream)
{
    return launch_assign(rows, dtype, n, d, centers, C, xnorm, cnorm, k, nearest, scratch, scratch_bytes, (cudaStream_t)stream);
}

int tkb_codes_to_native_dev(const uint64_t *codes, int64_t n_chunks, int M, void *native, void *stream)
{
    return launch_codes_to_native(codes, n_chunks, M, native, (cudaStream_t)stream);
}

int tkb_codes_from_native_dev(const void *native, int64_t n_chunks, int M, uint64_t *codes, void *stream)
{
    return launch_codes_from_native(native, n_chunks, M, codes, (cudaStream_t)stream);
}

int tkb_estimate_native_dev(const void *native, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                            uint8_t *est, int64_t est_stride, int order, int signd,
                            void *workspace, int64_t workspace_bytes, void *stream)
{
    return launch_estimate_native(native, n_chunks, M, tables, Q, est, est_stride, order, signd, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}

int tkb_ivf_scan_native_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                            const uint8_t *tables, const int32_t *probes, int Q, int P,
                            uint8_t *est, int64_t slot_stride, const int64_t *seg_off, int64_t max_chunks_per_query,
                            int order, int signd, void *workspace, int64_t workspace_bytes, void *stream)
{
    return launch_ivf_scan_native(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, est, slot_stride,
                                  seg_off, max_chunks_per_query, order, signd, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_ivf_scan_tc_supported(void)
{
    return tc_supported();
}

int tkb_ivf_scan_tc_workspace(int Q, int P, int n_lists, int64_t *bytes)
{
    return tc_workspace_bytes(Q, P, n_lists, bytes);
}

int tkb_ivf_scan_tc_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                        const uint8_t *tables, const int32_t *probes, int Q, int P,
                        uint8_t *est, const int64_t *seg_off, uint8_t *cmin, const int64_t *cm_home, int q_per_rank,
                        int64_t max_chunks_per_query, void *workspace, int64_t workspace_bytes, void *stream)
{
    return launch_ivf_scan_tc(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, est, seg_off, cmin, cm_home,
                              q_per_rank, max_chunks_per_query, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tkb_heap_fill_dev(int64_t *heap_idx, int32_t *heap_val, int64_t count, int signd, void *stream)
{
    return launch_heap_fill(heap_idx, heap_val, count, signd, (cudaStream_t)stream);
}

int tkb_replay_dev(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n,
                   int64_t *heap_idx, int32_t *heap_val, int Q, int R, int signd,
                   const int64_t *labels, void *stream)
{
    return launch_replay(est, est_stride, n_chunks, n, heap_idx, heap_val, Q, R, signd, labels, (cudaStream_t)stream);
}

int tkb_ivf_replay_dev(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                       const int32_t *list_size, int n_lists, const int64_t *ids,
                       const int32_t *probes, int Q, int P,
                       int64_t *heap_idx, int32_t *heap_val, int R, int signd, void *stream)
{
    return launch_ivf_replay(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx,
                             heap_val, R, signd, (cudaStream_t)stream);
}

int tkb_replay_fresh_dev(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n,
                         int64_t *heap_idx, int32_t *heap_val, int Q, int R, int signd, void *stream)
{
    return launch_replay_fresh(est, est_stride, n_chunks, n, heap_idx, heap_val, Q, R, signd, (cudaStream_t)stream);
}

int tkb_ivf_replay_fresh_dev(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                             const int32_t *list_size, int n_lists, const int64_t *ids,
                             const int32_t *probes, int Q, int P,
                             int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                             int unique_labels, int32_t *fallback, void *stream)
{
    return launch_ivf_replay_fresh(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx,
                                   heap_val, R, signd, unique_labels, fallback, (cudaStream_t)stream);
}

int tkb_ivf_scan_native_cm_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                               const uint8_t *tables, const int32_t *probes, int Q, int P,
                               uint8_t *est, const int64_t *seg_off, uint8_t *cmin, int64_t max_chunks_per_query,
                               int order, int signd, void *workspace, int64_t workspace_bytes, void *stream)
{
    TKB_REQUIRE(cmin && (uintptr_t)cmin % 16 == 0, "cmin must be a 16-byte aligned device buffer");
    return launch_ivf_scan_native(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, est, 0, seg_off,
                                  max_chunks_per_query, order, signd, workspace, workspace_bytes, (cudaStream_t)stream, cmin);
}

int tkb_ivf_scan_native_push_cm_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                                    const uint8_t *tables, const int32_t *probes, int Q, int P,
                                    const int64_t *seg_addr, const int64_t *cm_home, int q_per_rank, int64_t max_chunks_per_query,
                                    int order, int signd, void *workspace, int64_t workspace_bytes, void *stream)
{
    TKB_REQUIRE(seg_addr && cm_home, "null pointer");
    return launch_ivf_scan_native(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, nullptr, 0, seg_addr,
                                  max_chunks_per_query, order, signd, workspace, workspace_bytes, (cudaStream_t)stream, nullptr,
                                  cm_home, q_per_rank);
}

int tkb_ivf_replay_fresh_cm_dev(const uint8_t *est, const int64_t *seg_off, const uint8_t *cmin, const int64_t *list_chunk_off,
                                const int32_t *list_size, int n_lists, const int64_t *ids,
                                const int32_t *probes, int Q, int P,
                                int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                                int unique_labels, int32_t *fallback, void *stream)
{
    TKB_REQUIRE(cmin, "null pointer");
    return launch_ivf_replay_fresh(est, 0, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx,
                                   heap_val, R, signd, unique_labels, fallback, (cudaStream_t)stream, cmin);
}

int tkb_gather_dists_dev(const void *rows, int rows_dtype, int64_t n_rows, int d,
                         const float *queries, const int64_t *idx, int Q, int R,
                         void *dists, void *stream)
{
    return launch_gather_dists(rows, rows_dtype, n_rows, d, queries, idx, Q, R, dists, (cudaStream_t)stream);
}

int tkb_select_probes_dev(const int64_t *heap_idx, const void *dists, int dists_dtype, int Q, int R,
                          int P, int32_t *probes, void *stream)
{
    return launch_select_probes(heap_idx, dists, dists_dtype, Q, R, P, probes, (cudaStream_t)stream);
}

int tkb_coarse_probes_dev(const void *native_centers, int64_t n_chunks, int C, int M, const uint8_t *tables, int Q,
                          const float *centers, int d, const float *queries, int R, int P, int order, int32_t *probes,
                          int64_t *heap_idx, int32_t *heap_val, float *dists, void *stream)
{
    return launch_coarse_probes(native_centers, n_chunks, C, M, tables, Q, centers, d, queries, R, P, order, probes, heap_idx,
                                heap_val, dists, (cudaStream_t)stream);
}

int tkb_select_topk_dev(const int64_t *heap_idx, const void *dists, int dists_dtype, int Q, int R,
                        int k, int64_t *out_ids, void *out_dists, int32_t *out_count, void *stream)
{
    return launch_select_topk(heap_idx, dists, dists_dtype, Q, R, k, out_ids, out_dists, out_count, (cudaStream_t)stream);
}

int tkb_ivf_query_fused_workspace(int Q, int P, int R, int M, int order, int rows_dtype,
                                  int64_t max_list_chunks, int64_t *bytes)
{
    return fused_workspace_bytes(Q, P, R, M, order, rows_dtype, max_list_chunks, bytes);
}

int tkb_ivf_query_fused_dev(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists,
                            int M, const uint8_t *tables, const int32_t *probes, int Q, int P, const int64_t *ids,
                            const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                            int R, int k, int order, int64_t max_list_chunks,
                            int64_t *out_ids, void *out_dists, int32_t *out_count,
                            int64_t *heap_idx, int32_t *heap_val,
                            void *workspace, int64_t workspace_bytes, void *stream)
{
    return launch_ivf_query_fused(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, ids, rows, rows_dtype,
                                  n_rows, d, queries, R, k, order, max_list_chunks, out_ids, out_dists, out_count,
                                  heap_idx, heap_val, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
