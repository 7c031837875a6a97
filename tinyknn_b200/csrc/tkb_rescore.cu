// tkb_rescore.cu -- exact rescoring of heap candidates and the final selections.
//
// gather_dists replaces the arithmetic of knn_brute1 (ref: tinyknn/utils.py:89-92):
//     diff = Y[idx] - x ; dists = einsum('ij,ij->i', diff, diff)
// in the dtype numpy would use (f32 rows & f32 query -> f32; f64 rows -> f64). One warp per
// candidate row: a 512-byte f32 row is one coalesced 128-bit-per-lane read.
//
// select_probes / select_topk replace the argpartition-based picks of _FastDistanceTable.top
// (ref: tinyknn/fast_pq.py:307-312) and IVF.query (ref: tinyknn/ivf.py:154-163) with a
// deterministic device order: ascending distance, ties broken by heap slot. numpy's
// argpartition order is build/CPU specific; the host layer offers the numpy
// order as a parity mode on top of gather_dists.
#include "tkb_common.cuh"
#include "tkb_rescore_core.cuh"

namespace tkb {

constexpr int GD_WARPS = 8;
constexpr int GD_ROWS = 8;                            // candidate rows in flight per warp

// A warp takes GD_ROWS candidate rows of one query (blockIdx.y: no index division on the hot path): the row indices are
// read together, then all row reads are issued before any is consumed -- the gather is bound by memory-level parallelism.
template <typename T>
__global__ void __launch_bounds__(32 * GD_WARPS)
gather_dists_kernel(const T *__restrict__ rows, int64_t n_rows, int d, const float *__restrict__ queries,
                    const int64_t *__restrict__ idx, int R, T *__restrict__ dists)
{
    const int r0 = (blockIdx.x * GD_WARPS + (threadIdx.x >> 5)) * GD_ROWS;
    if (r0 >= R) return;
    const int lane = threadIdx.x & 31;
    const size_t w0 = (size_t)blockIdx.y * R + r0;
    int64_t mine = 0;
    if (lane < GD_ROWS && r0 + lane < R) mine = idx[w0 + lane];
    const T *y[GD_ROWS];
    bool live[GD_ROWS];
#pragma unroll
    for (int u = 0; u < GD_ROWS; u++) {
        int64_t row = __shfl_sync(FULL, mine, u);
        live[u] = r0 + u < R;
        if (row < 0) row += n_rows;                   // numpy negative indexing (heap padding -1 -> last row)
        y[u] = (live[u] && row >= 0 && row < n_rows) ? rows + row * d : nullptr;   // numpy would raise IndexError: NaN below
    }
    T acc[GD_ROWS];
    warp_rows_dispatch(y, queries + (size_t)blockIdx.y * d, d, lane, acc);
#pragma unroll
    for (int u = 0; u < GD_ROWS; u++) {
        const T a = warp_tree_sum<T>(acc[u]);
        if (lane == 0 && live[u]) dists[w0 + u] = y[u] ? a : (T)NAN;
    }
}

// One warp per query. Repeatedly extracts the smallest remaining (dist, slot) pair.
// `valid(slot)` decides which slots participate; NaNs sort last.
template <typename T>
__device__ __forceinline__ void warp_argmin(const T *d, const unsigned char *taken, int R, int lane,
                                            T &best_d, int &best_s)
{
    best_d = (T)INFINITY; best_s = INT32_MAX;
    for (int s = lane; s < R; s += 32) {
        if (taken[s]) continue;
        T v = d[s];
        if (v != v) v = (T)INFINITY;
        if (v < best_d || (v == best_d && s < best_s)) { best_d = v; best_s = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const T od = __shfl_xor_sync(FULL, best_d, o);
        const int os = __shfl_xor_sync(FULL, best_s, o);
        if (od < best_d || (od == best_d && os < best_s)) { best_d = od; best_s = os; }
    }
}

constexpr int SEL_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(32 * SEL_WARPS)
select_probes_kernel(const int64_t *__restrict__ heap_idx, const T *__restrict__ dists, int Q, int R,
                     int P, int32_t *__restrict__ probes)
{
    extern __shared__ unsigned char taken_all[];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * SEL_WARPS + wq;
    if (q >= Q) return;
    const int64_t *hi = heap_idx + (size_t)q * R;
    int32_t *out = probes + (size_t)q * P;
    if (R <= P) {                                     // ref: fast_pq.py:307-308 -- raw heap, no rescoring
        for (int s = lane; s < P; s += 32) out[s] = (s < R) ? (int32_t)hi[s] : PROBE_SKIP;
        return;
    }
    unsigned char *taken = taken_all + (size_t)wq * R;
    for (int s = lane; s < R; s += 32) taken[s] = 0;
    __syncwarp();
    const T *d = dists + (size_t)q * R;
    for (int k = 0; k < P; k++) {
        T bd; int bs;
        warp_argmin<T>(d, taken, R, lane, bd, bs);
        if (lane == 0) { taken[bs] = 1; out[k] = (int32_t)hi[bs]; }
        __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(32 * SEL_WARPS)
select_topk_kernel(const int64_t *__restrict__ heap_idx, const T *__restrict__ dists, int Q, int R, int k,
                   int64_t *__restrict__ out_ids, T *__restrict__ out_dists, int32_t *__restrict__ out_count)
{
    extern __shared__ unsigned char taken_all[];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * SEL_WARPS + wq;
    if (q >= Q) return;
    const int64_t *hi = heap_idx + (size_t)q * R;
    const T *d = dists + (size_t)q * R;
    int64_t *oi = out_ids + (size_t)q * k;
    T *od = out_dists ? out_dists + (size_t)q * k : nullptr;
    unsigned char *taken = taken_all + (size_t)wq * R;

    // drop heap padding (ref: ivf.py:154-155)
    int n_valid = 0;
    for (int s0 = 0; s0 < R; s0 += 32) {
        const int s = s0 + lane;
        const bool ok = (s < R) && hi[s] != -1;
        if (s < R) taken[s] = ok ? 0 : 1;
        n_valid += __popc(__ballot_sync(FULL, ok));
    }
    __syncwarp();
    if (n_valid <= k) {                               // ref: ivf.py:158-159 -- survivors in heap order
        int w = 0;
        for (int s0 = 0; s0 < R; s0 += 32) {
            const int s = s0 + lane;
            const bool ok = (s < R) && !taken[s];
            const unsigned b = __ballot_sync(FULL, ok);
            if (ok) {
                const int o = w + __popc(b & ((1u << lane) - 1));
                oi[o] = hi[s];
                if (od) od[o] = d[s];
            }
            w += __popc(b);
        }
        for (int o = n_valid + lane; o < k; o += 32) { oi[o] = -1; if (od) od[o] = (T)INFINITY; }
        if (lane == 0) out_count[q] = n_valid;
        return;
    }
    for (int j = 0; j < k; j++) {
        T bd; int bs;
        warp_argmin<T>(d, taken, R, lane, bd, bs);
        if (lane == 0) { taken[bs] = 1; oi[j] = hi[bs]; if (od) od[j] = d[bs]; }
        __syncwarp();
    }
    if (lane == 0) out_count[q] = k;
}

int launch_gather_dists(const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                        const int64_t *idx, int Q, int R, void *dists, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0, "negative extent");
    if (Q == 0 || R == 0) return TKB_OK;
    TKB_REQUIRE(rows && queries && idx && dists, "null pointer");
    TKB_REQUIRE(n_rows > 0 && d > 0, "empty rows");
    TKB_REQUIRE(rows_dtype == TKB_DTYPE_F32 || rows_dtype == TKB_DTYPE_F64, "rows dtype must be f32 or f64");
    const unsigned bx = (unsigned)((R + GD_WARPS * GD_ROWS - 1) / (GD_WARPS * GD_ROWS));
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        const dim3 grid(bx, (unsigned)qn);
        const size_t o = (size_t)q0 * R;
        if (rows_dtype == TKB_DTYPE_F32)
            gather_dists_kernel<float><<<grid, 32 * GD_WARPS, 0, st>>>((const float *)rows, n_rows, d, queries + (size_t)q0 * d, idx + o, R, (float *)dists + o);
        else
            gather_dists_kernel<double><<<grid, 32 * GD_WARPS, 0, st>>>((const double *)rows, n_rows, d, queries + (size_t)q0 * d, idx + o, R, (double *)dists + o);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

int launch_select_probes(const int64_t *heap_idx, const void *dists, int dtype, int Q, int R, int P,
                         int32_t *probes, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && P >= 0, "negative extent");
    if (Q == 0 || P == 0) return TKB_OK;
    TKB_REQUIRE(heap_idx && probes && (dists || R <= P), "null pointer");
    TKB_REQUIRE(dtype == TKB_DTYPE_F32 || dtype == TKB_DTYPE_F64, "dists dtype must be f32 or f64");
    const size_t smem = (size_t)SEL_WARPS * R;
    TKB_REQUIRE(smem <= 200 * 1024, "heap too large for device-side selection (R <= 51 200)");
    if (smem > 48 * 1024) {                      // large heaps (k = 100 with hundreds of probes): opt in to the SM's whole shared memory
        TKB_CUDA(cudaFuncSetAttribute(select_probes_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TKB_CUDA(cudaFuncSetAttribute(select_probes_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const unsigned blocks = (unsigned)((Q + SEL_WARPS - 1) / SEL_WARPS);
    if (dtype == TKB_DTYPE_F32)
        select_probes_kernel<float><<<blocks, 32 * SEL_WARPS, smem, st>>>(heap_idx, (const float *)dists, Q, R, P, probes);
    else
        select_probes_kernel<double><<<blocks, 32 * SEL_WARPS, smem, st>>>(heap_idx, (const double *)dists, Q, R, P, probes);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

int launch_select_topk(const int64_t *heap_idx, const void *dists, int dtype, int Q, int R, int k,
                       int64_t *out_ids, void *out_dists, int32_t *out_count, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && k >= 0, "negative extent");
    if (Q == 0 || k == 0) return TKB_OK;
    TKB_REQUIRE(heap_idx && dists && out_ids && out_count, "null pointer");
    TKB_REQUIRE(dtype == TKB_DTYPE_F32 || dtype == TKB_DTYPE_F64, "dists dtype must be f32 or f64");
    const size_t smem = (size_t)SEL_WARPS * R;
    TKB_REQUIRE(smem <= 200 * 1024, "heap too large for device-side selection (R <= 51 200)");
    if (smem > 48 * 1024) {
        TKB_CUDA(cudaFuncSetAttribute(select_topk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TKB_CUDA(cudaFuncSetAttribute(select_topk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const unsigned blocks = (unsigned)((Q + SEL_WARPS - 1) / SEL_WARPS);
    if (dtype == TKB_DTYPE_F32)
        select_topk_kernel<float><<<blocks, 32 * SEL_WARPS, smem, st>>>(heap_idx, (const float *)dists, Q, R, k, out_ids, (float *)out_dists, out_count);
    else
        select_topk_kernel<double><<<blocks, 32 * SEL_WARPS, smem, st>>>(heap_idx, (const double *)dists, Q, R, k, out_ids, (double *)out_dists, out_count);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
