// tkb_kmeans.cu -- Lloyd's k-means on the GPU for the two `fit` steps of the reference (build time).
//
// Replaces the arithmetic of sklearn.cluster.KMeans in `IVF.fit` (coarse centroids: KMeans(n_clusters, n_init=1) on the raw
// vectors; ref: tinyknn/ivf.py:19-51) and in `FastPQ._fit_code` (one KMeans(16, n_init=2) per block of dims_per_block
// dimensions; ref: tinyknn/fast_pq.py:106-145). PARITY UNPINNED by nature: sklearn seeds k-means++ from numpy's global random
// state, so two runs of the reference itself give different codebooks; what is tested is the quality of the clustering
// (inertia against sklearn's on the same data) and that the result is a deterministic function of (data, initial centres).
//
// Determinism: cluster sums are accumulated in FIXED POINT with integer atomics -- integer addition is associative, so the
// centroids do not depend on the order in which thread blocks or atomics happen to run (a float atomicAdd would; two ranks of a
// sharded job that fit the same data must end with the same index bit for bit).
//
//   coarse centroids : assignment = tkb_assign.cu's exact f32 FMA-chain kernel (k = 1), update = one warp per row adding the
//                      row to its cluster's int64 sums; the loop stops when no row changes its cluster
//   PQ codebooks     : all M blocks in one pass over the rows: the 16 centres of every block live in shared memory, a thread
//                      owns a row, the CTA accumulates in int32 shared-memory atomics and flushes to int64 global sums
#include <math.h>

#include "tkb_common.cuh"

namespace tkb {

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_PQ_ROWS = 2048;                 // rows per CTA of the PQ kernel (bounds its int32 partial sums)
constexpr int KM_MAX_DPB = 8;

// update step of the coarse k-means: sums[c][:] += fixed(x_i), counts[c] += 1, and how many rows changed their cluster
__global__ void __launch_bounds__(KM_THREADS)
km_accumulate_kernel(const float *__restrict__ rows, int64_t n, int d, const int32_t *__restrict__ assign,
                     const int32_t *__restrict__ prev, double scale, long long *__restrict__ sums, int *__restrict__ counts,
                     int *__restrict__ changed)
{
    const int64_t i = (int64_t)blockIdx.x * (KM_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int c = assign[i];
    const float *x = rows + (size_t)i * d;
    unsigned long long *s = reinterpret_cast<unsigned long long *>(sums + (size_t)c * d);
    for (int j = lane; j < d; j += 32)
        atomicAdd(s + j, (unsigned long long)__double2ll_rn((double)x[j] * scale));     // two's complement: adds signed values
    if (lane == 0) {
        atomicAdd(counts + c, 1);
        if (prev && prev[i] != c) atomicAdd(changed, 1);
    }
}

// centres of the non-empty clusters = sums / count (an empty cluster keeps its centre)
__global__ void km_finish_kernel(const long long *__restrict__ sums, const int *__restrict__ counts, int k, int d, double inv_scale,
                                 float *__restrict__ centers)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)k * d) return;
    const int cnt = counts[e / d];
    if (cnt > 0) centers[e] = (float)((double)sums[e] * inv_scale / (double)cnt);
}

// PQ codebooks: rows f32 [n][D]; centres in FastPQ's layout: centre c of block m at centers[c * D + m * dpb .. + dpb)
__global__ void __launch_bounds__(KM_THREADS)
km_pq_kernel(const float *__restrict__ rows, int64_t n, int D, int dpb, const float *__restrict__ centers, float fscale,
             long long *__restrict__ gsum /* [M][16][dpb] */, int *__restrict__ gcnt /* [M][16] */)
{
    extern __shared__ __align__(16) unsigned char km_sm[];
    const int M = D / dpb;
    float *cen = reinterpret_cast<float *>(km_sm);                  // [M][16][dpb]
    int *acc = reinterpret_cast<int *>(cen + (size_t)M * 16 * dpb);   // [M][16][dpb]
    int *cnt = acc + (size_t)M * 16 * dpb;                            // [M][16]
    for (int e = threadIdx.x; e < M * 16 * dpb; e += KM_THREADS) {
        const int j = e % dpb, c = (e / dpb) % 16, m = e / (16 * dpb);
        cen[e] = centers[(size_t)c * D + m * dpb + j];
        acc[e] = 0;
    }
    for (int e = threadIdx.x; e < M * 16; e += KM_THREADS) cnt[e] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * KM_PQ_ROWS, hi = min(n, lo + KM_PQ_ROWS);
    for (int64_t i = lo + threadIdx.x; i < hi; i += KM_THREADS) {
        const float *x = rows + (size_t)i * D;
        for (int m = 0; m < M; m++) {
            float v[KM_MAX_DPB];
#pragma unroll
            for (int j = 0; j < KM_MAX_DPB; j++) v[j] = j < dpb ? x[m * dpb + j] : 0.f;
            const float *cm = cen + (size_t)m * 16 * dpb;
            int best = 0;
            float bd = __int_as_float(0x7f800000);
            for (int c = 0; c < 16; c++) {
                float dist = 0.f;
#pragma unroll
                for (int j = 0; j < KM_MAX_DPB; j++)
                    if (j < dpb) { const float t = v[j] - cm[c * dpb + j]; dist = __fmaf_rn(t, t, dist); }
                if (dist < bd) { bd = dist; best = c; }               // first minimum
            }
#pragma unroll
            for (int j = 0; j < KM_MAX_DPB; j++)
                if (j < dpb) atomicAdd(acc + ((size_t)m * 16 + best) * dpb + j, __float2int_rn(v[j] * fscale));
            atomicAdd(cnt + m * 16 + best, 1);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < M * 16 * dpb; e += KM_THREADS)
        if (acc[e]) atomicAdd(reinterpret_cast<unsigned long long *>(gsum) + e, (unsigned long long)(long long)acc[e]);
    for (int e = threadIdx.x; e < M * 16; e += KM_THREADS)
        if (cnt[e]) atomicAdd(gcnt + e, cnt[e]);
}

__global__ void km_pq_finish_kernel(const long long *__restrict__ gsum, const int *__restrict__ gcnt, int D, int dpb, double inv_scale,
                                    float *__restrict__ centers)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;            // e = (m * 16 + c) * dpb + j
    const int M = D / dpb;
    if (e >= M * 16 * dpb) return;
    const int j = e % dpb, c = (e / dpb) % 16, m = e / (16 * dpb);
    const int cnt = gcnt[m * 16 + c];
    if (cnt > 0) centers[(size_t)c * D + m * dpb + j] = (float)((double)gsum[e] * inv_scale / (double)cnt);
}

// largest power of two s with rows_per_sum * absmax * s < 2^bits
double km_scale(double absmax, double rows_per_sum, int bits)
{
    if (!(absmax > 0)) absmax = 1.0;
    return exp2(floor((double)bits - ceil(log2(rows_per_sum)) - ceil(log2(absmax)) - 1.0));
}

size_t km_align(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

int kmeans_workspace_bytes(int64_t n, int d, int k, int64_t *bytes)
{
    TKB_REQUIRE(n >= 0 && d > 0 && k > 0 && bytes, "bad extent");
    // sums, counts + changed, previous assignment, norms for the assignment kernel
    *bytes = (int64_t)(km_align(8 * (size_t)k * d) + km_align(4 * ((size_t)k + 1)) + km_align(4 * (size_t)n) + km_align(4 * ((size_t)n + k)));
    return TKB_OK;
}

int launch_kmeans(const float *rows, int64_t n, int d, int k, float *centers, int max_iters, double absmax, int32_t *assign,
                  int *iters_done, void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    TKB_REQUIRE(n > 0 && d > 0 && k > 0 && max_iters >= 0, "bad extent");
    TKB_REQUIRE(rows && centers && assign && workspace, "null pointer");
    int64_t need = 0;
    kmeans_workspace_bytes(n, d, k, &need);
    TKB_REQUIRE(workspace_bytes >= need, "workspace too small (tkb_kmeans_workspace)");
    unsigned char *w = reinterpret_cast<unsigned char *>(workspace);
    long long *sums = reinterpret_cast<long long *>(w);
    int *counts = reinterpret_cast<int *>(w + km_align(8 * (size_t)k * d));
    int *changed = counts + k;
    int32_t *prev = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(counts) + km_align(4 * ((size_t)k + 1)));
    float *norms = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(prev) + km_align(4 * (size_t)n));
    const double scale = km_scale(absmax, (double)n, 62);
    int it = 0;
    for (; it < max_iters; it++) {
        if (it > 0) TKB_CUDA(cudaMemcpyAsync(prev, assign, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (int rc = launch_assign(rows, TKB_DTYPE_F32, n, d, centers, k, nullptr, nullptr, 1, assign, norms, 4 * (n + k), st)) return rc;
        TKB_CUDA(cudaMemsetAsync(sums, 0, 8 * (size_t)k * d, st));
        TKB_CUDA(cudaMemsetAsync(counts, 0, 4 * ((size_t)k + 1), st));
        km_accumulate_kernel<<<(unsigned)((n + KM_THREADS / 32 - 1) / (KM_THREADS / 32)), KM_THREADS, 0, st>>>(
            rows, n, d, assign, it > 0 ? prev : nullptr, scale, sums, counts, changed);
        TKB_LAUNCH_CHECK();
        int h_changed = 1;
        if (it > 0) {
            TKB_CUDA(cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            TKB_CUDA(cudaStreamSynchronize(st));
        }
        if (h_changed == 0) break;                                   // the centres are already the means of these clusters
        km_finish_kernel<<<(unsigned)(((int64_t)k * d + 255) / 256), 256, 0, st>>>(sums, counts, k, d, 1.0 / scale, centers);
        TKB_LAUNCH_CHECK();
    }
    if (iters_done) *iters_done = it;
    // the assignment that belongs to the returned centres
    return launch_assign(rows, TKB_DTYPE_F32, n, d, centers, k, nullptr, nullptr, 1, assign, norms, 4 * (n + k), st);
}

int launch_kmeans_pq(const float *rows, int64_t n, int D, int dpb, float *centers, int iters, double absmax, void *workspace,
                     int64_t workspace_bytes, cudaStream_t st)
{
    TKB_REQUIRE(n > 0 && D > 0 && dpb > 0 && dpb <= KM_MAX_DPB && D % dpb == 0 && iters >= 0, "bad extent (dims_per_block <= 8)");
    TKB_REQUIRE(rows && centers && workspace, "null pointer");
    const int M = D / dpb;
    const size_t nsum = (size_t)M * 16 * dpb, ncnt = (size_t)M * 16;
    TKB_REQUIRE(workspace_bytes >= (int64_t)(km_align(8 * nsum) + km_align(4 * ncnt)), "workspace too small (16 KB + 160 * D bytes)");
    long long *gsum = reinterpret_cast<long long *>(workspace);
    int *gcnt = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(workspace) + km_align(8 * nsum));
    const size_t smem = 8 * nsum + 4 * ncnt;
    TKB_REQUIRE(smem <= 200 * 1024, "too many blocks for the shared-memory codebooks");
    TKB_CUDA(cudaFuncSetAttribute(km_pq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double scale = km_scale(absmax, (double)KM_PQ_ROWS, 31);
    for (int it = 0; it < iters; it++) {
        TKB_CUDA(cudaMemsetAsync(gsum, 0, 8 * nsum, st));
        TKB_CUDA(cudaMemsetAsync(gcnt, 0, 4 * ncnt, st));
        km_pq_kernel<<<(unsigned)((n + KM_PQ_ROWS - 1) / KM_PQ_ROWS), KM_THREADS, smem, st>>>(rows, n, D, dpb, centers, (float)scale, gsum, gcnt);
        TKB_LAUNCH_CHECK();
        km_pq_finish_kernel<<<(unsigned)((nsum + 255) / 256), 256, 0, st>>>(gsum, gcnt, D, dpb, 1.0 / scale, centers);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

}  // namespace tkb
