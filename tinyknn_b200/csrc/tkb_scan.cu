// tkb_scan.cu -- 4-bit Quick-ADC PQ scan (estimate) kernels.
//
// Replaces compute_block_dists / estimate_pq_sse (ref: tinyknn/_fast_pq.pyx:209-236, :101-111) and
// compute_block_dists_avx / estimate_pq_avx (ref: tinyknn/_fast_pq_256.pyx:126-156, :52-62).
//
// GENERIC kernel (this file, round-1 baseline): any uint8 LUT, both accumulation orders, signed and
// unsigned. One thread owns one 16-vector chunk: it streams the chunk's M/2 16-byte code groups with
// 128-bit loads (byte v of a group = vector v: low nibble sub-quantizer 2p, high nibble 2p+1), looks
// the two nibbles up in the per-query LUT staged in shared memory (all lanes of a warp read the same
// 16-byte LUT row -> 4 banks, distinct words: conflict-free) and folds them with the reference's
// saturating adds, then stores the chunk's 16 estimates with one 128-bit store.
#include "tkb_common.cuh"

namespace tkb {

constexpr int SCAN_THREADS = 128;     // chunks per CTA tile

template <int ORDER, bool SIGNED>
__device__ __forceinline__ void scan_chunk(const uint4 *__restrict__ chunk, int M,
                                           const uint8_t *__restrict__ lut, uint4 &out)
{
    // acc[lane][v]; SSE order uses lane 0 only.
    int acc0[16], acc1[16];
#pragma unroll
    for (int v = 0; v < 16; v++) { acc0[v] = 0; acc1[v] = 0; }

    const int P = M >> 1;
    for (int p = 0; p < P; p += 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (p + h >= P) break;
            const uint4 w = ldg_nc_u4(chunk + p + h);
            const uint8_t *row_lo = lut + 32 * (p + h);      // LUT of sub-quantizer 2(p+h)
            const uint8_t *row_hi = row_lo + 16;             // LUT of sub-quantizer 2(p+h)+1
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int v = 0; v < 16; v++) {
                const uint32_t byte = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
                int t0 = row_lo[byte & 15u];
                int t1 = row_hi[byte >> 4];
                if (SIGNED) { t0 = (int)(int8_t)t0; t1 = (int)(int8_t)t1; }
                if (ORDER == TKB_ORDER_AVX && h == 1) {      // pair index odd -> (j & 2) != 0 -> lane 1
                    acc1[v] = sat_add8<SIGNED>(sat_add8<SIGNED>(acc1[v], t0), t1);
                } else {
                    acc0[v] = sat_add8<SIGNED>(sat_add8<SIGNED>(acc0[v], t0), t1);
                }
            }
        }
    }
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int v = 0; v < 16; v++) {
        int e = acc0[v];
        if (ORDER == TKB_ORDER_AVX) e = sat_add8<SIGNED>(acc0[v], acc1[v]);   // ref: _fast_pq_256.pyx:151-156
        o[v >> 2] |= (uint32_t)(e & 0xff) << (8 * (v & 3));
    }
    out = make_uint4(o[0], o[1], o[2], o[3]);
}

// grid: (chunk tiles, Q). est[q][16*chunk + v]
template <int ORDER, bool SIGNED>
__global__ void __launch_bounds__(SCAN_THREADS)
estimate_generic_kernel(const uint4 *__restrict__ codes, int64_t n_chunks, int M,
                        const uint8_t *__restrict__ tables, uint8_t *__restrict__ est,
                        int64_t est_stride)
{
    extern __shared__ __align__(16) uint8_t lut[];
    const int q = blockIdx.y;
    const uint4 *tq = reinterpret_cast<const uint4 *>(tables + (size_t)q * M * 16);
    for (int i = threadIdx.x; i < M; i += blockDim.x) reinterpret_cast<uint4 *>(lut)[i] = tq[i];
    __syncthreads();

    const int64_t c = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    if (c >= n_chunks) return;
    uint4 o;
    scan_chunk<ORDER, SIGNED>(codes + c * (M >> 1), M, lut, o);
    *reinterpret_cast<uint4 *>(est + (size_t)q * est_stride + 16 * c) = o;
}

// grid: (chunk tiles of the largest list, P, Q). est[q][s][16*chunk_in_list + v]
template <int ORDER, bool SIGNED>
__global__ void __launch_bounds__(SCAN_THREADS)
ivf_scan_generic_kernel(const uint4 *__restrict__ codes, const int64_t *__restrict__ list_chunk_off,
                        const int32_t *__restrict__ list_size, int n_lists, int M, const uint8_t *__restrict__ tables,
                        const int32_t *__restrict__ probes, int P, uint8_t *__restrict__ est,
                        int64_t slot_stride, const int64_t *__restrict__ seg_off)
{
    extern __shared__ __align__(16) uint8_t lut[];
    const int q = blockIdx.z, s = blockIdx.y;
    int l = probes[(size_t)q * P + s];
    if (l == PROBE_SKIP) return;
    if (l < 0) l += n_lists;                                   // Python list indexing (ref: ivf.py:141)
    const int64_t c0 = list_chunk_off[l];
    int64_t nc = list_chunk_off[l + 1] - c0;
    if (list_size) { const int64_t real = ((int64_t)list_size[l] + 15) >> 4; if (real < nc) nc = real; }
    const int64_t so = seg_off ? seg_off[(size_t)q * P + s] : ((int64_t)q * P + s) * slot_stride;
    if (so < 0 || (int64_t)blockIdx.x * SCAN_THREADS >= nc) return;

    const uint4 *tq = reinterpret_cast<const uint4 *>(tables + (size_t)q * M * 16);
    for (int i = threadIdx.x; i < M; i += blockDim.x) reinterpret_cast<uint4 *>(lut)[i] = tq[i];
    __syncthreads();

    const int64_t c = (int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    if (c >= nc) return;
    uint4 o;
    scan_chunk<ORDER, SIGNED>(codes + (c0 + c) * (M >> 1), M, lut, o);
    *reinterpret_cast<uint4 *>(est + so + 16 * c) = o;
}

static int check_scan_args(int M, int order)
{
    TKB_REQUIRE(order == TKB_ORDER_SSE || order == TKB_ORDER_AVX, "order must be TKB_ORDER_SSE or TKB_ORDER_AVX");
    TKB_REQUIRE(M > 0 && M % 2 == 0, "M (sub-quantizers) must be a positive multiple of 2");
    TKB_REQUIRE(order != TKB_ORDER_AVX || M % 4 == 0, "avx order needs M % 4 == 0 (ref: fast_pq.py:24 dpad)");
    TKB_REQUIRE(M * 16 <= 48 * 1024, "M too large for the shared-memory LUT");
    return TKB_OK;
}

#define TKB_DISPATCH_SCAN(KERNEL, grid, smem, st, ...)                                          \
    do {                                                                                        \
        if (order == TKB_ORDER_AVX) {                                                           \
            if (signd) KERNEL<TKB_ORDER_AVX, true><<<grid, SCAN_THREADS, smem, st>>>(__VA_ARGS__);  \
            else       KERNEL<TKB_ORDER_AVX, false><<<grid, SCAN_THREADS, smem, st>>>(__VA_ARGS__); \
        } else {                                                                                \
            if (signd) KERNEL<TKB_ORDER_SSE, true><<<grid, SCAN_THREADS, smem, st>>>(__VA_ARGS__);  \
            else       KERNEL<TKB_ORDER_SSE, false><<<grid, SCAN_THREADS, smem, st>>>(__VA_ARGS__); \
        }                                                                                       \
    } while (0)

int launch_estimate(const uint64_t *codes, int64_t n_chunks, int M, const uint8_t *tables, int Q,
                    uint8_t *est, int64_t est_stride, int order, int signd, cudaStream_t st)
{
    if (int rc = check_scan_args(M, order)) return rc;
    TKB_REQUIRE(n_chunks >= 0 && Q >= 0, "negative extent");
    if (n_chunks == 0 || Q == 0) return TKB_OK;
    TKB_REQUIRE(codes && tables && est, "null pointer");
    TKB_REQUIRE(est_stride >= 16 * n_chunks && est_stride % 16 == 0, "est_stride must be a multiple of 16 and >= 16*n_chunks");
    TKB_REQUIRE(((uintptr_t)codes % 16 == 0) && ((uintptr_t)est % 16 == 0) && ((uintptr_t)tables % 16 == 0),
                "device pointers must be 16-byte aligned");
    const int64_t tiles = (n_chunks + SCAN_THREADS - 1) / SCAN_THREADS;
    TKB_REQUIRE(tiles <= 0x7fffffff, "too many chunks for one launch");
    const uint4 *c4 = reinterpret_cast<const uint4 *>(codes);
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        dim3 grid((unsigned)tiles, (unsigned)qn);
        TKB_DISPATCH_SCAN(estimate_generic_kernel, grid, (size_t)M * 16, st, c4, n_chunks, M,
                          tables + (size_t)q0 * M * 16, est + (size_t)q0 * est_stride, est_stride);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

int launch_ivf_scan(const uint64_t *codes, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                    const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est,
                    int64_t slot_stride, const int64_t *seg_off, int64_t max_list_chunks, int order, int signd, cudaStream_t st)
{
    if (int rc = check_scan_args(M, order)) return rc;
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (max_list_chunks <= 0) max_list_chunks = slot_stride / 16;
    if (Q == 0 || P == 0 || max_list_chunks == 0) return TKB_OK;
    TKB_REQUIRE(codes && list_chunk_off && tables && probes && est, "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0, "slot_stride must be a multiple of 16");
    TKB_REQUIRE(P <= 65535, "too many probes");
    const int64_t tiles = (max_list_chunks + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(codes);
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        dim3 grid((unsigned)tiles, (unsigned)P, (unsigned)qn);
        TKB_DISPATCH_SCAN(ivf_scan_generic_kernel, grid, (size_t)M * 16, st, c4, list_chunk_off, list_size, n_lists, M,
                          tables + (size_t)q0 * M * 16, probes + (size_t)q0 * P, P,
                          seg_off ? est : est + (size_t)q0 * P * slot_stride, slot_stride,
                          seg_off ? seg_off + (size_t)q0 * P : nullptr);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

}  // namespace tkb
