// tkb_scan_fast.cu -- the B200-native PQ scan: register-resident LUTs looked up with PRMT, byte-packed
// deferred accumulation, exactness restored by a per-vector certificate + patch pass.
//
// Replaces the same reference functions as tkb_scan.cu (compute_block_dists*, estimate_pq_*;
// ref: tinyknn/_fast_pq.pyx:209-236, tinyknn/_fast_pq_256.pyx:126-156) but on a DEVICE-NATIVE code
// layout chosen at upload time (it round-trips to the reference layout, tkb_codes_from_native_dev):
//
//   tile   = 8 chunks = 128 vectors; tile t, pair p (sub-quantizers 2p, 2p+1), chunk slot s:
//            16 bytes at ((t * M/2 + p) * 8 + s) * 16  -> a warp reads four full 128-byte lines per load.
//   16 B   = 8 halfwords; halfword g (0..3)   = codes of sub-quantizer 2p   for vectors 4g..4g+3,
//                         halfword 4+g        = codes of sub-quantizer 2p+1 for vectors 4g..4g+3,
//            nibble i of a halfword = vector 4g+i. A halfword is directly a PRMT selector.
//
// Why this is exact although it does not clamp after every add (DESIGN.md "certificate"):
//   LUT rows are biased to t' = t - min_c t >= 0 (so the zero byte PRMT returns for the "other half"
//   of a 16-entry row is neutral and four vectors accumulate in one register without carries for 8
//   steps); per accumulation lane the plain sum S is exact. With N = sum_j max(0, -min_c t_j) the
//   reference's saturating fold equals S whenever N <= 128 (no prefix can drop below -128) and
//   S + N <= 127 (no prefix can exceed 127). Vectors failing the test are recomputed by the patch
//   kernel with the reference's step-by-step saturating fold; queries whose LUT fails the per-query
//   preconditions run the step-by-step fold for every vector.
#include <stdlib.h>

#include "tkb_scan_core.cuh"

namespace tkb {

// ------------------------------------------------------------------------------------------------
// layout conversion (upload time)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 group_to_native(uint4 r)
{
    // r: byte v = (code[v][2p] | code[v][2p+1] << 4), v = 0..15
    const uint32_t ws[4] = {r.x, r.y, r.z, r.w};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t w = ws[g];                          // vectors 4g..4g+3
        const uint32_t l = w & 0x0f0f0f0fu, h = (w >> 4) & 0x0f0f0f0fu;
        lo[g] = (l & 0xf) | ((l >> 4) & 0xf0) | ((l >> 8) & 0xf00) | ((l >> 12) & 0xf000);
        hi[g] = (h & 0xf) | ((h >> 4) & 0xf0) | ((h >> 8) & 0xf00) | ((h >> 12) & 0xf000);
    }
    return make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16));
}

__device__ __forceinline__ uint4 group_from_native(uint4 n)
{
    const uint32_t lo[4] = {n.x & 0xffffu, n.x >> 16, n.y & 0xffffu, n.y >> 16};
    const uint32_t hi[4] = {n.z & 0xffffu, n.z >> 16, n.w & 0xffffu, n.w >> 16};
    uint32_t out[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            w |= (((lo[g] >> (4 * i)) & 0xf) | (((hi[g] >> (4 * i)) & 0xf) << 4)) << (8 * i);
        out[g] = w;
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

__global__ void to_native_kernel(const uint4 *__restrict__ ref, int64_t n_chunks, int64_t n_chunks_pad, int Ph,
                                 uint4 *__restrict__ nat)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over n_chunks_pad * Ph groups
    if (i >= n_chunks_pad * Ph) return;
    const int64_t c = i / Ph;
    const int p = (int)(i - c * Ph);
    uint4 r = make_uint4(0, 0, 0, 0);
    if (c < n_chunks) r = ref[c * Ph + p];
    nat[native_off(c, p, Ph)] = group_to_native(r);
}

__global__ void from_native_kernel(const uint4 *__restrict__ nat, int64_t n_chunks, int Ph, uint4 *__restrict__ ref)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks * Ph) return;
    const int64_t c = i / Ph;
    const int p = (int)(i - c * Ph);
    ref[c * Ph + p] = group_from_native(nat[native_off(c, p, Ph)]);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
constexpr int PEND_CAP = 1024;                 // deferred chunks per CTA; more than that are recomputed on the spot

// Shared-memory carve-up of the scan kernels (P = 0 for the brute-force kernel).
struct ScanSmem {
    uint4 *rows, *raw;
    uint2 *sc;
    LutMeta *meta;
    int *scratch;
    int64_t *pend_off;                         // est byte offset of a deferred chunk
    uint32_t *pend_chunk;                      // its chunk in the code array
    int *n_pend;
    int64_t *seg_c0, *seg_o;
    int *seg_end;
    int keep;                                  // flagged chunks are pinned in L2 (evict_last) until the deferred pass
    uint8_t *cmin;                             // optional: smallest estimate of every chunk, at cmin[est offset / 16]
};

__host__ __device__ inline size_t scan_smem_carve(unsigned char *base, int M, int P, ScanSmem *out)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
    const size_t rows = take((size_t)M * 16), raw = take((size_t)M * 16), sc = take((size_t)M * 8);
    const size_t meta = take(sizeof(LutMeta)), scratch = take((size_t)M * 16);
    const size_t po = take((size_t)PEND_CAP * 8), pc = take((size_t)PEND_CAP * 4), np = take(16);
    const size_t c0 = take((size_t)P * 8), so = take((size_t)P * 8), se = take((size_t)P * 4);
    if (out) {
        out->rows = reinterpret_cast<uint4 *>(base + rows); out->raw = reinterpret_cast<uint4 *>(base + raw);
        out->sc = reinterpret_cast<uint2 *>(base + sc); out->meta = reinterpret_cast<LutMeta *>(base + meta);
        out->scratch = reinterpret_cast<int *>(base + scratch);
        out->pend_off = reinterpret_cast<int64_t *>(base + po); out->pend_chunk = reinterpret_cast<uint32_t *>(base + pc);
        out->n_pend = reinterpret_cast<int *>(base + np);
        out->seg_c0 = reinterpret_cast<int64_t *>(base + c0); out->seg_o = reinterpret_cast<int64_t *>(base + so);
        out->seg_end = reinterpret_cast<int *>(base + se);
        out->keep = 0;
        out->cmin = nullptr;
    }
    return o;
}

// The 16-byte groups of a chunk that failed the certificate will be read again by scan_deferred, a few hundred KB of
// streamed codes later: by then they have left L2 (the main loads are evict-first) and each 16-byte group costs a 32-byte
// DRAM sector a second time. Re-touching them with an evict_last prefetch while they are still in L2 keeps them there.
__device__ __forceinline__ void keep_chunk_in_l2(const uint4 *__restrict__ nat, int64_t c, int Ph)
{
    const uint4 *base = nat + native_off(c, 0, Ph);
#ifndef TKB_EMULATE                      // a cache hint: nothing to emulate on the CPU (tests/emulate)
    for (int p = 0; p < Ph; p++)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (size_t)p * TILE));
#else
    (void)base;
#endif
}

// Smallest of a chunk's 16 estimates (signed or unsigned like the estimates themselves): the heap replay of a long
// probe list tests this byte first and fetches the 16 estimates only when one of them can be a candidate.
template <bool SIGNED>
__device__ __forceinline__ uint8_t chunk_min(const uint4 o)
{
    uint32_t m;
    if (SIGNED) {
        m = __vmins4(__vmins4(o.x, o.y), __vmins4(o.z, o.w));
        m = __vmins4(m, m >> 16);
        m = __vmins4(m, m >> 8);
    } else {
        m = __vminu4(__vminu4(o.x, o.y), __vminu4(o.z, o.w));
        m = __vminu4(m, m >> 16);
        m = __vminu4(m, m >> 8);
    }
    return (uint8_t)(m & 0xffu);
}

// One chunk of one query. The fast path's certificate decides; a chunk that fails it is NOT recomputed by the thread
// that found it (its warp would wait) but queued in shared memory and recomputed by the whole CTA after the main loop
// (scan_deferred) with the step-by-step byte-SIMD fold -- the LUT of the query is still in shared memory then.
template <int ORDER, bool SIGNED, int PH>
__device__ __forceinline__ void scan_one(const uint4 *__restrict__ nat, int64_t c, int Ph, const ScanSmem &sm, const LutMeta &m,
                                         uint8_t *__restrict__ est, int64_t off)
{
    uint4 o;
    if (m.eligible) {
        bool flagged;
        o = scan_chunk_fast<SIGNED, PH>(nat, c, Ph, sm.rows, m, flagged);
        if (flagged) {
            const int i = atomicAdd(sm.n_pend, 1);
            if (i < PEND_CAP) {
                sm.pend_off[i] = off; sm.pend_chunk[i] = (uint32_t)c;
                if (sm.keep) keep_chunk_in_l2(nat, c, Ph);
                return;
            }
            o = scan_chunk_steps_cold<ORDER, SIGNED>(nat, c, Ph, sm.rows, sm.sc, m);
        }
    } else if (m.steps_ok) {
        o = scan_chunk_steps<ORDER, SIGNED>(nat, c, Ph, sm.rows, sm.sc, m);
    } else {
        o = scan_chunk_exact_cold<ORDER, SIGNED>(nat, c, Ph, reinterpret_cast<const uint8_t *>(sm.raw));
    }
    *reinterpret_cast<uint4 *>(est + off) = o;
    if (sm.cmin) sm.cmin[off >> 4] = chunk_min<SIGNED>(o);
}

template <int ORDER, bool SIGNED>
__device__ __forceinline__ void scan_deferred(const uint4 *__restrict__ nat, int Ph, const ScanSmem &sm, const LutMeta &m,
                                              uint8_t *__restrict__ est, unsigned long long *stat)
{
    __syncthreads();
    const int total = *sm.n_pend, n = total < PEND_CAP ? total : PEND_CAP;
    if (threadIdx.x == 0 && total && stat) atomicAdd(stat, (unsigned long long)total);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint4 o = scan_chunk_steps<ORDER, SIGNED>(nat, (int64_t)sm.pend_chunk[i], Ph, sm.rows, sm.sc, m);
        *reinterpret_cast<uint4 *>(est + sm.pend_off[i]) = o;
        if (sm.cmin) sm.cmin[sm.pend_off[i] >> 4] = chunk_min<SIGNED>(o);
    }
}

template <int ORDER, bool SIGNED, int PH = 0>
__global__ void __launch_bounds__(FAST_THREADS, 3)
estimate_fast_kernel(const uint4 *__restrict__ nat, int64_t n_chunks, int M, const uint8_t *__restrict__ tables,
                     uint8_t *__restrict__ est, int64_t est_stride, unsigned long long *stat, int Qn)
{
    extern __shared__ __align__(16) unsigned char smem[];
    ScanSmem sm;
    scan_smem_carve(smem, M, 0, &sm);
    // blockIdx.x = stripe * Qn + query: the CTAs that run together read the same stripe of codes for different queries,
    // so a code array larger than L2 is still fetched from HBM about once per batch, not once per query
    const int q = blockIdx.x % Qn, stripe = blockIdx.x / Qn, n_stripes = gridDim.x / Qn, Ph = M >> 1;
    if (threadIdx.x == 0) *sm.n_pend = 0;
    const bool fast_allowed = !(SIGNED && ORDER == TKB_ORDER_SSE);
    prepare_lut<ORDER, SIGNED>(tables + (size_t)q * M * 16, M, fast_allowed, sm.rows, sm.raw, sm.sc, sm.meta, sm.scratch);
    const LutMeta m = *sm.meta;
    for (int64_t c = (int64_t)stripe * blockDim.x + threadIdx.x; c < n_chunks; c += (int64_t)n_stripes * blockDim.x)
        scan_one<ORDER, SIGNED, PH>(nat, c, Ph, sm, m, est, (int64_t)q * est_stride + 16 * c);
    scan_deferred<ORDER, SIGNED>(nat, Ph, sm, m, est, stat);
}

// One CTA column per query: the P probed lists are walked as one flat range of chunks, so the LUT is
// prepared once per (query, split) and short lists do not leave lanes idle. Segment (q, s) is written at
// est + seg_off[q*P+s] (absent when negative) or, without a plan, at est + (q*P+s)*slot_stride. With
// list_size the walk covers only the chunks that hold real vectors (the reference's ceil(n/16)), not the
// tile padding of the native layout.
template <int ORDER, bool SIGNED, int PH = 0>
__global__ void __launch_bounds__(FAST_THREADS, 3)
ivf_scan_fast_kernel(const uint4 *__restrict__ nat, const int64_t *__restrict__ list_chunk_off,
                     const int32_t *__restrict__ list_size, int n_lists, int M,
                     const uint8_t *__restrict__ tables, const int32_t *__restrict__ probes, int P,
                     uint8_t *__restrict__ est, int64_t slot_stride, const int64_t *__restrict__ seg_off,
                     unsigned long long *stat, int keep, uint8_t *__restrict__ cmin,
                     const int64_t *__restrict__ cm_home, int q_per_rank, int q_base, const uint8_t *__restrict__ skip_q)
{
    if (skip_q && skip_q[blockIdx.y]) return;        // the query's segments are scanned by another kernel (tkb_scan_tc.cu)
    extern __shared__ __align__(16) unsigned char smem[];
    ScanSmem sm;
    scan_smem_carve(smem, M, P, &sm);
    sm.keep = keep;
    // push exchange: est offsets are absolute addresses inside the home rank's buffer; its minima region is addressed as
    // cm_home[home rank] + (address >> 4) (the table holds minima base - (est base >> 4), per home rank)
    sm.cmin = cm_home ? reinterpret_cast<uint8_t *>(cm_home[(q_base + (int)blockIdx.y) / q_per_rank]) : cmin;
    const int q = blockIdx.y, Ph = M >> 1;
    if (threadIdx.x == 0) {
        *sm.n_pend = 0;
        int run = 0;
        for (int s = 0; s < P; s++) {
            int l = probes[(size_t)q * P + s];
            int64_t c0 = 0, nc = 0;
            const int64_t o = seg_off ? seg_off[(size_t)q * P + s] : ((int64_t)q * P + s) * slot_stride;
            if (l != PROBE_SKIP && o >= 0) {
                if (l < 0) l += n_lists;
                c0 = list_chunk_off[l];
                nc = list_chunk_off[l + 1] - c0;
                if (list_size) { const int64_t real = ((int64_t)list_size[l] + 15) >> 4; if (real < nc) nc = real; }
            }
            sm.seg_c0[s] = c0;
            sm.seg_o[s] = o;
            run += (int)nc;
            sm.seg_end[s] = run;
        }
    }
    const bool fast_allowed = !(SIGNED && ORDER == TKB_ORDER_SSE);
    prepare_lut<ORDER, SIGNED>(tables + (size_t)q * M * 16, M, fast_allowed, sm.rows, sm.raw, sm.sc, sm.meta, sm.scratch);   // syncs
    const LutMeta m = *sm.meta;
    const int total = sm.seg_end[P - 1];
    int s = 0;                                       // f only grows: the slot search resumes where it stopped
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
        while (f >= sm.seg_end[s]) s++;
        const int local = f - (s ? sm.seg_end[s - 1] : 0);
        scan_one<ORDER, SIGNED, PH>(nat, sm.seg_c0[s] + local, Ph, sm, m, est, sm.seg_o[s] + 16 * (int64_t)local);
    }
    scan_deferred<ORDER, SIGNED>(nat, Ph, sm, m, est, stat);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static size_t fast_smem_bytes(int M, int P)
{
    return scan_smem_carve(nullptr, M, P, nullptr);
}

static int check_fast_args(int M, int order)
{
    TKB_REQUIRE(order == TKB_ORDER_SSE || order == TKB_ORDER_AVX, "order must be TKB_ORDER_SSE or TKB_ORDER_AVX");
    TKB_REQUIRE(M > 0 && M % 2 == 0, "M (sub-quantizers) must be a positive multiple of 2");
    TKB_REQUIRE(order != TKB_ORDER_AVX || M % 4 == 0, "avx order needs M % 4 == 0 (ref: fast_pq.py:24 dpad)");
    TKB_REQUIRE(M <= 1024, "M too large");
    return TKB_OK;
}

// Specialised instances of the default path (avx order, signed): compile-time pair count for the two shapes of
// BASELINE.json's configs (M = 52: GloVe-100, M = 32: 128-d rotated to 64). TKB_SCAN_GENERIC=1 forces the generic loop (A/B).
static bool scan_generic()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("TKB_SCAN_GENERIC"); v = (e && atoi(e)) ? 1 : 0; }
    return v != 0;
}

#define TKB_DISPATCH_FAST_AVXS(KERNEL, grid, block, smem, st, ...)                                          \
    do {                                                                                                    \
        const int ph_ = scan_generic() ? 0 : (M == 52 ? 26 : (M == 32 ? 16 : 0));                           \
        if (ph_ == 26)      KERNEL<TKB_ORDER_AVX, true, 26><<<grid, block, smem, st>>>(__VA_ARGS__);        \
        else if (ph_ == 16) KERNEL<TKB_ORDER_AVX, true, 16><<<grid, block, smem, st>>>(__VA_ARGS__);        \
        else                KERNEL<TKB_ORDER_AVX, true, 0><<<grid, block, smem, st>>>(__VA_ARGS__);         \
    } while (0)

#define TKB_DISPATCH_FAST(KERNEL, grid, block, smem, st, ...)                                     \
    do {                                                                                          \
        if (order == TKB_ORDER_AVX) {                                                             \
            if (signd) KERNEL<TKB_ORDER_AVX, true><<<grid, block, smem, st>>>(__VA_ARGS__);       \
            else       KERNEL<TKB_ORDER_AVX, false><<<grid, block, smem, st>>>(__VA_ARGS__);      \
        } else {                                                                                  \
            if (signd) KERNEL<TKB_ORDER_SSE, true><<<grid, block, smem, st>>>(__VA_ARGS__);       \
            else       KERNEL<TKB_ORDER_SSE, false><<<grid, block, smem, st>>>(__VA_ARGS__);      \
        }                                                                                         \
    } while (0)

int launch_codes_to_native(const uint64_t *ref, int64_t n_chunks, int M, void *native, cudaStream_t st)
{
    TKB_REQUIRE(M > 0 && M % 2 == 0 && n_chunks >= 0, "bad extent");
    if (n_chunks == 0) return TKB_OK;
    TKB_REQUIRE(ref && native, "null pointer");
    const int64_t pad = (n_chunks + TILE - 1) / TILE * TILE, total = pad * (M / 2);
    const int64_t blocks = (total + 255) / 256;
    TKB_REQUIRE(blocks <= 0x7fffffff, "too many chunks");
    to_native_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)ref, n_chunks, pad, M / 2, (uint4 *)native);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

int launch_codes_from_native(const void *native, int64_t n_chunks, int M, uint64_t *ref, cudaStream_t st)
{
    TKB_REQUIRE(M > 0 && M % 2 == 0 && n_chunks >= 0, "bad extent");
    if (n_chunks == 0) return TKB_OK;
    TKB_REQUIRE(ref && native, "null pointer");
    const int64_t total = n_chunks * (M / 2), blocks = (total + 255) / 256;
    TKB_REQUIRE(blocks <= 0x7fffffff, "too many chunks");
    from_native_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)native, n_chunks, M / 2, (uint4 *)ref);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

// The workspace only carries a statistic now: its first 8 bytes count the chunks whose certificate failed (they are
// recomputed inside the scan kernel). NULL is allowed.
static int stat_counter(void *workspace, int64_t workspace_bytes, unsigned long long **stat, cudaStream_t st)
{
    *stat = nullptr;
    if (!workspace || workspace_bytes < 16) return TKB_OK;
    TKB_REQUIRE((uintptr_t)workspace % 16 == 0, "workspace must be 16-byte aligned");
    *stat = reinterpret_cast<unsigned long long *>(workspace);
    TKB_CUDA(cudaMemsetAsync(workspace, 0, 16, st));
    return TKB_OK;
}

int launch_estimate_native(const void *native, int64_t n_chunks, int M, const uint8_t *tables, int Q, uint8_t *est,
                           int64_t est_stride, int order, int signd, void *workspace, int64_t workspace_bytes,
                           cudaStream_t st)
{
    if (int rc = check_fast_args(M, order)) return rc;
    TKB_REQUIRE(n_chunks >= 0 && Q >= 0, "negative extent");
    if (n_chunks == 0 || Q == 0) return TKB_OK;
    TKB_REQUIRE(native && tables && est, "null pointer");
    TKB_REQUIRE(n_chunks <= 0xffffffffLL, "too many chunks for one launch");
    TKB_REQUIRE(est_stride >= 16 * n_chunks && est_stride % 16 == 0, "est_stride must be a multiple of 16 and >= 16*n_chunks");
    TKB_REQUIRE(((uintptr_t)native % 16 == 0) && ((uintptr_t)est % 16 == 0) && ((uintptr_t)tables % 16 == 0),
                "device pointers must be 16-byte aligned");
    unsigned long long *stat;
    if (int rc = stat_counter(workspace, workspace_bytes, &stat, st)) return rc;
    // short code arrays (the PQ-encoded centroids of an IVF index): a CTA only as wide as the array
    const int threads = n_chunks >= FAST_THREADS ? FAST_THREADS : (int)((n_chunks + 31) / 32 * 32 < 64 ? 64 : (n_chunks + 31) / 32 * 32);
    int64_t tiles = (n_chunks + threads - 1) / threads;
    if (tiles > 148 * 6) tiles = 148 * 6;                          // grid-stride beyond that: the LUT prologue is paid per CTA
    if (tiles > 32768) tiles = 32768;                              // tiles * queries stays below 2^31
    const size_t smem = fast_smem_bytes(M, 0);
    const uint4 *n4 = reinterpret_cast<const uint4 *>(native);
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        const unsigned grid = (unsigned)(tiles * qn);
        if (order == TKB_ORDER_AVX && signd)
            TKB_DISPATCH_FAST_AVXS(estimate_fast_kernel, grid, threads, smem, st, n4, n_chunks, M,
                                   tables + (size_t)q0 * M * 16, est + (size_t)q0 * est_stride, est_stride, stat, qn);
        else
            TKB_DISPATCH_FAST(estimate_fast_kernel, grid, threads, smem, st, n4, n_chunks, M,
                              tables + (size_t)q0 * M * 16, est + (size_t)q0 * est_stride, est_stride, stat, qn);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

int launch_ivf_scan_native(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                           const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est,
                           int64_t slot_stride, const int64_t *seg_off, int64_t max_chunks_per_query, int order, int signd,
                           void *workspace, int64_t workspace_bytes, cudaStream_t st, uint8_t *cmin,
                           const int64_t *cm_home, int q_per_rank, const uint8_t *skip_q)
{
    if (int rc = check_fast_args(M, order)) return rc;
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    TKB_REQUIRE(!cmin || (est && seg_off), "chunk minima need a compact plan relative to est");
    TKB_REQUIRE(!cm_home || (!est && seg_off && q_per_rank > 0 && !cmin), "per-home minima tables belong to the push exchange (est == NULL)");
    if (Q == 0 || P == 0 || (slot_stride == 0 && !seg_off)) return TKB_OK;
    // est == NULL with a plan: the plan holds absolute addresses (TKB_PLAN_PUSH: segments land in the receive buffers of
    // the queries' home ranks, peer-mapped over NVLink)
    TKB_REQUIRE(native && list_chunk_off && tables && probes && (est || seg_off), "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0, "slot_stride must be a multiple of 16");
    TKB_REQUIRE(P <= 4096, "too many probes");
    TKB_REQUIRE((int64_t)Q * P <= 0xffffffffLL, "too many (query, probe) units for one launch");
    unsigned long long *stat;
    if (int rc = stat_counter(workspace, workspace_bytes, &stat, st)) return rc;
    // enough CTAs to fill the machine when Q is small; otherwise one CTA per query walks all its lists
    if (max_chunks_per_query <= 0) max_chunks_per_query = (int64_t)P * (slot_stride / 16);
    // 128-thread CTAs: a query's ~700 chunks quantise better over 128 threads than over 256 (0.58 -> 0.545 ms measured)
    static int scan_threads = 0;
    if (!scan_threads) { const char *e = getenv("TKB_SCAN_THREADS"); scan_threads = e ? atoi(e) : 128; if (scan_threads != 64 && scan_threads != 256) scan_threads = 128; }
    int64_t splits = (148 * 4 + Q - 1) / Q;
    const int64_t max_splits = (max_chunks_per_query + scan_threads - 1) / scan_threads;
    // long probe lists (100M-vector indexes: tens of thousands of chunks per query): several CTAs per query, so that the
    // grid is many waves of similar CTAs instead of a few waves whose length is the longest query
    const int64_t long_splits = max_chunks_per_query / (32 * scan_threads);
    if (splits < long_splits) splits = long_splits;
    if (splits > max_splits) splits = max_splits;
    if (splits > 1024) splits = 1024;
    // the launch that follows the tensor-core scan: almost every query is marked in skip_q and its CTAs leave at once; 620 000
    // empty CTAs cost 0.23 ms at 10 000 queries (1.8 ms at 80 000), so the few queries left get two CTAs each
    if (skip_q && splits > 2) splits = 2;
    if (splits < 1) splits = 1;
    const size_t smem = fast_smem_bytes(M, P);
    const uint4 *n4 = reinterpret_cast<const uint4 *>(native);
    // TKB_SCAN_KEEP=1 pins flagged chunks in L2 until the deferred pass. Measured on the 100M x 128 index (ncu, r1d): DRAM reads
    // 96.6 -> 87.7 GB per launch (= the algorithmic bytes), kernel 16.98 -> 16.43 ms in isolation, but no gain inside a step
    // (the evict_last lines pile up in L2 across launches): off by default.
    static int keep = -1;
    if (keep < 0) { const char *e = getenv("TKB_SCAN_KEEP"); keep = (e && atoi(e) == 1) ? 1 : 0; }
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        dim3 grid((unsigned)splits, (unsigned)qn);
        const int64_t *so = seg_off ? seg_off + (size_t)q0 * P : nullptr;          // offsets stay relative to `est`
        uint8_t *eb = seg_off ? est : est + (size_t)q0 * P * slot_stride;
        if (order == TKB_ORDER_AVX && signd)
            TKB_DISPATCH_FAST_AVXS(ivf_scan_fast_kernel, grid, scan_threads, smem, st, n4, list_chunk_off, list_size, n_lists, M,
                                   tables + (size_t)q0 * M * 16, probes + (size_t)q0 * P, P, eb, slot_stride, so, stat, keep, cmin, cm_home, q_per_rank, q0,
                                   skip_q ? skip_q + q0 : nullptr);
        else
            TKB_DISPATCH_FAST(ivf_scan_fast_kernel, grid, scan_threads, smem, st, n4, list_chunk_off, list_size, n_lists, M,
                              tables + (size_t)q0 * M * 16, probes + (size_t)q0 * P, P, eb, slot_stride, so, stat, keep, cmin, cm_home, q_per_rank, q0,
                                   skip_q ? skip_q + q0 : nullptr);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

}  // namespace tkb
