// tkb_coarse.cu -- probe selection of IVF.query as ONE kernel (opt-in until it has been timed on hardware).
//
// Replaces `dtable.top(pq_transformed_centers, active_centers, k=n_probes)` (ref: tinyknn/ivf.py:131 ->
// tinyknn/fast_pq.py:284-312) for a batch of queries: per query
//   1. the 4-bit scan of the PQ-encoded centroids (query_pq_*'s scan half; the chunk kernels of tkb_scan_core.cuh, so the
//      estimates are the reference's bit for bit),
//   2. the reference heap of R = min(2 * n_probes + 10, C) slots replayed exactly (init_heap + query_pq_*'s heap half:
//      bound frozen per 16-vector chunk, strict <, ascending position, replace-root + sift-down, label = position),
//   3. the exact distances of the R candidates to the raw centroids (knn_brute1's arithmetic, tkb_rescore_core.cuh: the same
//      bits as tkb_gather_dists_dev; a -1 heap slot reads the last centroid like numpy's negative index does),
//   4. the P nearest in device order (ascending distance, NaN last, ties by heap slot) = tkb_select_probes_dev.
// The staged path launches five kernels for this (estimate_fast, replay_rq2, gather_dists, select_probes, plus the
// estimates' round trip through global memory), 27 % of a GloVe-shape step for a problem of 68 chunks and a 30-slot heap per
// query. Here one CTA of three warps owns a query: the estimates, the heap and the distances never leave shared memory.
//
// Why a single lane replays the heap: with R ~ 30 an insert is <= 5 levels, the stream is ~1 100 vectors of which ~140 are
// inserted, and there are 10 000 independent queries in a batch -- thread-level parallelism across CTAs hides the
// shared-memory latency of the sift chain, which the big replay (tkb_heap.cu) has to hide by pipelining inside a query.
#include "tkb_rescore_core.cuh"
#include "tkb_scan_core.cuh"

namespace tkb {

namespace {

constexpr int CO_THREADS = 96;
constexpr int CO_MAX_R = 1024;

struct CoarseArgs {
    const uint4 *nat;            // PQ-encoded centroids, device-native layout
    int n_chunks, C, M;
    const uint8_t *tables;       // [Q][M][16]
    const float *centers;        // [C][d]
    int d;
    const float *queries;        // [Q][d]
    int R, P;
    int32_t *probes;             // [Q][P]
    int64_t *heap_idx;           // optional [Q][R]
    int32_t *heap_val;           // optional [Q][R]
    float *dists;                // optional [Q][R]
};

struct CoarseSmem {
    uint4 *rows, *raw;
    uint2 *sc;
    LutMeta *meta;
    int *scratch;
    uint4 *est;                  // [n_chunks] 16 estimates each
    int *hval, *hpos;            // [R]
    float *dist;                 // [R]
    unsigned char *taken;        // [R]
};

__host__ __device__ inline size_t coarse_carve(unsigned char *base, int M, int n_chunks, int R, CoarseSmem *out)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
    const size_t rows = take((size_t)M * 16), raw = take((size_t)M * 16), sc = take((size_t)M * 8);
    const size_t meta = take(sizeof(LutMeta)), scratch = take((size_t)M * 16);
    const size_t est = take((size_t)n_chunks * 16), hv = take((size_t)R * 4), hp = take((size_t)R * 4);
    const size_t ds = take((size_t)R * 4), tk = take((size_t)R);
    if (out) {
        out->rows = reinterpret_cast<uint4 *>(base + rows); out->raw = reinterpret_cast<uint4 *>(base + raw);
        out->sc = reinterpret_cast<uint2 *>(base + sc); out->meta = reinterpret_cast<LutMeta *>(base + meta);
        out->scratch = reinterpret_cast<int *>(base + scratch); out->est = reinterpret_cast<uint4 *>(base + est);
        out->hval = reinterpret_cast<int *>(base + hv); out->hpos = reinterpret_cast<int *>(base + hp);
        out->dist = reinterpret_cast<float *>(base + ds); out->taken = base + tk;
    }
    return o;
}

// replace the root by (pos, v) and sift down (ref: _fast_pq.pyx:290-307): a child moves up when STRICTLY greater than the
// value being placed; the left child is tried first, the right one wins only when strictly greater than the left.
__device__ __forceinline__ void coarse_sift(int *hval, int *hpos, int R, int pos, int v)
{
    int j = 0;
    for (;;) {
        const int l = 2 * j + 1, r = l + 1;
        int nxt = j, nv = v;
        if (l < R) { const int lv = hval[l]; if (lv > nv) { nxt = l; nv = lv; } }
        if (r < R) { const int rv = hval[r]; if (rv > nv) { nxt = r; nv = rv; } }
        if (nxt == j) break;
        hval[j] = nv; hpos[j] = hpos[nxt];
        j = nxt;
    }
    hval[j] = v; hpos[j] = pos;
}

// 4 bytes "est < bound" (signed) -> nibble mask, 16 estimates -> 16-bit mask
__device__ __forceinline__ uint32_t coarse_mask16(const uint4 e, int bound)
{
    const uint32_t b4 = (uint32_t)(bound & 0xff) * 0x01010101u;
    auto pack = [](uint32_t m) { m &= 0x80808080u; return ((m >> 7) | (m >> 14) | (m >> 21) | (m >> 28)) & 0xfu; };
    return pack(__vcmplts4(e.x, b4)) | (pack(__vcmplts4(e.y, b4)) << 4) | (pack(__vcmplts4(e.z, b4)) << 8) |
           (pack(__vcmplts4(e.w, b4)) << 12);
}

template <int ORDER, int PH>
__global__ void __launch_bounds__(CO_THREADS)
coarse_probes_kernel(CoarseArgs a)
{
    extern __shared__ __align__(16) unsigned char co_sm[];
    CoarseSmem sm;
    coarse_carve(co_sm, a.M, a.n_chunks, a.R, &sm);
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Ph = a.M >> 1;
    const int R = a.R;

    // ---- 1. scan of the encoded centroids into shared memory -----------------------------------------------------------
    prepare_lut<ORDER, true>(a.tables + (size_t)q * a.M * 16, a.M, ORDER == TKB_ORDER_AVX, sm.rows, sm.raw, sm.sc, sm.meta,
                             sm.scratch);                                                               // syncs
    const LutMeta m = *sm.meta;
    for (int c = tid; c < a.n_chunks; c += CO_THREADS) {
        uint4 o;
        if (m.eligible) {
            bool flagged;
            o = scan_chunk_fast<true, PH>(a.nat, c, Ph, sm.rows, m, flagged);
            if (flagged) o = scan_chunk_steps_cold<ORDER, true>(a.nat, c, Ph, sm.rows, sm.sc, m);
        } else if (m.steps_ok) {
            o = scan_chunk_steps<ORDER, true>(a.nat, c, Ph, sm.rows, sm.sc, m);
        } else {
            o = scan_chunk_exact_cold<ORDER, true>(a.nat, c, Ph, reinterpret_cast<const uint8_t *>(sm.raw));
        }
        sm.est[c] = o;
    }
    for (int j = tid; j < R; j += CO_THREADS) { sm.hval[j] = 127; sm.hpos[j] = -1; }                    // init_heap, signed
    __syncthreads();

    // ---- 2. exact replay by warp 0: 32 chunks are tested against the current bound at a time; the first chunk that holds a
    //         candidate is processed with the bound frozen (ref: _fast_pq.pyx:153-206), then the rest is tested again --------
    if (warp == 0) {
        int bound = 127;
        const int real_chunks = (a.C + 15) >> 4 < a.n_chunks ? (a.C + 15) >> 4 : a.n_chunks;
        for (int base = 0; base < real_chunks; base += 32) {
            const int c = base + lane;
            const bool live = c < real_chunks;
            uint4 e = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
            if (live) e = sm.est[c];
            const int rem = a.C - 16 * c;                                                               // vectors of the chunk
            const uint32_t valid = !live ? 0u : (rem >= 16 ? 0xffffu : ((1u << rem) - 1u));
            unsigned ball = __ballot_sync(FULL, (coarse_mask16(e, bound) & valid) != 0);
            while (ball) {
                const int src = __ffs(ball) - 1;
                if (lane == src) {                                                                      // the chunk's owner inserts
                    const int frozen = bound;
                    uint32_t mm = coarse_mask16(e, frozen) & valid;
                    while (mm) {
                        const int v = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const uint32_t w = v < 8 ? (v < 4 ? e.x : e.y) : (v < 12 ? e.z : e.w);      // selects, not a local array
                        const int ev = (int)(int8_t)((w >> (8 * (v & 3))) & 0xffu);
                        coarse_sift(sm.hval, sm.hpos, R, 16 * c + v, ev);
                    }
                }
                __syncwarp();
                bound = sm.hval[0];                                                                     // every lane re-reads the root
                ball = __ballot_sync(FULL, lane > src && (coarse_mask16(e, bound) & valid) != 0);
            }
        }
    }
    __syncthreads();

    // ---- 3. exact distances of the R candidates (skipped when R <= P: fast_pq.py:307-308 returns the raw heap) ----------------
    int32_t *out = a.probes + (size_t)q * a.P;
    if (a.heap_idx) for (int j = tid; j < R; j += CO_THREADS) a.heap_idx[(size_t)q * R + j] = sm.hpos[j];
    if (a.heap_val) for (int j = tid; j < R; j += CO_THREADS) a.heap_val[(size_t)q * R + j] = sm.hval[j];
    if (R <= a.P) {
        for (int s = tid; s < a.P; s += CO_THREADS) out[s] = s < R ? sm.hpos[s] : PROBE_SKIP;
        return;
    }
    const float *x = a.queries + (size_t)q * a.d;
    for (int j = warp; j < R; j += CO_THREADS / 32) {
        int row = sm.hpos[j];
        if (row < 0) row += a.C;                                                                        // numpy negative indexing
        const float dj = warp_row_dist<float>(a.centers + (size_t)row * a.d, x, a.d, lane);
        if (lane == 0) { sm.dist[j] = dj; sm.taken[j] = 0; }
    }
    __syncthreads();
    if (a.dists) for (int j = tid; j < R; j += CO_THREADS) a.dists[(size_t)q * R + j] = sm.dist[j];

    // ---- 4. the P nearest: ascending distance, NaN last, ties by heap slot (= select_probes_kernel) ----------------------------
    if (warp == 0) {
        for (int k = 0; k < a.P; k++) {
            float bd = INFINITY;
            int bs = INT32_MAX;
            for (int s = lane; s < R; s += 32) {
                if (sm.taken[s]) continue;
                float v = sm.dist[s];
                if (v != v) v = INFINITY;
                if (v < bd || (v == bd && s < bs)) { bd = v; bs = s; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(FULL, bd, o);
                const int os = __shfl_xor_sync(FULL, bs, o);
                if (od < bd || (od == bd && os < bs)) { bd = od; bs = os; }
            }
            if (lane == 0) { sm.taken[bs] = 1; out[k] = sm.hpos[bs]; }
            __syncwarp();
        }
    }
}

}  // namespace

int launch_coarse_probes(const void *native_centers, int64_t n_chunks, int C, int M, const uint8_t *tables, int Q,
                         const float *centers, int d, const float *queries, int R, int P, int order, int32_t *probes,
                         int64_t *heap_idx, int32_t *heap_val, float *dists, cudaStream_t st)
{
    TKB_REQUIRE(order == TKB_ORDER_SSE || order == TKB_ORDER_AVX, "order must be TKB_ORDER_SSE or TKB_ORDER_AVX");
    TKB_REQUIRE(M > 0 && M % 2 == 0 && M <= 1024, "M (sub-quantizers) must be a positive multiple of 2");
    TKB_REQUIRE(order != TKB_ORDER_AVX || M % 4 == 0, "avx order needs M % 4 == 0 (ref: fast_pq.py:24 dpad)");
    TKB_REQUIRE(Q >= 0 && C > 0 && d > 0 && P > 0 && R > 0, "bad extent");
    TKB_REQUIRE(R <= C && R <= CO_MAX_R, "R (candidates) must be <= the number of centroids and <= 1024");
    TKB_REQUIRE(n_chunks > 0 && n_chunks <= 4096 && 16 * n_chunks >= C, "centroid codes: 1..4096 chunks covering C");
    if (Q == 0) return TKB_OK;
    TKB_REQUIRE(native_centers && tables && centers && queries && probes, "null pointer");
    TKB_REQUIRE(((uintptr_t)native_centers % 16 == 0) && ((uintptr_t)tables % 16 == 0), "device pointers must be 16-byte aligned");
    const size_t smem = coarse_carve(nullptr, M, (int)n_chunks, R, nullptr);
    TKB_REQUIRE(smem <= 200 * 1024, "index too large for the one-kernel probe selection");
    CoarseArgs a{reinterpret_cast<const uint4 *>(native_centers), (int)n_chunks, C, M, tables, centers, d, queries, R, P, probes,
                 heap_idx, heap_val, dists};
#define TKB_COARSE_LAUNCH(ORDER, PH)                                                                                          \
    do {                                                                                                                       \
        TKB_CUDA(cudaFuncSetAttribute(coarse_probes_kernel<ORDER, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        coarse_probes_kernel<ORDER, PH><<<(unsigned)Q, CO_THREADS, smem, st>>>(a);                                             \
    } while (0)
    if (order == TKB_ORDER_AVX) {
        if (M == 52)      TKB_COARSE_LAUNCH(TKB_ORDER_AVX, 26);
        else if (M == 32) TKB_COARSE_LAUNCH(TKB_ORDER_AVX, 16);
        else              TKB_COARSE_LAUNCH(TKB_ORDER_AVX, 0);
    } else {
        TKB_COARSE_LAUNCH(TKB_ORDER_SSE, 0);
    }
#undef TKB_COARSE_LAUNCH
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
