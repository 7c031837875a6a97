// tkb_plan.cu -- segment planning for compact estimate buffers and for the multi-GPU exchange.
//
// The reference scans the probed lists one after the other inside IVF.query (ref: tinyknn/ivf.py:140-150)
// and never stores estimates. Here the scan of all (query, probed list) segments is one launch whose
// output, one byte per scanned vector, is consumed by the replay launch. This file decides WHERE each
// segment lives:
//   * one GPU: segments packed back to back in (query, probe slot) order (no slot_stride padding);
//   * lists sharded over ranks: a scanning rank packs the segments of the lists IT owns grouped by the home
//     rank of the query (the send buffer of one all-to-all), and a home rank lays the segments of ITS queries
//     out grouped by the rank that owns the list (the receive buffer). Inside a (scanning rank, home rank)
//     pair both sides order the segments by (query, probe slot), so no offsets ever travel.
// A segment is 16 * ceil(list_size / 16) bytes: the reference's padding of a list to whole chunks.
#include "tkb_common.cuh"

namespace tkb {

constexpr int PLAN_MAX_RANKS = 16;
constexpr int PLAN_CTA = 128;           // queries per CTA of the count / write kernels
constexpr int TKB_PLAN_PULL_ = 3;       // internal: the home side of the pull exchange (tkb_ivf_plan_pull_home_dev)

struct PlanArgs {
    const int32_t *probes;        // [Q][P] (global query numbering)
    const int32_t *list_size;     // [n_lists]
    const int32_t *list_owner;    // [n_lists] or null: every list is local
    int Q, P, n_lists, mode, rank, n_ranks, q_per_rank;
    const int64_t *home_base;     // TKB_PLAN_PUSH: [n_ranks] address of each home rank's receive buffer, as mapped here
                                  // pull, home side: [n_ranks] address of each OWNER's estimate buffer, as mapped here
    const int64_t *send_bases;    // pull, home side: [n_ranks][n_ranks + 1]; row r = (total, bases of the home groups) of owner r's buffer
    int64_t capacity;             // TKB_PLAN_SEND with capacity > 0: a segment that would end past it is dropped (offset -1)
};

__device__ __forceinline__ int64_t plan_seg(const PlanArgs &a, int q, int s, int &group, bool *mine = nullptr)
{
    // returns the segment's bytes (0 when absent from this rank's buffer) and its group
    group = 0;
    int l = a.probes[(size_t)q * a.P + s];
    if (l == PROBE_SKIP) return 0;
    if (l < 0) l += a.n_lists;
    const int owner = a.list_owner ? a.list_owner[l] : 0;
    if (a.mode == TKB_PLAN_PUSH) {
        // the home rank's buffer holds EVERY segment of its queries in (query, slot) order, whoever scans it;
        // this rank writes only the segments of the lists it owns
        group = a.n_ranks > 1 ? q / a.q_per_rank : 0;
        if (mine) *mine = !a.list_owner || owner == a.rank;
    } else if (a.mode == TKB_PLAN_SEND) {
        if (a.list_owner && owner != a.rank) return 0;
        group = a.n_ranks > 1 ? q / a.q_per_rank : 0;
    } else {
        group = owner;
    }
    return 16 * (((int64_t)a.list_size[l] + 15) >> 4);
}

// queries handled by this rank in `mode`: all of them when sending, the home block when receiving
__device__ __forceinline__ void plan_range(const PlanArgs &a, int &q_lo, int &q_n)
{
    if ((a.mode == TKB_PLAN_RECV || a.mode == TKB_PLAN_PULL_) && a.n_ranks > 1) {
        q_lo = a.rank * a.q_per_rank; q_n = min(a.q_per_rank, a.Q - q_lo); if (q_n < 0) q_n = 0;
    } else { q_lo = 0; q_n = a.Q; }
}

__device__ __forceinline__ void plan_query_totals(const PlanArgs &a, int q, int64_t (&tot)[PLAN_MAX_RANKS])
{
#pragma unroll
    for (int g = 0; g < PLAN_MAX_RANKS; g++) tot[g] = 0;
    for (int s = 0; s < a.P; s++) {
        int g;
        const int64_t b = plan_seg(a, q, s, g);
#pragma unroll
        for (int k = 0; k < PLAN_MAX_RANKS; k++) if (k == g) tot[k] += b;
    }
}

// phase 1: bytes per (query, group), left as the EXCLUSIVE prefix over the queries of the CTA (the first query of a CTA holds 0:
// phase 2 puts the CTA's base there)
__global__ void __launch_bounds__(PLAN_CTA) plan_count_kernel(PlanArgs a, int64_t *__restrict__ qtot /* [q_n][n_ranks] */)
{
    __shared__ int64_t s_w[PLAN_CTA / 32];
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int i = blockIdx.x * PLAN_CTA + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t tot[PLAN_MAX_RANKS];
    if (i < q_n) plan_query_totals(a, q_lo + i, tot);
    else {
#pragma unroll
        for (int g = 0; g < PLAN_MAX_RANKS; g++) tot[g] = 0;
    }
#pragma unroll
    for (int g = 0; g < PLAN_MAX_RANKS; g++) {
        if (g >= a.n_ranks) break;                               // uniform
        int64_t incl = tot[g];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int64_t before = 0;
        for (int w = 0; w < warp; w++) before += s_w[w];
        if (i < q_n) qtot[(size_t)i * a.n_ranks + g] = before + incl - tot[g];
        __syncthreads();
    }
}

// phase 2 (one CTA): exclusive scan over the CTAs of phase 1 per group, then the group bases. The total of a phase-1 CTA is the
// prefix of its last query plus that query's own bytes (recomputed: P probes); its base goes into the slot of its first query.
__global__ void __launch_bounds__(1024) plan_scan_kernel(PlanArgs a, int64_t *__restrict__ qtot, int64_t *__restrict__ group_bytes)
{
    __shared__ int64_t part[1024];
    __shared__ int64_t gbase[PLAN_MAX_RANKS + 1];
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int tid = threadIdx.x;
    const int n_blocks = (q_n + PLAN_CTA - 1) / PLAN_CTA;
    const int per = (n_blocks + 1023) / 1024;                      // phase-1 CTAs per thread (1 up to 131 072 queries)
    const int lo = min(n_blocks, tid * per), hi = min(n_blocks, lo + per);
    for (int g = 0; g < a.n_ranks; g++) {
        int64_t sum = 0;
        for (int b = lo; b < hi; b++) {
            const int last = min(q_n, (b + 1) * PLAN_CTA) - 1;
            int64_t t = qtot[(size_t)last * a.n_ranks + g];
            for (int s = 0; s < a.P; s++) { int gg; const int64_t by = plan_seg(a, q_lo + last, s, gg); if (gg == g) t += by; }
            sum += t;
        }
        part[tid] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {                      // Hillis-Steele inclusive scan
            const int64_t v = tid >= o ? part[tid - o] : 0;
            __syncthreads();
            part[tid] += v;
            __syncthreads();
        }
        int64_t run = part[tid] - sum;                            // exclusive prefix of this thread's CTAs
        if (tid == 1023) gbase[g + 1] = part[1023];
        for (int b = lo; b < hi; b++) {
            const int last = min(q_n, (b + 1) * PLAN_CTA) - 1;
            int64_t t = qtot[(size_t)last * a.n_ranks + g];
            for (int s = 0; s < a.P; s++) { int gg; const int64_t by = plan_seg(a, q_lo + last, s, gg); if (gg == g) t += by; }
            qtot[(size_t)b * PLAN_CTA * a.n_ranks + g] = run;
            run += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int64_t run = 0;
        for (int g = 0; g < a.n_ranks; g++) {
            const int64_t b = gbase[g + 1];
            group_bytes[g] = b;
            gbase[g] = run;
            run += b;
        }
        group_bytes[a.n_ranks] = run;                              // total
        for (int g = 0; g < a.n_ranks; g++) group_bytes[a.n_ranks + 1 + g] = gbase[g];
    }
}

// phase 3: segment offsets
__global__ void __launch_bounds__(PLAN_CTA)
plan_write_kernel(PlanArgs a, const int64_t *__restrict__ qbase, const int64_t *__restrict__ group_bytes,
                  int64_t *__restrict__ seg_off /* [q_n][P] */)
{
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int i = blockIdx.x * PLAN_CTA + threadIdx.x;
    if (i >= q_n) return;
    const bool push = a.mode == TKB_PLAN_PUSH, pull = a.mode == TKB_PLAN_PULL_;
    int64_t run[PLAN_MAX_RANKS];
#pragma unroll
    for (int g = 0; g < PLAN_MAX_RANKS; g++) {
        run[g] = 0;
        if (g < a.n_ranks) {
            // push: absolute addresses inside the home rank's buffer; pull: inside the owner's buffer, where the segments of
            // this rank's queries start at the owner's base of home group `rank`
            const int64_t base = push ? a.home_base[g]
                               : pull ? a.home_base[g] + a.send_bases[(size_t)g * (a.n_ranks + 1) + 1 + a.rank]
                                      : group_bytes[a.n_ranks + 1 + g];
            run[g] = base + qbase[(size_t)blockIdx.x * PLAN_CTA * a.n_ranks + g] + (threadIdx.x ? qbase[(size_t)i * a.n_ranks + g] : 0);
        }
    }
    const int64_t cap = (a.mode == TKB_PLAN_SEND && a.capacity > 0) ? a.capacity : INT64_MAX;
    for (int s = 0; s < a.P; s++) {
        int g;
        bool mine = true;
        const int64_t b = plan_seg(a, q_lo + i, s, g, &mine);
        int64_t off = -1;
        if (b > 0) {
#pragma unroll
            for (int k = 0; k < PLAN_MAX_RANKS; k++) if (k == g) { off = run[k]; run[k] += b; }
            if (!mine || off + b > cap) off = -1;
        }
        seg_off[(size_t)i * a.P + s] = off;
    }
}

int launch_ivf_plan(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                    int mode, int rank, int n_ranks, int q_per_rank, const int64_t *home_base, int64_t *seg_off,
                    int64_t *group_bytes, void *workspace, int64_t workspace_bytes, cudaStream_t st,
                    const int64_t *send_bases, int64_t capacity)
{
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    TKB_REQUIRE(mode == TKB_PLAN_SEND || mode == TKB_PLAN_RECV || mode == TKB_PLAN_PUSH || mode == TKB_PLAN_PULL_,
                "mode must be TKB_PLAN_SEND, _RECV or _PUSH");
    TKB_REQUIRE(mode != TKB_PLAN_PUSH || home_base, "TKB_PLAN_PUSH needs the receive-buffer addresses (home_base)");
    TKB_REQUIRE(mode != TKB_PLAN_PULL_ || (home_base && send_bases), "the pull plan needs the owners' buffer addresses and group bases");
    TKB_REQUIRE(n_ranks >= 1 && n_ranks <= PLAN_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank / n_ranks");
    TKB_REQUIRE(n_ranks == 1 || (q_per_rank > 0 && (int64_t)q_per_rank * n_ranks >= Q), "q_per_rank * n_ranks must cover Q");
    TKB_REQUIRE(n_ranks == 1 || list_owner, "list_owner is required when lists are sharded");
    TKB_REQUIRE(group_bytes, "null pointer");
    PlanArgs a{probes, list_size, list_owner, Q, P, n_lists, mode, rank, n_ranks, q_per_rank, home_base, send_bases, capacity};
    int q_n = Q;
    if ((mode == TKB_PLAN_RECV || mode == TKB_PLAN_PULL_) && n_ranks > 1) {
        q_n = Q - rank * q_per_rank; if (q_n > q_per_rank) q_n = q_per_rank; if (q_n < 0) q_n = 0;
    }
    if (q_n == 0 || P == 0) {
        TKB_CUDA(cudaMemsetAsync(group_bytes, 0, sizeof(int64_t) * (2 * n_ranks + 1), st));
        return TKB_OK;
    }
    TKB_REQUIRE(probes && list_size && seg_off, "null pointer");
    TKB_REQUIRE(workspace && workspace_bytes >= (int64_t)sizeof(int64_t) * q_n * n_ranks, "plan workspace too small (8 * queries * n_ranks bytes)");
    int64_t *qtot = reinterpret_cast<int64_t *>(workspace);
    const unsigned blocks = (unsigned)((q_n + PLAN_CTA - 1) / PLAN_CTA);
    plan_count_kernel<<<blocks, PLAN_CTA, 0, st>>>(a, qtot);
    TKB_LAUNCH_CHECK();
    plan_scan_kernel<<<1, 1024, 0, st>>>(a, qtot, group_bytes);
    TKB_LAUNCH_CHECK();
    plan_write_kernel<<<blocks, PLAN_CTA, 0, st>>>(a, qtot, group_bytes, seg_off);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

// ---- pull exchange: the chunk minima of a home rank's queries, fetched from the owners' buffers ---------------------------------
// One warp per (query, probe slot): the minima of the segment (one byte per chunk) are copied from the owner's minima region
// (addressed, like everywhere, by the estimate address >> 4 through a per-owner table) to their place in the home rank's own
// compact stream layout, where the replay reads them 16 chunks per load. Source and destination are byte-aligned only.
__global__ void __launch_bounds__(256)
pull_minima_kernel(const int32_t *__restrict__ probes, int64_t n_seg, const int32_t *__restrict__ list_size,
                   const int32_t *__restrict__ list_owner, int n_lists, const int64_t *__restrict__ seg_addr,
                   const int64_t *__restrict__ seg_local, const int64_t *__restrict__ cm_table, uint8_t *__restrict__ cmin_local)
{
    const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= n_seg) return;
    int l = probes[e];
    const int64_t src_a = seg_addr[e], dst_o = seg_local[e];
    if (l == PROBE_SKIP || src_a < 0 || dst_o < 0) return;
    if (l < 0) l += n_lists;
    const int n = (list_size[l] + 15) >> 4;
    const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(cm_table[list_owner ? list_owner[l] : 0] + (src_a >> 4)));
    uint8_t *dst = cmin_local + (dst_o >> 4);
    // head bytes up to the first 4-byte boundary of the destination, then destination-aligned words assembled from the two
    // aligned source words that hold their bytes (the second is the next lane's first: one 128-byte request per 32 words), then the tail
    const int head = min(n, (int)((4 - ((uintptr_t)dst & 3)) & 3));
    if (lane < head) dst[lane] = src[lane];
    const int n_words = (n - head) >> 2;
    const uint8_t *sa = src + head;
    const int mis = (int)((uintptr_t)sa & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sa - mis);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
    constexpr int U = 8;                                           // 1 KB of a segment in flight per warp (the loads cross NVLink)
    for (int j0 = 0; j0 < n_words; j0 += 32 * U) {
        // every source word is loaded ONCE: the upper word a destination word needs is the next lane's (the next row's first
        // lane's) lower word, passed by shuffle -- loads from a peer GPU are not cached in L1, so a second, shifted load of the
        // same line would cross NVLink again
        uint32_t lo[U + 1];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            lo[u] = j < n_words + (mis ? 1 : 0) ? sw[j] : 0u;        // (word n_words holds the last bytes when the source is shifted)
        }
        {
            const int j = j0 + 32 * U;                              // first word of the next block: the last lane's upper word
            lo[U] = (mis && j <= n_words) ? sw[j] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            const uint32_t nxt_lane = __shfl_down_sync(FULL, lo[u], 1);
            const uint32_t nxt_row = __shfl_sync(FULL, lo[u + 1], 0);
            const uint32_t hi = lane == 31 ? nxt_row : nxt_lane;
            if (j < n_words) dw[j] = mis ? (lo[u] >> (8 * mis)) | (hi << (32 - 8 * mis)) : lo[u];
        }
    }
    const int done = head + 4 * n_words;
    if (lane < n - done) dst[done + lane] = src[done + lane];
}

int launch_pull_minima(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                       const int64_t *seg_addr, const int64_t *seg_local, const int64_t *cm_table, uint8_t *cmin_local,
                       cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || P == 0) return TKB_OK;
    TKB_REQUIRE(probes && list_size && seg_addr && seg_local && cm_table && cmin_local, "null pointer");
    const int64_t n_seg = (int64_t)Q * P;
    pull_minima_kernel<<<(unsigned)((n_seg + 7) / 8), 256, 0, st>>>(probes, n_seg, list_size, list_owner, n_lists, seg_addr,
                                                                    seg_local, cm_table, cmin_local);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
