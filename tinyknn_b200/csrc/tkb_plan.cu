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

struct PlanArgs {
    const int32_t *probes;        // [Q][P] (global query numbering)
    const int32_t *list_size;     // [n_lists]
    const int32_t *list_owner;    // [n_lists] or null: every list is local
    int Q, P, n_lists, mode, rank, n_ranks, q_per_rank;
    const int64_t *home_base;     // TKB_PLAN_PUSH: [n_ranks] address of each home rank's receive buffer, as mapped here
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
    if (a.mode == TKB_PLAN_RECV && a.n_ranks > 1) { q_lo = a.rank * a.q_per_rank; q_n = min(a.q_per_rank, a.Q - q_lo); if (q_n < 0) q_n = 0; }
    else { q_lo = 0; q_n = a.Q; }
}

// phase 1: bytes per (query, group)
__global__ void plan_count_kernel(PlanArgs a, int64_t *__restrict__ qtot /* [q_n][n_ranks] */)
{
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q_n) return;
    int64_t tot[PLAN_MAX_RANKS];
#pragma unroll
    for (int g = 0; g < PLAN_MAX_RANKS; g++) tot[g] = 0;
    for (int s = 0; s < a.P; s++) {
        int g;
        const int64_t b = plan_seg(a, q_lo + i, s, g);
#pragma unroll
        for (int k = 0; k < PLAN_MAX_RANKS; k++) if (k == g) tot[k] += b;
    }
    for (int g = 0; g < a.n_ranks; g++) qtot[(size_t)i * a.n_ranks + g] = tot[g];
}

// phase 2 (one CTA): exclusive scan over the queries per group, then the group bases; qtot becomes qbase
__global__ void __launch_bounds__(1024) plan_scan_kernel(PlanArgs a, int64_t *__restrict__ qtot, int64_t *__restrict__ group_bytes)
{
    __shared__ int64_t part[1024];
    __shared__ int64_t gbase[PLAN_MAX_RANKS + 1];
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int tid = threadIdx.x;
    const int per = (q_n + 1023) / 1024;
    const int lo = min(q_n, tid * per), hi = min(q_n, lo + per);
    for (int g = 0; g < a.n_ranks; g++) {
        int64_t sum = 0;
        for (int i = lo; i < hi; i++) sum += qtot[(size_t)i * a.n_ranks + g];
        part[tid] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {                      // Hillis-Steele inclusive scan
            const int64_t v = tid >= o ? part[tid - o] : 0;
            __syncthreads();
            part[tid] += v;
            __syncthreads();
        }
        int64_t run = part[tid] - sum;                            // exclusive prefix of this thread's block
        if (tid == 1023) gbase[g + 1] = part[1023];
        for (int i = lo; i < hi; i++) {
            const int64_t b = qtot[(size_t)i * a.n_ranks + g];
            qtot[(size_t)i * a.n_ranks + g] = run;
            run += b;
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
__global__ void plan_write_kernel(PlanArgs a, const int64_t *__restrict__ qbase, const int64_t *__restrict__ group_bytes,
                                  int64_t *__restrict__ seg_off /* [q_n][P] */)
{
    int q_lo, q_n;
    plan_range(a, q_lo, q_n);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q_n) return;
    const bool push = a.mode == TKB_PLAN_PUSH;
    int64_t run[PLAN_MAX_RANKS];
#pragma unroll
    for (int g = 0; g < PLAN_MAX_RANKS; g++)                      // push: absolute addresses inside the home rank's buffer
        run[g] = g < a.n_ranks ? (push ? a.home_base[g] : group_bytes[a.n_ranks + 1 + g]) + qbase[(size_t)i * a.n_ranks + g] : 0;
    for (int s = 0; s < a.P; s++) {
        int g;
        bool mine = true;
        const int64_t b = plan_seg(a, q_lo + i, s, g, &mine);
        int64_t off = -1;
        if (b > 0) {
#pragma unroll
            for (int k = 0; k < PLAN_MAX_RANKS; k++) if (k == g) { off = run[k]; run[k] += b; }
            if (!mine) off = -1;
        }
        seg_off[(size_t)i * a.P + s] = off;
    }
}

int launch_ivf_plan(const int32_t *probes, int Q, int P, const int32_t *list_size, const int32_t *list_owner, int n_lists,
                    int mode, int rank, int n_ranks, int q_per_rank, const int64_t *home_base, int64_t *seg_off,
                    int64_t *group_bytes, void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    TKB_REQUIRE(mode == TKB_PLAN_SEND || mode == TKB_PLAN_RECV || mode == TKB_PLAN_PUSH, "mode must be TKB_PLAN_SEND, _RECV or _PUSH");
    TKB_REQUIRE(mode != TKB_PLAN_PUSH || home_base, "TKB_PLAN_PUSH needs the receive-buffer addresses (home_base)");
    TKB_REQUIRE(n_ranks >= 1 && n_ranks <= PLAN_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank / n_ranks");
    TKB_REQUIRE(n_ranks == 1 || (q_per_rank > 0 && (int64_t)q_per_rank * n_ranks >= Q), "q_per_rank * n_ranks must cover Q");
    TKB_REQUIRE(n_ranks == 1 || list_owner, "list_owner is required when lists are sharded");
    TKB_REQUIRE(group_bytes, "null pointer");
    PlanArgs a{probes, list_size, list_owner, Q, P, n_lists, mode, rank, n_ranks, q_per_rank, home_base};
    int q_n = Q;
    if (mode == TKB_PLAN_RECV && n_ranks > 1) { q_n = Q - rank * q_per_rank; if (q_n > q_per_rank) q_n = q_per_rank; if (q_n < 0) q_n = 0; }
    if (q_n == 0 || P == 0) {
        TKB_CUDA(cudaMemsetAsync(group_bytes, 0, sizeof(int64_t) * (2 * n_ranks + 1), st));
        return TKB_OK;
    }
    TKB_REQUIRE(probes && list_size && seg_off, "null pointer");
    TKB_REQUIRE(workspace && workspace_bytes >= (int64_t)sizeof(int64_t) * q_n * n_ranks, "plan workspace too small (8 * queries * n_ranks bytes)");
    int64_t *qtot = reinterpret_cast<int64_t *>(workspace);
    const unsigned blocks = (unsigned)((q_n + 127) / 128);
    plan_count_kernel<<<blocks, 128, 0, st>>>(a, qtot);
    TKB_LAUNCH_CHECK();
    plan_scan_kernel<<<1, 1024, 0, st>>>(a, qtot, group_bytes);
    TKB_LAUNCH_CHECK();
    plan_write_kernel<<<blocks, 128, 0, st>>>(a, qtot, group_bytes, seg_off);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
