// tkb_heap.cu -- exact replay of the reference's thresholded heap (one warp per query).
//
// Replaces the heap half of query_pq_sse / query_pq_avx (ref: tinyknn/_fast_pq.pyx:153-206,
// tinyknn/_fast_pq_256.pyx:73-123) and insert (ref: tinyknn/_fast_pq.pyx:274-307).
//
// The reference heap is NOT a top-R selection: the admission bound is vals[0] frozen at the start
// of each 16-vector chunk, admitted lanes are inserted in ascending position, and insert() always
// overwrites the root. The result therefore depends on the visiting order and must be replayed
// sequentially. What makes that cheap on a GPU: the bound is monotone non-increasing, so a warp
// can test 32 chunks (512 estimates) against the current bound with byte-SIMD compares and one
// ballot, jump to the first chunk that admits anything, process exactly that chunk with the frozen
// bound, refresh the bound and re-test only the chunks after it. Cost ~ number of inserts, not N.
#include <stdlib.h>

#include "tkb_common.cuh"

namespace tkb {

// byte-wise "est < bound" over 4 packed bytes -> 0xff / 0x00 per byte
template <bool SIGNED>
__device__ __forceinline__ uint32_t cmp_lt4(uint32_t est4, uint32_t bound4)
{
    return SIGNED ? __vcmplts4(est4, bound4) : __vcmpltu4(est4, bound4);
}

template <bool SIGNED>
__device__ __forceinline__ int load_bound(const int32_t *hval)
{
    // ref: _mm_set1_epi8(vals[0]) -- the int is truncated to 8 bits (_fast_pq.pyx:153, :206)
    const int v = *reinterpret_cast<const volatile int32_t *>(hval);
    return SIGNED ? (int)(int8_t)v : (int)(uint8_t)v;
}

// warp-cooperative insert: parallel dedupe over all R slots, then lane 0 replaces the root.
__device__ __forceinline__ void warp_insert(int64_t *hidx, int32_t *hval, int R, int64_t label,
                                            int v, int lane)
{
    bool found = false;
    for (int j = lane; j < R; j += 32) found |= (hidx[j] == label);
    if (__any_sync(FULL, found)) return;
    if (lane == 0) heap_sift_from_root(hidx, hval, R, label, v);
    __syncwarp();
}

// One segment == one query_pq call of the reference.
template <bool SIGNED>
__device__ void replay_segment(const uint8_t *__restrict__ est, int64_t n_chunks, int n,
                               const int64_t *__restrict__ labels, int64_t *hidx, int32_t *hval,
                               int R, int lane)
{
    if (R <= 0) return;
    int bound = load_bound<SIGNED>(hval);
    const uint4 *est4 = reinterpret_cast<const uint4 *>(est);
    for (int64_t base = 0; base < n_chunks; base += 32) {
        const int64_t c = base + lane;
        uint4 e = make_uint4(0, 0, 0, 0);
        const bool live = c < n_chunks;
        if (live) e = est4[c];
        uint32_t b4 = (uint32_t)(bound & 0xff) * 0x01010101u;
        bool any = live && ((cmp_lt4<SIGNED>(e.x, b4) | cmp_lt4<SIGNED>(e.y, b4) |
                             cmp_lt4<SIGNED>(e.z, b4) | cmp_lt4<SIGNED>(e.w, b4)) != 0);
        unsigned ball = __ballot_sync(FULL, any);
        while (ball) {
            const int src = __ffs(ball) - 1;
            // the chunk's 16 estimates, broadcast to the whole warp
            const uint32_t w0 = __shfl_sync(FULL, e.x, src), w1 = __shfl_sync(FULL, e.y, src);
            const uint32_t w2 = __shfl_sync(FULL, e.z, src), w3 = __shfl_sync(FULL, e.w, src);
            const int64_t cpos = 16 * (base + src);
            const int frozen = bound;                     // bound is frozen for the whole chunk
#pragma unroll 1
            for (int v = 0; v < 16; v++) {
                const uint32_t w = (v < 8) ? ((v < 4) ? w0 : w1) : ((v < 12) ? w2 : w3);
                const uint32_t byte = (w >> (8 * (v & 3))) & 0xffu;
                const int ev = SIGNED ? (int)(int8_t)byte : (int)byte;
                const int64_t pos = cpos + v;
                if (ev < frozen && pos < n) {
                    const int64_t label = labels ? labels[pos] : pos;
                    warp_insert(hidx, hval, R, label, ev, lane);
                }
            }
            __syncwarp();
            bound = load_bound<SIGNED>(hval);
            b4 = (uint32_t)(bound & 0xff) * 0x01010101u;
            any = live && lane > src &&
                  ((cmp_lt4<SIGNED>(e.x, b4) | cmp_lt4<SIGNED>(e.y, b4) |
                    cmp_lt4<SIGNED>(e.z, b4) | cmp_lt4<SIGNED>(e.w, b4)) != 0);
            ball = __ballot_sync(FULL, any);
        }
    }
}

constexpr int REPLAY_WARPS = 4;

template <bool SIGNED>
__global__ void __launch_bounds__(32 * REPLAY_WARPS)
replay_kernel(const uint8_t *__restrict__ est, int64_t est_stride, int64_t n_chunks, int n,
              int64_t *heap_idx, int32_t *heap_val, int Q, int R, const int64_t *__restrict__ labels)
{
    const int q = blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5);
    if (q >= Q) return;
    replay_segment<SIGNED>(est + (size_t)q * est_stride, n_chunks, n, labels,
                           heap_idx + (size_t)q * R, heap_val + (size_t)q * R, R, threadIdx.x & 31);
}

// ref: tinyknn/ivf.py:140-150 -- the probed lists are visited in `top` order with labels=ids[cl]
template <bool SIGNED>
__global__ void __launch_bounds__(32 * REPLAY_WARPS)
ivf_replay_kernel(const uint8_t *__restrict__ est, int64_t slot_stride, const int64_t *__restrict__ seg_off,
                  const int64_t *__restrict__ list_chunk_off, const int32_t *__restrict__ list_size,
                  int n_lists, const int64_t *__restrict__ ids, const int32_t *__restrict__ probes,
                  int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R)
{
    const int q = blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int lane = threadIdx.x & 31;
    for (int s = 0; s < P; s++) {
        int l = probes[(size_t)q * P + s];
        if (l == PROBE_SKIP) continue;
        if (l < 0) l += n_lists;
        const int64_t c0 = list_chunk_off[l];
        const int64_t nc = ((int64_t)list_size[l] + 15) >> 4;              // real chunks (tile padding is not stored)
        const int64_t so = seg_off ? seg_off[(size_t)q * P + s] : ((int64_t)q * P + s) * slot_stride;
        if (so < 0) continue;
        replay_segment<SIGNED>(est + so, nc, list_size[l], ids + 16 * c0,
                               heap_idx + (size_t)q * R, heap_val + (size_t)q * R, R, lane);
    }
}

template <bool SIGNED>
__device__ __forceinline__ uint32_t cand_mask16(const uint4 e, int bound)
{
    const uint32_t b4 = (uint32_t)(bound & 0xff) * 0x01010101u;
    const uint32_t m0 = cmp_lt4<SIGNED>(e.x, b4) & 0x80808080u, m1 = cmp_lt4<SIGNED>(e.y, b4) & 0x80808080u;
    const uint32_t m2 = cmp_lt4<SIGNED>(e.z, b4) & 0x80808080u, m3 = cmp_lt4<SIGNED>(e.w, b4) & 0x80808080u;
    // gather the 4 sign bits of each word into a nibble: bit 7,15,23,31 -> 0..3
    auto pack = [](uint32_t m) { return ((m >> 7) | (m >> 14) | (m >> 21) | (m >> 28)) & 0xfu; };
    return pack(m0) | (pack(m1) << 4) | (pack(m2) << 8) | (pack(m3) << 12);
}

// ------------------------------------------------------------------------------------------------
// Queue replay ("rq"): the fresh-heap replay used by the query path.
//
// The sift-down of the reference heap is sequential by nature; giving it a whole warp (above) leaves 31
// lanes idle, giving every query one lane for everything makes each lane walk ALL of its query's chunks
// with 32 unrelated 16-byte loads per step. Here a CTA owns QPC queries and alternates two phases per round:
//   produce : each warp takes whole queries; its 32 lanes read 32 consecutive chunks of the query's
//             estimate stream (one coalesced 512-byte read), compare them with the query's bound AS OF
//             THE START OF THE ROUND and append the surviving (value, payload) records, in stream
//             order, to the query's queue in shared memory. The bound is monotone non-increasing
//             (SURVEY.md H1), so a stale bound admits a superset of the true candidates, never misses one.
//   consume : one LANE per query walks its queue in order and applies the reference rule exactly (bound
//             frozen when the first record of a new chunk arrives, strict <, replace the root, sift
//             down), heaps interleaved by query in shared memory.
// Round windows double (the admission rate after n vectors is ~R/n), so a query needs ~log2(n/R) rounds
// and the lanes only ever touch records that had a real chance. payload = position in the query's stream
// of real chunks; labels are resolved once, at the end. Preconditions: fresh heap (filled here); labels
// unique across a query's segments, so that the reference's label dedupe (ref: _fast_pq.pyx:284-287)
// can never fire -- the host certifies this for the index, and a query whose probe list contains negative
// (Python-wrapped) entries is handed to the warp-per-query kernel through `fallback`.
// mode 0: one segment per query (est row q, n vectors, label = position)      [probe selection, top()]
// mode 1: P segments per query taken from the probe list, labels from `ids`   [IVF.query]
// ------------------------------------------------------------------------------------------------
constexpr int RQ_THREADS = 256;
constexpr uint32_t RQ_EMPTY = 0xffffffffu;        // payload of a heap slot that was never filled

// heap entry / queue record: .x = payload (stream position: 16 * stream chunk + lane), .y = value (int32)
template <bool SIGNED>
__global__ void __launch_bounds__(RQ_THREADS)
replay_rq_kernel(int mode, const uint8_t *__restrict__ est, int64_t stride, const int64_t *__restrict__ seg_off,
                 int64_t n_chunks0, int n0, const int64_t *__restrict__ list_chunk_off,
                 const int32_t *__restrict__ list_size, int n_lists, const int64_t *__restrict__ ids,
                 const int32_t *__restrict__ probes, int Q, int P, int64_t *__restrict__ heap_idx,
                 int32_t *__restrict__ heap_val, int R, int *__restrict__ fallback, int QPC, int QCAP, int LPW)
{
    extern __shared__ __align__(16) unsigned char rq_sm[];
    uint2 *H = reinterpret_cast<uint2 *>(rq_sm);                           // [R+1][QPC] slot j of query t at j*QPC+t; slot R = sentinel
    uint2 *QU = H + (size_t)(R + 1) * QPC;                                 // [QPC][QCAP+1]
    int *cum = reinterpret_cast<int *>(QU + (size_t)QPC * (QCAP + 1));     // [QPC][P+1] real chunks before segment s
    int *s_cursor = cum + (size_t)QPC * (P + 1);
    int *s_seg = s_cursor + QPC, *s_bound = s_seg + QPC, *s_count = s_bound + QPC, *s_round = s_count + QPC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = RQ_THREADS / 32;
    const int q0 = blockIdx.x * QPC;
    const int init = SIGNED ? 127 : 255;

    // ---- set-up: heaps, segment tables ------------------------------------------------------------
    for (int i = tid; i < (R + 1) * QPC; i += RQ_THREADS)
        H[i] = i < R * QPC ? make_uint2(RQ_EMPTY, (uint32_t)init) : make_uint2(RQ_EMPTY, (uint32_t)INT32_MIN);
    for (int t = tid; t < QPC; t += RQ_THREADS) {
        const int q = q0 + t;
        int *c = cum + (size_t)t * (P + 1);
        bool ok = q < Q;
        int run = 0;
        c[0] = 0;
        for (int s = 0; s < P; s++) {
            int nc = 0;
            if (ok) {
                if (mode == 1) {
                    const int l = probes[(size_t)q * P + s];
                    if (l < 0 && l != PROBE_SKIP) ok = false;              // Python-wrapped index: lists may repeat
                    else if (l != PROBE_SKIP) nc = (list_size[l] + 15) >> 4;
                } else {
                    int64_t r = ((int64_t)n0 + 15) >> 4;
                    nc = (int)(r < n_chunks0 ? r : n_chunks0);
                }
            }
            run += nc;
            c[s + 1] = run;
        }
        if (q < Q && fallback) fallback[q] = ok ? 0 : 1;
        if (!ok) for (int s = 0; s <= P; s++) c[s] = 0;                    // nothing to do here
        s_cursor[t] = 0; s_seg[t] = 0; s_bound[t] = init; s_count[t] = 0; s_round[t] = 0;
    }
    __syncthreads();

    for (;;) {
        // ---- produce ------------------------------------------------------------------------------
        bool more = false;
        for (int t = warp; t < QPC; t += n_warps) {
            const int *c = cum + (size_t)t * (P + 1);
            const int total = c[P];
            const int cursor = s_cursor[t];
            if (cursor >= total) { if (lane == 0) s_count[t] = 0; continue; }
            const int q = q0 + t;
            const int bound = s_bound[t];
            // first window: just enough to fill the heap; then the stream position doubles every round
            int W = s_round[t] == 0 ? ((R + 15) >> 4) + 1 : (cursor < 32 ? 32 : cursor);
            if (W > (1 << 16)) W = 1 << 16;
            int end = (total - cursor < W) ? total : cursor + W;
            int count = 0, sg = s_seg[t];
            uint2 *qu = QU + (size_t)t * (QCAP + 1);
            for (int base = cursor; base < end; base += 32) {
                const int cc = base + lane;
                const bool act = cc < end;
                uint32_t m = 0;
                uint4 e = make_uint4(0, 0, 0, 0);
                int sl = sg;
                if (act) {
                    while (cc >= c[sl + 1]) sl++;
                    const int local = cc - c[sl];
                    int n;
                    const uint8_t *ep;
                    if (mode == 1) {
                        n = list_size[probes[(size_t)q * P + sl]];
                        ep = est + (seg_off ? seg_off[(size_t)q * P + sl] : ((int64_t)q * P + sl) * stride);
                    } else {
                        n = n0;
                        ep = est + (int64_t)q * stride;
                    }
                    e = ldg_nc_u4(reinterpret_cast<const uint4 *>(ep) + local);
                    const int rem = n - 16 * local;                        // >= 1 (only real chunks are in the stream)
                    m = cand_mask16<SIGNED>(e, bound) & (rem >= 16 ? 0xffffu : ((1u << rem) - 1u));
                }
                const int last = end - 1 - base;
                sg = __shfl_sync(FULL, sl, last < 31 ? last : 31);         // segment of the last active lane
                const int cnt = __popc(m);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                const int tot = __shfl_sync(FULL, incl, 31);
                bool cut = false;
                int keep_tot = tot;
                if (count + tot > QCAP) {                                  // queue full: stop at a chunk boundary
                    const unsigned over = __ballot_sync(FULL, count + incl > QCAP);
                    const int cl = __ffs(over) - 1;                        // first chunk that does not fit
                    keep_tot = __shfl_sync(FULL, incl - cnt, cl);
                    sg = __shfl_sync(FULL, sl, cl);                        // the next round resumes at chunk base+cl
                    if (lane >= cl) m = 0;
                    end = base + cl;
                    cut = true;
                }
                int k = count + incl - cnt;
                const uint32_t ws[4] = {e.x, e.y, e.z, e.w};
                while (m) {
                    const int v = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t byte = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
                    qu[k++] = make_uint2(16u * (uint32_t)cc + v, SIGNED ? (uint32_t)(int)(int8_t)byte : byte);
                }
                count += keep_tot;
                if (cut) break;
            }
            if (lane == 0) { s_cursor[t] = end; s_seg[t] = sg; s_count[t] = count; s_round[t] = 1; }
            more = true;
        }
        if (!__syncthreads_or(more)) break;
        // ---- consume: warp w replays queries [w*LPW, (w+1)*LPW), one lane each (LPW = a power of two) ----
        if (warp * LPW < QPC) {
            const int t = warp * LPW + lane;
            const bool mine = lane < LPW && t < QPC;
            const int cnt = mine ? s_count[t] : 0;
            int mx = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
            if (mine) {
                const uint2 *qu = QU + (size_t)t * (QCAP + 1);
                unsigned char *hb = reinterpret_cast<unsigned char *>(H + t);   // slot j at hb + j*S
                const uint32_t S = (uint32_t)QPC * 8u, lim = (uint32_t)R * S;
                uint32_t cur_chunk = RQ_EMPTY;
                int frozen = 0;
                for (int i = 0; i < mx; i++) {
                    if (i >= cnt) continue;
                    const uint2 rec = qu[i];
                    const int ev = (int)rec.y;
                    if ((rec.x >> 4) != cur_chunk) {                       // first candidate of a chunk: freeze the bound
                        cur_chunk = rec.x >> 4;
                        frozen = (int)reinterpret_cast<const uint2 *>(hb)->y;
                    }
                    if (ev >= frozen) continue;
                    // replace the root and sift down (ref: _fast_pq.pyx:290-307): the larger child moves up while
                    // it is strictly greater than the new value; the right child wins only when strictly greater
                    // than the left one. Slot R is a sentinel (INT32_MIN), so a right child always exists.
                    uint32_t jo = 0;
                    for (;;) {
                        const uint32_t lo = 2 * jo + S;
                        if (lo >= lim) break;
                        const uint2 el = *reinterpret_cast<const uint2 *>(hb + lo);
                        const uint2 er = *reinterpret_cast<const uint2 *>(hb + lo + S);
                        const bool pr = (int)er.y > (int)el.y;
                        const uint2 ce = pr ? er : el;
                        if ((int)ce.y <= ev) break;
                        *reinterpret_cast<uint2 *>(hb + jo) = ce;
                        jo = pr ? lo + S : lo;
                    }
                    *reinterpret_cast<uint2 *>(hb + jo) = rec;
                }
                s_bound[t] = (int)reinterpret_cast<const uint2 *>(hb)->y;
            }
        }
        __syncthreads();
    }

    // ---- resolve labels and write the heap arrays --------------------------------------------------
    for (int i = tid; i < R * QPC; i += RQ_THREADS) {
        const int t = i / R, j = i - t * R;
        const int q = q0 + t;
        if (q >= Q) continue;
        const uint2 e = H[(size_t)j * QPC + t];
        int64_t label = -1;
        if (e.x != RQ_EMPTY) {
            const int cc = (int)(e.x >> 4), v = (int)(e.x & 15u);
            if (mode == 1) {
                const int *c = cum + (size_t)t * (P + 1);
                int lo = 0, hi = P;                                        // segment s with c[s] <= cc < c[s+1]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c[mid] <= cc) lo = mid; else hi = mid; }
                const int l = probes[(size_t)q * P + lo];
                label = ids[16 * (list_chunk_off[l] + (cc - c[lo])) + v];
            } else {
                label = 16 * (int64_t)cc + v;
            }
        }
        heap_idx[(size_t)q * R + j] = label;
        heap_val[(size_t)q * R + j] = (int)e.y;
    }
}

// ------------------------------------------------------------------------------------------------
// Queue replay, pipelined ("rq2"): same rounds as replay_rq_kernel, but
//   * heap entries and queue records are ONE 32-bit word: value in the top byte (stored so that a signed compare of
//     the words' top bytes orders them: int8 as is, uint8 ^ 0x80), stream position in the low 24 bits. A query's heap
//     is a contiguous array (node i at word i+1, so that the two children of a node are one aligned 8-byte load),
//     padded with sentinels (value -128: never "strictly greater" than anything), so leaves need no bounds test;
//   * the sift-downs of CONSECUTIVE inserts of a query are pipelined over L lanes: an insert replaces the root and
//     walks down one level per step; the next insert may start two steps later, because by then the previous one has
//     finished writing levels 0 and 1, and from there on the two stay two levels apart (a step reads level s+1 and
//     writes level s). The admission test of the next record only needs the root, which an insert fixes in its first
//     step. So a query retires one insert every two steps instead of one every ~depth steps, and the result is the
//     reference's array, slot for slot (ref: _fast_pq.pyx:274-307; the bound is still frozen per 16-vector chunk).
// Preconditions on top of replay_rq_kernel's: stream positions < 2^24 - 1, depth of the heap <= 2L steps.
// ------------------------------------------------------------------------------------------------
constexpr int RQ_CM_BLOCK = 512;                  // chunks a warp tests per step of the chunk-minimum path (16 per lane)
constexpr int RQ_CM_MIN = 1024;                   // stream position (chunks) from which a round uses the chunk minima
constexpr uint32_t RQ2_EMPTY = 0x00ffffffu;       // payload of a heap slot that was never filled
constexpr uint32_t RQ2_SENTINEL = 0x80ffffffu;    // value -128

template <bool SIGNED, int L, bool CM = false>
__global__ void __launch_bounds__(RQ_THREADS, CM ? 4 : 5)   // 5 CTAs/SM: 10 000 queries at 16 per CTA are one wave
replay_rq2_kernel(int mode, const uint8_t *__restrict__ est, int64_t stride, const int64_t *__restrict__ seg_off,
                  int64_t n_chunks0, int n0, const int64_t *__restrict__ list_chunk_off,
                  const int32_t *__restrict__ list_size, int n_lists, const int64_t *__restrict__ ids,
                  const int32_t *__restrict__ probes, int Q, int P, int64_t *__restrict__ heap_idx,
                  int32_t *__restrict__ heap_val, int R, int *__restrict__ fallback, int QPC, int QCAP,
                  const uint8_t *__restrict__ cmin, int qpw, const int64_t *__restrict__ cm_seg, int ovl_arg)
{
    // cm_seg (pull exchange): the compact layout the chunk minima are addressed by, while `seg_off` holds absolute addresses of
    // segments that live in other GPUs' buffers (est == null); null: the minima follow seg_off
    extern __shared__ __align__(16) unsigned char rq_sm[];
    const int HS = 2 * R + 4;                                              // words per heap: slot i at word i+1, children of R-1 included
    uint32_t *H = reinterpret_cast<uint32_t *>(rq_sm);                     // [QPC][HS]
    // ovl_arg: two queues of QCAP records per query (see the main loop); else one
    uint32_t *QU = H + (size_t)QPC * HS;                                   // [1 or 2][QPC][QCAP+2]
    int *cum = reinterpret_cast<int *>(QU + (size_t)(ovl_arg ? 2 : 1) * QPC * (QCAP + 2));     // [QPC][P+1] real chunks before segment s
    int *s_cursor = cum + (size_t)QPC * (P + 1);
    int *s_seg = s_cursor + QPC, *s_bound = s_seg + QPC, *s_count = s_bound + 2 * QPC, *s_round = s_count + 2 * QPC;   // s_bound, s_count: [2][QPC]
    // chunk-minimum path (cmin != null): est offset of the query's first chunk, or -1 when its segments are not back to back;
    // one list of flagged chunks per warp
    long long *s_qoff = reinterpret_cast<long long *>(rq_sm + (((size_t)((unsigned char *)(s_round + QPC) - rq_sm) + 15) & ~(size_t)15));   // (s_round is the last int array)
    uint32_t *LST = reinterpret_cast<uint32_t *>(s_qoff + QPC);            // [n_warps][RQ_CM_BLOCK]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NTH = blockDim.x, n_warps = NTH / 32;   // 64, 128 or 256 threads
    const int q0 = blockIdx.x * QPC;
    const uint32_t flip = SIGNED ? 0u : 0x80u;                             // stored byte = value ^ flip
    const int init = 127;                                                  // stored form of the reference's 127 / 255
    // qpw < 0: FREE-RUNNING warps. Warp w owns the queries [w |qpw|, (w + 1) |qpw|) of the CTA in both halves of a round, so
    // a round ends with __syncwarp instead of two CTA barriers: no warp waits for the slowest producer of the CTA, and every
    // warp (not just QPC / (32 / L) of them) has queries to consume. The ncu capture of round 2 showed 21 of 27 stall cycles
    // per issue at those barriers -- but see rq_qpw_arg: measured slower, off by default.
    const bool indep = qpw < 0;
    const int QPW = indep ? -qpw : ((qpw > 0 && qpw <= 32 / L) ? qpw : 32 / L);

    for (int i = tid; i < QPC * HS; i += NTH) {
        const int j = i % HS;
        H[i] = (j >= 1 && j <= R) ? (((uint32_t)init << 24) | RQ2_EMPTY) : RQ2_SENTINEL;
    }
    for (int t = tid; t < QPC; t += NTH) {
        const int q = q0 + t;
        int *c = cum + (size_t)t * (P + 1);
        bool ok = q < Q;
        int run = 0;
        c[0] = 0;
        for (int s = 0; s < P; s++) {
            int nc = 0;
            if (ok) {
                if (mode == 1) {
                    const int l = probes[(size_t)q * P + s];
                    if (l < 0 && l != PROBE_SKIP) ok = false;              // Python-wrapped index: lists may repeat
                    else if (l != PROBE_SKIP) nc = (list_size[l] + 15) >> 4;
                } else {
                    int64_t r = ((int64_t)n0 + 15) >> 4;
                    nc = (int)(r < n_chunks0 ? r : n_chunks0);
                }
            }
            run += nc;
            c[s + 1] = run;
        }
        if (run >= (1 << 20) - 1) ok = false;                              // positions must fit 24 bits
        if (q < Q && fallback) fallback[q] = ok ? 0 : 1;
        if (!ok) for (int s = 0; s <= P; s++) c[s] = 0;
        s_cursor[t] = 0; s_seg[t] = 0; s_bound[t] = init; s_bound[QPC + t] = init; s_count[t] = 0; s_count[QPC + t] = 0; s_round[t] = 0;
        if (CM && cmin) {
            // the query's stream is one contiguous byte range of `est` when its segments lie back to back (the compact
            // single-GPU plan): chunk cc of the stream is est[qoff + 16 cc ..] and its minimum is cmin[qoff / 16 + cc]
            long long qoff = -1;
            const int64_t *lay = cm_seg ? cm_seg : seg_off;
            bool contig = ok && mode == 1 && lay != nullptr;
            for (int s = 0; contig && s < P; s++) {
                if (c[s + 1] == c[s]) continue;
                const long long o = lay[(size_t)q * P + s];
                if (qoff < 0) qoff = o - 16LL * c[s];
                if (o < 0 || o != qoff + 16LL * c[s]) contig = false;
            }
            s_qoff[t] = (contig && qoff >= 0) ? qoff : -1;
        }
    }
    __syncthreads();

    // OVERLAPPED ROUNDS (ovl): the ncu capture of round 2 shows 21 of 27 stall cycles per issued instruction at the two CTA
    // barriers of a round -- the consumer warps walk the sift chain while the producers wait, and vice versa. With two queues per
    // query the n_cons consumer warps replay the queues filled in the previous iteration WHILE the other warps filter the next
    // windows into the other queues, against the bound as it was when the iteration began. Exact: the bound only falls, so a
    // stale bound admits a superset, and the consumer applies the reference's own test to every record.
    const int n_cons = (QPC + QPW - 1) / QPW;                              // warps that hold all the CTA's queries when consuming
    const bool ovl = ovl_arg && !indep && n_cons < n_warps;
    for (int r = 0;; r++) {
        // ---- produce (as in replay_rq_kernel; records are packed words) -----------------------------
        const int pb = ovl ? (r & 1) : 0, cb = ovl ? (pb ^ 1) : 0;          // queue filled / queue replayed in this iteration
        const bool prod_warp = !ovl || r == 0 || warp >= n_cons;
        const int p_first = !prod_warp ? 0 : (indep ? warp * QPW : ((ovl && r > 0) ? warp - n_cons : warp));
        const int p_step = indep ? 1 : ((ovl && r > 0) ? n_warps - n_cons : n_warps);
        const int p_end = !prod_warp ? 0 : (indep ? min(QPC, (warp + 1) * QPW) : QPC);
        bool more = false;
        for (int t = p_first; t < p_end; t += p_step) {
            const int *c = cum + (size_t)t * (P + 1);
            const int total = c[P];
            const int cursor = s_cursor[t];
            if (cursor >= total) { if (lane == 0) s_count[pb * QPC + t] = 0; continue; }
            const int q = q0 + t;
            // overlapped rounds: the bound the consumer left at the end of the PREVIOUS iteration (its own slot: the consumer of
            // this iteration writes the other one, so there is no concurrent access and the filter does not depend on timing)
            const int bound = s_bound[(ovl ? (r & 1) : 0) * QPC + t];      // stored form
            int W = s_round[t] == 0 ? ((R + 15) >> 4) + 1 : (cursor < 32 ? 32 : cursor);
            if (W > (1 << 16)) W = 1 << 16;
            int end = (total - cursor < W) ? total : cursor + W;
            int count = 0, sg = s_seg[t];
            uint32_t *qu = QU + ((size_t)pb * QPC + t) * (QCAP + 2);
            // PF steps of 32 chunks are fetched before the first is examined: a window is thousands of chunks long once the
            // bound has settled and almost nothing survives the filter, so the walk is a chain of load latencies otherwise
            constexpr int PF = 4;
            const bool use_cm = CM && cmin != nullptr && cursor >= RQ_CM_MIN && s_qoff[t] >= 0;
            if (use_cm) {
                // Long streams: once the bound has settled almost no chunk holds a candidate. The scan left one byte per
                // chunk (its smallest estimate); a lane tests 16 chunks with one 16-byte load of those, the chunks that may
                // hold a candidate are listed in stream order and only they are fetched and examined as above.
                const long long qoff = s_qoff[t];
                const uint8_t *cmq = cmin + (qoff >> 4);
                const uint8_t *eq = est + qoff;
                uint32_t *lst = LST + (size_t)warp * RQ_CM_BLOCK;
                const int skew = (int)((uintptr_t)(cmq + cursor) & 15);
                bool cut = false;
                auto load_cm = [&](int lc) {
                    uint4 e = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);      // stored form of "no candidate"
                    if (lc < end && lc + 16 > cursor) {
                        e = ldg_nc_u4(reinterpret_cast<const uint4 *>(cmq + lc));
                        if (!SIGNED) { e.x ^= 0x80808080u; e.y ^= 0x80808080u; e.z ^= 0x80808080u; e.w ^= 0x80808080u; }
                    }
                    return e;
                };
                uint4 nxt = load_cm(cursor - skew + 16 * lane);
                for (int b0 = cursor - skew; b0 < end && !cut; b0 += RQ_CM_BLOCK) {
                    const int lc = b0 + 16 * lane;                           // this lane's 16 chunks: lc .. lc+15
                    const uint4 e = nxt;
                    nxt = load_cm(lc + RQ_CM_BLOCK);                         // the next block is in flight while this one is examined
                    uint32_t m = 0;
                    if (lc < end && lc + 16 > cursor) {
                        m = cand_mask16<true>(e, bound);
                        if (lc < cursor) m &= ~((1u << (cursor - lc)) - 1u);
                        if (lc + 16 > end) m &= (1u << (end - lc)) - 1u;
                    }
                    if (__ballot_sync(FULL, m != 0) == 0) continue;
                    const int fc = __popc(m);
                    int fi = fc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, fi, o); if (lane >= o) fi += v; }
                    const int F = __shfl_sync(FULL, fi, 31);
                    {
                        int k = fi - fc;
                        while (m) { const int v = __ffs(m) - 1; m &= m - 1; lst[k++] = (uint32_t)(lc + v); }
                    }
                    __syncwarp();
                    for (int j0 = 0; j0 < F; j0 += 32) {
                        const bool act = j0 + lane < F;
                        const int cc = act ? (int)lst[j0 + lane] : 0;
                        uint32_t mm = 0;
                        uint4 ee = make_uint4(0, 0, 0, 0);
                        if (act) {
                            int lo = 0, hi = P;                              // segment with c[lo] <= cc < c[lo+1]
                            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c[mid] <= cc) lo = mid; else hi = mid; }
                            const int rem = list_size[probes[(size_t)q * P + lo]] - 16 * (cc - c[lo]);
                            const uint8_t *ea = cm_seg ? est + seg_off[(size_t)q * P + lo] + 16LL * (cc - c[lo]) : eq + 16LL * cc;
                            ee = ldg_nc_u4(reinterpret_cast<const uint4 *>(ea));
                            if (!SIGNED) { ee.x ^= 0x80808080u; ee.y ^= 0x80808080u; ee.z ^= 0x80808080u; ee.w ^= 0x80808080u; }
                            mm = cand_mask16<true>(ee, bound) & (rem >= 16 ? 0xffffu : ((1u << rem) - 1u));
                        }
                        if (__ballot_sync(FULL, mm != 0) == 0) continue;
                        const int cnt = __popc(mm);
                        int incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                        const int tot = __shfl_sync(FULL, incl, 31);
                        int keep_tot = tot;
                        if (count + tot > QCAP) {                            // queue full: the next round resumes at the first chunk that does not fit
                            const unsigned over = __ballot_sync(FULL, count + incl > QCAP);
                            const int cl = __ffs(over) - 1;
                            keep_tot = __shfl_sync(FULL, incl - cnt, cl);
                            end = __shfl_sync(FULL, cc, cl);
                            if (lane >= cl) mm = 0;
                            cut = true;
                        }
                        int k = count + incl - cnt;
                        const uint32_t ws[4] = {ee.x, ee.y, ee.z, ee.w};
                        while (mm) {
                            const int v = __ffs(mm) - 1;
                            mm &= mm - 1;
                            const uint32_t byte = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
                            qu[k++] = (byte << 24) | (16u * (uint32_t)cc + v);
                        }
                        count += keep_tot;
                        if (cut) break;
                    }
                    __syncwarp();
                }
            } else
            for (int base0 = cursor; base0 < end; base0 += 32 * PF) {
                uint4 ev[PF];
                int slv[PF], remv[PF];
                {
                    int sl = sg;
#pragma unroll
                    for (int u = 0; u < PF; u++) {
                        const int cc = base0 + 32 * u + lane;
                        ev[u] = make_uint4(0, 0, 0, 0); remv[u] = 0;
                        if (cc < end) {
                            while (cc >= c[sl + 1]) sl++;
                            const int local = cc - c[sl];
                            int n;
                            const uint8_t *ep;
                            if (mode == 1) {
                                n = list_size[probes[(size_t)q * P + sl]];
                                ep = est + (seg_off ? seg_off[(size_t)q * P + sl] : ((int64_t)q * P + sl) * stride);
                            } else {
                                n = n0;
                                ep = est + (int64_t)q * stride;
                            }
                            ev[u] = ldg_nc_u4(reinterpret_cast<const uint4 *>(ep) + local);
                            remv[u] = n - 16 * local;
                        }
                        slv[u] = sl;
                    }
                }
                bool cut = false;
#pragma unroll
                for (int u = 0; u < PF; u++) {
                    const int base = base0 + 32 * u;
                    if (base >= end) break;
                    const int cc = base + lane;
                    const bool act = cc < end;
                    uint32_t m = 0;
                    uint4 e = ev[u];
                    const int sl = slv[u];
                    if (act) {
                        if (!SIGNED) { e.x ^= 0x80808080u; e.y ^= 0x80808080u; e.z ^= 0x80808080u; e.w ^= 0x80808080u; }
                        const int rem = remv[u];
                        m = cand_mask16<true>(e, bound) & (rem >= 16 ? 0xffffu : ((1u << rem) - 1u));
                    }
                    const int last = end - 1 - base;
                    sg = __shfl_sync(FULL, sl, last < 31 ? last : 31);
                    if (__ballot_sync(FULL, m != 0) == 0) continue;          // nothing survives in these 32 chunks: the common case
                    const int cnt = __popc(m);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                    const int tot = __shfl_sync(FULL, incl, 31);
                    int keep_tot = tot;
                    if (count + tot > QCAP) {
                        const unsigned over = __ballot_sync(FULL, count + incl > QCAP);
                        const int cl = __ffs(over) - 1;
                        keep_tot = __shfl_sync(FULL, incl - cnt, cl);
                        sg = __shfl_sync(FULL, sl, cl);
                        if (lane >= cl) m = 0;
                        end = base + cl;
                        cut = true;
                    }
                    int k = count + incl - cnt;
                    const uint32_t ws[4] = {e.x, e.y, e.z, e.w};
                    while (m) {
                        const int v = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t byte = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
                        qu[k++] = (byte << 24) | (16u * (uint32_t)cc + v);
                    }
                    count += keep_tot;
                    if (cut) break;
                }
                if (cut) break;
            }
            if (lane == 0) { s_cursor[t] = end; s_seg[t] = sg; s_count[pb * QPC + t] = count; s_round[t] = 1; qu[count] = RQ2_SENTINEL; qu[count + 1] = RQ2_SENTINEL; }
            more = true;
        }
        if (indep) { __syncwarp(); if (!more) break; }                      // `more` is warp-uniform
        else if (!ovl && !__syncthreads_or(more)) break;
        // ---- consume: L lanes per query, 32/L queries per warp. The step is branch-free (the queries of a warp are in
        //      different states; divergent code would be issued once per state and the loop is issue-bound) -----------
        {
            // qpw queries per consumer warp (default 32 / L; fewer = less lock-step waste, more warps busy: the replay model
            // of tools/replay_model.py puts 8 lock-stepped queries at +26 % steps over one query per warp)
            const int role = lane % L;
            for (int tb = (ovl && (r == 0 || warp >= n_cons)) ? QPC : warp * QPW; tb < QPC; tb += n_warps * QPW) {
                const int t = tb + lane / L;
                const bool mine = lane / L < QPW && t < QPC;
                uint32_t *hw = H + (size_t)(mine ? t : 0) * HS;
                const uint32_t *qu = QU + ((size_t)cb * QPC + (mine ? t : 0)) * (QCAP + 2);
                const int cnt = mine ? s_count[cb * QPC + t] : 0;
                // State replicated over the query's L lanes: queue cursor, bound frozen for the current chunk, last record seen.
                // Values stay in the top byte of their 32-bit words and are compared there: for words a, b with the value in
                // bits 24..31 (signed) and a 24-bit payload below, value(a) < value(b)  <=>  (int)a < (int)(b & 0xff000000),
                // so no per-step byte extraction; "same chunk" is a masked xor of two records. Cursors are pointers, not
                // indices; the heap node an insert looks at is a 32-bit word index nw (slot j = word j + 1, children = words 2 nw, 2 nw + 1).
                const uint32_t *qp = qu, *const qend = qu + cnt;
                int next_role = 0, cooldown = 0;
                uint32_t frm = 0, cur = 0xffffffffu;                       // chunk 0xfffff never occurs (positions < 2^24 - 16)
                bool active = false;
                uint32_t rw = 0;
                uint32_t nw = 1;
                // The root is read AFTER the barrier that ends a step and carried into the next one: the lane that starts an
                // insert rewrites hw[1] in its first step, and the L lanes of the query must all see the value from before
                // that write (they replicate the admission state) without relying on lock-step execution inside a step.
                uint32_t rootm = hw[1] & 0xff000000u;
                while (__any_sync(FULL, active || qp < qend)) {
                    // admission: at most two records per step (the queue is padded with two sentinels)
                    const uint32_t r0 = qp[0], r1 = qp[1];
                    const bool ok0 = cooldown == 0 && qp < qend;
                    const uint32_t fr0 = (ok0 && ((r0 ^ cur) & 0x00fffff0u) != 0) ? rootm : frm;   // first record of a chunk freezes the bound
                    const bool acc0 = ok0 && (int)r0 < (int)fr0;
                    const bool ok1 = ok0 && !acc0 && qp + 1 < qend;
                    const uint32_t fr1 = (ok1 && ((r1 ^ r0) & 0x00fffff0u) != 0) ? rootm : fr0;
                    const bool acc1 = ok1 && (int)r1 < (int)fr1;
                    cur = ok1 ? r1 : (ok0 ? r0 : cur);
                    frm = fr1;
                    qp += (ok0 ? 1 : 0) + (ok1 ? 1 : 0);
                    const bool start = acc0 || acc1;
                    if (start && role == next_role) { active = true; rw = acc0 ? r0 : r1; nw = 1; }
                    next_role = start ? (next_role + 1 == L ? 0 : next_role + 1) : next_role;
                    cooldown = start ? 1 : (cooldown ? cooldown - 1 : 0);                 // next admission two steps from now
                    // one level of the sift-down (ref: _fast_pq.pyx:290-307)
                    const uint32_t kw = 2 * nw;
                    const uint2 ch2 = *reinterpret_cast<const uint2 *>(hw + kw);
                    const bool pr = (int)ch2.x < (int)(ch2.y & 0xff000000u);              // the right child wins only when strictly greater
                    const uint32_t cw = pr ? ch2.y : ch2.x;
                    const bool stop = !((int)rw < (int)(cw & 0xff000000u));               // no child strictly greater: the record stays here
                    if (active) hw[nw] = stop ? rw : cw;
                    nw = kw + (pr ? 1u : 0u);
                    active = active && !stop;
                    if (!active) nw = 1;
                    __syncwarp();
                    rootm = hw[1] & 0xff000000u;
                }
                if (mine && role == 0) s_bound[(ovl ? ((r + 1) & 1) : 0) * QPC + t] = (int)rootm >> 24;
            }
        }
        if (indep) __syncwarp();
        else if (ovl) { if (!__syncthreads_or(more)) break; }              // nothing was produced: the last queues have just been replayed
        else __syncthreads();
    }
    if (indep) __syncthreads();                                            // every warp's queries are finished

    // ---- resolve labels and write the heap arrays --------------------------------------------------
    for (int i = tid; i < R * QPC; i += NTH) {
        const int t = i / R, j = i - t * R;
        const int q = q0 + t;
        if (q >= Q) continue;
        const uint32_t w = H[(size_t)t * HS + j + 1];
        const uint32_t pay = w & 0xffffffu;
        int64_t label = -1;
        if (pay != RQ2_EMPTY) {
            const int cc = (int)(pay >> 4), v = (int)(pay & 15u);
            if (mode == 1) {
                const int *c = cum + (size_t)t * (P + 1);
                int lo = 0, hi = P;
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c[mid] <= cc) lo = mid; else hi = mid; }
                const int l = probes[(size_t)q * P + lo];
                label = ids[16 * (list_chunk_off[l] + (cc - c[lo])) + v];
            } else {
                label = 16 * (int64_t)cc + v;
            }
        }
        const uint32_t b = (w >> 24) ^ flip;
        heap_idx[(size_t)q * R + j] = label;
        heap_val[(size_t)q * R + j] = SIGNED ? (int)(int8_t)b : (int)b;
    }
}

// warp-per-query IVF replay restricted to the queries flagged by the queue-replay kernel
template <bool SIGNED>
__global__ void __launch_bounds__(32 * REPLAY_WARPS)
ivf_replay_fallback_kernel(const uint8_t *__restrict__ est, int64_t slot_stride, const int64_t *__restrict__ seg_off,
                           const int64_t *__restrict__ list_chunk_off, const int32_t *__restrict__ list_size,
                           int n_lists, const int64_t *__restrict__ ids, const int32_t *__restrict__ probes,
                           int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, const int *__restrict__ fallback)
{
    const int q = blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5);
    if (q >= Q || !fallback[q]) return;
    const int lane = threadIdx.x & 31;
    for (int j = lane; j < R; j += 32) { heap_idx[(size_t)q * R + j] = -1; heap_val[(size_t)q * R + j] = SIGNED ? 127 : 255; }
    __syncwarp();
    for (int s = 0; s < P; s++) {
        int l = probes[(size_t)q * P + s];
        if (l == PROBE_SKIP) continue;
        if (l < 0) l += n_lists;
        const int64_t c0 = list_chunk_off[l];
        const int64_t nc = ((int64_t)list_size[l] + 15) >> 4;              // real chunks (tile padding is not stored)
        const int64_t so = seg_off ? seg_off[(size_t)q * P + s] : ((int64_t)q * P + s) * slot_stride;
        if (so < 0) continue;
        replay_segment<SIGNED>(est + so, nc, list_size[l], ids + 16 * c0,
                               heap_idx + (size_t)q * R, heap_val + (size_t)q * R, R, lane);
    }
}

__global__ void heap_fill_kernel(int64_t *heap_idx, int32_t *heap_val, int64_t count, int init_val)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { heap_idx[i] = -1; heap_val[i] = init_val; }
}

int launch_heap_fill(int64_t *heap_idx, int32_t *heap_val, int64_t count, int signd, cudaStream_t st)
{
    TKB_REQUIRE(count >= 0, "negative count");
    if (count == 0) return TKB_OK;
    TKB_REQUIRE(heap_idx && heap_val, "null pointer");
    const int64_t blocks = (count + 255) / 256;
    TKB_REQUIRE(blocks <= 0x7fffffff, "heap too large");
    heap_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(heap_idx, heap_val, count, signd ? 127 : 255);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

int launch_replay(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n, int64_t *heap_idx,
                  int32_t *heap_val, int Q, int R, int signd, const int64_t *labels, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && n_chunks >= 0, "negative extent");
    if (Q == 0 || R == 0 || n_chunks == 0) return TKB_OK;
    TKB_REQUIRE(est && heap_idx && heap_val, "null pointer");
    TKB_REQUIRE(est_stride % 16 == 0 && (uintptr_t)est % 16 == 0, "est must be 16-byte aligned/strided");
    const unsigned blocks = (unsigned)((Q + REPLAY_WARPS - 1) / REPLAY_WARPS);
    if (signd) replay_kernel<true><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, est_stride, n_chunks, n, heap_idx, heap_val, Q, R, labels);
    else       replay_kernel<false><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, est_stride, n_chunks, n, heap_idx, heap_val, Q, R, labels);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

int launch_ivf_replay(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                      const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                      int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                      cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || R == 0 || P == 0) return TKB_OK;
    TKB_REQUIRE(est && list_chunk_off && list_size && ids && probes && heap_idx && heap_val, "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0 && (uintptr_t)est % 16 == 0, "est must be 16-byte aligned/strided");
    const unsigned blocks = (unsigned)((Q + REPLAY_WARPS - 1) / REPLAY_WARPS);
    if (signd) ivf_replay_kernel<true><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val, R);
    else       ivf_replay_kernel<false><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val, R);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

// Queue-replay launch geometry: QPC queries per CTA (a power of two <= 16), queue capacity QCAP records,
// LPW lanes (queries) per consumer warp.
struct RqGeom { int qpc, qcap, lpw; size_t smem; int v2, lanes, threads; };

static size_t rq_smem(int R, int P, int qpc, int qcap)
{
    return (size_t)qpc * (8 * ((size_t)R + 1) + 8 * ((size_t)qcap + 1) + 4 * ((size_t)P + 1) + 20) + 16;
}

static size_t rq2_smem(int R, int P, int qpc, int qcap, bool cm = false, int threads = RQ_THREADS)
{
    // (qcap + 4: two queues of qcap / 2 + 2 words when the rounds overlap; 28: five per-query ints, bound and count twice)
    const size_t base = (size_t)qpc * (4 * (2 * (size_t)R + 4) + 4 * ((size_t)qcap + 4) + 4 * ((size_t)P + 1) + 28) + 16;
    // chunk-minimum path: s_qoff[qpc] + one list of RQ_CM_BLOCK chunk numbers per warp
    return cm ? base + 16 + 8 * (size_t)qpc + 4 * (size_t)(threads / 32) * RQ_CM_BLOCK : base;
}

// Launch geometry knobs of the pipelined replay (A/B switches; results do not depend on them): threads per CTA (TKB_RQ_THREADS:
// 64 / 128 / 256), queue capacity in multiples of R (TKB_RQ_QCAP: 1..8, default 4), shared memory per CTA that decides how many
// queries share one (TKB_RQ_SMEM_KB, default 45)
static int rq_env(const char *name, int dflt, int lo, int hi)
{
    const char *e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    return (v < lo || v > hi) ? dflt : v;
}

static bool rq_geometry(int Q, int R, int P, RqGeom &g, int64_t stream_chunks = 0, bool cm = false, bool remote = false)
{
    if (R <= 0 || P <= 0) return false;
    static int v2_env = -1;
    if (v2_env < 0) { const char *e = getenv("TKB_RQ2"); v2_env = e ? atoi(e) : 1; }
    g.v2 = 0; g.lanes = 0; g.threads = RQ_THREADS;
    if (v2_env && R <= 65535 && stream_chunks < (1 << 20) - 1) {                    // pipelined kernel: an insert takes <= floor(log2 R) + 1 steps <= 2 * lanes
        g.v2 = 1;
        g.lanes = R <= 255 ? 4 : 8;
        // TKB_RQ_LANES=8 forces 8 lanes per query for small heaps too: half as many queries per consumer warp, twice as many
        // consumer warps to hide the shared-memory latency of the sift chain (an A/B for round 2; results are identical)
        static int lanes_env = -1;
        if (lanes_env < 0) { const char *e = getenv("TKB_RQ_LANES"); lanes_env = e ? atoi(e) : 0; }
        if (lanes_env == 8) g.lanes = 8;
        static int qcap_mult = 0, smem_kb = 0, threads = 0;
        if (!qcap_mult) {
            qcap_mult = rq_env("TKB_RQ_QCAP", 4, 0, 8); smem_kb = rq_env("TKB_RQ_SMEM_KB", 45, 8, 200);
            threads = rq_env("TKB_RQ_THREADS", 256, 64, 256);
            if (threads != 64 && threads != 128) threads = 256;
        }
        // Long streams (the chunk-minimum path: 100M-vector indexes). The launch is a few waves of CTAs and every wave lasts as
        // long as the sift chain of its queries, so what counts is how many queries are RESIDENT per SM, i.e. shared memory per
        // query: queues of R records instead of 4 R, 128-thread CTAs of 8 queries (two consumer warps, two producer warps; the
        // per-warp list of flagged chunks is 2 KB). Measured at 100M x 128, R = 331: 3.82 -> 2.68 ms (ncu: 2 500 CTAs of 4 queries,
        // four per SM = 4.2 waves, one consumer warp per CTA busy). The environment switches still override.
        const bool long_streams = cm;
        int th = threads, qm = qcap_mult, kb = smem_kb;
        if (long_streams) {
            // estimates in other GPUs' memory (pull exchange): a producer waits microseconds for every fetch, so six producer
            // warps per CTA instead of two (same residency: the shared memory still admits four CTAs); 2 GPUs: 3.47 -> 3.21 ms
            if (!getenv("TKB_RQ_THREADS")) th = remote ? 256 : 128;
            if (!getenv("TKB_RQ_QCAP")) qm = 1;
            if (!getenv("TKB_RQ_SMEM_KB")) kb = 36;
        }
        g.threads = th;
        g.qcap = qm * R < 128 ? 128 : qm * R;
        // CTAs wanted before queries are packed 16 to a CTA (TKB_RQ_MIN_CTAS, default 2 x 148). The launch list of round 1
        // shows a 5 000-query launch (313 CTAs) taking almost as long as a 10 000-query one: worth an A/B at 4 x 148.
        static int min_ctas = 0;
        if (!min_ctas) { const char *e = getenv("TKB_RQ_MIN_CTAS"); min_ctas = e ? atoi(e) : 0; if (min_ctas <= 0) min_ctas = 2 * 148; }
        int qpc = 16;
        while (qpc > 1 && (Q + qpc - 1) / qpc < min_ctas) qpc >>= 1;
        while (qpc > 1 && rq2_smem(R, P, qpc, g.qcap) > (size_t)kb * 1024) qpc >>= 1;
        g.qpc = qpc; g.lpw = 0;
        g.smem = rq2_smem(R, P, qpc, g.qcap, cm, g.threads);
        if (g.smem <= 200 * 1024) return true;
        g.v2 = 0;
    }
    g.qcap = 2 * R < 64 ? 64 : 2 * R;
    // enough CTAs to cover the machine about twice, as many queries per CTA as that allows, and at most
    // ~45 KB of shared memory so that several CTAs share an SM
    int qpc = 16;
    while (qpc > 1 && (Q + qpc - 1) / qpc < 2 * 148) qpc >>= 1;
    while (qpc > 1 && rq_smem(R, P, qpc, g.qcap) > 45 * 1024) qpc >>= 1;
    g.qpc = qpc;
    // few lanes per consumer warp = little lock-step waste, more warps to issue (measured flat from 2 to 16)
    static int lpw_env = -1;
    if (lpw_env < 0) { const char *e = getenv("TKB_RQ_LPW"); lpw_env = e ? atoi(e) : 0; }
    g.lpw = (lpw_env == 1 || lpw_env == 2 || lpw_env == 4 || lpw_env == 8 || lpw_env == 16) ? lpw_env : 4;
    while (g.lpw * (RQ_THREADS / 32) < qpc) g.lpw <<= 1;
    g.smem = rq_smem(R, P, qpc, g.qcap);
    return g.smem <= 200 * 1024;
}

// TKB_RQ_QPW: queries per consumer warp of the pipelined replay (1, 2, 4, 8; default 32 / lanes). An A/B switch for round 2.
static int rq_qpw()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("TKB_RQ_QPW"); v = e ? atoi(e) : 0; if (v != 1 && v != 2 && v != 4 && v != 8) v = 0; }
    return v;
}

// Free-running warps (TKB_RQ_INDEP=1, off by default): every warp of the CTA owns qpc / n_warps queries for both halves of a
// round. Measured on a B200 (100M x 128): no CTA barrier stalls any more, but eight warps each walk the sift chain for two
// queries where two warps walked it for eight, and the consume step is issue-bound: probe selection 0.32 -> 0.58 ms, lists
// 4.00 -> 4.14 ms. Kept as an A/B switch. Returns the kernel's qpw argument: negative = free-running with that many queries per warp.
static int rq_qpw_arg(const RqGeom &g, int lanes)
{
    static int indep = -1;
    if (indep < 0) { const char *e = getenv("TKB_RQ_INDEP"); indep = e ? atoi(e) : 0; }
    const int n_warps = g.threads / 32;
    if (!indep || rq_qpw() || g.qpc % n_warps != 0) return rq_qpw();
    const int per = g.qpc / n_warps;
    return (per >= 1 && per <= 32 / lanes) ? -per : rq_qpw();
}

// Overlapped rounds of the pipelined replay (see the kernel's main loop). TKB_RQ_OVERLAP unset: for long streams only (the
// chunk-minimum launches of 100M-vector indexes: 4.15 -> 3.82 ms); short streams keep all warps producing and the full queues
// (GloVe shape: 0.283 ms against 0.296 overlapped, probe selection 0.101 against 0.120). 1 / 0 force it on / off.
static int rq_ovl(bool long_streams)
{
    static int v = -2;
    if (v == -2) { const char *e = getenv("TKB_RQ_OVERLAP"); v = e ? (atoi(e) != 0) : -1; }
    return v < 0 ? (long_streams ? 1 : 0) : v;
}

template <bool SIGNED>
static int launch_rq(int mode, const uint8_t *est, int64_t stride, const int64_t *seg_off, int64_t n_chunks0, int n0,
                     const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, const int64_t *ids,
                     const int32_t *probes, int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int *fallback,
                     const RqGeom &g, cudaStream_t st, const uint8_t *cmin = nullptr, const int64_t *cm_seg = nullptr)
{
    const unsigned blocks = (unsigned)((Q + g.qpc - 1) / g.qpc);
    if (g.v2) {
#define TKB_RQ2_LAUNCH(LANES, CMV)                                                                                                   \
        do {                                                                                                                         \
            TKB_CUDA(cudaFuncSetAttribute(replay_rq2_kernel<SIGNED, LANES, CMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem)); \
            replay_rq2_kernel<SIGNED, LANES, CMV><<<blocks, g.threads, g.smem, st>>>(mode, est, stride, seg_off, n_chunks0, n0, list_chunk_off, \
                                                                                     list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val,  \
                                                                                     R, fallback, g.qpc, rq_ovl(cmin != nullptr) ? g.qcap / 2 : g.qcap, cmin, rq_qpw_arg(g, LANES), cm_seg, rq_ovl(cmin != nullptr)); \
        } while (0)
        if (cmin) { if (g.lanes == 4) TKB_RQ2_LAUNCH(4, true); else TKB_RQ2_LAUNCH(8, true); }
        else      { if (g.lanes == 4) TKB_RQ2_LAUNCH(4, false); else TKB_RQ2_LAUNCH(8, false); }
#undef TKB_RQ2_LAUNCH
        TKB_LAUNCH_CHECK();
        return TKB_OK;
    }
    TKB_CUDA(cudaFuncSetAttribute(replay_rq_kernel<SIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    replay_rq_kernel<SIGNED><<<blocks, RQ_THREADS, g.smem, st>>>(mode, est, stride, seg_off, n_chunks0, n0, list_chunk_off,
                                                                 list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val,
                                                                 R, fallback, g.qpc, g.qcap, g.lpw);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

// Fresh-heap replays: queue-replay kernel when its preconditions hold, the warp-per-query kernel otherwise.
int launch_replay_fresh(const uint8_t *est, int64_t est_stride, int64_t n_chunks, int n, int64_t *heap_idx,
                        int32_t *heap_val, int Q, int R, int signd, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && n_chunks >= 0, "negative extent");
    if (Q == 0 || R == 0) return TKB_OK;
    TKB_REQUIRE(heap_idx && heap_val, "null pointer");
    RqGeom g;
    if (n_chunks == 0 || n_chunks >= (1LL << 27) || !rq_geometry(Q, R, 1, g, n_chunks)) {
        if (int rc = launch_heap_fill(heap_idx, heap_val, (int64_t)Q * R, signd, st)) return rc;
        return launch_replay(est, est_stride, n_chunks, n, heap_idx, heap_val, Q, R, signd, nullptr, st);
    }
    TKB_REQUIRE(est && est_stride % 16 == 0 && (uintptr_t)est % 16 == 0, "est must be 16-byte aligned/strided");
    if (signd) return launch_rq<true>(0, est, est_stride, nullptr, n_chunks, n, nullptr, nullptr, 0, nullptr, nullptr, Q, 1, heap_idx, heap_val, R, nullptr, g, st);
    return launch_rq<false>(0, est, est_stride, nullptr, n_chunks, n, nullptr, nullptr, 0, nullptr, nullptr, Q, 1, heap_idx, heap_val, R, nullptr, g, st);
}

template <bool SIGNED>
static int ivf_replay_fresh_t(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                              const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                              int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int *fallback,
                              const RqGeom &g, cudaStream_t st, const uint8_t *cmin, const int64_t *cm_seg)
{
    if (int rc = launch_rq<SIGNED>(1, est, slot_stride, seg_off, 0, 0, list_chunk_off, list_size, n_lists, ids, probes,
                                   Q, P, heap_idx, heap_val, R, fallback, g, st, cmin, cm_seg)) return rc;
    // queries whose probe list holds Python-wrapped (negative) entries: warp-per-query kernel with label dedupe
    ivf_replay_fallback_kernel<SIGNED><<<(unsigned)((Q + REPLAY_WARPS - 1) / REPLAY_WARPS), 32 * REPLAY_WARPS, 0, st>>>(
        est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val, R, fallback);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

int launch_ivf_replay_fresh(const uint8_t *est, int64_t slot_stride, const int64_t *seg_off, const int64_t *list_chunk_off,
                            const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                            int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                            int unique_labels, int *fallback, cudaStream_t st, const uint8_t *cmin, const int64_t *cm_seg)
{
    // est == null with a segment plan: seg_off holds absolute addresses (pull exchange; cm_seg = the layout of the minima)
    TKB_REQUIRE(Q >= 0 && R >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || R == 0) return TKB_OK;
    TKB_REQUIRE(heap_idx && heap_val, "null pointer");
    TKB_REQUIRE(!cmin || ((uintptr_t)cmin % 16 == 0 && seg_off), "cmin must be 16-byte aligned and needs a segment plan");
    RqGeom g;
    if (!unique_labels || P == 0 || P >= 0xffff || !fallback || !rq_geometry(Q, R, P, g, 0, cmin != nullptr, cm_seg != nullptr)) {
        if (int rc = launch_heap_fill(heap_idx, heap_val, (int64_t)Q * R, signd, st)) return rc;
        return launch_ivf_replay(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx,
                                 heap_val, R, signd, st);
    }
    TKB_REQUIRE((est || seg_off) && list_chunk_off && list_size && ids && probes, "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0 && (uintptr_t)est % 16 == 0, "est must be 16-byte aligned/strided");
    if (!g.v2) cmin = nullptr;                                             // only the pipelined kernel reads the chunk minima
    if (!cmin) cm_seg = nullptr;
    if (signd) return ivf_replay_fresh_t<true>(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P,
                                               heap_idx, heap_val, R, fallback, g, st, cmin, cm_seg);
    return ivf_replay_fresh_t<false>(est, slot_stride, seg_off, list_chunk_off, list_size, n_lists, ids, probes, Q, P,
                                     heap_idx, heap_val, R, fallback, g, st, cmin, cm_seg);
}

}  // namespace tkb
