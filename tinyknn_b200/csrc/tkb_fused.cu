// tkb_fused.cu -- the per-query IVF pipeline after probe selection as ONE kernel.
//
// Replaces, for a batch of queries, the body of IVF.query from the list loop on (ref: tinyknn/ivf.py:135-163):
//     heap of pass_1 slots -> for cl in top: query_pq(list cl, labels=ids[cl]) -> drop -1 -> knn_brute1 -> k ids
// i.e. the scan half and the heap half of query_pq_avx / query_pq_sse (ref: tinyknn/_fast_pq_256.pyx:65-123,
// tinyknn/_fast_pq.pyx:114-206, insert :274-307), and knn_brute1 / bottom_k (ref: tinyknn/utils.py:22-25, :89-92).
//
// Why one kernel: the scan is integer-issue bound, the exact heap replay is a chain of dependent shared-memory
// accesses, the rescoring is a latency-bound row gather. As separate launches they run one after the other and
// each leaves most of the SM idle; here a CTA owns a GROUP of G queries and walks them through the phases
//     LUT preparation -> scan of the probed lists (estimates to an L2-resident scratch of the CTA) -> exact
//     recomputation of the chunks whose certificate failed -> queue replay of the reference heap -> labels ->
//     exact distances of the pass_1 candidates -> k nearest
// while the other CTAs of the SM are in other phases, so the pipes overlap. The estimates never leave the chip's
// L2 as a (Q x scanned) array and no segment plan is needed: a query's estimates are one contiguous stream.
//
// Every phase reuses the device code of the stand-alone kernels (tkb_scan_core.cuh; the replay follows
// replay_rq_kernel in tkb_heap.cu slot for slot), so results are identical to the unfused path bit for bit.
// Probe lists with Python-wrapped (negative) entries can visit a list twice; the reference's label dedupe
// (ref: _fast_pq.pyx:284-287) is then reproduced on canonical stream positions (labels are unique across
// lists, so equal labels <=> same list and same position).
#include <stdlib.h>

#include "tkb_scan_core.cuh"
#include "tkb_rescore_core.cuh"

namespace tkb {

constexpr int FZ_THREADS = 256;
constexpr int FZ_WARPS = FZ_THREADS / 32;
constexpr uint32_t FZ_EMPTY = 0xffffffffu;

struct FusedArgs {
    const uint4 *nat; const int64_t *list_chunk_off; const int32_t *list_size; int n_lists; int M;
    const uint8_t *tables; const int32_t *probes; int Q; int P;
    const int64_t *ids;
    const void *rows; int64_t n_rows; int d; const float *queries;
    int R; int k;
    unsigned char *scratch;                 // per-CTA estimate streams
    int64_t scratch_per_cta;                // bytes
    int *work_counter;                      // dynamic group scheduling
    unsigned long long *flagged;            // statistics: chunks recomputed exactly
    int64_t *out_ids; void *out_dists; int32_t *out_count;
    int64_t *heap_idx; int32_t *heap_val;   // optional (Q x R): the reference's final heap arrays
    int G, QCAP, LPW, patch_cap, n_groups;
};

struct FusedSmem {                          // byte offsets into dynamic shared memory
    size_t rows, H, QU, taken, meta, seg_c0, seg_n, seg_list, seg_first, cum, qtot, st, patch_f, patch_c, misc, total;
};

__host__ __device__ inline size_t fz_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline FusedSmem fused_layout(int G, int R, int P, int M, int QCAP, int patch_cap)
{
    FusedSmem L;
    size_t o = 0;
    L.rows = o;      o += (size_t)G * M * 16;
    L.H = o;         o += (size_t)(R + 1) * G * 8;
    // queue region; reused for LUT statistics (3 ints per row) before the scan and for labels + distances after the replay
    size_t qu = (size_t)G * (QCAP + 1) * 8;
    const size_t lutstat = (size_t)G * M * 12, fin = (size_t)G * R * 16;
    if (qu < lutstat) qu = lutstat;
    if (qu < fin) qu = fin;
    L.QU = o;        o += fz_align(qu, 16);
    L.taken = o;     o += fz_align((size_t)G * R, 16);
    L.meta = o;      o += (size_t)G * sizeof(LutMeta);
    L.seg_c0 = o;    o += (size_t)G * P * 8;
    L.seg_n = o;     o += (size_t)G * P * 4;
    L.seg_list = o;  o += (size_t)G * P * 4;
    L.seg_first = o; o += (size_t)G * P * 4;
    L.cum = o;       o += (size_t)G * (P + 1) * 4;
    L.qtot = o;      o += fz_align((size_t)(G + 1) * 4, 8);
    L.st = o;        o += (size_t)G * 6 * 4;               // cursor, seg, bound, count, round, dup
    L.patch_f = o;   o += (size_t)patch_cap * 4;
    L.patch_c = o;   o += (size_t)patch_cap * 4;
    L.misc = o;      o += 16;                              // [0] patch count, [1] current group
    L.total = fz_align(o, 16);
    return L;
}

__device__ __forceinline__ uint4 ldg_cg_u4(const uint4 *p)
{
#ifdef TKB_EMULATE                       // tests/emulate: the source compiled for the CPU, no PTX
    return *p;
#else
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#endif
}

__device__ __forceinline__ uint32_t fz_cmp_lt4s(uint32_t est4, uint32_t bound4) { return __vcmplts4(est4, bound4); }

__device__ __forceinline__ uint32_t fz_cand_mask16(const uint4 e, int bound)
{
    const uint32_t b4 = (uint32_t)(bound & 0xff) * 0x01010101u;
    const uint32_t m0 = fz_cmp_lt4s(e.x, b4) & 0x80808080u, m1 = fz_cmp_lt4s(e.y, b4) & 0x80808080u;
    const uint32_t m2 = fz_cmp_lt4s(e.z, b4) & 0x80808080u, m3 = fz_cmp_lt4s(e.w, b4) & 0x80808080u;
    auto pack = [](uint32_t m) { return ((m >> 7) | (m >> 14) | (m >> 21) | (m >> 28)) & 0xfu; };
    return pack(m0) | (pack(m1) << 4) | (pack(m2) << 8) | (pack(m3) << 12);
}

// exact distances of U rows at a time with the arithmetic of gather_dists_kernel (tkb_rescore_core.cuh): the partial
// sums of all U rows are formed before the first shuffle, so U row reads are in flight per warp.
template <typename T, int U>
__device__ __forceinline__ void row_dists(const T *const (&y)[U], const float *__restrict__ x, int d, int lane, T (&out)[U])
{
    T acc[U];
    warp_rows_dispatch(y, x, d, lane, acc);
#pragma unroll
    for (int u = 0; u < U; u++) out[u] = warp_tree_sum<T>(acc[u]);
}

template <int ORDER, typename T>
__global__ void __launch_bounds__(FZ_THREADS, 3)
ivf_fused_kernel(const FusedArgs a)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int G = a.G, R = a.R, P = a.P, M = a.M, Ph = a.M >> 1, QCAP = a.QCAP;
    const FusedSmem L = fused_layout(G, R, P, M, QCAP, a.patch_cap);
    uint4 *rows = reinterpret_cast<uint4 *>(sm + L.rows);
    uint2 *H = reinterpret_cast<uint2 *>(sm + L.H);                 // [R+1][G]: slot j of query t at j*G+t; slot R = sentinel
    uint2 *QU = reinterpret_cast<uint2 *>(sm + L.QU);               // [G][QCAP+1]
    int *lutstat = reinterpret_cast<int *>(sm + L.QU);              // [G*M][3]   (before the scan)
    int64_t *lab = reinterpret_cast<int64_t *>(sm + L.QU);          // [G][R]     (after the replay)
    T *dist = reinterpret_cast<T *>(sm + L.QU + (size_t)G * R * 8); // [G][R]
    unsigned char *taken = sm + L.taken;                            // [G][R]
    LutMeta *meta = reinterpret_cast<LutMeta *>(sm + L.meta);
    int64_t *seg_c0 = reinterpret_cast<int64_t *>(sm + L.seg_c0);   // first chunk of the list in the code array
    int *seg_n = reinterpret_cast<int *>(sm + L.seg_n);             // true list size
    int *seg_list = reinterpret_cast<int *>(sm + L.seg_list);       // resolved list id, -1 = slot does not exist
    int *seg_first = reinterpret_cast<int *>(sm + L.seg_first);     // first slot that visits the same list
    int *cum = reinterpret_cast<int *>(sm + L.cum);                 // [G][P+1] real chunks before slot s
    int *qtot = reinterpret_cast<int *>(sm + L.qtot);               // [G+1] chunks before query t (CTA stream)
    int *s_cursor = reinterpret_cast<int *>(sm + L.st);
    int *s_seg = s_cursor + G, *s_bound = s_seg + G, *s_count = s_bound + G, *s_round = s_count + G, *s_dup = s_round + G;
    uint32_t *patch_f = reinterpret_cast<uint32_t *>(sm + L.patch_f);
    uint32_t *patch_c = reinterpret_cast<uint32_t *>(sm + L.patch_c);
    int *misc = reinterpret_cast<int *>(sm + L.misc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *est = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
    const bool fast_allowed = ORDER != TKB_ORDER_SSE;               // signed SSE order: every prefix clamps, step-by-step fold only

    for (;;) {
        // ---- next group of queries (dynamic: groups differ a lot in cost) -------------------------------
        __syncthreads();
        if (tid == 0) { misc[1] = atomicAdd(a.work_counter, 1); misc[0] = 0; }
        __syncthreads();
        const int group = misc[1];
        if (group >= a.n_groups) break;
        const int q0 = group * G;

        // ---- phase 0: segment tables, heaps, LUT rows -----------------------------------------------------
        for (int i = tid; i < (R + 1) * G; i += FZ_THREADS)
            H[i] = i < R * G ? make_uint2(FZ_EMPTY, 127u) : make_uint2(FZ_EMPTY, (uint32_t)INT32_MIN);
        for (int t = tid; t < G; t += FZ_THREADS) {
            const int q = q0 + t;
            int *c = cum + (size_t)t * (P + 1);
            int run = 0, dup = 0;
            for (int s = 0; s < P; s++) {
                int l = q < a.Q ? a.probes[(size_t)q * P + s] : PROBE_SKIP;
                int nc = 0, n = 0;
                int64_t c0 = 0;
                if (l != PROBE_SKIP) {
                    if (l < 0) { l += a.n_lists; dup = 1; }           // Python-wrapped index (ref: ivf.py:141): lists may repeat
                    c0 = a.list_chunk_off[l];
                    n = a.list_size[l];
                    nc = (n + 15) >> 4;
                } else {
                    l = -1;
                }
                seg_c0[t * P + s] = c0; seg_n[t * P + s] = n; seg_list[t * P + s] = l; seg_first[t * P + s] = s;
                c[s] = run;
                run += nc;
            }
            c[P] = run;
            if (dup) {
                dup = 0;
                for (int s = 1; s < P; s++)
                    for (int s2 = 0; s2 < s; s2++)
                        if (seg_list[t * P + s] >= 0 && seg_list[t * P + s2] == seg_list[t * P + s]) {
                            seg_first[t * P + s] = s2; dup = 1; break;
                        }
            }
            s_cursor[t] = 0; s_seg[t] = 0; s_bound[t] = 127; s_count[t] = 0; s_round[t] = 0; s_dup[t] = dup;
        }
        for (int i = tid; i < G * M; i += FZ_THREADS) {              // LUT rows, biased to >= 0 (see prepare_lut)
            const int t = i / M, j = i - t * M, q = q0 + t;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (q < a.Q) r = reinterpret_cast<const uint4 *>(a.tables + (size_t)q * M * 16)[j];
            const uint32_t ws[4] = {r.x, r.y, r.z, r.w};
            int mn = 1 << 30, mx = -(1 << 30);
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const int tv = (int)(int8_t)((ws[c >> 2] >> (8 * (c & 3))) & 0xffu);
                mn = min(mn, tv); mx = max(mx, tv);
            }
            const int bias = -mn;
            lutstat[3 * i + 0] = bias; lutstat[3 * i + 1] = max(0, -mn); lutstat[3 * i + 2] = mx + bias;
            uint32_t o[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t v = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int tv = (int)(int8_t)((ws[w] >> (8 * c)) & 0xffu);
                    v |= (uint32_t)((tv + bias) & 0xff) << (8 * c);
                }
                o[w] = v;
            }
            rows[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        if (tid < G) {
            const int t = tid;
            int bias[2] = {0, 0}, N[2] = {0, 0}, range = 0;
            for (int j = 0; j < M; j++) {
                const int l = (j >> 1) & 1, *st = lutstat + 3 * (t * M + j);
                bias[l] += st[0]; N[l] += st[1]; range = max(range, st[2]);
            }
            LutMeta m;
            m.eligible = fast_allowed && range <= 31 && N[0] <= 128 && N[1] <= 128;
            m.bias_tot = bias[0] + bias[1];
            int kk[2] = {127, 127};                                  // certificate thresholds as in prepare_lut (tkb_scan_core.cuh)
            {
                int Pm[2] = {0, 0}, suf[2] = {N[0], N[1]};
                bool found[2] = {false, false};
                for (int j = 0; j < M; j++) {
                    const int l = (j >> 1) & 1, *st = lutstat + 3 * (t * M + j);
                    suf[l] -= st[1];
                    Pm[l] += max(0, st[2] - st[0]);
                    if (!found[l] && Pm[l] > 127) { found[l] = true; kk[l] = 127 - suf[l]; }
                }
            }
            m.k0 = kk[0] + bias[0];
            m.k1 = kk[1] + bias[1];
            meta[t] = m;
        }
        if (tid == 0) {
            int run = 0;
            for (int t = 0; t < G; t++) { qtot[t] = run; run += cum[(size_t)t * (P + 1) + P]; }
            qtot[G] = run;
        }
        __syncthreads();

        // ---- phase 1: scan. One thread per 16-vector chunk of the CTA's flat (query, list, chunk) stream ------
        {
            const int total = qtot[G];
            int t = 0, s = 0;
            for (int f = tid; f < total; f += FZ_THREADS) {
                while (f >= qtot[t + 1]) { t++; s = 0; }
                const int cc = f - qtot[t];
                const int *c = cum + (size_t)t * (P + 1);
                while (cc >= c[s + 1]) s++;
                const int64_t ch = seg_c0[t * P + s] + (cc - c[s]);
                const LutMeta m = meta[t];
                uint4 o;
                if (m.eligible) {
                    bool flagged;
                    o = scan_chunk_fast<true>(a.nat, ch, Ph, rows + (size_t)t * M, m, flagged);
                    if (flagged) {
                        const int i = atomicAdd(&misc[0], 1);
                        if (i < a.patch_cap) { patch_f[i] = (uint32_t)f; patch_c[i] = (uint32_t)ch; }
                        else o = scan_chunk_exact_cold<ORDER, true>(a.nat, ch, Ph, a.tables + (size_t)(q0 + t) * M * 16);
                    }
                } else {
                    o = scan_chunk_exact<ORDER, true>(a.nat, ch, Ph, a.tables + (size_t)(q0 + t) * M * 16);
                }
                *reinterpret_cast<uint4 *>(est + 16 * (size_t)f) = o;
            }
        }
        __syncthreads();
        // ---- phase 1b: chunks whose certificate failed, recomputed with the reference's step-by-step fold ----
        {
            const int cnt = min(misc[0], a.patch_cap);
            if (tid == 0 && misc[0] > 0 && a.flagged) atomicAdd(a.flagged, (unsigned long long)misc[0]);
            for (int i = tid >> 4; i < cnt; i += FZ_THREADS / 16) {
                const uint32_t f = patch_f[i];
                int t = 0;
                while ((int)f >= qtot[t + 1]) t++;
                const int e = exact_vector<ORDER, true>(a.nat, (int64_t)patch_c[i], Ph, tid & 15,
                                                        a.tables + (size_t)(q0 + t) * M * 16);
                est[16 * (size_t)f + (tid & 15)] = (unsigned char)e;
            }
        }
        __syncthreads();

        // ---- phase 2: exact queue replay of the reference heap (see replay_rq_kernel, tkb_heap.cu) -----------
        for (;;) {
            bool more = false;
            for (int t = warp; t < G; t += FZ_WARPS) {
                const int *c = cum + (size_t)t * (P + 1);
                const int total = c[P];
                const int cursor = s_cursor[t];
                if (cursor >= total) { if (lane == 0) s_count[t] = 0; continue; }
                const int bound = s_bound[t];
                int W = s_round[t] == 0 ? ((R + 15) >> 4) + 1 : (cursor < 32 ? 32 : cursor);
                if (W > (1 << 16)) W = 1 << 16;
                int end = (total - cursor < W) ? total : cursor + W;
                int count = 0, sg = s_seg[t];
                uint2 *qu = QU + (size_t)t * (QCAP + 1);
                const uint4 *ep = reinterpret_cast<const uint4 *>(est) + qtot[t];
                for (int base = cursor; base < end; base += 32) {
                    const int cc = base + lane;
                    const bool act = cc < end;
                    uint32_t m = 0;
                    uint4 e = make_uint4(0, 0, 0, 0);
                    int sl = sg;
                    if (act) {
                        while (cc >= c[sl + 1]) sl++;
                        const int local = cc - c[sl];
                        e = ldg_cg_u4(ep + cc);
                        const int rem = seg_n[t * P + sl] - 16 * local;         // >= 1: only real chunks are in the stream
                        m = fz_cand_mask16(e, bound) & (rem >= 16 ? 0xffffu : ((1u << rem) - 1u));
                    }
                    const int last = end - 1 - base;
                    sg = __shfl_sync(FULL, sl, last < 31 ? last : 31);
                    const int cnt = __popc(m);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                    const int tot = __shfl_sync(FULL, incl, 31);
                    bool cut = false;
                    int keep_tot = tot;
                    if (count + tot > QCAP) {                                    // queue full: stop at a chunk boundary
                        const unsigned over = __ballot_sync(FULL, count + incl > QCAP);
                        const int cl = __ffs(over) - 1;
                        keep_tot = __shfl_sync(FULL, incl - cnt, cl);
                        sg = __shfl_sync(FULL, sl, cl);
                        if (lane >= cl) m = 0;
                        end = base + cl;
                        cut = true;
                    }
                    int kq = count + incl - cnt;
                    const uint32_t ws[4] = {e.x, e.y, e.z, e.w};
                    while (m) {
                        const int v = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t byte = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
                        qu[kq++] = make_uint2(16u * (uint32_t)cc + v, (uint32_t)(int)(int8_t)byte);
                    }
                    count += keep_tot;
                    if (cut) break;
                }
                if (lane == 0) { s_cursor[t] = end; s_seg[t] = sg; s_count[t] = count; s_round[t] = 1; }
                more = true;
            }
            if (!__syncthreads_or(more)) break;
            if (warp * a.LPW < G) {
                const int t = warp * a.LPW + lane;
                const bool mine = lane < a.LPW && t < G;
                const int cnt = mine ? s_count[t] : 0;
                int mx = cnt;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
                if (mine) {
                    const uint2 *qu = QU + (size_t)t * (QCAP + 1);
                    unsigned char *hb = reinterpret_cast<unsigned char *>(H + t);       // slot j at hb + j*S
                    const uint32_t S = (uint32_t)G * 8u, lim = (uint32_t)R * S;
                    const bool dupq = s_dup[t] != 0;
                    const int *c = cum + (size_t)t * (P + 1);
                    uint32_t cur_chunk = FZ_EMPTY;
                    int frozen = 0;
                    for (int i = 0; i < mx; i++) {
                        if (i >= cnt) continue;
                        uint2 rec = qu[i];
                        const int ev = (int)rec.y;
                        if ((rec.x >> 4) != cur_chunk) {                       // first candidate of a chunk: freeze the bound
                            cur_chunk = rec.x >> 4;
                            frozen = (int)reinterpret_cast<const uint2 *>(hb)->y;
                        }
                        if (ev >= frozen) continue;
                        if (dupq) {                                            // a list visited twice: the reference dedupes by label
                            int s = 0;
                            while ((int)cur_chunk >= c[s + 1]) s++;
                            const int fs = seg_first[t * P + s];
                            if (fs != s) {
                                const uint32_t canon = rec.x - 16u * (uint32_t)c[s] + 16u * (uint32_t)c[fs];
                                bool found = false;
                                for (uint32_t o = 0; o < lim; o += S)
                                    if (reinterpret_cast<const uint2 *>(hb + o)->x == canon) { found = true; break; }
                                if (found) continue;
                                rec.x = canon;
                            }
                        }
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

        // ---- phase 3a: labels of the heap slots (ids are read once, here) -------------------------------------
        for (int i = tid; i < R * G; i += FZ_THREADS) {
            const int t = i / R, j = i - t * R, q = q0 + t;
            const uint2 e = H[(size_t)j * G + t];
            int64_t label = -1;
            if (e.x != FZ_EMPTY && q < a.Q) {
                const int cc = (int)(e.x >> 4), v = (int)(e.x & 15u);
                const int *c = cum + (size_t)t * (P + 1);
                int lo = 0, hi = P;
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (c[mid] <= cc) lo = mid; else hi = mid; }
                label = a.ids[16 * (seg_c0[t * P + lo] + (cc - c[lo])) + v];
            }
            // lab aliases the queue region: every queue was drained before the loop above ended
            lab[i] = label;
            taken[i] = label == -1 ? 1 : 0;                               // ref: ivf.py:154-155
            if (a.heap_idx && q < a.Q) { a.heap_idx[(size_t)q * R + j] = label; a.heap_val[(size_t)q * R + j] = (int)e.y; }
        }
        __syncthreads();
        // ---- phase 3b: exact distances of the candidates (ref: utils.py:89-91), 4 rows in flight per warp -----
        {
            const T *rowsT = reinterpret_cast<const T *>(a.rows);
            constexpr int U = 4;
            const int nquad = (R + U - 1) / U;
            for (int qi = warp; qi < G * nquad; qi += FZ_WARPS) {
                const int t = qi / nquad, base = (qi - t * nquad) * U, q = q0 + t;
                if (q >= a.Q) continue;
                const T *y[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    y[u] = nullptr;
                    if (base + u < R) {
                        const int64_t row = lab[(size_t)t * R + base + u];
                        if (row >= 0 && row < a.n_rows) y[u] = rowsT + row * a.d;
                    }
                }
                T out[U];
                row_dists<T, U>(y, a.queries + (size_t)q * a.d, a.d, lane, out);
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (base + u < R) dist[(size_t)t * R + base + u] = y[u] ? out[u] : (T)NAN;
                }
            }
        }
        __syncthreads();
        // ---- phase 3c: the k nearest (ref: ivf.py:158-163; device order: ascending distance, ties by heap slot) -
        for (int t = warp; t < G; t += FZ_WARPS) {
            const int q = q0 + t;
            if (q >= a.Q) continue;
            const int64_t *hi = lab + (size_t)t * R;
            const T *dq = dist + (size_t)t * R;
            unsigned char *tk = taken + (size_t)t * R;
            int64_t *oi = a.out_ids + (size_t)q * a.k;
            T *od = a.out_dists ? reinterpret_cast<T *>(a.out_dists) + (size_t)q * a.k : nullptr;
            int n_valid = 0;
            for (int s0 = 0; s0 < R; s0 += 32) {
                const int s = s0 + lane;
                n_valid += __popc(__ballot_sync(FULL, s < R && !tk[s]));
            }
            if (n_valid <= a.k) {                                         // ref: ivf.py:158-159 -- survivors in heap order
                int w = 0;
                for (int s0 = 0; s0 < R; s0 += 32) {
                    const int s = s0 + lane;
                    const bool ok = (s < R) && !tk[s];
                    const unsigned b = __ballot_sync(FULL, ok);
                    if (ok) {
                        const int o = w + __popc(b & ((1u << lane) - 1));
                        oi[o] = hi[s];
                        if (od) od[o] = dq[s];
                    }
                    w += __popc(b);
                }
                for (int o = n_valid + lane; o < a.k; o += 32) { oi[o] = -1; if (od) od[o] = (T)INFINITY; }
                if (lane == 0) a.out_count[q] = n_valid;
                continue;
            }
            for (int j = 0; j < a.k; j++) {
                T bd = (T)INFINITY; int bs = INT32_MAX;
                for (int s = lane; s < R; s += 32) {
                    if (tk[s]) continue;
                    T v = dq[s];
                    if (v != v) v = (T)INFINITY;
                    if (v < bd || (v == bd && s < bs)) { bd = v; bs = s; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const T od2 = __shfl_xor_sync(FULL, bd, o);
                    const int os = __shfl_xor_sync(FULL, bs, o);
                    if (od2 < bd || (od2 == bd && os < bs)) { bd = od2; bs = os; }
                }
                if (lane == 0) { tk[bs] = 1; oi[j] = hi[bs]; if (od) od[j] = dq[bs]; }
                __syncwarp();
            }
            if (lane == 0) a.out_count[q] = a.k;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct FusedGeom { int G, QCAP, LPW, patch_cap, ctas_per_sm, n_sms; size_t smem; };

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int ORDER, typename T>
static int fused_occupancy(size_t smem, int &ctas)
{
    TKB_CUDA(cudaFuncSetAttribute(ivf_fused_kernel<ORDER, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, ivf_fused_kernel<ORDER, T>, FZ_THREADS, smem));
    return TKB_OK;
}

static int fused_geometry(int Q, int R, int P, int M, int order, int rows_dtype, FusedGeom &g)
{
    int dev = 0;
    TKB_CUDA(cudaGetDevice(&dev));
    TKB_CUDA(cudaDeviceGetAttribute(&g.n_sms, cudaDevAttrMultiProcessorCount, dev));
    g.QCAP = 2 * R < 64 ? 64 : 2 * R;
    g.patch_cap = 256;
    int G = env_int("TKB_FUSED_G", 8);
    if (G < 1) G = 1;
    if (G > 32) G = 32;
    // groups must cover the machine a few times over, and the group state should leave room for 3 CTAs per SM
    while (G > 1 && (Q + G - 1) / G < 3 * g.n_sms) G >>= 1;
    while (G > 1 && fused_layout(G, R, P, M, g.QCAP, g.patch_cap).total > 72 * 1024) G >>= 1;
    g.G = G;
    g.smem = fused_layout(G, R, P, M, g.QCAP, g.patch_cap).total;
    if (g.smem > 200 * 1024) return set_err(TKB_ERR_INVALID, "invalid argument: heap/probe state does not fit shared memory for the fused query kernel");
    int lpw = env_int("TKB_FUSED_LPW", 4);
    if (lpw != 1 && lpw != 2 && lpw != 4 && lpw != 8 && lpw != 16 && lpw != 32) lpw = 4;
    while (lpw * FZ_WARPS < G) lpw <<= 1;
    g.LPW = lpw;
    int rc;
    if (order == TKB_ORDER_AVX)
        rc = rows_dtype == TKB_DTYPE_F32 ? fused_occupancy<TKB_ORDER_AVX, float>(g.smem, g.ctas_per_sm)
                                         : fused_occupancy<TKB_ORDER_AVX, double>(g.smem, g.ctas_per_sm);
    else
        rc = rows_dtype == TKB_DTYPE_F32 ? fused_occupancy<TKB_ORDER_SSE, float>(g.smem, g.ctas_per_sm)
                                         : fused_occupancy<TKB_ORDER_SSE, double>(g.smem, g.ctas_per_sm);
    if (rc) return rc;
    if (g.ctas_per_sm < 1) return set_err(TKB_ERR_CUDA, "fused query kernel does not fit on an SM");
    return TKB_OK;
}

static int64_t fused_per_cta(int G, int P, int64_t max_list_chunks)
{
    return 16 * (int64_t)G * P * (max_list_chunks > 0 ? max_list_chunks : 1);
}

int fused_workspace_bytes(int Q, int P, int R, int M, int order, int rows_dtype, int64_t max_list_chunks, int64_t *bytes)
{
    TKB_REQUIRE(bytes, "null pointer");
    TKB_REQUIRE(Q >= 0 && P >= 0 && R >= 0 && M > 0 && max_list_chunks >= 0, "bad extent");
    if (Q == 0 || P == 0 || R == 0) { *bytes = 64; return TKB_OK; }
    FusedGeom g;
    if (int rc = fused_geometry(Q, R, P, M, order, rows_dtype, g)) return rc;
    int64_t ctas = (int64_t)g.n_sms * g.ctas_per_sm;
    const int64_t groups = (Q + g.G - 1) / g.G;
    if (ctas > groups) ctas = groups;
    *bytes = 64 + ctas * fused_per_cta(g.G, P, max_list_chunks);
    return TKB_OK;
}

int launch_ivf_query_fused(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                           const uint8_t *tables, const int32_t *probes, int Q, int P, const int64_t *ids,
                           const void *rows, int rows_dtype, int64_t n_rows, int d, const float *queries,
                           int R, int k, int order, int64_t max_list_chunks,
                           int64_t *out_ids, void *out_dists, int32_t *out_count, int64_t *heap_idx, int32_t *heap_val,
                           void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    TKB_REQUIRE(order == TKB_ORDER_SSE || order == TKB_ORDER_AVX, "order must be TKB_ORDER_SSE or TKB_ORDER_AVX");
    TKB_REQUIRE(M > 0 && M % 2 == 0 && M <= 1024, "M (sub-quantizers) must be a positive multiple of 2");
    TKB_REQUIRE(order != TKB_ORDER_AVX || M % 4 == 0, "avx order needs M % 4 == 0 (ref: fast_pq.py:24 dpad)");
    TKB_REQUIRE(Q >= 0 && P >= 0 && R >= 0 && k >= 0 && n_lists > 0 && d > 0 && n_rows > 0, "bad extent");
    if (Q == 0) return TKB_OK;
    TKB_REQUIRE(P > 0 && R > 0 && k > 0, "P, R and k must be positive");
    TKB_REQUIRE(native && list_chunk_off && list_size && tables && probes && ids && rows && queries, "null pointer");
    TKB_REQUIRE(out_ids && out_count && workspace, "null pointer");
    TKB_REQUIRE((heap_idx == nullptr) == (heap_val == nullptr), "heap_idx and heap_val go together");
    TKB_REQUIRE(rows_dtype == TKB_DTYPE_F32 || rows_dtype == TKB_DTYPE_F64, "rows dtype must be f32 or f64");
    TKB_REQUIRE((uintptr_t)workspace % 16 == 0 && (uintptr_t)native % 16 == 0 && (uintptr_t)tables % 16 == 0,
                "device pointers must be 16-byte aligned");
    TKB_REQUIRE((int64_t)P * max_list_chunks < (1LL << 27), "too many scanned vectors per query for the fused kernel");
    FusedGeom g;
    if (int rc = fused_geometry(Q, R, P, M, order, rows_dtype, g)) return rc;
    const int64_t per_cta = fused_per_cta(g.G, P, max_list_chunks);
    TKB_REQUIRE((int64_t)g.G * P * max_list_chunks < (1LL << 31), "group too large");
    const int groups = (Q + g.G - 1) / g.G;
    int64_t ctas = (int64_t)g.n_sms * g.ctas_per_sm;
    if (ctas > groups) ctas = groups;
    if (ctas > (workspace_bytes - 64) / per_cta) ctas = (workspace_bytes - 64) / per_cta;
    TKB_REQUIRE(ctas >= 1, "workspace too small for the fused query kernel (see tkb_ivf_query_fused_workspace)");
    FusedArgs a;
    a.nat = reinterpret_cast<const uint4 *>(native); a.list_chunk_off = list_chunk_off; a.list_size = list_size;
    a.n_lists = n_lists; a.M = M; a.tables = tables; a.probes = probes; a.Q = Q; a.P = P; a.ids = ids;
    a.rows = rows; a.n_rows = n_rows; a.d = d; a.queries = queries; a.R = R; a.k = k;
    a.scratch = reinterpret_cast<unsigned char *>(workspace) + 64; a.scratch_per_cta = per_cta;
    a.work_counter = reinterpret_cast<int *>(workspace);
    a.flagged = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(workspace) + 8);
    a.out_ids = out_ids; a.out_dists = out_dists; a.out_count = out_count; a.heap_idx = heap_idx; a.heap_val = heap_val;
    a.G = g.G; a.QCAP = g.QCAP; a.LPW = g.LPW; a.patch_cap = g.patch_cap; a.n_groups = groups;
    TKB_CUDA(cudaMemsetAsync(workspace, 0, 64, st));
    const unsigned grid = (unsigned)ctas;
    if (order == TKB_ORDER_AVX) {
        if (rows_dtype == TKB_DTYPE_F32) ivf_fused_kernel<TKB_ORDER_AVX, float><<<grid, FZ_THREADS, g.smem, st>>>(a);
        else                             ivf_fused_kernel<TKB_ORDER_AVX, double><<<grid, FZ_THREADS, g.smem, st>>>(a);
    } else {
        if (rows_dtype == TKB_DTYPE_F32) ivf_fused_kernel<TKB_ORDER_SSE, float><<<grid, FZ_THREADS, g.smem, st>>>(a);
        else                             ivf_fused_kernel<TKB_ORDER_SSE, double><<<grid, FZ_THREADS, g.smem, st>>>(a);
    }
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
