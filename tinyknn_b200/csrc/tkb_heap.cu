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
ivf_replay_kernel(const uint8_t *__restrict__ est, int64_t slot_stride,
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
        const int64_t nc = list_chunk_off[l + 1] - c0;
        replay_segment<SIGNED>(est + ((size_t)q * P + s) * slot_stride, nc, list_size[l], ids + 16 * c0,
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

int launch_ivf_replay(const uint8_t *est, int64_t slot_stride, const int64_t *list_chunk_off,
                      const int32_t *list_size, int n_lists, const int64_t *ids, const int32_t *probes,
                      int Q, int P, int64_t *heap_idx, int32_t *heap_val, int R, int signd,
                      cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0 && R >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || R == 0 || P == 0) return TKB_OK;
    TKB_REQUIRE(est && list_chunk_off && list_size && ids && probes && heap_idx && heap_val, "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0 && (uintptr_t)est % 16 == 0, "est must be 16-byte aligned/strided");
    const unsigned blocks = (unsigned)((Q + REPLAY_WARPS - 1) / REPLAY_WARPS);
    if (signd) ivf_replay_kernel<true><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, slot_stride, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val, R);
    else       ivf_replay_kernel<false><<<blocks, 32 * REPLAY_WARPS, 0, st>>>(est, slot_stride, list_chunk_off, list_size, n_lists, ids, probes, Q, P, heap_idx, heap_val, R);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
