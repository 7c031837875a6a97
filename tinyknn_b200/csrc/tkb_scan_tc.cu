// tkb_scan_tc.cu -- the LIST-MAJOR PQ scan on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces, for batches in which several queries probe the same inverted list, the same reference functions as
// tkb_scan_fast.cu (compute_block_dists_avx + the scan half of query_pq_avx; ref: tinyknn/_fast_pq_256.pyx:65-156,
// tinyknn/ivf.py:140-150), and produces the same bytes: one estimate per (query, scanned vector) in the compact segment
// layout of tkb_ivf_plan_dev, plus the optional chunk minima.
//
// Why a GEMM is exact here. Per accumulation lane l (sub-quantizers with (j >> 1) & 1 == l) the
// reference folds M/2 LUT entries with a saturating add after every step. The UNSATURATED lane sums are a linear map
//     S_l[vector, query] = sum_j  onehot(code[vector][j])[0..16) . T_query[j][0..16)          (j in lane l)
// i.e. (one-hot codes, 128 x 16*M/2, int8) x (LUT slab, 16*M/2 x queries, int8) -> int32, exactly what
// `tcgen05.mma.kind::i8` computes. With N_l = sum_j max(0, -min_c T[j][c]) per (query, lane): N_l <= 128 means no
// prefix of the fold can go below -128, S_l + N_l <= 127 means none can exceed 127, and then the reference's lane value
// IS S_l and its estimate is clamp(S_0 + S_1). The certificate is checked per (vector, query); a pair that fails it is
// refolded step by step (the reference's own recurrence) from the LUT slab in shared memory before the tile is written;
// a query with N_l > 128 is left to the CUDA-core kernel (tkb_scan_fast.cu) altogether.
//
// Why list-major. The CUDA-core scan reads a list's codes once per (query, list) and spends ~1.7 ALU instructions per
// lookup: on the 100M x 128 index it is bound by HBM *and* by the integer pipe at the same time. Here a 128-vector tile
// of a list is expanded to one-hot form ONCE (by the CUDA cores, straight into tensor memory as the A operand) and
// multiplied with the LUTs of ALL the queries of the batch that probe this list (B operand, shared memory): the code bytes
// are read once per list and batch, and the per-(vector, query) work left on the CUDA cores is the epilogue.
//
// Work decomposition: a counting sort of the batch's (query, probe slot) pairs by list (tc_count / tc_offsets / tc_fill),
// then a persistent kernel, one CTA per SM, pulls items = (list, group of <= NT queries, range of tiles).
//
//   A (TMEM, K-major)  : lane r = vector r of the tile, 32-bit column 4j + w = bytes 4w..4w+3 of onehot(code[r][j])
//   B (SMEM, K-major, no swizzle): unit (j, i) = the 16 LUT bytes T_i[j][0..16) at (j * NT + (i / 8) * 8 + i % 8) * 16:
//                        8-query core matrices of 128 B, stride-byte-offset 128 B, leading-byte-offset NT * 16 B
//   D (TMEM)           : two accumulators (lane 0 / lane 1), column n = query n of the group, int32
//   one MMA per sub-quantizer pair p (K = 32 bytes = 2 units), accumulating into D[p & 1].
#include <stdlib.h>

#include "tkb_scan_core.cuh"

namespace tkb {

constexpr int TC_NT = 64;                       // queries per group (UMMA N <= 64: 2 x 2 x 64 accumulator columns + 256 of A)
constexpr int TC_TILES_PER_ITEM = 128;          // tiles (of 128 vectors) per work item: long lists are split
constexpr int TC_OUT_STRIDE = TC_NT / 4 + 1;    // words per row of the transposed output tile (see tc_out_addr)

struct TcQueryMeta { int16_t elig, k0, k1, pad; };

// Word (row = vector of the tile, c = four queries) of the transposed output tile. Row stride 17 keeps the epilogue's stores (32
// consecutive rows, one column) on 32 banks; rotating the column by 4 per 32 rows does the same for the loads of the copy-out
// (rows 16 apart, four consecutive columns).
__host__ __device__ constexpr int tc_out_addr(int row, int c) { return row * TC_OUT_STRIDE + ((c + 4 * (row >> 5)) & 15); }
// cycle counters of one representative warp per role (lane 0), summed over all CTAs
enum TcClock { CK_E_WAIT = 0, CK_E_WORK, CK_M_WAIT_A, CK_M_WAIT_D, CK_M_ISSUE, CK_P_WAIT, CK_P_EPI, CK_P_BAR, CK_P_COPY, CK_P_FLUSH,
               CK_L_WAIT, CK_L_STAGE, CK_TOTAL, CK_COUNT };

// ---- work list -------------------------------------------------------------------------------------------------------
// ws (int32 words): [0] total items, [1] next item (persistent kernel's counter), [2] refolded pairs, [3] tiles multiplied,
// [4] sum over those tiles of the group's N / 16 (statistics), [5..7] pad; then 16 int64 cycle counters (where the roles of
// the kernel spend their time, summed over the CTAs: see TcClock)
// then cnt[n_lists], cursor[n_lists], bucket_off[n_lists + 1], item_off[n_lists + 1], bucket[Q * P].
struct TcWork {
    int *hdr, *cnt, *cursor, *bucket_off, *item_off, *bucket;
    unsigned long long *clk;
    TcQueryMeta *qmeta;
    int64_t *seg_rest;
    uint8_t *skip_q;
};

__host__ __device__ inline size_t tc_carve(void *base, int Q, int P, int n_lists, TcWork *w)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
    const size_t hdr = take(32 + 8 * 16), cnt = take(4 * (size_t)n_lists), cur = take(4 * (size_t)n_lists);
    const size_t bo = take(4 * ((size_t)n_lists + 1)), io = take(4 * ((size_t)n_lists + 1));
    const size_t bk = take(4 * (size_t)Q * P), qm = take(sizeof(TcQueryMeta) * (size_t)Q), sr = take(8 * (size_t)Q * P);
    const size_t sk = take((size_t)Q);
    if (w) {
        unsigned char *b = reinterpret_cast<unsigned char *>(base);
        w->hdr = reinterpret_cast<int *>(b + hdr); w->cnt = reinterpret_cast<int *>(b + cnt);
        w->clk = reinterpret_cast<unsigned long long *>(b + hdr + 32);
        w->cursor = reinterpret_cast<int *>(b + cur); w->bucket_off = reinterpret_cast<int *>(b + bo);
        w->item_off = reinterpret_cast<int *>(b + io); w->bucket = reinterpret_cast<int *>(b + bk);
        w->qmeta = reinterpret_cast<TcQueryMeta *>(b + qm); w->seg_rest = reinterpret_cast<int64_t *>(b + sr);
        w->skip_q = b + sk;
    }
    return o;
}

// one warp per query: N_l = sum over the lane's rows of max(0, -min_c T[j][c]); eligible iff both <= 128 (no prefix of the fold
// can go below -128). Threshold of the certificate (S_l <= k_l => no prefix can exceed 127, so the fold is the plain sum):
// a prefix that ends at row k is at most Pmax_k = sum of the rows' largest positive entries so far, and at most
// S_l + (negatives the rows AFTER k can still contribute). It can pass 127 only from the first row k* with Pmax_k* > 127 on, where
// the second bound is largest: k_l = 127 - sum_{j > k*} max(0, -min_c T[j][c])  (127 when Pmax never passes 127: then S_l cannot either).
// With the usual tables (row range ~25, minimum ~ -3) that is ~120 instead of 127 - N_l ~ 80: 500 times fewer pairs to refold.
__global__ void tc_query_meta_kernel(const uint8_t *__restrict__ tables, int Q, int M, TcQueryMeta *__restrict__ qmeta,
                                     uint8_t *__restrict__ skip_q)
{
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Q) return;
    int n0 = 0, n1 = 0, rneg = 0, rmax = 0;
    for (int j = lane; j < M; j += 32) {
        const uint4 r = reinterpret_cast<const uint4 *>(tables + (size_t)q * M * 16)[j];
        const uint32_t m4 = __vmins4(__vmins4(r.x, r.y), __vmins4(r.z, r.w));
        const uint32_t m2 = __vmins4(m4, m4 >> 16);
        const int mn = (int)(int8_t)(__vmins4(m2, m2 >> 8) & 0xffu);
        const uint32_t x4 = __vmaxs4(__vmaxs4(r.x, r.y), __vmaxs4(r.z, r.w));
        const uint32_t x2 = __vmaxs4(x4, x4 >> 16);
        const int mx = (int)(int8_t)(__vmaxs4(x2, x2 >> 8) & 0xffu);
        const int c = mn < 0 ? -mn : 0;
        if ((j >> 1) & 1) n1 += c; else n0 += c;
        rneg = c; rmax = mx > 0 ? mx : 0;                          // (row `lane` when M <= 32)
    }
    for (int o = 16; o > 0; o >>= 1) { n0 += __shfl_xor_sync(FULL, n0, o); n1 += __shfl_xor_sync(FULL, n1, o); }
    int k0 = 127 - n0, k1 = 127 - n1;
    if (M <= 32) {                                                   // every lane walks the rows in fold order (shuffles broadcast them)
        int P0 = 0, P1 = 0, s0 = n0, s1 = n1;
        bool f0 = false, f1 = false;
        k0 = k1 = 127;
        for (int j = 0; j < M; j++) {
            const int rm = __shfl_sync(FULL, rmax, j), rn = __shfl_sync(FULL, rneg, j);
            if ((j >> 1) & 1) { s1 -= rn; P1 += rm; if (!f1 && P1 > 127) { f1 = true; k1 = 127 - s1; } }
            else              { s0 -= rn; P0 += rm; if (!f0 && P0 > 127) { f0 = true; k0 = 127 - s0; } }
        }
    }
    if (lane == 0) {
        TcQueryMeta m;
        m.elig = (n0 <= 128 && n1 <= 128) ? 1 : 0;
        m.k0 = (int16_t)k0; m.k1 = (int16_t)k1; m.pad = 0;
        qmeta[q] = m;
        skip_q[q] = (uint8_t)m.elig;             // the CUDA-core kernel skips the queries this kernel handles
    }
}

__global__ void tc_count_kernel(const int32_t *__restrict__ probes, const int64_t *__restrict__ seg_off, int64_t QP, int P,
                                int n_lists, const TcQueryMeta *__restrict__ qmeta, int *__restrict__ cnt,
                                int64_t *__restrict__ seg_rest)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= QP) return;
    int l = probes[e];
    const int64_t so = seg_off[e];
    const bool valid = l != PROBE_SKIP && so >= 0;
    if (valid && l < 0) l += n_lists;                 // Python-wrapped list index (ref: ivf.py:141 indexes a list)
    const bool tc = valid && l >= 0 && l < n_lists && qmeta[e / P].elig;
    if (tc) atomicAdd(cnt + l, 1);
    seg_rest[e] = tc ? -2 : so;                      // -2: taken by the tensor-core path (any negative = absent for the CUDA-core kernel)
}

// one CTA: exclusive prefix sums over the lists of (pairs) and (work items)
__global__ void __launch_bounds__(1024)
tc_offsets_kernel(const int *__restrict__ cnt, const int32_t *__restrict__ list_size, int n_lists, int *__restrict__ bucket_off,
                  int *__restrict__ item_off, int *__restrict__ hdr, int tpi /* tiles per work item */)
{
    __shared__ int s_a[1024], s_b[1024];
    const int tid = threadIdx.x, per = (n_lists + 1023) / 1024;
    const int lo = tid * per, hi = min(n_lists, lo + per);
    int a = 0, b = 0;
    for (int l = lo; l < hi; l++) {
        const int c = cnt[l];
        const int tiles = (((list_size[l] + 15) >> 4) + 7) >> 3;
        a += c;
        b += ((c + TC_NT - 1) / TC_NT) * ((tiles + tpi - 1) / tpi);
    }
    s_a[tid] = a; s_b[tid] = b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int va = tid >= o ? s_a[tid - o] : 0, vb = tid >= o ? s_b[tid - o] : 0;
        __syncthreads();
        s_a[tid] += va; s_b[tid] += vb;
        __syncthreads();
    }
    int ra = s_a[tid] - a, rb = s_b[tid] - b;
    for (int l = lo; l < hi; l++) {
        const int c = cnt[l];
        const int tiles = (((list_size[l] + 15) >> 4) + 7) >> 3;
        bucket_off[l] = ra; item_off[l] = rb;
        ra += c;
        rb += ((c + TC_NT - 1) / TC_NT) * ((tiles + tpi - 1) / tpi);
    }
    if (tid == 1023) { bucket_off[n_lists] = s_a[1023]; item_off[n_lists] = s_b[1023]; hdr[0] = s_b[1023]; }
}

__global__ void tc_fill_kernel(const int32_t *__restrict__ probes, const int64_t *__restrict__ seg_rest, int64_t QP, int n_lists,
                               const int *__restrict__ bucket_off, int *__restrict__ cursor, int *__restrict__ bucket)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= QP || seg_rest[e] != -2) return;
    int l = probes[e];
    if (l < 0) l += n_lists;
    bucket[bucket_off[l] + atomicAdd(cursor + l, 1)] = (int)e;
}

#ifndef TKB_EMULATE                      // tensor memory and tcgen05 have no CPU stand-in (tests/emulate): the entry point reports that
// ---- tcgen05 / mbarrier wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// backoff_ns > 0: a role that usually waits long sleeps between polls instead of taking issue slots from the working warps
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned backoff_ns = 0)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity), "r"(1000000u) : "memory");
        if (done) break;
        if (backoff_ns) __nanosleep(backoff_ns);
    }
}

// one lane of a converged warp (the tcgen05 instructions take their operands from uniform registers: issued from divergent
// code -- `if (lane == 0)` -- the compiler has to wrap every one of them in a broadcast loop)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}\n" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem_d] (+)= A[tmem_a] (128 x 32 int8, TMEM) * B[desc_b] (N x 32 int8, SMEM, K-major)
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

// 8 accumulator columns, the low 16 bits of each, two per register (.pack::16b): exactly the s16x2 operands of the epilogue
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t taddr, uint32_t (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.pack::16b.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}

// K-major, no swizzle: start address, leading-byte-offset (between the two 16-byte K chunks of one MMA), stride-byte-offset
// (between 8-row core matrices), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes & 0x3ffffu) >> 4) << 16) |
           ((uint64_t)((sbo_bytes & 0x3ffffu) >> 4) << 32) | (1ull << 46);
}

// kind::i8 instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_i8(int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// The 16-byte one-hot unit of a 4-bit code c: byte k = (k == c). One PRMT per 32-bit word: the selector nibble of byte b of word
// w is 0 (-> source byte 0 = 0x01) exactly when c == 4w + b; any other nibble value selects a zero byte (or, with bit 3 set, the
// replicated sign bit of a byte that is 0x00 or 0x01, i.e. zero). For words 0..2 the selector is c * 0x1111 - (nibbles 4w..4w+3):
// ONE multiply-add on the FMA pipe (a borrow between nibbles can never fake a zero nibble there -- checked exhaustively); for
// word 3 the borrow chain of c = 0 would, so its selector is the XOR form. The integer pipe, which binds this kernel, is left
// with the four PRMTs, one XOR and the two instructions that extract the code.
__device__ __forceinline__ void onehot_unit(uint32_t c, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3)
{
    const uint32_t cr = c * 0x1111u;
    w0 = prmt(1u, 0u, c * 0x1111u - 0x3210u); w1 = prmt(1u, 0u, c * 0x1111u - 0x7654u);
    w2 = prmt(1u, 0u, c * 0x1111u - 0xba98u); w3 = prmt(1u, 0u, cr ^ 0xfedcu);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- roles of the persistent kernel -------------------------------------------------------------------------------------
//   warps 0..15  workers   : two halves of 8 warps; half h owns the tiles of parity h and the buffers A[h], D[h] in tensor memory.
//                            A thread is one vector (TMEM lane 32 * (warp % 4) + lane) of the tile; the two warps that share a lane
//                            quadrant split the sub-quantizer pairs (expansion) and the group's queries (epilogue) between them.
//                            Per tile: one-hot units into A[h] (tcgen05.st) -> [the MMA warp multiplies; meanwhile the warp writes
//                            out its part of the half's previous tile] -> tcgen05.ld of D[h], certificate + clamp (s16x2 SIMD),
//                            transposed tile in shared memory. While a half waits for its MMA or for tensor memory the other half
//                            works: every scheduler holds four busy warps.
//   warps 16, 17 multiply  : one per half; an elected lane issues the tile's PH tcgen05.mma and commits them to the half's mbarrier
//   warp 18      loads     : fetches the next work item, stages its LUT slab (double-buffered)
// A half reads D[h] and rewrites A[h] in program order, so the MMA warp needs no "empty" barriers: a_full[h] implies both.
constexpr int TC_WORKERS = 16;
constexpr int TC_WARPS = TC_WORKERS + 3;      // + one MMA issuer per half + the loader
constexpr int TC_THREADS2 = 32 * TC_WARPS;
constexpr int TC_QUEUE = 4096;                  // refold queue of one work item

struct TcItem {
    int valid, list, nq, N, t0, t1, n_real, pad;
    uint32_t nk0u, nk1u;                        // the SMALLEST thresholds of the group, negated, in both halves of an s16x2: the hot loop's test
    long long tile0;                            // first tile of the list in the code array
    int q_of[TC_NT];                            // query of group member i (-1: padding column)
    uint2 kq2[TC_NT / 2];                       // NEGATED certificate thresholds of query pairs: .x = (-k0[2i], -k0[2i+1]), .y = (-k1[..]), s16x2
    long long dst[TC_NT];                       // est offset of its segment (an absolute address inside the push exchange)
    unsigned long long cmb[TC_NT];              // base of the chunk minima its segment belongs to (0: none)
};

struct TcShared {
    uint64_t item_full[2], item_empty[2], a_full[2], d_full[2];
    uint32_t tmem_base;
    int n_queue;
    TcItem item[2];
    uint32_t outT[2][2][128 * TC_OUT_STRIDE];   // [half][tile parity, WIDE only]: row = vector, byte n = query n of the group
    uint32_t queue[TC_QUEUE];                   // (tile << 16) | (row << 8) | query column of the pairs whose certificate failed
    uint8_t refold_mark[2][8][TC_NT];           // push exchange: (half, chunk of the tile, query) holds a pair that will be refolded
};

// the reference's fold of one (vector, query): codes from the code array, LUT rows from the slab
template <int PH>
__device__ __noinline__ int tc_refold(const uint32_t *__restrict__ nat32, long long tile, int r, const uint8_t *Bq)
{
    const int s = r >> 4, v = r & 15, gq = v >> 2, sh = 16 * (gq & 1) + 4 * (v & 3);
    const uint32_t *tb = nat32 + ((size_t)tile * PH * 8 + s) * 4 + (gq >> 1);
    int a0 = 0, a1 = 0;
    uint32_t wa[PH], wb[PH];                                        // all 2 PH code words in flight before the first lookup
#pragma unroll
    for (int p = 0; p < PH; p++) { wa[p] = __ldg(tb + (size_t)p * 32); wb[p] = __ldg(tb + (size_t)p * 32 + 2); }
#pragma unroll
    for (int p = 0; p < PH; p++) {
        const uint32_t ca = (wa[p] >> sh) & 15u, cb = (wb[p] >> sh) & 15u;
        const int ta = (int)(int8_t)Bq[(size_t)(2 * p) * TC_NT * 16 + ca];
        const int tb2 = (int)(int8_t)Bq[(size_t)(2 * p + 1) * TC_NT * 16 + cb];
        if (p & 1) a1 = sat_add8<true>(sat_add8<true>(a1, ta), tb2);
        else       a0 = sat_add8<true>(sat_add8<true>(a0, ta), tb2);
    }
    return sat_add8<true>(a0, a1);
}

// signed minimum into one byte of global memory (chunk minima after a refold: the value can only go down)
__device__ __forceinline__ void atomic_min_s8(uint8_t *addr, int v)
{
    unsigned int *word = reinterpret_cast<unsigned int *>(reinterpret_cast<uintptr_t>(addr) & ~(uintptr_t)3);
    const int sh = 8 * (int)(reinterpret_cast<uintptr_t>(addr) & 3);
    unsigned int old = *reinterpret_cast<volatile unsigned int *>(word);
    for (;;) {
        if ((int)(int8_t)((old >> sh) & 0xffu) <= v) return;
        const unsigned int want = (old & ~(0xffu << sh)) | ((unsigned int)(v & 0xff) << sh);
        const unsigned int seen = atomicCAS(word, old, want);
        if (seen == old) return;
        old = seen;
    }
}

// Slow path of the epilogue (kept out of line: the four roles of the kernel run different code at the same time and share one
// instruction cache): among these 8 queries of the warp's 32 vectors some certificate failed. Every lane builds the bit mask of its
// failing queries, one warp scan and ONE shared-memory atomic reserve the queue slots, the lanes write their (tile, vector, query)
// records; when the queue is full the pairs are refolded on the spot and the corrected byte replaces the provisional one in
// (o0, o1). Called by whole warps.
template <int PH>
__device__ __noinline__ void tc_flagged8(TcShared &S, const uint2 *kq2, uint4 pa, uint4 pc, int n0, int nq, int t_rel, int row,
                                         uint32_t &o0, uint32_t &o1, const uint32_t *__restrict__ nat32, long long tile,
                                         const uint8_t *B, uint8_t *mark /* [8][TC_NT] of this half, or null */)
{
    const uint32_t pas[4] = {pa.x, pa.y, pa.z, pa.w}, pcs[4] = {pc.x, pc.y, pc.z, pc.w};
    const int lane = threadIdx.x & 31;
    uint32_t mask = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint2 k = kq2[(n0 >> 1) + u];
        const uint32_t d = __vmaxs2(__vadd2(pas[u], k.x), __vadd2(pcs[u], k.y));        // > 0 in a half: that query's certificate failed (k negated)
        if ((int)(int16_t)(d & 0xffffu) > 0) mask |= 1u << (2 * u);
        if ((int)(int16_t)(d >> 16) > 0) mask |= 2u << (2 * u);
    }
    const int valid = nq - n0;                                                          // padding columns of the group never count
    if (valid < 8) mask &= (1u << (valid < 0 ? 0 : valid)) - 1u;
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(&S.n_queue, total);
    base = __shfl_sync(FULL, base, 31) + incl - cnt;
    while (mask) {
        const int u = __ffs(mask) - 1;
        mask &= mask - 1;
        if (base < TC_QUEUE) {
            S.queue[base] = ((uint32_t)t_rel << 16) | ((uint32_t)row << 8) | (uint32_t)(n0 + u);
            if (mark) mark[(row >> 4) * TC_NT + n0 + u] = 1;
        } else {
            const int e = tc_refold<PH>(nat32, tile, row, B + (size_t)(n0 + u) * 16);
            uint32_t &o = u < 4 ? o0 : o1;
            o = (o & ~(0xffu << (8 * (u & 3)))) | ((uint32_t)(e & 0xff) << (8 * (u & 3)));
        }
        base++;
    }
}

// WIDE: the half writes a tile out together (one barrier per tile), 8 consecutive lanes = the 128 contiguous bytes a tile holds for
// one query -- full lines for the peer-mapped buffers of the push exchange, where 32-byte stores waste NVLink; otherwise every warp
// writes out what it computed itself (32-byte pieces, no barrier), which is faster into local memory.
// CLK: the cycle counters of the roles (TcClock) are kept only in the instance bench.py's stage pass asks for (TKB_TC_CLOCKS=1): they
// cost a dozen registers in a kernel that sits at its register limit.
// DBG (tools/tc_probe.py, TKB_TC_DBG=<bits>): an instance whose phases can be switched off one by one to see what the others cost
// (results are then wrong): 1 no copy-out, 2 copy-out without its global stores, 4 no epilogue arithmetic, 8 no expansion,
// 16 no MMAs, 32 certificate failures ignored.
template <int PH, bool WIDE, bool CLK, bool DBG = false>
__global__ void __launch_bounds__(TC_THREADS2, 1)
ivf_scan_tc_kernel(const uint32_t *__restrict__ nat32, const int64_t *__restrict__ list_chunk_off,
                   const int32_t *__restrict__ list_size, int n_lists, const uint8_t *__restrict__ tables, int P,
                   uint8_t *__restrict__ est, const int64_t *__restrict__ seg_off, uint8_t *__restrict__ cmin,
                   const int64_t *__restrict__ cm_home, int q_per_rank, TcWork W, int dbg, int tpi)
{
    constexpr int M = 2 * PH;
    constexpr int A_COLS = 8 * PH;                                   // 32-bit columns of one one-hot tile
    constexpr int D_COL0 = 2 * A_COLS;                               // accumulator (buffer b, lane l) at D_COL0 + (2 b + l) * TC_NT
    static_assert(2 * A_COLS + 4 * TC_NT <= 512, "tensor memory: two one-hot tiles + two pairs of accumulators");
    extern __shared__ __align__(128) unsigned char tc_smem[];
    TcShared &S = *reinterpret_cast<TcShared *>(tc_smem);
    uint8_t *Bslab = tc_smem + ((sizeof(TcShared) + 127) / 128) * 128;   // two LUT slabs of M * TC_NT units of 16 bytes
    constexpr size_t SLAB = (size_t)M * TC_NT * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 32) {
        for (int b = 0; b < 2; b++) {
            mbar_init(&S.item_full[b], 32);
            mbar_init(&S.item_empty[b], TC_WORKERS + 2);
            mbar_init(&S.a_full[b], TC_WORKERS / 2);
            mbar_init(&S.d_full[b], 1);
        }
        S.n_queue = 0;
        for (int i = 0; i < 2 * 8 * TC_NT; i++) (&S.refold_mark[0][0][0])[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == TC_WARPS - 1) {
        // ================================ loader =====================================================================
        long long lk[2] = {0, 0};
        for (uint32_t it = 0;; it++) {
            const int par = it & 1;
            const long long c0_ = (CLK ? clock64() : 0LL);
            mbar_wait(&S.item_empty[par], ((it >> 1) & 1) ^ 1, 1000);
            const long long c1_ = (CLK ? clock64() : 0LL);
            lk[0] += c1_ - c0_;
            TcItem &I = S.item[par];
            int item = 0;
            if (lane == 0) item = atomicAdd(W.hdr + 1, 1);
            item = __shfl_sync(FULL, item, 0);
            if (item >= W.hdr[0]) {
                if (lane == 0) I.valid = 0;
                __syncwarp();
                mbar_arrive(&S.item_full[par]);
                break;
            }
            int lo = 0, hi = n_lists;                                // largest l with item_off[l] <= item
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (W.item_off[mid] <= item) lo = mid; else hi = mid; }
            const int l = lo;
            const int cnt = W.cnt[l], n_real = (list_size[l] + 15) >> 4, tiles = (n_real + 7) >> 3;
            const int splits = (tiles + tpi - 1) / tpi;
            const int local = item - W.item_off[l], g = local / splits, sp = local - g * splits;
            const int per = (tiles + splits - 1) / splits, t0 = sp * per, t1 = min(tiles, t0 + per);
            const int nq = min(TC_NT, cnt - g * TC_NT);
            const int N = max(16, (nq + 15) & ~15);
            if (lane == 0) {
                I.valid = 1; I.list = l; I.nq = nq; I.N = N; I.t0 = t0; I.t1 = t1; I.n_real = n_real;
                I.tile0 = list_chunk_off[l] >> 3;                    // lists start on a tile
            }
            int kmin0 = 32767, kmin1 = 32767;
            for (int i = lane; i < TC_NT; i += 32) {
                int q = -1;
                long long d = 0;
                int k0 = 0, k1 = 0;
                if (i < nq) {
                    const int e = W.bucket[W.bucket_off[l] + g * TC_NT + i];
                    q = e / P;
                    d = seg_off[e];
                    const TcQueryMeta m = W.qmeta[q];
                    k0 = m.k0; k1 = m.k1;
                    kmin0 = min(kmin0, k0); kmin1 = min(kmin1, k1);
                }
                I.q_of[i] = q; I.dst[i] = d;
                // push exchange: the minima region of the query's home rank, addressed like the estimates (address >> 4)
                I.cmb[i] = q < 0 ? 0ull : (cm_home ? (unsigned long long)cm_home[q / q_per_rank] : (unsigned long long)(uintptr_t)cmin);
                reinterpret_cast<uint16_t *>(I.kq2)[4 * (i >> 1) + (i & 1)] = (uint16_t)(int16_t)(-k0);          // negated: the epilogue adds
                reinterpret_cast<uint16_t *>(I.kq2)[4 * (i >> 1) + 2 + (i & 1)] = (uint16_t)(int16_t)(-k1);
            }
            for (int o = 16; o > 0; o >>= 1) { kmin0 = min(kmin0, __shfl_xor_sync(FULL, kmin0, o)); kmin1 = min(kmin1, __shfl_xor_sync(FULL, kmin1, o)); }
            if (lane == 0) {
                I.nk0u = (uint32_t)((-kmin0) & 0xffff) * 0x00010001u;
                I.nk1u = (uint32_t)((-kmin1) & 0xffff) * 0x00010001u;
            }
            __syncwarp();
            uint8_t *B = Bslab + (size_t)par * SLAB;
            for (int u0 = 0; u0 < M * N; u0 += 32 * 8) {             // unit (j, i): LUT row j of group member i; 8 loads in flight
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int u = u0 + 32 * k + lane;
                    v[k] = make_uint4(0, 0, 0, 0);
                    if (u < M * N) {
                        const int q = I.q_of[u % N];
                        if (q >= 0) v[k] = __ldg(reinterpret_cast<const uint4 *>(tables + ((size_t)q * M + u / N) * 16));
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int u = u0 + 32 * k + lane;
                    if (u < M * N) *reinterpret_cast<uint4 *>(B + ((size_t)(u / N) * TC_NT + u % N) * 16) = v[k];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            mbar_arrive(&S.item_full[par]);
            lk[1] += (CLK ? clock64() : 0LL) - c1_;
        }
        if (CLK && lane == 0) { atomicAdd(W.clk + CK_L_WAIT, (unsigned long long)lk[0]); atomicAdd(W.clk + CK_L_STAGE, (unsigned long long)lk[1]); }
    } else if (warp >= TC_WORKERS) {
        // ================================ MMA issuers ================================================================
        // One per half. A single issuer took the tiles in order, so a half whose tile was expanded first still waited for the other
        // half's a_full (measured: ~400 cycles per tile between the end of the copy-out and d_full); with its own issuer a half's
        // MMAs are issued the moment its tile is expanded, and the tensor pipe interleaves the two streams.
        const int h = warp - TC_WORKERS;
        uint32_t g = 0, k_tile = 0;                                  // tiles seen by the kernel / tiles of this half
        long long mk[3] = {0, 0, 0};
        const long long k0_ = (CLK ? clock64() : 0LL);
        for (uint32_t it = 0;; it++) {
            const int par = it & 1;
            mbar_wait(&S.item_full[par], (it >> 1) & 1);
            const TcItem &I = S.item[par];
            if (!I.valid) break;
            {                                                         // the whole warp runs the loop, one elected lane issues
                const uint32_t idesc = umma_idesc_i8(I.N);
                const uint32_t b0 = smem_u32(Bslab + (size_t)par * SLAB);
                const int first = I.t0 + (int)((h - g) & 1);
                for (int t = first; t < I.t1; t += 2, k_tile++) {
                    const uint32_t b = (uint32_t)h, ph = k_tile & 1;
                    const long long c0_ = (CLK ? clock64() : 0LL);
                    // A[b] written AND D[b] read by the half (program order of its warps). The issuer waits most of the time in a
                    // kernel whose workers are issue-bound: it sleeps between polls instead of taking their slots
                    mbar_wait(&S.a_full[b], ph, 32);
                    tc_fence_after();
                    const long long c1_ = (CLK ? clock64() : 0LL), c2_ = c1_;
                    if (elect_one()) {
                        if (!(DBG && (dbg & 16)))
#pragma unroll
                        for (int p = 0; p < PH; p++)
                            umma_i8_ts(tmem + D_COL0 + (2 * b + (p & 1)) * TC_NT, tmem + b * A_COLS + 8 * p,
                                       umma_desc_kmajor(b0 + (uint32_t)(2 * p) * TC_NT * 16, TC_NT * 16, 128), idesc, p >= 2 ? 1u : 0u);
                        umma_commit(&S.d_full[b]);
                    }
                    __syncwarp();
                    mk[0] += c1_ - c0_; mk[1] += c2_ - c1_; mk[2] += (CLK ? clock64() : 0LL) - c2_;
                }
                g += (uint32_t)(I.t1 - I.t0);
            }
            if (lane == 0) {
                if (h == 0) {
                    atomicAdd(W.hdr + 3, I.t1 - I.t0);
                    atomicAdd(W.hdr + 4, (I.t1 - I.t0) * (I.N >> 4));
                }
                mbar_arrive(&S.item_empty[par]);
            }
        }
        if (CLK && lane == 0 && h == 0) {                            // (half 0's issuer: its waits and issues are counted per tile of the kernel)
            atomicAdd(W.clk + CK_M_WAIT_A, (unsigned long long)(2 * mk[0])); atomicAdd(W.clk + CK_M_WAIT_D, (unsigned long long)(2 * mk[1]));
            atomicAdd(W.clk + CK_M_ISSUE, (unsigned long long)(2 * mk[2])); atomicAdd(W.clk + CK_TOTAL, (unsigned long long)((CLK ? clock64() : 0LL) - k0_));
        }
    } else {
        // ================================ workers ====================================================================
        const int h = warp >> 3, sub = (warp >> 2) & 1;              // half (tile parity, buffers), which of the two warps of a lane quadrant
        const int row = 32 * (warp & 3) + lane;
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        const int s = row >> 4, v = row & 15, gq = v >> 2, sh = 16 * (gq & 1) + 4 * (v & 3);
        constexpr int PW = PH / 2;                                   // sub-quantizer pairs per warp
        uint32_t g = 0, k_tile = 0;                                  // tiles seen by the kernel / tiles of this half
        long long ck[5] = {0, 0, 0, 0, 0};
        for (uint32_t it = 0;; it++) {
            const int par = it & 1;
            mbar_wait(&S.item_full[par], (it >> 1) & 1, 100);
            const TcItem &I = S.item[par];
            if (!I.valid) break;
            const int t0 = I.t0, t1 = I.t1, N = I.N, nq = I.nq, n_real = I.n_real, nh = N >> 1;
            const uint32_t nk0u = I.nk0u, nk1u = I.nk1u;
            const long long tile0 = I.tile0;
            const uint8_t *B = Bslab + (size_t)par * SLAB;
            const int first = t0 + (int)((h - g) & 1);               // this half's first tile of the item: (g + first - t0) & 1 == h
            const uint32_t *tb = nat32 + ((size_t)tile0 * PH * 8 + s) * 4 + (gq >> 1) + (size_t)sub * PW * 32;
            const int htid = tid - 256 * h;                          // 0..255 inside the half
            // The estimates a warp computed (its 32 vectors x its half of the queries) lie transposed in shared memory; the same warp
            // writes them out: lane = (query pair, one of the warp's two chunks): 16 words in, 2 x 16 bytes out. No barrier: a warp
            // reads back only what it wrote itself.
            auto copy_out = [&](int t, const uint32_t *outT) {
                if (WIDE) {
                    // the whole half, after its barrier: a task = 2 queries x one chunk, the chunk fastest: 8 lanes write 128 contiguous bytes
                    for (int task = htid; task < ((nq + 1) >> 1) * 8; task += 256) {
                        const int n = 2 * (task >> 3), sc = task & 7;
                        const int chunk = t * 8 + sc;
                        if (chunk >= n_real) continue;
                        uint32_t w[16];
#pragma unroll
                        for (int kk = 0; kk < 16; kk++) w[kk] = outT[tc_out_addr(16 * sc + kk, n >> 2)];
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            if (n + j >= nq) break;
                            const uint32_t sel = 0x0040u + 0x0011u * (uint32_t)((n & 2) + j);
                            uint32_t o[4];
#pragma unroll
                            for (int q4 = 0; q4 < 4; q4++) {
                                const uint32_t lo = prmt(w[4 * q4], w[4 * q4 + 1], sel), hi = prmt(w[4 * q4 + 2], w[4 * q4 + 3], sel);
                                o[q4] = prmt(lo, hi, 0x5410u);
                            }
                            const long long off = I.dst[n + j] + 16LL * chunk;
                            *reinterpret_cast<uint4 *>(est + off) = make_uint4(o[0], o[1], o[2], o[3]);
                            uint8_t *cm = reinterpret_cast<uint8_t *>((uintptr_t)I.cmb[n + j]);
                            if (cm) {
                                uint32_t m = __vmins4(__vmins4(o[0], o[1]), __vmins4(o[2], o[3]));
                                m = __vmins4(m, m >> 16);
                                m = __vmins4(m, m >> 8);
                                // the minima live in another GPU's memory, where the atomic minimum of the refold pass would be a round
                                // trip over NVLink per pair: a chunk with a pair still to be refolded gets the smallest possible minimum
                                // instead, which only makes the replay look at its 16 estimates (exact by then)
                                if (cm_home && S.refold_mark[h][sc][n + j]) { m = 0x80u; S.refold_mark[h][sc][n + j] = 0; }
                                cm[off >> 4] = (uint8_t)(m & 0xffu);
                            }
                        }
                    }
                    return;
                }
                __syncwarp();
                if (lane < nh) {                                      // lane = (query pair, one of the warp's two chunks)
                    const int n = sub * nh + 2 * (lane >> 1), sc = 2 * (warp & 3) + (lane & 1);
                    const int chunk = t * 8 + sc;
                    if (chunk < n_real && n < nq) {
                        uint32_t w[16];
#pragma unroll
                        for (int kk = 0; kk < 16; kk++) w[kk] = outT[tc_out_addr(16 * sc + kk, n >> 2)];
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            if (n + j >= nq) break;
                            const uint32_t sel = 0x0040u + 0x0011u * (uint32_t)((n & 2) + j);   // (w0.byte b, w1.byte b), b = (n + j) % 4
                            uint32_t o[4];
#pragma unroll
                            for (int q4 = 0; q4 < 4; q4++) {          // that byte of four consecutive rows
                                const uint32_t lo = prmt(w[4 * q4], w[4 * q4 + 1], sel), hi = prmt(w[4 * q4 + 2], w[4 * q4 + 3], sel);
                                o[q4] = prmt(lo, hi, 0x5410u);
                            }
                            const long long off = I.dst[n + j] + 16LL * chunk;
                            if (DBG && (dbg & 2)) { if ((o[0] ^ o[1] ^ o[2] ^ o[3]) == 0x12345678u) est[off] = 1; continue; }
                            *reinterpret_cast<uint4 *>(est + off) = make_uint4(o[0], o[1], o[2], o[3]);
                            uint8_t *cm = reinterpret_cast<uint8_t *>((uintptr_t)I.cmb[n + j]);
                            if (cm) {
                                uint32_t m = __vmins4(__vmins4(o[0], o[1]), __vmins4(o[2], o[3]));
                                m = __vmins4(m, m >> 16);
                                m = __vmins4(m, m >> 8);
                                cm[off >> 4] = (uint8_t)(m & 0xffu);
                            }
                        }
                    }
                }
                __syncwarp();
            };
            // the code words of the half's NEXT tile are fetched while the current one is multiplied and written out: the expansion
            // never waits for global memory except at the first tile of an item
            uint32_t nxt[2 * PW];
            auto fetch = [&](int t) {
#pragma unroll
                for (int p = 0; p < PW; p++) {
                    nxt[2 * p] = __ldg(tb + ((size_t)t * PH + p) * 32);
                    nxt[2 * p + 1] = __ldg(tb + ((size_t)t * PH + p) * 32 + 2);
                }
            };
            if (first < t1) fetch(first);
            for (int t = first; t < t1; t += 2, k_tile++) {
                const uint32_t ph = (k_tile & 1);                    // the half's buffers complete one phase per own tile
                const long long c0_ = (CLK ? clock64() : 0LL);
                // ---- expand: this thread's vector, this warp's half of the sub-quantizer pairs ------------------------------
                {
                    uint32_t cur[2 * PW];
#pragma unroll
                    for (int p = 0; p < 2 * PW; p++) cur[p] = nxt[p];
                    if (!(DBG && (dbg & 8)))
#pragma unroll
                    for (int p = 0; p < PW; p++) {
                        uint32_t r[8];
                        onehot_unit((cur[2 * p] >> sh) & 15u, r[0], r[1], r[2], r[3]);
                        onehot_unit((cur[2 * p + 1] >> sh) & 15u, r[4], r[5], r[6], r[7]);
                        tmem_st8(tmem + lane_base + h * A_COLS + 8 * (sub * PW + p), r);
                    }
                    if (t + 2 < t1) fetch(t + 2);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.a_full[h]);
                }
                const long long c1_ = (CLK ? clock64() : 0LL);
                // ---- while the MMA warp multiplies: write out (this warp's part of) the half's previous tile --------------------------
                if (WIDE) asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory");   // everybody's epilogue of that tile is in outT
                if (t != first && !(DBG && (dbg & 1))) copy_out(t - 2, S.outT[h][WIDE ? (k_tile - 1) & 1 : 0]);
                const long long c2_ = (CLK ? clock64() : 0LL);
                mbar_wait(&S.d_full[h], ph);
                tc_fence_after();
                const long long c3_ = (CLK ? clock64() : 0LL);
                // ---- epilogue: this warp's half of the group's queries, 8 per step, loads one step ahead --------------------
                if (!(DBG && (dbg & 4))) {
                    const int nb = sub * nh;
                    uint32_t *outT = S.outT[h][WIDE ? k_tile & 1 : 0];
                    uint32_t la[4], lc[4];                            // lane sums of two queries per register (s16x2: |S| <= 16 * 128)
                    tmem_ld8_pack16(tmem + lane_base + D_COL0 + (2 * h) * TC_NT + nb, la);
                    tmem_ld8_pack16(tmem + lane_base + D_COL0 + (2 * h + 1) * TC_NT + nb, lc);
#pragma unroll 1
                    for (int n0 = nb; n0 < nb + nh; n0 += 8) {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        uint32_t pa[4], pc[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) { pa[u] = la[u]; pc[u] = lc[u]; }
                        if (n0 + 8 < nb + nh) {
                            tmem_ld8_pack16(tmem + lane_base + D_COL0 + (2 * h) * TC_NT + n0 + 8, la);
                            tmem_ld8_pack16(tmem + lane_base + D_COL0 + (2 * h + 1) * TC_NT + n0 + 8, lc);
                        }
                        uint32_t e2[4], fl = 0x80008000u;
#pragma unroll
                        for (int u = 0; u < 4; u++) {                 // two queries per step. The test uses the group's smallest thresholds
                            // (registers, no shared-memory load per pair): it can only flag more; tc_flagged8 applies each query's own
                            e2[u] = __vimin3_s16x2(__viaddmax_s16x2(pa[u], pc[u], 0xff80ff80u), 0x007f007fu, 0x007f007fu);
                            fl = __viaddmax_s16x2(pa[u], nk0u, fl);
                            fl = __viaddmax_s16x2(pc[u], nk1u, fl);
                        }
                        uint32_t o0 = prmt(e2[0], e2[1], 0x6420u), o1 = prmt(e2[2], e2[3], 0x6420u);
                        if (!(DBG && (dbg & 32)) && __any_sync(FULL, (int)(int16_t)(fl & 0xffffu) > 0 || (int)(int16_t)(fl >> 16) > 0))
                            tc_flagged8<PH>(S, I.kq2, make_uint4(pa[0], pa[1], pa[2], pa[3]), make_uint4(pc[0], pc[1], pc[2], pc[3]), n0, nq,
                                            t - t0, row, o0, o1, nat32, tile0 + t, B, cm_home ? &S.refold_mark[h][0][0] : nullptr);
                        outT[tc_out_addr(row, n0 >> 2)] = o0;
                        outT[tc_out_addr(row, (n0 >> 2) + 1)] = o1;
                    }
                    tc_fence_before();                                // (orders the TMEM reads before the next tile's a_full arrive)
                }
                if (DBG && (dbg & 4)) tc_fence_before();
                ck[0] += c1_ - c0_; ck[4] += c2_ - c1_; ck[1] += c3_ - c2_; ck[2] += (CLK ? clock64() : 0LL) - c3_;
            }
            if (first < t1) {                                         // the half's last tile of the item
                if (WIDE) asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory");
                if (!(DBG && (dbg & 1))) copy_out(first + 2 * ((t1 - 1 - first) >> 1), S.outT[h][WIDE ? (k_tile - 1) & 1 : 0]);
            }
            g += (uint32_t)(t1 - t0);
            // the item's refold queue (both halves together): the reference's recurrence for the pairs whose certificate failed,
            // bytes patched in place (the refolded value is never above the provisional one: a chunk minimum can only go down)
            const long long f0_ = (CLK ? clock64() : 0LL);
            asm volatile("bar.sync 3, %0;" ::"n"(32 * TC_WORKERS) : "memory");
            {
                const int nqd = min(S.n_queue, TC_QUEUE);
                for (int i = tid; i < nqd; i += 32 * TC_WORKERS) {
                    const uint32_t en = S.queue[i];
                    const int tt = t0 + (int)(en >> 16), r = (int)((en >> 8) & 0xffu), n = (int)(en & 0xffu);
                    if (tt * 8 + (r >> 4) >= n_real) continue;       // a padding vector of the last tile: never written
                    const int e = tc_refold<PH>(nat32, tile0 + tt, r, B + (size_t)n * 16);
                    const long long off = I.dst[n] + 128LL * tt + r;
                    est[off] = (uint8_t)e;
                    uint8_t *cm = reinterpret_cast<uint8_t *>((uintptr_t)I.cmb[n]);
                    if (cm && !cm_home) atomic_min_s8(cm + (off >> 4), e);
                }
                asm volatile("bar.sync 3, %0;" ::"n"(32 * TC_WORKERS) : "memory");
                if (tid == 0) { if (S.n_queue) atomicAdd(W.hdr + 2, S.n_queue); S.n_queue = 0; }
                asm volatile("bar.sync 3, %0;" ::"n"(32 * TC_WORKERS) : "memory");
            }
            if (CLK && warp == 0 && lane == 0) atomicAdd(W.clk + CK_P_FLUSH, (unsigned long long)((CLK ? clock64() : 0LL) - f0_));
            if (lane == 0) mbar_arrive(&S.item_empty[par]);
        }
        if (CLK && warp == 0 && lane == 0) {
            atomicAdd(W.clk + CK_E_WORK, (unsigned long long)ck[0]); atomicAdd(W.clk + CK_P_WAIT, (unsigned long long)ck[1]);
            atomicAdd(W.clk + CK_P_EPI, (unsigned long long)ck[2]); atomicAdd(W.clk + CK_P_BAR, (unsigned long long)ck[3]);
            atomicAdd(W.clk + CK_P_COPY, (unsigned long long)ck[4]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

#endif  // TKB_EMULATE

// ---- host side -------------------------------------------------------------------------------------------------------
// 1 when this build can run the tensor-core scan on the current device (an sm_100 GPU; never on the CPU emulator)
int tc_supported()
{
#ifdef TKB_EMULATE
    return 0;
#else
    static int ok = -1;
    if (ok < 0) {
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
        ok = major == 10 ? 1 : 0;
    }
    return ok;
#endif
}

static int n_sm_hint()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int tc_workspace_bytes(int Q, int P, int n_lists, int64_t *bytes)
{
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists >= 0 && bytes, "bad extent");
    *bytes = (int64_t)tc_carve(nullptr, Q, P, n_lists, nullptr) + 256;
    return TKB_OK;
}

int launch_ivf_scan_tc(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                       const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est, const int64_t *seg_off,
                       uint8_t *cmin, const int64_t *cm_home, int q_per_rank, int64_t max_chunks_per_query, void *workspace,
                       int64_t workspace_bytes, cudaStream_t st)
{
#ifdef TKB_EMULATE
    return set_err(TKB_ERR_INVALID, "the tensor-core scan needs an sm_100a device (not available on the CPU emulator)");
#else
    TKB_REQUIRE(M == 32, "the tensor-core scan is built for M = 32 sub-quantizers (d = 128 rotated to 64)");
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || P == 0) return TKB_OK;
    // est == NULL: the plan holds absolute addresses (push exchange: segments land in the home ranks' peer-mapped buffers)
    TKB_REQUIRE(native && list_chunk_off && list_size && tables && probes && seg_off && workspace, "null pointer");
    TKB_REQUIRE(!cm_home || (!est && !cmin && q_per_rank > 0), "per-home minima tables belong to the push exchange (est == NULL)");
    TKB_REQUIRE((int64_t)Q * P <= 0x7fffffffLL, "too many (query, probe) units for one launch");
    TKB_REQUIRE((uintptr_t)workspace % 16 == 0 && (uintptr_t)est % 16 == 0 && (uintptr_t)tables % 16 == 0, "pointers must be 16-byte aligned");
    int64_t need = 0;
    tc_workspace_bytes(Q, P, n_lists, &need);
    TKB_REQUIRE(workspace_bytes >= need, "workspace too small (tkb_ivf_scan_tc_workspace)");
    TcWork W;
    tc_carve(workspace, Q, P, n_lists, &W);
    const int64_t QP = (int64_t)Q * P;
    // hdr, cnt, cursor are contiguous at the start of the workspace
    TKB_CUDA(cudaMemsetAsync(workspace, 0, (size_t)((unsigned char *)W.bucket_off - (unsigned char *)workspace), st));
    tc_query_meta_kernel<<<(Q + 7) / 8, 256, 0, st>>>(tables, Q, M, W.qmeta, W.skip_q);
    TKB_LAUNCH_CHECK();
    tc_count_kernel<<<(unsigned)((QP + 255) / 256), 256, 0, st>>>(probes, seg_off, QP, P, n_lists, W.qmeta, W.cnt, W.seg_rest);
    TKB_LAUNCH_CHECK();
    // Tiles per work item: 128 keeps the per-item costs (slab staging, pipeline fill) small on an index of thousands of lists.
    // ONE list that every query probes (a brute-force scan; the encoded centroids of probe selection) has only Q / 64 groups
    // x tiles / 128 items: smaller items there, about six per SM, so that the last round of items does not idle half the GPU.
    int tpi = TC_TILES_PER_ITEM;
    if (n_lists == 1 && P == 1) {
        const int64_t tiles = (max_chunks_per_query + 7) / 8, groups = (Q + TC_NT - 1) / TC_NT;
        int64_t want = (tiles * groups + 6LL * n_sm_hint() - 1) / (6LL * n_sm_hint());
        tpi = (int)(want < 8 ? 8 : (want > TC_TILES_PER_ITEM ? TC_TILES_PER_ITEM : want));
    }
    tc_offsets_kernel<<<1, 1024, 0, st>>>(W.cnt, list_size, n_lists, W.bucket_off, W.item_off, W.hdr, tpi);
    TKB_LAUNCH_CHECK();
    tc_fill_kernel<<<(unsigned)((QP + 255) / 256), 256, 0, st>>>(probes, W.seg_rest, QP, n_lists, W.bucket_off, W.cursor, W.bucket);
    TKB_LAUNCH_CHECK();
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        TKB_CUDA(cudaGetDevice(&dev));
        TKB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    // one CTA per SM (the kernel owns all 512 TMEM columns): the slab + the rest of the shared memory is padded past half an SM's
    const size_t smem = ((sizeof(TcShared) + 127) / 128) * 128 + 2 * (size_t)M * TC_NT * 16 + 128;
    const size_t smem_req = smem > 120 * 1024 ? smem : 120 * 1024;
    const char *ce = getenv("TKB_TC_CLOCKS");
    const bool clk = ce && ce[0] == '1';
    const char *de = getenv("TKB_TC_DBG");
    const int dbg = de ? atoi(de) : 0;
#define TKB_TC_LAUNCH(WIDE_, CLK_)                                                                                                     \
    do {                                                                                                                               \
        TKB_CUDA(cudaFuncSetAttribute(ivf_scan_tc_kernel<16, WIDE_, CLK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req)); \
        ivf_scan_tc_kernel<16, WIDE_, CLK_><<<n_sm, TC_THREADS2, smem_req, st>>>(reinterpret_cast<const uint32_t *>(native),          \
            list_chunk_off, list_size, n_lists, tables, P, est, seg_off, cmin, cm_home, q_per_rank, W, 0, tpi);                        \
    } while (0)
    if (dbg && est != nullptr) {                  // probe instance (wrong results by design)
        TKB_CUDA(cudaFuncSetAttribute(ivf_scan_tc_kernel<16, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
        ivf_scan_tc_kernel<16, false, false, true><<<n_sm, TC_THREADS2, smem_req, st>>>(reinterpret_cast<const uint32_t *>(native),
            list_chunk_off, list_size, n_lists, tables, P, est, seg_off, cmin, cm_home, q_per_rank, W, dbg, tpi);
    } else if (est == nullptr) {                         // push exchange: full 128-byte lines into the peer-mapped buffers
        if (clk) TKB_TC_LAUNCH(true, true); else TKB_TC_LAUNCH(true, false);
    } else {
        if (clk) TKB_TC_LAUNCH(false, true); else TKB_TC_LAUNCH(false, false);
    }
#undef TKB_TC_LAUNCH
    TKB_LAUNCH_CHECK();
    // the (query, list) pairs the tensor-core path does not take (queries whose LUT fails the per-query precondition):
    // the CUDA-core kernel, which skips every query marked in skip_q
    return launch_ivf_scan_native(native, list_chunk_off, list_size, n_lists, M, tables, probes, Q, P, est, 0, W.seg_rest,
                                  max_chunks_per_query, TKB_ORDER_AVX, 1, nullptr, 0, st, cmin, cm_home, q_per_rank, W.skip_q);
#endif
}

}  // namespace tkb
