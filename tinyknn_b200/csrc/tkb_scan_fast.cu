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
#include "tkb_common.cuh"

namespace tkb {

constexpr int FAST_THREADS = 256;
constexpr int TILE = 8;                        // chunks per tile

__device__ __forceinline__ size_t native_off(int64_t chunk, int p, int Ph)     // in uint4 units
{
    return ((size_t)(chunk >> 3) * Ph + p) * TILE + (chunk & 7);
}

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
// per-query LUT preparation (CTA prologue)
// ------------------------------------------------------------------------------------------------
struct LutMeta {
    int eligible;          // fast path allowed for this query
    int bias_tot;          // bias_0 + bias_1                         (signed fast path)
    int k0, k1;            // certificate thresholds on the biased lane sums: S'_l <= k_l
};

// smem layout: uint4 rows[M] (biased when eligible, raw otherwise) | raw copy uint4 raw[M] | LutMeta | scratch
template <bool SIGNED>
__device__ void prepare_lut(const uint8_t *__restrict__ tq, int M, bool fast_allowed, uint4 *rows, uint4 *raw,
                            LutMeta *meta, int *scratch /* 4*M ints */)
{
    const int tid = threadIdx.x;
    for (int j = tid; j < M; j += blockDim.x) {
        const uint4 r = reinterpret_cast<const uint4 *>(tq)[j];
        raw[j] = r;
        const uint32_t ws[4] = {r.x, r.y, r.z, r.w};
        int mn = 1 << 30, mx = -(1 << 30);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t b = (ws[c >> 2] >> (8 * (c & 3))) & 0xffu;
            const int t = SIGNED ? (int)(int8_t)b : (int)b;
            mn = min(mn, t); mx = max(mx, t);
        }
        const int bias = SIGNED ? -mn : 0;                 // unsigned rows are used as they are
        scratch[4 * j + 0] = bias;
        scratch[4 * j + 1] = SIGNED ? max(0, -mn) : 0;     // contribution to N
        scratch[4 * j + 2] = mx + bias;                    // largest biased entry
        uint32_t o[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t v = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t b = (ws[w] >> (8 * c)) & 0xffu;
                const int t = SIGNED ? (int)(int8_t)b : (int)b;
                v |= (uint32_t)((t + bias) & 0xff) << (8 * c);
            }
            o[w] = v;
        }
        rows[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    if (tid == 0) {
        int bias[2] = {0, 0}, N[2] = {0, 0}, range = 0;
        for (int j = 0; j < M; j++) {
            const int l = (j >> 1) & 1;
            bias[l] += scratch[4 * j + 0];
            N[l] += scratch[4 * j + 1];
            range = max(range, scratch[4 * j + 2]);
        }
        LutMeta m;
        // 8 steps of one lane accumulate in a byte: 8 * range <= 255; PRMT zero trick needs entries < 128
        m.eligible = fast_allowed && range <= 31 && (!SIGNED || (N[0] <= 128 && N[1] <= 128));
        m.bias_tot = bias[0] + bias[1];
        m.k0 = 127 - N[0] + bias[0];
        m.k1 = 127 - N[1] + bias[1];
        *meta = m;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// the fast chunk kernel: 16 vectors, M sub-quantizers, AVX lane split (pair p -> lane p & 1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

// one sub-quantizer (LUT row L = 16 biased bytes), two code words = 4 groups of 4 vectors
__device__ __forceinline__ void lookup_step(const uint4 L, uint32_t wa, uint32_t wb, uint32_t (&acc)[4])
{
    const uint32_t xa = wa ^ 0x88888888u, xb = wb ^ 0x88888888u;
    // entries 0..7 live in (L.x, L.y), entries 8..15 in (L.z, L.w). A selector nibble with bit 3 set
    // makes PRMT return the replicated sign bit of the addressed byte, i.e. 0 for our entries < 128.
    acc[0] += prmt(L.x, L.y, wa) + prmt(L.z, L.w, xa);
    acc[1] += prmt(L.x, L.y, wa >> 16) + prmt(L.z, L.w, xa >> 16);
    acc[2] += prmt(L.x, L.y, wb) + prmt(L.z, L.w, xb);
    acc[3] += prmt(L.x, L.y, wb >> 16) + prmt(L.z, L.w, xb >> 16);
}

// Returns the 16 estimates (one byte per vector, vector order) and whether any vector failed the
// certificate (then the caller queues the chunk for the patch kernel).
template <bool SIGNED>
__device__ __forceinline__ uint4 scan_chunk_fast(const uint4 *__restrict__ nat, int64_t chunk, int Ph,
                                                 const uint4 *__restrict__ rows, const LutMeta &meta,
                                                 bool &flagged)
{
    uint32_t wide[2][4][2];                    // [lane][group][even/odd] packed s16x2 biased sums
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
        for (int g = 0; g < 4; g++) { wide[l][g][0] = 0; wide[l][g][1] = 0; }

    const uint4 *base = nat + native_off(chunk, 0, Ph);
    for (int p0 = 0; p0 < Ph; p0 += 8) {
        uint4 w[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (p0 + i < Ph) w[i] = ldg_nc_u4(base + (size_t)(p0 + i) * TILE);
        uint32_t acc[2][4];
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
            for (int g = 0; g < 4; g++) acc[l][g] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (p0 + i < Ph) {
                const int j = 2 * (p0 + i);
                lookup_step(rows[j], w[i].x, w[i].y, acc[i & 1]);          // sub-quantizer 2p
                lookup_step(rows[j + 1], w[i].z, w[i].w, acc[i & 1]);      // sub-quantizer 2p+1
            }
        }
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
            for (int g = 0; g < 4; g++) {
                wide[l][g][0] += prmt(acc[l][g], 0u, 0x4240u);             // vectors 4g+0, 4g+2
                wide[l][g][1] += prmt(acc[l][g], 0u, 0x4341u);             // vectors 4g+1, 4g+3
            }
    }

    uint32_t outw[4];
    uint32_t flag = 0x80008000u;               // running max of (S'_l - k_l), starts at the most negative s16x2
    if (SIGNED) {
        const uint32_t nbias = (uint32_t)((-meta.bias_tot) & 0xffff) * 0x00010001u;
        const uint32_t nk0 = (uint32_t)((-meta.k0) & 0xffff) * 0x00010001u;
        const uint32_t nk1 = (uint32_t)((-meta.k1) & 0xffff) * 0x00010001u;
        const uint32_t lo128 = 0xff80ff80u, hi127 = 0x007f007fu;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t e[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t s = __vadd2(wide[0][g][h], wide[1][g][h]);                  // no overflow: < 2^12
                e[h] = __vimin3_s16x2(__viaddmax_s16x2(s, nbias, lo128), hi127, hi127);   // clamp(S0+S1, -128, 127)
                const uint32_t d0 = __vadd2(wide[0][g][h], nk0);
                flag = __vimax3_s16x2(flag, d0, __vadd2(wide[1][g][h], nk1));
            }
            outw[g] = prmt(e[0], e[1], 0x6240u);                                          // bytes v0 v1 v2 v3
        }
        flagged = ((int)(int16_t)(flag & 0xffffu) > 0) || ((int)(int16_t)(flag >> 16) > 0);
    } else {
        const uint32_t hi255 = 0x00ff00ffu;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t e[2];
#pragma unroll
            for (int h = 0; h < 2; h++)
                e[h] = __vimin3_u16x2(__vadd2(wide[0][g][h], wide[1][g][h]), hi255, hi255);   // min(255, sum)
            outw[g] = prmt(e[0], e[1], 0x6240u);
        }
        flagged = false;
    }
    return make_uint4(outw[0], outw[1], outw[2], outw[3]);
}

// step-by-step saturating fold on the native layout (ineligible LUTs, signed SSE order, patch kernel)
template <int ORDER, bool SIGNED>
__device__ __forceinline__ int exact_vector(const uint4 *__restrict__ nat, int64_t chunk, int Ph, int v,
                                            const uint8_t *__restrict__ raw /* M*16 bytes */)
{
    const int g = v >> 2, sh = 4 * (v & 3) + 16 * (g & 1);
    int a0 = 0, a1 = 0;
    for (int p = 0; p < Ph; p++) {
        const uint4 w = nat[native_off(chunk, p, Ph)];
        const uint32_t c0 = (((g < 2) ? w.x : w.y) >> sh) & 15u;
        const uint32_t c1 = (((g < 2) ? w.z : w.w) >> sh) & 15u;
        int t0 = raw[32 * p + c0], t1 = raw[32 * p + 16 + c1];
        if (SIGNED) { t0 = (int)(int8_t)t0; t1 = (int)(int8_t)t1; }
        if (ORDER == TKB_ORDER_AVX && (p & 1)) a1 = sat_add8<SIGNED>(sat_add8<SIGNED>(a1, t0), t1);
        else                                  a0 = sat_add8<SIGNED>(sat_add8<SIGNED>(a0, t0), t1);
    }
    return (ORDER == TKB_ORDER_AVX) ? sat_add8<SIGNED>(a0, a1) : a0;
}

template <int ORDER, bool SIGNED>
__device__ __forceinline__ uint4 scan_chunk_exact(const uint4 *__restrict__ nat, int64_t chunk, int Ph,
                                                  const uint8_t *__restrict__ raw)
{
    uint32_t o[4] = {0, 0, 0, 0};
    for (int v = 0; v < 16; v++) {
        const int e = exact_vector<ORDER, SIGNED>(nat, chunk, Ph, v, raw);
        o[v >> 2] |= (uint32_t)(e & 0xff) << (8 * (v & 3));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// cold path (patch list full): same fold, kept out of line so that it does not cost the hot loop registers
template <int ORDER, bool SIGNED>
__device__ __noinline__ uint4 scan_chunk_exact_cold(const uint4 *__restrict__ nat, int64_t chunk, int Ph,
                                                    const uint8_t *__restrict__ raw)
{
    return scan_chunk_exact<ORDER, SIGNED>(nat, chunk, Ph, raw);
}

struct PatchList {
    unsigned long long *count;     // number of flagged chunks (may exceed `cap`: the excess was recomputed inline)
    uint2 *entry;                  // .x = unit (query, or query * P + probe slot), .y = chunk inside the unit's segment
    unsigned long long cap;        // entries that fit
};

// queue a flagged chunk for the patch pass; false = list full, the caller recomputes the chunk itself
__device__ __forceinline__ bool patch_push(const PatchList &pl, uint32_t unit, uint32_t local)
{
    const unsigned long long i = atomicAdd(pl.count, 1ULL);
    if (i >= pl.cap) return false;
    pl.entry[i] = make_uint2(unit, local);
    return true;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <int ORDER, bool SIGNED>
__global__ void __launch_bounds__(FAST_THREADS, 3)
estimate_fast_kernel(const uint4 *__restrict__ nat, int64_t n_chunks, int M, const uint8_t *__restrict__ tables,
                     uint8_t *__restrict__ est, int64_t est_stride, PatchList patch)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint4 *rows = reinterpret_cast<uint4 *>(smem);
    uint4 *raw = rows + M;
    LutMeta *meta = reinterpret_cast<LutMeta *>(raw + M);
    int *scratch = reinterpret_cast<int *>(meta + 1);
    const int q = blockIdx.y, Ph = M >> 1;
    const bool fast_allowed = !(SIGNED && ORDER == TKB_ORDER_SSE);
    prepare_lut<SIGNED>(tables + (size_t)q * M * 16, M, fast_allowed, rows, raw, meta, scratch);
    const LutMeta m = *meta;
    for (int64_t c = (int64_t)blockIdx.x * FAST_THREADS + threadIdx.x; c < n_chunks;
         c += (int64_t)gridDim.x * FAST_THREADS) {
        const int64_t off = (int64_t)q * est_stride + 16 * c;
        uint4 o;
        if (m.eligible) {
            bool flagged;
            o = scan_chunk_fast<SIGNED>(nat, c, Ph, rows, m, flagged);
            if (flagged && !patch_push(patch, (uint32_t)q, (uint32_t)c))
                o = scan_chunk_exact_cold<ORDER, SIGNED>(nat, c, Ph, reinterpret_cast<const uint8_t *>(raw));
        } else {
            o = scan_chunk_exact<ORDER, SIGNED>(nat, c, Ph, reinterpret_cast<const uint8_t *>(raw));
        }
        *reinterpret_cast<uint4 *>(est + off) = o;
    }
}

// One CTA column per query: the P probed lists are walked as one flat range of chunks, so the LUT is
// prepared once per (query, split) and short lists do not leave lanes idle. Segment (q, s) is written at
// est + seg_off[q*P+s] (absent when negative) or, without a plan, at est + (q*P+s)*slot_stride. With
// list_size the walk covers only the chunks that hold real vectors (the reference's ceil(n/16)), not the
// tile padding of the native layout.
template <int ORDER, bool SIGNED>
__global__ void __launch_bounds__(FAST_THREADS, 3)
ivf_scan_fast_kernel(const uint4 *__restrict__ nat, const int64_t *__restrict__ list_chunk_off,
                     const int32_t *__restrict__ list_size, int n_lists, int M,
                     const uint8_t *__restrict__ tables, const int32_t *__restrict__ probes, int P,
                     uint8_t *__restrict__ est, int64_t slot_stride, const int64_t *__restrict__ seg_off, PatchList patch)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint4 *rows = reinterpret_cast<uint4 *>(smem);
    uint4 *raw = rows + M;
    LutMeta *meta = reinterpret_cast<LutMeta *>(raw + M);
    int *scratch = reinterpret_cast<int *>(meta + 1);
    int64_t *seg_c0 = reinterpret_cast<int64_t *>(scratch + 4 * M);                       // 16-byte aligned offset
    int64_t *seg_o = seg_c0 + P;                                                         // est offset of the segment
    int *seg_end = reinterpret_cast<int *>(seg_o + P);                                   // inclusive prefix of chunk counts
    const int q = blockIdx.y, Ph = M >> 1;
    if (threadIdx.x == 0) {
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
            seg_c0[s] = c0;
            seg_o[s] = o;
            run += (int)nc;
            seg_end[s] = run;
        }
    }
    const bool fast_allowed = !(SIGNED && ORDER == TKB_ORDER_SSE);
    prepare_lut<SIGNED>(tables + (size_t)q * M * 16, M, fast_allowed, rows, raw, meta, scratch);   // syncs
    const LutMeta m = *meta;
    const int total = seg_end[P - 1];
    int s = 0;                                       // f only grows: the slot search resumes where it stopped
    for (int f = blockIdx.x * FAST_THREADS + threadIdx.x; f < total; f += gridDim.x * FAST_THREADS) {
        while (f >= seg_end[s]) s++;
        const int local = f - (s ? seg_end[s - 1] : 0);
        const int64_t c = seg_c0[s] + local;
        const int64_t off = seg_o[s] + 16 * (int64_t)local;
        uint4 o;
        if (m.eligible) {
            bool flagged;
            o = scan_chunk_fast<SIGNED>(nat, c, Ph, rows, m, flagged);
            if (flagged && !patch_push(patch, (uint32_t)(q * P + s), (uint32_t)local))
                o = scan_chunk_exact_cold<ORDER, SIGNED>(nat, c, Ph, reinterpret_cast<const uint8_t *>(raw));
        } else {
            o = scan_chunk_exact<ORDER, SIGNED>(nat, c, Ph, reinterpret_cast<const uint8_t *>(raw));
        }
        *reinterpret_cast<uint4 *>(est + off) = o;
    }
}

// Patch pass: one half-warp per queued chunk recomputes its 16 estimates with the reference's fold.
// mode 0: brute force (unit = q, chunk = local); mode 1: IVF (unit = q*P+s, chunk = list start + local).
template <int ORDER, bool SIGNED>
__global__ void __launch_bounds__(256)
patch_kernel(const uint4 *__restrict__ nat, int M, const uint8_t *__restrict__ tables, uint8_t *__restrict__ est,
             PatchList patch, int mode, int64_t stride /* est_stride or slot_stride */, const int64_t *__restrict__ seg_off,
             const int64_t *__restrict__ list_chunk_off, int n_lists, const int32_t *__restrict__ probes, int P)
{
    unsigned long long count = *patch.count;
    if (count > patch.cap) count = patch.cap;
    const int Ph = M >> 1;
    const int hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, v = threadIdx.x & 15;
    const int n_hw = (gridDim.x * blockDim.x) >> 4;
    for (unsigned long long i = hw; i < count; i += n_hw) {
        const uint2 en = patch.entry[i];
        int64_t q, c, off;
        if (mode == 0) {
            q = en.x;
            c = en.y;
            off = q * stride + 16 * c;
        } else {
            q = en.x / (uint32_t)P;
            int l = probes[en.x];
            if (l < 0) l += n_lists;
            c = list_chunk_off[l] + en.y;
            off = (seg_off ? seg_off[en.x] : (int64_t)en.x * stride) + 16 * (int64_t)en.y;
        }
        const int e = exact_vector<ORDER, SIGNED>(nat, c, Ph, v, tables + (size_t)q * M * 16);
        est[off + v] = (uint8_t)e;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static size_t fast_smem_bytes(int M, int P)
{
    return (size_t)M * 32 + sizeof(LutMeta) + sizeof(int) * (4 * (size_t)M + 4) + (size_t)P * 20 + 16;
}

static int check_fast_args(int M, int order)
{
    TKB_REQUIRE(order == TKB_ORDER_SSE || order == TKB_ORDER_AVX, "order must be TKB_ORDER_SSE or TKB_ORDER_AVX");
    TKB_REQUIRE(M > 0 && M % 2 == 0, "M (sub-quantizers) must be a positive multiple of 2");
    TKB_REQUIRE(order != TKB_ORDER_AVX || M % 4 == 0, "avx order needs M % 4 == 0 (ref: fast_pq.py:24 dpad)");
    TKB_REQUIRE(M <= 1024, "M too large");
    return TKB_OK;
}

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

// workspace = 16-byte header (flagged-chunk counter) + 8-byte entries; whatever does not fit is recomputed inline
static int split_workspace(void *workspace, int64_t workspace_bytes, PatchList &pl)
{
    TKB_REQUIRE(workspace && workspace_bytes >= 16 + 8, "scan workspace too small (need at least 24 bytes)");
    TKB_REQUIRE((uintptr_t)workspace % 16 == 0, "workspace must be 16-byte aligned");
    pl.count = reinterpret_cast<unsigned long long *>(workspace);
    pl.entry = reinterpret_cast<uint2 *>(reinterpret_cast<char *>(workspace) + 16);
    pl.cap = (unsigned long long)((workspace_bytes - 16) / 8);
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
    PatchList pl;
    if (int rc = split_workspace(workspace, workspace_bytes, pl)) return rc;
    int64_t tiles = (n_chunks + FAST_THREADS - 1) / FAST_THREADS;
    if (tiles > 148 * 8 && Q > 1) tiles = 148 * 8;                 // grid-stride beyond that
    TKB_REQUIRE(tiles <= 0x7fffffff, "too many chunks for one launch");
    const size_t smem = fast_smem_bytes(M, 0);
    const uint4 *n4 = reinterpret_cast<const uint4 *>(native);
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        dim3 grid((unsigned)tiles, (unsigned)qn);
        TKB_CUDA(cudaMemsetAsync(pl.count, 0, 16, st));
        // patch entries are relative to this launch's block of queries: both kernels get the shifted pointers
        TKB_DISPATCH_FAST(estimate_fast_kernel, grid, FAST_THREADS, smem, st, n4, n_chunks, M,
                          tables + (size_t)q0 * M * 16, est + (size_t)q0 * est_stride, est_stride, pl);
        TKB_LAUNCH_CHECK();
        TKB_DISPATCH_FAST(patch_kernel, 148 * 4, 256, 0, st, n4, M, tables + (size_t)q0 * M * 16,
                          est + (size_t)q0 * est_stride, pl, 0, est_stride, nullptr, nullptr, 0, nullptr, 1);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

int launch_ivf_scan_native(const void *native, const int64_t *list_chunk_off, const int32_t *list_size, int n_lists, int M,
                           const uint8_t *tables, const int32_t *probes, int Q, int P, uint8_t *est,
                           int64_t slot_stride, const int64_t *seg_off, int64_t max_chunks_per_query, int order, int signd,
                           void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    if (int rc = check_fast_args(M, order)) return rc;
    TKB_REQUIRE(Q >= 0 && P >= 0 && n_lists > 0, "bad extent");
    if (Q == 0 || P == 0 || (slot_stride == 0 && !seg_off)) return TKB_OK;
    TKB_REQUIRE(native && list_chunk_off && tables && probes && est, "null pointer");
    TKB_REQUIRE(slot_stride % 16 == 0, "slot_stride must be a multiple of 16");
    TKB_REQUIRE(P <= 4096, "too many probes");
    TKB_REQUIRE((int64_t)Q * P <= 0xffffffffLL, "too many (query, probe) units for one launch");
    PatchList pl;
    if (int rc = split_workspace(workspace, workspace_bytes, pl)) return rc;
    // enough CTAs to fill the machine when Q is small; otherwise one CTA per query walks all its lists
    if (max_chunks_per_query <= 0) max_chunks_per_query = (int64_t)P * (slot_stride / 16);
    int64_t splits = (148 * 4 + Q - 1) / Q;
    const int64_t max_splits = (max_chunks_per_query + FAST_THREADS - 1) / FAST_THREADS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const size_t smem = fast_smem_bytes(M, P);
    const uint4 *n4 = reinterpret_cast<const uint4 *>(native);
    for (int q0 = 0; q0 < Q; q0 += 65535) {
        const int qn = (Q - q0 < 65535) ? (Q - q0) : 65535;
        dim3 grid((unsigned)splits, (unsigned)qn);
        const int64_t *so = seg_off ? seg_off + (size_t)q0 * P : nullptr;          // offsets stay relative to `est`
        uint8_t *eb = seg_off ? est : est + (size_t)q0 * P * slot_stride;
        TKB_CUDA(cudaMemsetAsync(pl.count, 0, 16, st));
        TKB_DISPATCH_FAST(ivf_scan_fast_kernel, grid, FAST_THREADS, smem, st, n4, list_chunk_off, list_size, n_lists, M,
                          tables + (size_t)q0 * M * 16, probes + (size_t)q0 * P, P, eb, slot_stride, so, pl);
        TKB_LAUNCH_CHECK();
        TKB_DISPATCH_FAST(patch_kernel, 148 * 4, 256, 0, st, n4, M, tables + (size_t)q0 * M * 16, eb, pl, 1, slot_stride,
                          so, list_chunk_off, n_lists, probes + (size_t)q0 * P, P);
        TKB_LAUNCH_CHECK();
    }
    return TKB_OK;
}

}  // namespace tkb
