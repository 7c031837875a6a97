// tkb_scan_core.cuh -- device-side core of the fast PQ scan, shared by the stand-alone scan kernels
// (tkb_scan_fast.cu) and the fused per-query pipeline (tkb_fused.cu). See tkb_scan_fast.cu for the layout and the
// exactness argument (certificate + exact recomputation).
#pragma once
#include "tkb_common.cuh"

namespace tkb {
constexpr int FAST_THREADS = 256;
constexpr int TILE = 8;                        // chunks per tile

__device__ __forceinline__ size_t native_off(int64_t chunk, int p, int Ph)     // in uint4 units
{
    return ((size_t)(chunk >> 3) * Ph + p) * TILE + (chunk & 7);
}

// ------------------------------------------------------------------------------------------------
// per-query LUT preparation (CTA prologue)
// ------------------------------------------------------------------------------------------------
struct LutMeta {
    int eligible;          // fast path allowed for this query
    int bias_tot;          // bias_0 + bias_1                         (signed fast path)
    int k0, k1;            // certificate thresholds on the biased lane sums: S'_l <= k_l
    int steps_ok;          // the byte-SIMD step-by-step fold (scan_chunk_steps) is usable: biased entries < 128
    int pad_[3];
};

// smem: uint4 rows[M] (rows biased to >= 0) | raw copy uint4 raw[M] | uint2 sc[M] (clamp bounds of the step-by-step fold
// after each row, in the biased domain) | LutMeta | scratch. ORDER decides which rows share an accumulator.
template <int ORDER, bool SIGNED>
__device__ void prepare_lut(const uint8_t *__restrict__ tq, int M, bool fast_allowed, uint4 *rows, uint4 *raw, uint2 *sc,
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
        // Certificate thresholds (raw domain: S_l <= kk_l => no prefix of lane l's fold can exceed 127). A prefix ending at row k
        // is at most Pmax_k (the rows' largest positive entries so far) and at most S_l + the negatives the rows after k can
        // still add; it can pass 127 only from the first row k* with Pmax_k* > 127 on, where the second bound is largest:
        // kk_l = 127 - sum_{j > k*} max(0, -min_c t_j) (127 when Pmax never passes 127). 127 - N_l, the bound over ALL rows, flagged
        // ~2 % of the chunks of the benchmark indexes; this one flags almost none.
        int kk[2] = {127, 127};
        {
            int Pm[2] = {0, 0}, suf[2] = {N[0], N[1]};
            bool found[2] = {false, false};
            for (int j = 0; j < M; j++) {
                const int l = (j >> 1) & 1;
                suf[l] -= scratch[4 * j + 1];
                Pm[l] += max(0, scratch[4 * j + 2] - scratch[4 * j + 0]);
                if (!found[l] && Pm[l] > 127) { found[l] = true; kk[l] = 127 - suf[l]; }
            }
        }
        m.k0 = kk[0] + bias[0];
        m.k1 = kk[1] + bias[1];
        // step-by-step fold in the biased domain: A_j = a_j + B_j with B_j the bias accumulated so far in the row's
        // accumulator, so that a_j = clamp(a_{j-1} + t_j) becomes A_j = clamp(A_{j-1} + t'_j, -128 + B_j, 127 + B_j)
        int B[2] = {0, 0};
        for (int j = 0; j < M; j++) {
            const int l = ORDER == TKB_ORDER_AVX ? (j >> 1) & 1 : 0;
            B[l] += scratch[4 * j + 0];
            const int lo = SIGNED ? -128 + B[l] : 0, hi = SIGNED ? 127 + B[l] : 255;
            sc[j] = make_uint2((uint32_t)(lo & 0xffff) * 0x00010001u, (uint32_t)(hi & 0xffff) * 0x00010001u);
        }
        m.steps_ok = range <= 127 && B[0] <= 30000 && B[1] <= 30000;
        m.pad_[0] = m.pad_[1] = m.pad_[2] = 0;
        *meta = m;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// the fast chunk kernel: 16 vectors, M sub-quantizers, AVX lane split (pair p -> lane p & 1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
#ifdef TKB_EMULATE                       // tests/emulate: the source compiled for the CPU, no PTX
    return emu::prmt(a, b, s);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
#endif
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
// PH > 0: the number of sub-quantizer pairs is a compile-time constant (26: GloVe-100 shape, 16: 128-d rotated to 64), the
// pair loop unrolls completely and the tail predicates of the generic loop disappear (scan 0.65 -> 0.58 ms on the
// GloVe-shape bench). Moving the selector shifts to the FMA pipe as IMAD.HI was measured too: no gain (tools/ubench.cu).
template <bool SIGNED, int PH = 0>
__device__ __forceinline__ uint4 scan_chunk_fast(const uint4 *__restrict__ nat, int64_t chunk, int Ph_rt,
                                                 const uint4 *__restrict__ rows, const LutMeta &meta,
                                                 bool &flagged)
{
    const int Ph = PH > 0 ? PH : Ph_rt;
    uint32_t wide[2][4][2];                    // [lane][group][even/odd] packed s16x2 biased sums
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
        for (int g = 0; g < 4; g++) { wide[l][g][0] = 0; wide[l][g][1] = 0; }

    const uint4 *base = nat + native_off(chunk, 0, Ph);
#pragma unroll
    for (int p0 = 0; p0 < (PH > 0 ? PH : Ph); p0 += 8) {
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
                lookup_step(rows[j], w[i].x, w[i].y, acc[i & 1]);      // sub-quantizer 2p
                lookup_step(rows[j + 1], w[i].z, w[i].w, acc[i & 1]);  // sub-quantizer 2p+1
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

// ------------------------------------------------------------------------------------------------
// The reference's fold step by step, 16 vectors per thread, byte-SIMD: every row is looked up with PRMT like in the
// fast path, widened to s16x2 at once and added with the clamp of THAT step (VIADDMNMX + VIMNMX), so the result is the
// reference's for any table whose biased entries stay below 128 -- no certificate needed. About twice the work of
// scan_chunk_fast; used for the chunks whose certificate failed, for LUTs that are not eligible for the fast path and
// for the signed SSE order (one accumulator: every prefix matters).
// ------------------------------------------------------------------------------------------------
template <bool SIGNED>
__device__ __forceinline__ void steps_row(const uint4 L, const uint2 c, uint32_t wa, uint32_t wb, uint32_t (&A)[4][2])
{
    const uint32_t xa = wa ^ 0x88888888u, xb = wb ^ 0x88888888u;
    uint32_t b[4];
    b[0] = prmt(L.x, L.y, wa) + prmt(L.z, L.w, xa);
    b[1] = prmt(L.x, L.y, wa >> 16) + prmt(L.z, L.w, xa >> 16);
    b[2] = prmt(L.x, L.y, wb) + prmt(L.z, L.w, xb);
    b[3] = prmt(L.x, L.y, wb >> 16) + prmt(L.z, L.w, xb >> 16);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t ev = prmt(b[g], 0u, 0x4240u), od = prmt(b[g], 0u, 0x4341u);      // vectors 4g+0,4g+2 / 4g+1,4g+3
        if (SIGNED) {
            A[g][0] = __vimin3_s16x2(__viaddmax_s16x2(A[g][0], ev, c.x), c.y, c.y);
            A[g][1] = __vimin3_s16x2(__viaddmax_s16x2(A[g][1], od, c.x), c.y, c.y);
        } else {
            A[g][0] = __viaddmin_u16x2(A[g][0], ev, 0x00ff00ffu);
            A[g][1] = __viaddmin_u16x2(A[g][1], od, 0x00ff00ffu);
        }
    }
}

template <int ORDER, bool SIGNED>
__device__ __forceinline__ uint4 scan_chunk_steps(const uint4 *__restrict__ nat, int64_t chunk, int Ph,
                                                  const uint4 *__restrict__ rows, const uint2 *__restrict__ sc,
                                                  const LutMeta &meta)
{
    uint32_t A0[4][2], A1[4][2];               // accumulators of rows with j&2 == 0 / != 0 (avx); sse uses A0 only
#pragma unroll
    for (int g = 0; g < 4; g++) { A0[g][0] = A0[g][1] = A1[g][0] = A1[g][1] = 0; }
    const uint4 *base = nat + native_off(chunk, 0, Ph);
    for (int p = 0; p < Ph; p += 2) {
        const uint4 w0 = ldg_nc_u4(base + (size_t)p * TILE);
        uint4 w1 = make_uint4(0, 0, 0, 0);
        const bool two = p + 1 < Ph;
        if (two) w1 = ldg_nc_u4(base + (size_t)(p + 1) * TILE);
        steps_row<SIGNED>(rows[2 * p], sc[2 * p], w0.x, w0.y, A0);
        steps_row<SIGNED>(rows[2 * p + 1], sc[2 * p + 1], w0.z, w0.w, A0);
        if (two) {
            if (ORDER == TKB_ORDER_AVX) {
                steps_row<SIGNED>(rows[2 * p + 2], sc[2 * p + 2], w1.x, w1.y, A1);
                steps_row<SIGNED>(rows[2 * p + 3], sc[2 * p + 3], w1.z, w1.w, A1);
            } else {
                steps_row<SIGNED>(rows[2 * p + 2], sc[2 * p + 2], w1.x, w1.y, A0);
                steps_row<SIGNED>(rows[2 * p + 3], sc[2 * p + 3], w1.z, w1.w, A0);
            }
        }
    }
    uint32_t outw[4];
    if (SIGNED) {
        const uint32_t nbias = (uint32_t)((-meta.bias_tot) & 0xffff) * 0x00010001u;
        const uint32_t lo128 = 0xff80ff80u, hi127 = 0x007f007fu;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t e[2];
#pragma unroll
            for (int h = 0; h < 2; h++)        // sat(a0 + a1): the final add of the avx order; a no-op clamp for sse (A1 == 0)
                e[h] = __vimin3_s16x2(__viaddmax_s16x2(__vadd2(A0[g][h], A1[g][h]), nbias, lo128), hi127, hi127);
            outw[g] = prmt(e[0], e[1], 0x6240u);
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t e[2];
#pragma unroll
            for (int h = 0; h < 2; h++) e[h] = __vimin3_u16x2(__vadd2(A0[g][h], A1[g][h]), 0x00ff00ffu, 0x00ff00ffu);
            outw[g] = prmt(e[0], e[1], 0x6240u);
        }
    }
    return make_uint4(outw[0], outw[1], outw[2], outw[3]);
}

template <int ORDER, bool SIGNED>
__device__ __noinline__ uint4 scan_chunk_steps_cold(const uint4 *__restrict__ nat, int64_t chunk, int Ph,
                                                    const uint4 *__restrict__ rows, const uint2 *__restrict__ sc,
                                                    const LutMeta &meta)
{
    return scan_chunk_steps<ORDER, SIGNED>(nat, chunk, Ph, rows, sc, meta);
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

// The same fold for the patch kernel, where nothing else hides latency: the code words and then the table bytes of 8
// pairs are fetched before the (sequential) adds consume them, so a block costs two load round trips instead of sixteen.
template <int ORDER, bool SIGNED>
__device__ __forceinline__ int exact_vector_batched(const uint4 *__restrict__ nat, int64_t chunk, int Ph, int v,
                                                    const uint8_t *__restrict__ raw /* M*16 bytes */)
{
    const int g = v >> 2, sh = 4 * (v & 3) + 16 * (g & 1);
    int a0 = 0, a1 = 0;
    for (int p0 = 0; p0 < Ph; p0 += 8) {
        uint32_t c0[8], c1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (p0 + i < Ph) {
                const uint4 w = nat[native_off(chunk, p0 + i, Ph)];
                c0[i] = (((g < 2) ? w.x : w.y) >> sh) & 15u;
                c1[i] = (((g < 2) ? w.z : w.w) >> sh) & 15u;
            }
        }
        int t0[8], t1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (p0 + i < Ph) {
                t0[i] = raw[32 * (p0 + i) + c0[i]];
                t1[i] = raw[32 * (p0 + i) + 16 + c1[i]];
                if (SIGNED) { t0[i] = (int)(int8_t)t0[i]; t1[i] = (int)(int8_t)t1[i]; }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (p0 + i < Ph) {                                    // (p0 + i) & 1 == i & 1
                if (ORDER == TKB_ORDER_AVX && (i & 1)) a1 = sat_add8<SIGNED>(sat_add8<SIGNED>(a1, t0[i]), t1[i]);
                else                                  a0 = sat_add8<SIGNED>(sat_add8<SIGNED>(a0, t0[i]), t1[i]);
            }
        }
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

}  // namespace tkb
