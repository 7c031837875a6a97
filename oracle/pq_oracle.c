/*
 * pq_oracle.c -- CPU restatement of tinyknn's 4-bit Quick-ADC scan + heap (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the CUDA hot path. It is plain scalar C (no SIMD): it
 * restates WHAT the reference's SSE/AVX kernels compute, element by element, so that a test can
 * compare the GPU result with something a reader can check against the reference by eye.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. Nothing under tinyknn_b200/ links or calls it.
 *
 * Parity pinned: validated bit-for-bit against the compiled reference (oracle/_ref, built from
 * /root/reference/tinyknn/_fast_pq.pyx and _fast_pq_256.pyx) and against the reference's own
 * known-answer tests (tests/test_oracle_pinned.py).
 *
 * Layouts (reference: tinyknn/_transform.py:53-77 and :114-138)
 *   codes : uint64[n_chunks][M]   chunk = 16 vectors. The 16 bytes at &codes[c][2p] hold one byte
 *           per vector v = 0..15 of the chunk; low nibble = code of sub-quantizer 2p,
 *           high nibble = code of sub-quantizer 2p+1.
 *   tables: uint64[2M] = uint8[M][16]; row j = LUT of sub-quantizer j, byte c = entry for code c
 *           (int8 two's complement when signd).
 *   out   : uint64[2*n_chunks] = one byte per vector, in vector order.
 */
#include <stdint.h>
#include <stddef.h>

#define TKO_ORDER_SSE 0
#define TKO_ORDER_AVX 1

static inline int sat_add(int acc, int t, int signd)
{
    /* _mm_adds_epi8 / _mm_adds_epu8 semantics (_fast_pq.pyx:230-232, _fast_pq_256.pyx:143-145) */
    int s = acc + t;
    if (signd) {
        if (s > 127) s = 127;
        if (s < -128) s = -128;
    } else {
        if (s > 255) s = 255;
    }
    return s;
}

static inline int code_of(const uint8_t *chunk_bytes, int j, int v)
{
    /* byte (2*(j/2))*8 + v ... : the 16 bytes of pair p = j/2 start at offset 16*p */
    uint8_t b = chunk_bytes[16 * (j >> 1) + v];
    return (j & 1) ? (b >> 4) : (b & 15);
}

static inline int entry_of(const uint8_t *tab, int j, int c, int signd)
{
    uint8_t e = tab[16 * j + c];
    return signd ? (int)(int8_t)e : (int)e;
}

/* One chunk: 16 estimates. est[v] is in [-128,127] (signd) or [0,255].
 * SSE order (_fast_pq.pyx:209-236): one accumulator, sub-quantizers folded in ascending j.
 * AVX order (_fast_pq_256.pyx:126-156): 256-bit register = two 128-bit halves; half 0 folds
 *   sub-quantizers 4i,4i+1, half 1 folds 4i+2,4i+3 (i.e. (j & 2) == 0 / != 0, ascending j), and
 *   the two halves are combined with one more saturating add (:151-156). */
void tko_chunk_estimates(const uint64_t *chunk, int M, const uint64_t *tables,
                         int order, int signd, int *est)
{
    const uint8_t *cb = (const uint8_t *)chunk;
    const uint8_t *tab = (const uint8_t *)tables;
    for (int v = 0; v < 16; v++) {
        if (order == TKO_ORDER_SSE) {
            int a = 0;
            for (int j = 0; j < M; j++)
                a = sat_add(a, entry_of(tab, j, code_of(cb, j, v), signd), signd);
            est[v] = a;
        } else {
            int a0 = 0, a1 = 0;
            for (int j = 0; j < M; j++) {
                int t = entry_of(tab, j, code_of(cb, j, v), signd);
                if ((j & 2) == 0) a0 = sat_add(a0, t, signd);
                else              a1 = sat_add(a1, t, signd);
            }
            est[v] = sat_add(a0, a1, signd);
        }
    }
}

/* estimate_pq_{sse,avx} (_fast_pq.pyx:101-111, _fast_pq_256.pyx:52-62) */
void tko_estimate_pq(const uint64_t *codes, int64_t n_chunks, int M, const uint64_t *tables,
                     uint64_t *out, int order, int signd)
{
    uint8_t *o = (uint8_t *)out;
    int est[16];
    for (int64_t c = 0; c < n_chunks; c++) {
        tko_chunk_estimates(codes + c * M, M, tables, order, signd, est);
        for (int v = 0; v < 16; v++) o[16 * c + v] = (uint8_t)est[v];
    }
}

/* init_heap (_fast_pq.pyx:240-252) */
void tko_init_heap(int64_t *indices, int32_t *vals, int R, int signd)
{
    for (int i = 0; i < R; i++) { indices[i] = -1; vals[i] = signd ? 127 : 255; }
}

/* insert (_fast_pq.pyx:274-307): dedupe on label over ALL R slots, then overwrite the root and
 * sift down; a child replaces the hole when strictly greater (left is tried first, right only
 * wins when strictly greater than what was chosen so far). */
void tko_insert(int64_t *indices, int32_t *vals, int R, int64_t label, int v)
{
    for (int j = 0; j < R; j++)
        if (indices[j] == label) return;
    int j = 0;
    for (;;) {
        int nxt = j, nxt_val = v;
        int l = 2 * j + 1, r = 2 * j + 2;
        if (l < R && vals[l] > nxt_val) { nxt = l; nxt_val = vals[l]; }
        if (r < R && vals[r] > nxt_val) { nxt = r; nxt_val = vals[r]; }
        if (nxt == j) { vals[j] = v; indices[j] = label; return; }
        vals[j] = vals[nxt]; indices[j] = indices[nxt];
        j = nxt;
    }
}

/* insert_is (_fast_pq.pyx:256-271): insertion-sort variant (unused by the library itself). */
void tko_insert_is(int64_t *indices, int32_t *vals, int R, int64_t label, int v)
{
    for (int j = 0; j < R; j++)
        if (indices[j] == label) return;
    int j = 0;
    while (j + 1 != R && vals[j + 1] > v) {
        indices[j] = indices[j + 1]; vals[j] = vals[j + 1];
        j++;
    }
    indices[j] = label; vals[j] = v;
}

/* query_pq_{sse,avx} (_fast_pq.pyx:114-206, _fast_pq_256.pyx:65-123).
 * bound = vals[0] truncated to 8 bits and FROZEN for a chunk (:153, :206); a lane is a candidate
 * when est < bound (signed compare, or unsigned compare emulated at :167-169), lanes are visited
 * in ascending position (:176-203), padding positions >= n are skipped (:193), the label is
 * labels[pos] or pos (:197), and the bound is refreshed after the chunk only if it had any
 * candidate bit (:174, :206) -- which is equivalent to refreshing always, since vals[0] cannot
 * change without an insert. */
void tko_query_pq(const uint64_t *codes, int64_t n_chunks, int M, int n, const uint64_t *tables,
                  int64_t *indices, int32_t *vals, int R, int order, int signd,
                  const int64_t *labels)
{
    int est[16];
    if (R <= 0) return;
    int bound = signd ? (int)(int8_t)vals[0] : (int)(uint8_t)vals[0];
    for (int64_t c = 0; c < n_chunks; c++) {
        tko_chunk_estimates(codes + c * M, M, tables, order, signd, est);
        int any = 0;
        for (int v = 0; v < 16; v++) {
            if (est[v] < bound) {
                any = 1;
                int64_t pos = 16 * c + v;
                if (pos < n)
                    tko_insert(indices, vals, R, labels ? labels[pos] : pos, est[v]);
            }
        }
        if (any) bound = signd ? (int)(int8_t)vals[0] : (int)(uint8_t)vals[0];
    }
}

/* Same replay, but fed with precomputed estimates (one byte per padded position). Used to
 * check the GPU replay kernel independently of the GPU scan kernel. */
void tko_replay(const uint8_t *est8, int64_t n_chunks, int n, int64_t *indices, int32_t *vals,
                int R, int signd, const int64_t *labels)
{
    if (R <= 0) return;
    int bound = signd ? (int)(int8_t)vals[0] : (int)(uint8_t)vals[0];
    for (int64_t c = 0; c < n_chunks; c++) {
        int any = 0;
        for (int v = 0; v < 16; v++) {
            int e = signd ? (int)(int8_t)est8[16 * c + v] : (int)est8[16 * c + v];
            if (e < bound) {
                any = 1;
                int64_t pos = 16 * c + v;
                if (pos < n)
                    tko_insert(indices, vals, R, labels ? labels[pos] : pos, e);
            }
        }
        if (any) bound = signd ? (int)(int8_t)vals[0] : (int)(uint8_t)vals[0];
    }
}
