/*
 * chain_oracle.c -- scalar C statement of the floating-point arithmetic the build-time CUDA kernels implement
 * (TEST INFRASTRUCTURE ONLY; loaded by tests/ only).
 *
 * tkb_encode.cu (FastPQ.transform, ref: tinyknn/fast_pq.py:147-184) and tkb_assign.cu (knn_brute in IVF.build,
 * ref: tinyknn/ivf.py:84-86, tinyknn/utils.py:66-86) claim to be bit-identical to the reference because numpy's `@`
 * (OpenBLAS gemm) is a sequential FMA chain over the inner dimension starting from 0, and np.einsum('ij,ij->i') is
 * separately rounded products added left to right. This file states exactly that arithmetic one element at a time; the CPU
 * tests compare it with numpy itself (tests/test_oracle_pinned.py), which pins the assumption on the machine that runs
 * the tests, and the GPU tests compare the kernels with numpy. Compile with -ffp-contract=off: every rounding below is
 * explicit.
 */
#include <math.h>
#include <stdint.h>

/* out[i][j] = fma chain over k of a[i][k] * b[j][k], acc starts at 0 */
void tko_dchain(const double *a, const double *b, double *out, int64_t n, int m, int K)
{
    for (int64_t i = 0; i < n; i++)
        for (int j = 0; j < m; j++) {
            double acc = 0.0;
            for (int k = 0; k < K; k++) acc = fma(a[i * K + k], b[(int64_t)j * K + k], acc);
            out[i * m + j] = acc;
        }
}

void tko_schain(const float *a, const float *b, float *out, int64_t n, int m, int K)
{
    for (int64_t i = 0; i < n; i++)
        for (int j = 0; j < m; j++) {
            float acc = 0.0f;
            for (int k = 0; k < K; k++) acc = fmaf(a[i * K + k], b[(int64_t)j * K + k], acc);
            out[i * m + j] = acc;
        }
}

/* Nearest-of-16 codes of n (already padded / rotated) rows, compute type double: codes[i][m] in 0..15.
 * x [n][Dp] f64, centers f32 [16][Dp], cnorm f32 [M][16] (|c|^2 in f32), dpb dims per block. */
void tko_encode_f64(const double *x, int64_t n, int Dp, int dpb, const float *centers, const float *cnorm, uint8_t *codes)
{
    const int M = Dp / dpb;
    for (int64_t i = 0; i < n; i++)
        for (int m = 0; m < M; m++) {
            const double *xv = x + i * Dp + m * dpb;
            double xn = xv[0] * xv[0];
            for (int k = 1; k < dpb; k++) xn = xn + xv[k] * xv[k];
            int best = 0;
            double bv = 0;
            for (int c = 0; c < 16; c++) {
                const float *cc = centers + (int64_t)c * Dp + m * dpb;
                double dot2 = 0.0;
                for (int k = 0; k < dpb; k++) dot2 = fma(xv[k] + xv[k], (double)cc[k], dot2);
                const double part = (xn + (double)cnorm[m * 16 + c]) - dot2;
                if (c == 0 || part < bv) { bv = part; best = c; }
            }
            codes[i * M + m] = (uint8_t)best;
        }
}

void tko_encode_f32(const float *x, int64_t n, int Dp, int dpb, const float *centers, const float *cnorm, uint8_t *codes)
{
    const int M = Dp / dpb;
    for (int64_t i = 0; i < n; i++)
        for (int m = 0; m < M; m++) {
            const float *xv = x + i * Dp + m * dpb;
            float xn = xv[0] * xv[0];
            for (int k = 1; k < dpb; k++) xn = xn + xv[k] * xv[k];
            int best = 0;
            float bv = 0;
            for (int c = 0; c < 16; c++) {
                const float *cc = centers + (int64_t)c * Dp + m * dpb;
                float dot2 = 0.0f;
                for (int k = 0; k < dpb; k++) dot2 = fmaf(xv[k] + xv[k], cc[k], dot2);
                const float part = (xn + cnorm[m * 16 + c]) - dot2;
                if (c == 0 || part < bv) { bv = part; best = c; }
            }
            codes[i * M + m] = (uint8_t)best;
        }
}

/* Nearest centroid of every row (f32): part = (xnorm + cnorm) - chain(2x, c); first minimum. */
void tko_assign_f32(const float *x, int64_t n, int d, const float *c, int C, const float *xnorm, const float *cnorm, int32_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        int best = 0;
        float bv = 0;
        for (int j = 0; j < C; j++) {
            float acc = 0.0f;
            for (int k = 0; k < d; k++) acc = fmaf(x[i * d + k] + x[i * d + k], c[(int64_t)j * d + k], acc);
            const float part = (xnorm[i] + cnorm[j]) - acc;
            if (j == 0 || part < bv) { bv = part; best = j; }
        }
        out[i] = best;
    }
}

/* numpy's VECTOR @ MATRIX.T for float64 (`q @ R.T` in FastPQ.distance_table, ref: tinyknn/fast_pq.py:203-204) is OpenBLAS's
 * dgemv, not gemm: four accumulators (k mod 4), each a sequential FMA chain, combined as (a0 + a2) + (a1 + a3).
 * out[j] = that over k of a[k] * b[j][k]; K % 4 == 0 (every padded dimension of the avx build). tkb_lut.cu mirrors it. */
void tko_dgemv4(const double *a, const double *b, double *out, int m, int K)
{
    for (int j = 0; j < m; j++) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < K; k++) acc[k & 3] = fma(a[k], b[(int64_t)j * K + k], acc[k & 3]);
        out[j] = (acc[0] + acc[2]) + (acc[1] + acc[3]);
    }
}
