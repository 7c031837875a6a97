// tkb_lut.cu -- batched per-query distance-table (LUT) construction.
//
// Replaces FastPQ.distance_table (ref: tinyknn/fast_pq.py:186-222), FastPQ.udistance_table
// (ref: tinyknn/fast_pq.py:224-252), pad1 (ref: tinyknn/utils.py:6-11) and transform_tables
// (ref: tinyknn/_transform.py:114-138), for Q queries per launch (the reference is single-query).
//
// Bit-exactness with the reference depends on numpy's float semantics, which this kernel mirrors
// operation by operation:
//   * rotated path (R given): q is promoted to f64 by `q @ R.T`, everything after is f64;
//   * unrotated path: diff/dists/shift/max stay f32, `scale` and `dists*scale` are f64 (NEP 50);
//   * einsum('ijk,ijk->ij') over a dpb-long axis: products rounded separately, summed left to
//     right, no FMA contraction;
//   * _mean: numpy's pairwise summation (8 accumulators, blocks of <=128, recursive halving on
//     multiples of 8) over the flattened C-order (16, M) array, then dtype(f64(sum) / count);
//   * np.round = round-half-to-even; astype(uint8) of a negative value wraps modulo 256.
// One CTA per query; all staging in shared memory. The mean is numpy's pairwise sum computed by the whole CTA.
#include "tkb_common.cuh"

namespace tkb {

constexpr int LUT_THREADS = 128;

// numpy pairwise sum (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum):
//   n < 8      : res = 0; res += a[i] in order
//   n <= 128   : r[k] = a[k] (k < 8); r[k] += a[i + k] for i = 8, 16, .. < n - n % 8;
//                res = ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)); then res += a[i] for the n % 8 tail
//   otherwise  : n2 = n / 2; n2 -= n2 % 8; pairwise(a, n2) + pairwise(a + n2, n - n2)
// Computed by the whole CTA, bit for bit. numpy's recursion splits [0, n) into leaves of <= 128 elements
// (n2 = n/2 rounded down to a multiple of 8); inside a leaf the 8 strided partial sums are independent chains and are
// combined in a fixed tree, then the leaves are added pairwise in recursion order. Here thread 0 lists the leaves, 8
// lanes per leaf form the partial sums, one lane per leaf combines them (plus the <8 tail elements), and thread 0 adds
// the leaf sums in recursion order -- the operations and their order are exactly numpy's.
struct PwScratch { int *off, *len; int n_leaves; };     // leaves are >= 60 elements long: at most n / 32 + 2 of them

__device__ void pw_list_leaves(int off, int n, PwScratch *ps)
{
    if (n <= 128) { const int i = ps->n_leaves++; ps->off[i] = off; ps->len[i] = n; return; }
    int n2 = n / 2;
    n2 -= n2 % 8;
    pw_list_leaves(off, n2, ps);
    pw_list_leaves(off + n2, n - n2, ps);
}

template <typename T>
__device__ T pw_combine(int n, const T *leaf_sum, int &next)
{
    if (n <= 128) return leaf_sum[next++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const T l = pw_combine<T>(n2, leaf_sum, next);
    const T r = pw_combine<T>(n - n2, leaf_sum, next);
    return l + r;
}

// a: n values in shared memory; part: 8 * PW_MAX_LEAVES scratch (T); returns the sum in every thread of the CTA.
template <typename T>
__device__ T cta_pairwise_sum(const T *a, int n, PwScratch *ps, T *part, T *leaf_sum, T *result)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { ps->n_leaves = 0; pw_list_leaves(0, n, ps); }
    __syncthreads();
    const int nl = ps->n_leaves;
    for (int i = tid; i < 8 * nl; i += nt) {                   // partial sum k of leaf l: a[k], a[8+k], ... (ascending)
        const int l = i >> 3, k = i & 7, len = ps->len[l];
        const T *p = a + ps->off[l];
        if (len >= 8) {
            T r = p[k];
            for (int j = 8; j < len - (len % 8); j += 8) r += p[j + k];
            part[i] = r;
        }
    }
    __syncthreads();
    for (int l = tid; l < nl; l += nt) {
        const int len = ps->len[l];
        const T *p = a + ps->off[l];
        T res;
        if (len < 8) {
            res = (T)0;
            for (int j = 0; j < len; j++) res += p[j];
        } else {
            const T *r = part + 8 * l;
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (int j = len - (len % 8); j < len; j++) res += p[j];
        }
        leaf_sum[l] = res;
    }
    __syncthreads();
    if (tid == 0) { int next = 0; *result = pw_combine<T>(n, leaf_sum, next); }
    __syncthreads();
    return *result;
}

// CTA-wide min / max of one value per thread (warp shuffles, then one value per warp through shared memory)
template <typename T, bool MAX>
__device__ T cta_minmax(T v, T *red)
{
    for (int o = 16; o > 0; o >>= 1) { const T u = __shfl_xor_sync(FULL, v, o); v = MAX ? fmax(v, u) : fmin(v, u); }
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    T r = red[0];
    for (int w = 1; w < nw; w++) r = MAX ? fmax(r, red[w]) : fmin(r, red[w]);
    return r;
}

__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// T = double when rotated (R != nullptr), float otherwise.
// dynamic smem: T dists[16*M] | T qv[Dp] | red[LUT_THREADS] (T) | float qpad[Dpad] | pairwise-sum scratch
template <typename T>
__global__ void __launch_bounds__(LUT_THREADS)
lut_build_kernel(const float *__restrict__ queries, int d, int normalize, float *__restrict__ q_out,
                 const float *__restrict__ centers, int Dp, int dpb, const double *__restrict__ R,
                 int Dpad, double sqrt_n_blocks, double log_n_blocks, int signd,
                 uint8_t *__restrict__ tables, double *__restrict__ q_rot, double *__restrict__ shift_out,
                 double *__restrict__ scale_out, int pw_off)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = Dp / dpb;
    T *dists = reinterpret_cast<T *>(smem_raw);
    T *qv = dists + 16 * M;
    T *red = qv + Dp;
    float *qpad = reinterpret_cast<float *>(red + LUT_THREADS);
    const int ML = 16 * M / 32 + 2;
    T *pw_part = reinterpret_cast<T *>(smem_raw + pw_off);                 // [8 * ML] partial sums, [ML] leaf sums, then off/len
    __shared__ PwScratch s_pw;
    if (threadIdx.x == 0) { s_pw.off = reinterpret_cast<int *>(pw_part + 9 * ML); s_pw.len = s_pw.off + ML; }
    __shared__ T s_sum;
    __shared__ float s_norm;

    const int q = blockIdx.x, tid = threadIdx.x;
    const float *qin = queries + (size_t)q * d;

    // ---- pad1 (+ optional angular normalisation, ref: ivf.py:126-127) --------------------------
    for (int i = tid; i < Dpad; i += LUT_THREADS) qpad[i] = (i < d) ? qin[i] : 0.0f;
    __syncthreads();
    if (normalize) {
        // `np.linalg.norm(q)` = sqrt(q.dot(q)) in f32, and q.dot(q) is OpenBLAS's sdot. Its arithmetic on the hosts of this pool
        // (OpenBLAS 0.3.30, SkylakeX kernels -- threadpoolctl reports that core here and on the GPU boxes), established against
        // numpy itself (tests/test_oracle_pinned.py, oracle/restate.py:sdot_openblas_skylakex; 100 % of 4 400 random vectors of
        // eleven lengths): the first n & ~31 elements in a vector kernel -- blocks of 64 into four 16-lane FMA accumulators, each
        // folded low half + high half into 8 lanes; a last block of 32 into those four 8-lane accumulators; ((a0 + a1) + a2) + a3;
        // low 4 lanes + high 4 lanes; (h0 + h1) + (h2 + h3) -- then the remaining elements as separately rounded f32 products
        // summed in DOUBLE, plus the kernel's value, rounded to f32 once. Round 2's bench found 1 LUT byte in 2 000 GloVe-shape
        // queries off with a plain shuffle tree; this reproduces numpy's norm bit for bit on such hosts.
        if (tid < 32) {
            const int n1 = d & ~31, n64 = n1 & ~63;
            float r0 = 0.0f, r1 = 0.0f;                         // 16-lane accumulators 0,1 (lanes tid) and 2,3 (lanes tid + 32)
            for (int i = 0; i < n64; i += 64) {
                r0 = fmaf(qpad[i + tid], qpad[i + tid], r0);
                r1 = fmaf(qpad[i + 32 + tid], qpad[i + 32 + tid], r1);
            }
            float f0 = __fadd_rn(r0, __shfl_down_sync(FULL, r0, 8));     // lanes 0..7: accumulator 0, lanes 16..23: accumulator 1
            float f1 = __fadd_rn(r1, __shfl_down_sync(FULL, r1, 8));     // the same for accumulators 2 and 3
            if (n1 > n64 && (tid & 15) < 8) {                   // the last block of 32: element 8 k + l into lane l of accumulator k
                const int k0 = tid >> 4, l = tid & 7;
                const float xa = qpad[n64 + 8 * k0 + l], xb = qpad[n64 + 16 + 8 * k0 + l];
                f0 = fmaf(xa, xa, f0);
                f1 = fmaf(xb, xb, f1);
            }
            const float a1 = __shfl_sync(FULL, f0, (tid & 7) + 16), a3 = __shfl_sync(FULL, f1, (tid & 7) + 16);
            const float s8 = __fadd_rn(__fadd_rn(__fadd_rn(f0, a1), f1), a3);          // lanes 0..7
            const float h4 = __fadd_rn(s8, __shfl_down_sync(FULL, s8, 4));             // lanes 0..3
            const float p2 = __fadd_rn(h4, __shfl_down_sync(FULL, h4, 1));             // lane 0: h0 + h1, lane 2: h2 + h3
            const float vec = __fadd_rn(p2, __shfl_down_sync(FULL, p2, 2));            // lane 0
            if (tid == 0) {
                double t = 0.0;
                for (int i = n1; i < d; i++) t += (double)__fmul_rn(qpad[i], qpad[i]);
                s_norm = sqrtf((float)(t + (double)(n1 ? vec : 0.0f)));
            }
        }
        __syncthreads();
        const float nrm = s_norm;
        for (int i = tid; i < d; i += LUT_THREADS) {
            const float v = qpad[i] / nrm;            // q /= norm : f32 division like numpy
            qpad[i] = v;
        }
        __syncthreads();
    }
    if (q_out) for (int i = tid; i < d; i += LUT_THREADS) q_out[(size_t)q * d + i] = qpad[i];

    // ---- rotation q @ R.T (f64) or identity ----------------------------------------------------
    if (R) {
        // numpy's `q @ R.T` (f32 vector promoted to f64, times the f64 matrix) is OpenBLAS's dgemv: four accumulators
        // (k mod 4), each a sequential FMA chain over k, combined as (a0 + a2) + (a1 + a3) -- established against numpy itself
        // (numpy 2.3.5 / OpenBLAS 0.3.30, every Dpad % 4 == 0 from 16 to 1024: tests/test_oracle_pinned.py). The same
        // operations in the same order here, four lanes per output, so q_rot equals numpy's bit for bit; a tail of Dpad % 4
        // elements (sse build with odd block counts only) is summed separately and added last.
        const int K4 = Dpad & ~3;
        for (int u0 = 0; u0 < 4 * Dp; u0 += LUT_THREADS) {
            const int u = u0 + tid, i = u >> 2, l = u & 3;
            const bool valid = u < 4 * Dp;
            double acc = 0.0;
            if (valid) {
                const double *Ri = R + (size_t)i * Dpad;
                for (int k = l; k < K4; k += 4) acc = fma((double)qpad[k], Ri[k], acc);
            }
            const double s = add_rn(acc, __shfl_xor_sync(FULL, acc, 2));        // lanes 0,2: a0 + a2; lanes 1,3: a1 + a3
            double r = add_rn(s, __shfl_xor_sync(FULL, s, 1));                  // lane 0: (a0 + a2) + (a1 + a3)
            if (valid && l == 0) {
                if (K4 < Dpad) {
                    const double *Ri = R + (size_t)i * Dpad;
                    double t = 0.0;
                    for (int k = K4; k < Dpad; k++) t = add_rn(t, mul_rn((double)qpad[k], Ri[k]));
                    r = add_rn(r, t);
                }
                qv[i] = (T)r;
            }
        }
    } else {
        for (int i = tid; i < Dp; i += LUT_THREADS) qv[i] = (T)qpad[i];
    }
    __syncthreads();
    if (q_rot) for (int i = tid; i < Dp; i += LUT_THREADS) q_rot[(size_t)q * Dp + i] = (double)qv[i];

    // ---- dists[c][m] = sum_k (centers[c][m*dpb+k] - q[m*dpb+k])^2 ------------------------------
    {
        int c = tid / M, m = tid - c * M;                                  // e = c * M + m advances by LUT_THREADS
        const int dc = LUT_THREADS / M, dm = LUT_THREADS - dc * M;
        for (int e = tid; e < 16 * M; e += LUT_THREADS) {
            const float *cen = centers + (size_t)c * Dp + m * dpb;
            const T *qq = qv + m * dpb;
            T acc;
            {
                const T df = (T)cen[0] - qq[0];
                acc = mul_rn(df, df);
            }
            for (int k = 1; k < dpb; k++) {
                const T df = (T)cen[k] - qq[k];
                acc = add_rn(acc, mul_rn(df, df));
            }
            dists[e] = acc;
            c += dc; m += dm;
            if (m >= M) { m -= M; c++; }
        }
    }
    __syncthreads();

    // ---- shift ---------------------------------------------------------------------------------
    T shift;
    if (signd) {
        const T sum = cta_pairwise_sum<T>(dists, 16 * M, &s_pw, pw_part, pw_part + 8 * ML, &s_sum);
        const T mean = (T)((double)sum / (double)(16 * M));         // dtype(f64(sum)/intp(count))
        shift = mul_rn(mean, (T)0.6931471806);                      // python float is "weak": multiply in T
    } else {
        T mn = (T)INFINITY;
        for (int e = tid; e < 16 * M; e += LUT_THREADS) mn = fmin(mn, dists[e]);
        shift = cta_minmax<T, false>(mn, red);
    }

    // ---- dists -= shift ; max ------------------------------------------------------------------
    T mx = -(T)INFINITY;
    for (int e = tid; e < 16 * M; e += LUT_THREADS) {
        const T v = dists[e] - shift;
        dists[e] = v;
        mx = fmax(mx, v);
    }
    const double amax = (double)cta_minmax<T, true>(mx, red);

    // ---- scale (f64 on both paths) -------------------------------------------------------------
    double scale;
    if (signd) scale = 128.0 / (amax * sqrt_n_blocks);                       // fast_pq.py:216
    else       scale = 255.0 / ((amax * log_n_blocks) * sqrt_n_blocks);      // fast_pq.py:248
    if (tid == 0) {
        if (shift_out) shift_out[q] = (double)shift;
        if (scale_out) scale_out[q] = scale;
    }

    // ---- round, wrap to u8, transpose to [M][16] -----------------------------------------------
    uint8_t *tq = tables + (size_t)q * M * 16;
    for (int e = tid; e < 16 * M; e += LUT_THREADS) {
        const int m = e >> 4, c = e & 15;
        const double r = rint(__dmul_rn((double)dists[c * M + m], scale));
        tq[e] = (uint8_t)(int)r;                                             // C-cast wrap: -3.0 -> 253
    }
}

int launch_lut_build(const float *queries, int Q, int d, int normalize, float *q_out,
                     const float *centers, int Dp, int dpb, const double *R, int Dpad,
                     double sqrt_n_blocks, double log_n_blocks, int signd, uint8_t *tables,
                     double *q_rot, double *shift, double *scale, cudaStream_t st)
{
    TKB_REQUIRE(Q >= 0, "negative Q");
    if (Q == 0) return TKB_OK;
    TKB_REQUIRE(queries && centers && tables, "null pointer");
    TKB_REQUIRE(d > 0 && dpb > 0 && Dp > 0 && Dp % dpb == 0, "bad dimensions");
    TKB_REQUIRE(Dpad >= d, "Dpad must be >= d");
    TKB_REQUIRE(R != nullptr || Dp == Dpad, "without a rotation Dp must equal the padded dimension");
    const int M = Dp / dpb;
    const size_t tsz = R ? sizeof(double) : sizeof(float);
    size_t pw_off = tsz * (16 * (size_t)M + Dp + LUT_THREADS) + sizeof(float) * Dpad;
    pw_off = (pw_off + 15) / 16 * 16;
    const size_t ML = 16 * (size_t)M / 32 + 2;
    const size_t smem = pw_off + (tsz * 9 + 8) * ML;
    TKB_REQUIRE(smem <= 200 * 1024, "dimension too large for the LUT kernel");
    if (R) {
        TKB_CUDA(cudaFuncSetAttribute(lut_build_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lut_build_kernel<double><<<Q, LUT_THREADS, smem, st>>>(queries, d, normalize, q_out, centers, Dp, dpb, R, Dpad,
                                                                sqrt_n_blocks, log_n_blocks, signd, tables, q_rot, shift, scale, (int)pw_off);
    } else {
        TKB_CUDA(cudaFuncSetAttribute(lut_build_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lut_build_kernel<float><<<Q, LUT_THREADS, smem, st>>>(queries, d, normalize, q_out, centers, Dp, dpb, R, Dpad,
                                                               sqrt_n_blocks, log_n_blocks, signd, tables, q_rot, shift, scale, (int)pw_off);
    }
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
