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
// One CTA per query; all staging in shared memory.
#include "tkb_common.cuh"

namespace tkb {

constexpr int LUT_THREADS = 128;

// numpy pairwise sum (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum), executed by
// one thread. The 8 partial sums per leaf block are independent, so the FP pipe stays busy.
template <typename T>
__device__ T np_pairwise_sum(const T *a, int n)
{
    if (n < 8) {
        T res = (T)0;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    if (n <= 128) {
        T r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
            r0 += a[i + 0]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
            r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
        }
        T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// T = double when rotated (R != nullptr), float otherwise.
// dynamic smem: T dists[16*M] | T qv[Dp] | float qpad[Dpad] | red[LUT_THREADS] (T)
template <typename T>
__global__ void __launch_bounds__(LUT_THREADS)
lut_build_kernel(const float *__restrict__ queries, int d, int normalize, float *__restrict__ q_out,
                 const float *__restrict__ centers, int Dp, int dpb, const double *__restrict__ R,
                 int Dpad, double sqrt_n_blocks, double log_n_blocks, int signd,
                 uint8_t *__restrict__ tables, double *__restrict__ q_rot, double *__restrict__ shift_out,
                 double *__restrict__ scale_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = Dp / dpb;
    T *dists = reinterpret_cast<T *>(smem_raw);
    T *qv = dists + 16 * M;
    T *red = qv + Dp;
    float *qpad = reinterpret_cast<float *>(red + LUT_THREADS);
    __shared__ T s_shift;
    __shared__ float s_norm;

    const int q = blockIdx.x, tid = threadIdx.x;
    const float *qin = queries + (size_t)q * d;

    // ---- pad1 (+ optional angular normalisation, ref: ivf.py:126-127) --------------------------
    for (int i = tid; i < Dpad; i += LUT_THREADS) qpad[i] = (i < d) ? qin[i] : 0.0f;
    __syncthreads();
    if (normalize) {
        if (tid < 32) {
            float acc = 0.0f;
            for (int i = tid; i < d; i += 32) acc = fmaf(qpad[i], qpad[i], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            if (tid == 0) s_norm = sqrtf(acc);
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
        const int warp = tid >> 5, lane = tid & 31;
        for (int i = warp; i < Dp; i += LUT_THREADS / 32) {
            const double *Ri = R + (size_t)i * Dpad;
            double acc = 0.0;
            for (int k = lane; k < Dpad; k += 32) acc = fma((double)qpad[k], Ri[k], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            if (lane == 0) qv[i] = (T)acc;
        }
    } else {
        for (int i = tid; i < Dp; i += LUT_THREADS) qv[i] = (T)qpad[i];
    }
    __syncthreads();
    if (q_rot) for (int i = tid; i < Dp; i += LUT_THREADS) q_rot[(size_t)q * Dp + i] = (double)qv[i];

    // ---- dists[c][m] = sum_k (centers[c][m*dpb+k] - q[m*dpb+k])^2 ------------------------------
    for (int e = tid; e < 16 * M; e += LUT_THREADS) {
        const int c = e / M, m = e - c * M;
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
    }
    __syncthreads();

    // ---- shift ---------------------------------------------------------------------------------
    if (signd) {
        if (tid == 0) {
            const T sum = np_pairwise_sum<T>(dists, 16 * M);
            const T mean = (T)((double)sum / (double)(16 * M));     // dtype(f64(sum)/intp(count))
            s_shift = mul_rn(mean, (T)0.6931471806);                // python float is "weak": multiply in T
        }
    } else {
        T mn = (T)INFINITY;
        for (int e = tid; e < 16 * M; e += LUT_THREADS) mn = fmin(mn, dists[e]);
        red[tid] = mn;
        __syncthreads();
        for (int o = LUT_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] = fmin(red[tid], red[tid + o]);
            __syncthreads();
        }
        if (tid == 0) s_shift = red[0];
    }
    __syncthreads();
    const T shift = s_shift;

    // ---- dists -= shift ; max ------------------------------------------------------------------
    T mx = -(T)INFINITY;
    for (int e = tid; e < 16 * M; e += LUT_THREADS) {
        const T v = dists[e] - shift;
        dists[e] = v;
        mx = fmax(mx, v);
    }
    red[tid] = mx;
    __syncthreads();
    for (int o = LUT_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] = fmax(red[tid], red[tid + o]);
        __syncthreads();
    }
    const double amax = (double)red[0];

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
    const size_t smem = tsz * (16 * (size_t)M + Dp + LUT_THREADS) + sizeof(float) * Dpad;
    TKB_REQUIRE(smem <= 200 * 1024, "dimension too large for the LUT kernel");
    if (R) {
        TKB_CUDA(cudaFuncSetAttribute(lut_build_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lut_build_kernel<double><<<Q, LUT_THREADS, smem, st>>>(queries, d, normalize, q_out, centers, Dp, dpb, R, Dpad,
                                                                sqrt_n_blocks, log_n_blocks, signd, tables, q_rot, shift, scale);
    } else {
        TKB_CUDA(cudaFuncSetAttribute(lut_build_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lut_build_kernel<float><<<Q, LUT_THREADS, smem, st>>>(queries, d, normalize, q_out, centers, Dp, dpb, R, Dpad,
                                                               sqrt_n_blocks, log_n_blocks, signd, tables, q_rot, shift, scale);
    }
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
