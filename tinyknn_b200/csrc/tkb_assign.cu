// tkb_assign.cu -- nearest coarse centroids of every database row (build time).
//
// Replaces the arithmetic of `knn_brute(data, self.all_centers, k=n_probes)` in IVF.build (ref: tinyknn/ivf.py:84-86,
// tinyknn/utils.py:66-86): part[i][c] = (|x_i|^2 + |c|^2) - (2 x_i) . c, the k smallest per row. This is the one dense
// rows x centroids contraction of the reference. It runs on the CUDA cores in f32 (f64 for f64 rows) on purpose: the
// reference's `2 * Xchunk @ Y.T` is an sgemm whose every output is a sequential FMA chain over the dimension starting from
// 0 (verified against numpy 2.3.5 / OpenBLAS 0.3.30 for d <= 384), and a register-tiled GEMM whose accumulators walk k in
// ascending order computes exactly that chain -- a bf16/tf32 tensor-core product would not, and the assignment (hence every
// inverted list) would differ near ties. |x|^2 and |c|^2 are taken from the caller when given (numpy's einsum sums long
// rows in SIMD lanes, an order this kernel does not try to guess), else computed here left to right.
//
// CTA tile: 128 rows x 64 centroids, 256 threads, 8 x 4 accumulators per thread, k-tiles of 16 staged k-major in shared
// memory. Each thread keeps the best KSEL (1 or 2) candidates of its rows among its centroid columns; the 16 threads that
// share a row merge with shuffles at the end. Order of the output: ascending (part, centroid index).
#include "tkb_common.cuh"

namespace tkb {

namespace {

constexpr int AS_BM = 128, AS_BN = 64, AS_BK = 16, AS_THREADS = 256;

template <typename T> __device__ __forceinline__ T a_fma(T a, T b, T c);
template <> __device__ __forceinline__ float a_fma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double a_fma<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <typename T> __device__ __forceinline__ T a_add(T a, T b);
template <> __device__ __forceinline__ float a_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double a_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T a_mul(T a, T b);
template <> __device__ __forceinline__ float a_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double a_mul<double>(double a, double b) { return __dmul_rn(a, b); }

template <typename T> __device__ __forceinline__ T a_inf();
template <> __device__ __forceinline__ float a_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double a_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

// (v, i) < (w, j) in the output order
template <typename T> __device__ __forceinline__ bool a_less(T v, int i, T w, int j) { return v < w || (v == w && i < j); }

template <typename T, int KSEL>
__device__ __forceinline__ void a_insert(T (&bv)[KSEL], int (&bi)[KSEL], T v, int i)
{
    if (KSEL == 1) {
        if (a_less(v, i, bv[0], bi[0])) { bv[0] = v; bi[0] = i; }
    } else {
        if (a_less(v, i, bv[KSEL - 1], bi[KSEL - 1])) {
            bv[KSEL - 1] = v; bi[KSEL - 1] = i;
#pragma unroll
            for (int s = KSEL - 1; s > 0; s--)
                if (a_less(bv[s], bi[s], bv[s - 1], bi[s - 1])) {
                    const T tv = bv[s]; bv[s] = bv[s - 1]; bv[s - 1] = tv;
                    const int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
                }
        }
    }
}

template <typename T>
__global__ void row_sqnorm_kernel(const T *__restrict__ x, int64_t n, int d, T *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T *r = x + i * d;
    T s = a_mul<T>(r[0], r[0]);
    for (int k = 1; k < d; k++) s = a_add<T>(s, a_mul<T>(r[k], r[k]));
    out[i] = s;
}

template <typename T, int KSEL>
__global__ void __launch_bounds__(AS_THREADS)
assign_kernel(const T *__restrict__ rows, int64_t n, int d, const T *__restrict__ centers, int C,
              const T *__restrict__ xnorm, const T *__restrict__ cnorm, int32_t *__restrict__ nearest /* [n][KSEL] */)
{
    __shared__ __align__(16) T As[AS_BK][AS_BM];        // (2 x) k-major
    __shared__ __align__(16) T Bs[AS_BK][AS_BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * AS_BM;
    T bv[8][KSEL];
    int bi[8][KSEL];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int s = 0; s < KSEL; s++) { bv[i][s] = a_inf<T>(); bi[i][s] = 0x7fffffff; }
    T xn[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { const int64_t r = row0 + ty * 8 + i; xn[i] = r < n ? xnorm[r] : (T)0; }

    for (int c0 = 0; c0 < C; c0 += AS_BN) {
        T acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = (T)0;
        for (int k0 = 0; k0 < d; k0 += AS_BK) {
            __syncthreads();
            {   // rows tile: thread -> (row = tid / 2, 8 consecutive k)
                const int r = tid >> 1, kq = (tid & 1) * 8;
                const int64_t gr = row0 + r;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int k = k0 + kq + u;
                    const T v = (gr < n && k < d) ? rows[gr * d + k] : (T)0;
                    As[kq + u][r] = a_add<T>(v, v);                       // 2 * X is exact
                }
                // centroid tile: thread -> (centroid = tid / 4, 4 consecutive k)
                const int c = tid >> 2, kc = (tid & 3) * 4;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = k0 + kc + u;
                    Bs[kc + u][c] = (c0 + c < C && k < d) ? centers[(size_t)(c0 + c) * d + k] : (T)0;
                }
            }
            __syncthreads();
            const int kn = (d - k0 < AS_BK) ? (d - k0) : AS_BK;
            for (int k = 0; k < kn; k++) {                                 // ascending k: the sgemm kernel's FMA chain
                T a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = As[k][ty * 8 + i];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = a_fma<T>(a[i], b[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c0 + tx * 4 + j;
            if (c >= C) continue;
            const T cn = cnorm[c];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const T part = a_add<T>(a_add<T>(xn[i], cn), -acc[i][j]);    // utils.py:83
                a_insert<T, KSEL>(bv[i], bi[i], part, c);
            }
        }
    }
    // merge over the 16 threads (tx) that share the rows ty*8..ty*8+7: they are 16 consecutive lanes
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            T ov[KSEL];
            int oi[KSEL];
#pragma unroll
            for (int s = 0; s < KSEL; s++) {
                ov[s] = __shfl_xor_sync(0xffffffffu, bv[i][s], o);
                oi[s] = __shfl_xor_sync(0xffffffffu, bi[i][s], o);
            }
#pragma unroll
            for (int s = 0; s < KSEL; s++) a_insert<T, KSEL>(bv[i], bi[i], ov[s], oi[s]);
        }
        const int64_t r = row0 + ty * 8 + i;
        if (tx == 0 && r < n) {
#pragma unroll
            for (int s = 0; s < KSEL; s++) nearest[r * KSEL + s] = bi[i][s];
        }
    }
}

template <typename T>
int assign_t(const T *rows, int64_t n, int d, const T *centers, int C, const T *xnorm, const T *cnorm, int k,
             int32_t *nearest, T *scratch, cudaStream_t st)
{
    if (!xnorm) {
        row_sqnorm_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows, n, d, scratch);
        TKB_LAUNCH_CHECK();
        xnorm = scratch;
    }
    if (!cnorm) {
        row_sqnorm_kernel<T><<<(unsigned)((C + 255) / 256), 256, 0, st>>>(centers, C, d, scratch + n);
        TKB_LAUNCH_CHECK();
        cnorm = scratch + n;
    }
    const unsigned grid = (unsigned)((n + AS_BM - 1) / AS_BM);
    if (k == 1) assign_kernel<T, 1><<<grid, AS_THREADS, 0, st>>>(rows, n, d, centers, C, xnorm, cnorm, nearest);
    else        assign_kernel<T, 2><<<grid, AS_THREADS, 0, st>>>(rows, n, d, centers, C, xnorm, cnorm, nearest);
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}


}  // namespace

int launch_assign(const void *rows, int dtype, int64_t n, int d, const void *centers, int C, const void *xnorm,
                  const void *cnorm, int k, int32_t *nearest, void *scratch, int64_t scratch_bytes, cudaStream_t st)
{
    TKB_REQUIRE(dtype == TKB_DTYPE_F32 || dtype == TKB_DTYPE_F64, "dtype must be TKB_DTYPE_F32 or TKB_DTYPE_F64");
    TKB_REQUIRE(n >= 0 && d > 0 && C > 0, "bad extent");
    TKB_REQUIRE(k == 1 || k == 2, "k (lists per point) must be 1 or 2");
    TKB_REQUIRE(k <= C, "k exceeds the number of centroids");
    if (n == 0) return TKB_OK;
    TKB_REQUIRE(rows && centers && nearest, "null pointer");
    TKB_REQUIRE((n + AS_BM - 1) / AS_BM <= 0x7fffffffLL, "too many rows for one launch");
    const int64_t esz = dtype == TKB_DTYPE_F64 ? 8 : 4;
    TKB_REQUIRE((xnorm && cnorm) || (scratch && scratch_bytes >= esz * (n + C)), "scratch too small ((n + C) elements) for the norms");
    if (dtype == TKB_DTYPE_F64)
        return assign_t<double>((const double *)rows, n, d, (const double *)centers, C, (const double *)xnorm,
                                (const double *)cnorm, k, nearest, (double *)scratch, st);
    return assign_t<float>((const float *)rows, n, d, (const float *)centers, C, (const float *)xnorm, (const float *)cnorm,
                           k, nearest, (float *)scratch, st);
}

}  // namespace tkb
