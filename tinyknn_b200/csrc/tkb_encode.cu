// tkb_encode.cu -- batched PQ encoder: vectors -> packed 4-bit codes in the reference's chunk layout.
//
// Replaces FastPQ.transform (ref: tinyknn/fast_pq.py:147-184) and what it calls: pad2 (ref: tinyknn/utils.py:14-19),
// the optional rotation `data @ self.R.T`, per block `knn_brute(col, code, 1)` (ref: tinyknn/utils.py:66-86: the expansion
// |x|^2 + |c|^2 - 2 x.c and the index of its minimum) and transform_data (ref: tinyknn/_transform.py:4-77, the nibble
// packing pinned by tests/test_transform.py:80-101). Not on the query path: it is what builds the codes the scan reads
// (SURVEY.md 8(f)1), and what IVF.build runs once per point.
//
// The arithmetic mirrors numpy operation by operation, so the codes equal the reference's bit for bit whenever the minimum
// is unique (verified against numpy 2.3.5 / OpenBLAS 0.3.30 in the build container, tests/golden/encode.npz):
//   rotation      f64; out[j] = fma chain over k = 0..Dpad-1 starting from 0 (what the BLAS dgemm kernel computes)
//   |x|^2, |c|^2  np.einsum('ij,ij->i'): separately rounded products added left to right, no FMA; |c|^2 in f32 (the
//                 codebook is f32), |x|^2 in the compute type T (f64 when rotated or the rows are f64, else f32)
//   2 x.c         `2 * X @ Y.T` = (2X) @ Y.T: fma chain over the block's dims starting from 0, in T
//   part          (|x|^2 + |c|^2) - 2 x.c, two roundings in T; np.argpartition(part, 1)[:, :1] = the first minimum
//
// One CTA encodes tiles of TV output positions (TV/16 chunks). Rotated case: the tile's rows are staged k-major in shared
// memory as f64, R (transposed, k-major) next to it, every thread owns a 4 vectors x 4 outputs register tile per 64 output
// columns; the rotated tile stays in shared memory for the nearest-of-16 search. Output bytes are staged and written as one
// contiguous span of the codes array.
#include "tkb_common.cuh"

namespace tkb {

namespace {

constexpr int ENC_THREADS = 256;
constexpr int ENC_TV = 64;              // vectors per tile in the rotated case

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T fma_rn(T a, T b, T c);
template <> __device__ __forceinline__ float fma_rn<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fma_rn<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

struct EncArgs {
    const void *rows;            // [n_rows][d], f32 or f64
    int rows_f64;
    int64_t n_rows;
    int d;
    const int64_t *row_index;    // [n_out] or null
    int64_t n_out;               // multiple of 16
    const float *centers;        // [16][Dp]
    const float *cnorm;          // [M][16]
    int Dp, dpb, M;
    const double *R;             // [Dp][Dpad] or null
    int Dpad;
    uint8_t *codes;              // [n_out/16][M*8] bytes
    int TV;                      // vectors per tile
};

__device__ __forceinline__ int64_t enc_row(const EncArgs &a, int64_t pos)
{
    if (pos >= a.n_out) return -1;
    int64_t r = a.row_index ? a.row_index[pos] : pos;
    if (r < 0 || r >= a.n_rows) return -1;         // padding position: the zero vector (ref: fast_pq.py:165 pad2)
    return r;
}

__device__ __forceinline__ double enc_load(const EncArgs &a, int64_t row, int k)
{
    if (row < 0 || k >= a.d) return 0.0;
    return a.rows_f64 ? reinterpret_cast<const double *>(a.rows)[row * a.d + k]
                      : (double)reinterpret_cast<const float *>(a.rows)[row * a.d + k];
}

// nearest-of-16 for the TV x M/2 (vector, pair of blocks) items of a tile; x: the tile in shared memory, row stride xs
template <typename T>
__device__ __forceinline__ void encode_tile(const EncArgs &a, const T *x, int xs, const float *cs, const float *cn,
                                            uint8_t *stage, int tv)
{
    const int Ph = a.M >> 1, dpb = a.dpb;
    for (int i = threadIdx.x; i < tv * Ph; i += ENC_THREADS) {
        const int v = i / Ph, p = i - v * Ph;
        uint32_t byte = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = 2 * p + h;
            const T *xv = x + (size_t)v * xs + m * dpb;
            T xn = mul_rn<T>(xv[0], xv[0]);
            for (int k = 1; k < dpb; k++) xn = add_rn<T>(xn, mul_rn<T>(xv[k], xv[k]));
            int best = 0;
            T bv = 0;
            for (int c = 0; c < 16; c++) {
                const float *cc = cs + (size_t)c * a.Dp + m * dpb;
                T dot2 = 0;
                for (int k = 0; k < dpb; k++) dot2 = fma_rn<T>(add_rn<T>(xv[k], xv[k]), (T)cc[k], dot2);
                const T part = add_rn<T>(add_rn<T>(xn, (T)cn[m * 16 + c]), -dot2);
                if (c == 0 || part < bv) { bv = part; best = c; }
            }
            byte |= (uint32_t)best << (4 * h);
        }
        stage[((size_t)(v >> 4) * Ph + p) * 16 + (v & 15)] = (uint8_t)byte;      // _transform.py:53-77
    }
}

__device__ __forceinline__ void store_tile(const EncArgs &a, const uint8_t *stage, int64_t pos0, int tv)
{
    const int64_t chunk0 = pos0 >> 4;
    const int n16 = (tv >> 4) * a.M * 8 / 16;                       // uint4 per tile
    uint4 *dst = reinterpret_cast<uint4 *>(a.codes + chunk0 * a.M * 8);
    const uint4 *src = reinterpret_cast<const uint4 *>(stage);
    for (int i = threadIdx.x; i < n16; i += ENC_THREADS) dst[i] = src[i];
}

// ---- unrotated: the (padded) rows are the vectors ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ENC_THREADS) encode_plain_kernel(EncArgs a)
{
    extern __shared__ __align__(16) unsigned char enc_sm[];
    const int xs = a.Dp + 2;
    T *x = reinterpret_cast<T *>(enc_sm);                                        // [TV][Dp + 2]
    float *cs = reinterpret_cast<float *>(x + (size_t)a.TV * xs);                // [16][Dp]
    float *cn = cs + 16 * a.Dp;                                                  // [M][16]
    uint8_t *stage = reinterpret_cast<uint8_t *>(cn + a.M * 16);                 // [TV/16][M/2][16]
    for (int i = threadIdx.x; i < 16 * a.Dp; i += ENC_THREADS) cs[i] = a.centers[i];
    for (int i = threadIdx.x; i < 16 * a.M; i += ENC_THREADS) cn[i] = a.cnorm[i];
    const int64_t n_tiles = (a.n_out + a.TV - 1) / a.TV;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pos0 = tile * a.TV;
        const int tv = (int)((a.n_out - pos0 < a.TV) ? (a.n_out - pos0) : a.TV);
        __syncthreads();
        for (int v = threadIdx.x >> 5; v < tv; v += ENC_THREADS / 32) {
            const int64_t row = enc_row(a, pos0 + v);
            for (int k = threadIdx.x & 31; k < a.Dp; k += 32) x[(size_t)v * xs + k] = (T)enc_load(a, row, k);
        }
        __syncthreads();
        encode_tile<T>(a, x, xs, cs, cn, stage, tv);
        __syncthreads();
        store_tile(a, stage, pos0, tv);
    }
}

// ---- rotated: x' = x @ R.T in f64, then the same search -----------------------------------------------------------
template <int NH>                                                               // NH = ceil(Dp / 64) in {1, 2}
__global__ void __launch_bounds__(ENC_THREADS) encode_rot_kernel(EncArgs a)
{
    extern __shared__ __align__(16) unsigned char enc_sm[];
    constexpr int DpP = 64 * NH;                       // padded output columns
    constexpr int KT = NH == 1 ? 128 : 64;             // k-tile: KT * DpP * 8 = 64 KB of R
    constexpr int XS = ENC_TV + 2;                     // row stride of the k-major x tile (doubles)
    constexpr int RS = DpP + 2;                        // row stride of the rotated tile
    double *Rt = reinterpret_cast<double *>(enc_sm);                             // [KT][DpP]   R transposed, k-major
    double *xk = Rt + (size_t)KT * DpP;                                          // [KT][XS]    tile rows, k-major
    double *xr = xk + (size_t)KT * XS;                                           // [TV][RS]    rotated tile
    float *cs = reinterpret_cast<float *>(xr + (size_t)ENC_TV * RS);             // [16][Dp]
    float *cn = cs + 16 * a.Dp;
    uint8_t *stage = reinterpret_cast<uint8_t *>(cn + a.M * 16);
    for (int i = threadIdx.x; i < 16 * a.Dp; i += ENC_THREADS) cs[i] = a.centers[i];
    for (int i = threadIdx.x; i < 16 * a.M; i += ENC_THREADS) cn[i] = a.cnorm[i];
    const int og = threadIdx.x & 15, vg = threadIdx.x >> 4;                     // columns og*4.. (+64h), vectors vg*4..
    const int n_kt = (a.Dpad + KT - 1) / KT;
    const int64_t n_tiles = (a.n_out + ENC_TV - 1) / ENC_TV;
    bool r_loaded = false;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t pos0 = tile * ENC_TV;
        const int tv = (int)((a.n_out - pos0 < ENC_TV) ? (a.n_out - pos0) : ENC_TV);
        double acc[NH][4][4];
#pragma unroll
        for (int h = 0; h < NH; h++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[h][i][j] = 0.0;
        for (int kt = 0; kt < n_kt; kt++) {
            const int k0 = kt * KT, kn = (a.Dpad - k0 < KT) ? (a.Dpad - k0) : KT;
            __syncthreads();
            if (!(r_loaded && n_kt == 1)) {                                      // a single k-tile of R stays resident
                for (int i = threadIdx.x; i < kn * DpP; i += ENC_THREADS) {
                    const int k = i / DpP, j = i - k * DpP;
                    Rt[i] = j < a.Dp ? a.R[(size_t)j * a.Dpad + k0 + k] : 0.0;
                }
                r_loaded = true;
            }
            for (int v = threadIdx.x >> 5; v < ENC_TV; v += ENC_THREADS / 32) {
                const int64_t row = v < tv ? enc_row(a, pos0 + v) : -1;
                for (int k = threadIdx.x & 31; k < kn; k += 32) xk[(size_t)k * XS + v] = enc_load(a, row, k0 + k);
            }
            __syncthreads();
            for (int k = 0; k < kn; k++) {                                       // k ascending: the dgemm kernel's fma chain
                const double2 xa = *reinterpret_cast<const double2 *>(xk + (size_t)k * XS + vg * 4);
                const double2 xb = *reinterpret_cast<const double2 *>(xk + (size_t)k * XS + vg * 4 + 2);
                const double xv[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const double2 ra = *reinterpret_cast<const double2 *>(Rt + (size_t)k * DpP + 64 * h + og * 4);
                    const double2 rb = *reinterpret_cast<const double2 *>(Rt + (size_t)k * DpP + 64 * h + og * 4 + 2);
                    const double rv[4] = {ra.x, ra.y, rb.x, rb.y};
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[h][i][j] = __fma_rn(xv[i], rv[j], acc[h][i][j]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < NH; h++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) xr[(size_t)(vg * 4 + i) * RS + 64 * h + og * 4 + j] = acc[h][i][j];
        __syncthreads();
        encode_tile<double>(a, xr, RS, cs, cn, stage, tv);
        __syncthreads();
        store_tile(a, stage, pos0, tv);
    }
}

size_t enc_tail_bytes(int Dp, int M, int TV) { return (size_t)16 * Dp * 4 + (size_t)M * 16 * 4 + (size_t)(TV / 16) * (M / 2) * 16; }

}  // namespace

int launch_encode(const void *rows, int rows_dtype, int64_t n_rows, int d, const int64_t *row_index, int64_t n_out,
                  const float *centers, const float *cnorm, int Dp, int dpb, const double *R, int Dpad, uint64_t *codes,
                  cudaStream_t st)
{
    TKB_REQUIRE(rows_dtype == TKB_DTYPE_F32 || rows_dtype == TKB_DTYPE_F64, "rows_dtype must be TKB_DTYPE_F32 or TKB_DTYPE_F64");
    TKB_REQUIRE(n_rows >= 0 && n_out >= 0 && n_out % 16 == 0, "n_out must be a non-negative multiple of 16");
    TKB_REQUIRE(d > 0 && dpb > 0 && Dp > 0 && Dp % (2 * dpb) == 0, "Dp must be a positive multiple of 2 * dims_per_block");
    TKB_REQUIRE(Dpad >= d, "Dpad (padded dimension) must be >= d");
    TKB_REQUIRE(R || Dp == Dpad, "without a rotation the codebook dimension must equal the padded dimension");
    if (n_out == 0) return TKB_OK;
    TKB_REQUIRE((rows || n_rows == 0) && centers && cnorm && codes, "null pointer");
    TKB_REQUIRE((uintptr_t)codes % 16 == 0, "codes must be 16-byte aligned");
    const int M = Dp / dpb;
    EncArgs a{rows, rows_dtype == TKB_DTYPE_F64, n_rows, d, row_index, n_out, centers, cnorm, Dp, dpb, M, R, Dpad,
              reinterpret_cast<uint8_t *>(codes), ENC_TV};
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (R) {
        TKB_REQUIRE(Dp <= 128, "rotated encoder supports rotate_dim <= 128");
        const int NH = Dp <= 64 ? 1 : 2, DpP = 64 * NH, KT = NH == 1 ? 128 : 64;
        const size_t smem = ((size_t)KT * DpP + (size_t)KT * (ENC_TV + 2) + (size_t)ENC_TV * (DpP + 2)) * 8 + enc_tail_bytes(Dp, M, ENC_TV);
        const int64_t n_tiles = (n_out + ENC_TV - 1) / ENC_TV;
        const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);       // one persistent CTA per SM (R stays in shared memory)
        if (NH == 1) {
            TKB_CUDA(cudaFuncSetAttribute(encode_rot_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            encode_rot_kernel<1><<<grid, ENC_THREADS, smem, st>>>(a);
        } else {
            TKB_CUDA(cudaFuncSetAttribute(encode_rot_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            encode_rot_kernel<2><<<grid, ENC_THREADS, smem, st>>>(a);
        }
        TKB_LAUNCH_CHECK();
        return TKB_OK;
    }
    const bool f64 = rows_dtype == TKB_DTYPE_F64;
    const size_t esz = f64 ? 8 : 4;
    int TV = 64;
    while (TV > 16 && (size_t)TV * (Dp + 2) * esz + enc_tail_bytes(Dp, M, TV) > 96 * 1024) TV >>= 1;
    const size_t smem = (size_t)TV * (Dp + 2) * esz + enc_tail_bytes(Dp, M, TV);
    TKB_REQUIRE(smem <= 200 * 1024, "dimension too large for the encoder");
    a.TV = TV;
    const int64_t n_tiles = (n_out + TV - 1) / TV;
    const int64_t want = (int64_t)sms * 8;
    const unsigned grid = (unsigned)(n_tiles < want ? n_tiles : want);
    if (f64) {
        TKB_CUDA(cudaFuncSetAttribute(encode_plain_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_plain_kernel<double><<<grid, ENC_THREADS, smem, st>>>(a);
    } else {
        TKB_CUDA(cudaFuncSetAttribute(encode_plain_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_plain_kernel<float><<<grid, ENC_THREADS, smem, st>>>(a);
    }
    TKB_LAUNCH_CHECK();
    return TKB_OK;
}

}  // namespace tkb
