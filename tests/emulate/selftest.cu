// selftest.cu -- kernels that check the EMULATOR itself (tests/emulate; TEST INFRASTRUCTURE ONLY). Translated and compiled like
// the library's sources; each entry point returns what the kernel observed so that the Python test can compare it with the
// documented CUDA semantics, or provokes a condition the emulator must report (divergent barrier, shared-memory overrun).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void collectives_kernel(uint32_t *out)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *o = out + (size_t)threadIdx.x * 8;
    o[0] = __ballot_sync(0xffffffffu, (lane % 3) == 0);
    o[1] = __shfl_sync(0xffffffffu, 100 * warp + lane, 5);
    o[2] = __shfl_xor_sync(0xffffffffu, lane, 16);
    o[3] = __shfl_up_sync(0xffffffffu, lane, 3);
    o[4] = __shfl_down_sync(0xffffffffu, lane, 30);
    o[5] = __any_sync(0xffffffffu, warp == 1 && lane == 31);
    o[6] = __all_sync(0xffffffffu, lane < 32);
    int v = (int)lane;                                              // inclusive prefix sum, the way the kernels write it
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v += u; }
    o[7] = (uint32_t)v;
}

__global__ void barrier_kernel(int *out, int rounds)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    int *sm = reinterpret_cast<int *>(sm_raw);
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    for (int r = 0; r < rounds; r++) {
        sm[threadIdx.x] = (int)threadIdx.x + r;
        __syncthreads();
        const int other = sm[(threadIdx.x + 1) % blockDim.x];       // written by another thread before the barrier
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) atomicAdd(&total, other);
    }
    const int any = __syncthreads_or(threadIdx.x == 77 && rounds == 3);
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = total; out[2 * blockIdx.x + 1] = any; }
}

// half of warp 0 waits at a warp collective, the rest of the block at __syncthreads: nobody can make progress
__global__ void divergent_kernel(int *out)
{
    if (threadIdx.x < 16) out[0] = (int)__ballot_sync(0xffffffffu, 1);
    else __syncthreads();
}

__global__ void early_exit_kernel(int *out, int n)
{
    if ((int)threadIdx.x >= n) return;                              // exited threads do not hold up the barrier
    __shared__ int s[256];
    s[threadIdx.x] = 1;
    __syncthreads();
    int sum = 0;
    for (int i = 0; i < n; i++) sum += s[i];
    out[threadIdx.x] = sum;
}

__global__ void overrun_kernel(int bytes)
{
    extern __shared__ unsigned char sm_o[];
    if (threadIdx.x == 0) sm_o[bytes] = 1;                          // one byte past the end
}

__global__ void simd_kernel(const uint32_t *a, const uint32_t *b, const uint32_t *c, uint32_t *out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t *o = out + (size_t)i * 12;
    o[0] = __vadd2(a[i], b[i]);
    o[1] = __vimin3_s16x2(a[i], b[i], c[i]);
    o[2] = __vimax3_s16x2(a[i], b[i], c[i]);
    o[3] = __vimin3_u16x2(a[i], b[i], c[i]);
    o[4] = __viaddmax_s16x2(a[i], b[i], c[i]);
    o[5] = __viaddmin_u16x2(a[i], b[i], c[i]);
    o[6] = __vmins4(a[i], b[i]);
    o[7] = __vminu4(a[i], b[i]);
    o[8] = __vcmplts4(a[i], b[i]);
    o[9] = __vcmpltu4(a[i], b[i]);
    o[10] = (uint32_t)__viaddmin_s32((int)a[i], (int)b[i], (int)c[i]);
    o[11] = emu::prmt(a[i], b[i], c[i]);
}

// a deliberate race: every thread reads its neighbour's slot without a barrier after the writes
__global__ void racy_kernel(int *out)
{
    __shared__ int s[64];
    s[threadIdx.x] = (int)threadIdx.x + 1;
    out[threadIdx.x] = s[(threadIdx.x + 1) % 64];
}

// a 16-byte vector load from an address the caller chooses (TKB_EMU_UBSAN=1 builds abort when it is misaligned)
__global__ void vector_load_kernel(const unsigned char *p, uint32_t *out)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    out[0] = v.x + v.y + v.z + v.w;
}

}  // namespace

extern "C" {
int emu_selftest_vector_load(const unsigned char *p, uint32_t *out)
{
    vector_load_kernel<<<1, 1, 0, 0>>>(p, out);
    return cudaGetLastError();
}
int emu_selftest_racy(int *out)
{
    racy_kernel<<<1, 64, 0, 0>>>(out);
    return cudaGetLastError();
}
int emu_selftest_collectives(uint32_t *out, int threads)
{
    collectives_kernel<<<1, threads, 0, 0>>>(out);
    return cudaGetLastError();
}
int emu_selftest_barrier(int *out, int blocks, int threads, int rounds)
{
    barrier_kernel<<<blocks, threads, threads * sizeof(int), 0>>>(out, rounds);
    return cudaGetLastError();
}
int emu_selftest_divergent(int *out)
{
    divergent_kernel<<<1, 64, 0, 0>>>(out);
    return cudaGetLastError();
}
int emu_selftest_early_exit(int *out, int threads, int n)
{
    early_exit_kernel<<<1, threads, 0, 0>>>(out, n);
    return cudaGetLastError();
}
int emu_selftest_overrun(int bytes)
{
    overrun_kernel<<<1, 32, bytes, 0>>>(bytes);
    return cudaGetLastError();
}
int emu_selftest_simd(const uint32_t *a, const uint32_t *b, const uint32_t *c, uint32_t *out, int n)
{
    simd_kernel<<<(n + 127) / 128, 128, 0, 0>>>(a, b, c, out, n);
    return cudaGetLastError();
}
}
