// tools/ubench.cu -- instruction-throughput calibration for the scan kernel design (B200, sm_100a).
// Prints lane-ops per clock per SM for the integer / half ops the PQ scan can be built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define CHAINS 8

enum Op { PRMT, LOP3, IADD3, SHF, VIADDMIN32, VIADDMIN16X2, VIMAX16X2, IMAD, HFMA2SAT, HADD2, LDS8, LDS128, PRMT_VIADD, PRMT_HFMA, IMADHI, PRMT_IMADHI, PRMT_IMAD, PRMT_SHF, NOPS };
static const char *names[] = {"PRMT", "LOP3", "IADD3", "SHF", "VIADDMNMX.s32", "VIADDMNMX.s16x2", "VIADDMNMX.s16x2.RELU", "IMAD",
                              "HFMA2.SAT", "HADD2", "LDS.U8 (row bcast)", "LDS.128 (bcast)", "PRMT+VIADDMNMX16x2 mix", "PRMT+HFMA2.SAT mix",
                              "IMAD.HI (mul.hi.u32 reg)", "PRMT+IMAD.HI mix", "PRMT+IMAD mix", "PRMT+SHF mix"};

template <int OP>
__global__ void __launch_bounds__(512) bench(uint32_t *out, uint32_t seed, long long *cycles)
{
    __shared__ __align__(16) uint8_t lut[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = (uint8_t)(i * 7 + seed);
    __syncthreads();
    uint32_t a[CHAINS], b = seed * 2654435761u + threadIdx.x, c = seed ^ 0x5a5a5a5a;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = b + k * 0x01010101u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) {
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[k]) : "r"(b));
            if (OP == VIADDMIN32) a[k] = __viaddmin_s32(a[k], b, c);
            if (OP == VIADDMIN16X2) a[k] = __viaddmin_s16x2(a[k], b, c);
            if (OP == VIMAX16X2) a[k] = __viaddmin_s16x2_relu(a[k], b, c);
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            if (OP == HFMA2SAT) asm volatile("fma.rn.sat.f16x2 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
            if (OP == LDS8) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(lut) + ((a[k] & 15) + 16 * k))); a[k] += v; }
            if (OP == LDS128) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(lut) + 16 * ((it + k) & 63))); a[k] ^= v.x ^ v.w; }
            if (OP == PRMT_VIADD) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(c), "r"(a[k])); a[k] = __viaddmin_s16x2(a[k], t, c); }
            if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
            if (OP == PRMT_IMADHI) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(c), "r"(a[k])); asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[k]) : "r"(t), "r"(b)); }
            if (OP == PRMT_IMAD) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(c), "r"(a[k])); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(t), "r"(b)); }
            if (OP == PRMT_SHF) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(c), "r"(a[k])); asm volatile("shf.r.wrap.b32 %0, %1, %2, 16;" : "=r"(a[k]) : "r"(t), "r"(b)); }
            if (OP == PRMT_HFMA) { uint32_t t; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(c), "r"(a[k])); asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(t), "r"(c)); }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, uint32_t *out, long long *cyc)
{
    const int blocks = sms * 4, threads = 512;        // 2048 threads/SM resident (4 x 512)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<blocks, threads>>>(out, 1, cyc);
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads>>>(out, 2, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double ops_per_thread = (double)ITERS * CHAINS * ((OP == PRMT_VIADD || OP == PRMT_HFMA || OP == PRMT_IMADHI || OP == PRMT_IMAD || OP == PRMT_SHF) ? 2 : 1);
    const double per_sm_clk = ops_per_thread * 2048.0 / (double)h[0];
    const double total = ops_per_thread * blocks * threads;
    printf("%-28s %8.1f lane-ops/clk/SM   %8.2f Tlane-ops/s   (%.3f ms, %lld cycles)\n", names[OP], per_sm_clk,
           total / (ms * 1e-3) / 1e12, ms, h[0]);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 512 * 4); cudaMalloc(&cyc, 8 * 1024 * 8);
    int s = p.multiProcessorCount;
    run<PRMT>(s, out, cyc); run<LOP3>(s, out, cyc); run<IADD3>(s, out, cyc); run<SHF>(s, out, cyc);
    run<VIADDMIN32>(s, out, cyc); run<VIADDMIN16X2>(s, out, cyc); run<VIMAX16X2>(s, out, cyc); run<IMAD>(s, out, cyc);
    run<HFMA2SAT>(s, out, cyc); run<HADD2>(s, out, cyc); run<LDS8>(s, out, cyc); run<LDS128>(s, out, cyc);
    run<PRMT_VIADD>(s, out, cyc); run<PRMT_HFMA>(s, out, cyc);
    run<IMADHI>(s, out, cyc); run<PRMT_IMADHI>(s, out, cyc); run<PRMT_IMAD>(s, out, cyc); run<PRMT_SHF>(s, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
