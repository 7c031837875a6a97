// assign_emu.cpp -- runs the SOURCE of tkb_assign.cu's kernels on the CPU through cuda_emu.h (TEST INFRASTRUCTURE ONLY).
//   g++ -std=c++20 -O1 -pthread -ffp-contract=off -mfma -DTKB_EMULATE -I tests/emulate -shared -fPIC ...
#define TKB_EMULATE 1
#include "../../tinyknn_b200/csrc/tkb_assign.cu"

using namespace tkb;

template <typename T, int KSEL>
static void run(const T *rows, int64_t n, int d, const T *centers, int C, const T *xnorm, const T *cnorm, int32_t *nearest)
{
    const unsigned grid = (unsigned)((n + AS_BM - 1) / AS_BM);
    emu_launch(grid, AS_THREADS, [&] { assign_kernel<T, KSEL>(rows, n, d, centers, C, xnorm, cnorm, nearest); });
}

extern "C" void emu_assign_f32(const float *rows, int64_t n, int d, const float *centers, int C, const float *xnorm,
                               const float *cnorm, int k, int32_t *nearest)
{
    if (k == 1) run<float, 1>(rows, n, d, centers, C, xnorm, cnorm, nearest);
    else        run<float, 2>(rows, n, d, centers, C, xnorm, cnorm, nearest);
}

extern "C" void emu_assign_f64(const double *rows, int64_t n, int d, const double *centers, int C, const double *xnorm,
                               const double *cnorm, int k, int32_t *nearest)
{
    if (k == 1) run<double, 1>(rows, n, d, centers, C, xnorm, cnorm, nearest);
    else        run<double, 2>(rows, n, d, centers, C, xnorm, cnorm, nearest);
}

extern "C" void emu_row_sqnorm_f32(const float *x, int64_t n, int d, float *out)
{
    emu_launch((unsigned)((n + 255) / 256), 256, [&] { row_sqnorm_kernel<float>(x, n, d, out); });
}
