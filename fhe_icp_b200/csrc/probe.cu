// probe.cu -- live measurement of the FP64 FMA peak of this GPU, the denominator of the
// blind-rotation roofline (SURVEY.md section 7.2 asks for it next to the HBM copy peak).
#include "kernels.h"

namespace fhe {

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

cudaError_t probe_fp64(int sm_count, double* tflops, cudaStream_t s) {
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sm_count * 8, iters = 1 << 14;
    dfma_probe_kernel<<<blocks, 256, 0, s>>>(d, 256, 0.999999, 1e-9);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0, s);
        dfma_probe_kernel<<<blocks, 256, 0, s>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, s);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        count_launch();
    }
    count_launch();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return e;
}

}  // namespace fhe
