// dfma_probe.cu -- what FP64 issue rate does sm_100a sustain on butterfly-shaped code?  (the bound of the
// blind-rotation kernels.)  Per-SM DFMA warp-instructions per clock, measured with clock64() inside one resident CTA per
// SM, for: (0) x = fma(x, a, b) with two operands shared by every chain (what fhe_b200_probe_fp64 measures: the nominal
// peak of 2 warp-instructions per clock per SM); (1) x = fma(x, y_c, z_c), three distinct register pairs per instruction;
// (2) the fused 6-instruction radix-2 butterfly on 16 complex points in registers with register twiddles (dit16's inner
// shape); (3) the same with the twiddles read from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_probe tools/dfma_probe.cu && tools/dfma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int MODE>
__global__ void probe(double* out, long long* cycles, double a, double b, int warps) {
    __shared__ double2 tw[64];
    if (threadIdx.x < 64) tw[threadIdx.x] = make_double2(0.999 + 1e-6 * threadIdx.x, 0.01 + 1e-6 * threadIdx.x);
    __syncthreads();
    double re[16], im[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[i] = threadIdx.x + i; im[i] = threadIdx.x - i; }
    double y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = a + 1e-9 * i; z[i] = b + 1e-12 * i; }
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { re[i] = fma(re[i], a, b); im[i] = fma(im[i], a, b); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { re[i] = fma(re[i], y[i & 7], z[(i + 3) & 7]); im[i] = fma(im[i], y[(i + 1) & 7], z[(i + 5) & 7]); }
        } else {
#pragma unroll
            for (int half = 1; half <= 8; half <<= 1) {
#pragma unroll
                for (int base = 0; base < 16; base += 2 * half) {
#pragma unroll
                    for (int j = 0; j < half; ++j) {
                        const int p = base + j, q = base + j + half;
                        double wr, wi;
                        if (MODE == 2) { wr = y[(j + half) & 7]; wi = z[(j + half) & 7]; }
                        else { const double2 w = tw[(j + half + (threadIdx.x & 31)) & 63]; wr = w.x; wi = w.y; }
                        const double ar = re[p], ai = im[p], br = re[q], bi = im[q];
                        const double y0r = fma(-bi, wi, fma(br, wr, ar));
                        const double y0i = fma(bi, wr, fma(br, wi, ai));
                        re[p] = y0r; im[p] = y0i;
                        re[q] = fma(2.0, ar, -y0r);
                        im[q] = fma(2.0, ai, -y0i);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { re[i] *= 0.25; im[i] *= 0.25; }   // keep the values bounded (counted below)
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += re[i] + im[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int threads) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* d; long long* c;
    cudaMalloc(&d, 8);
    cudaMalloc(&c, 8 * p.multiProcessorCount);
    const int grid = p.multiProcessorCount;
    probe<MODE><<<grid, threads>>>(d, c, 0.999999, 1e-9, threads / 32);
    probe<MODE><<<grid, threads>>>(d, c, 0.999999, 1e-9, threads / 32);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, c, 8 * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += h[i];
    avg /= grid;
    const double rate = (double)instr_per_iter * ITERS * (threads / 32) / avg;
    printf("%-64s %2d warps/SM: %5.3f FP64 warp-instr/clk/SM = %5.1f %% of 2.0\n", name, threads / 32, rate, 50.0 * rate);
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int threads : {256, 512}) {
        run<0>("fma(x, a, b): two shared operands", 32, threads);
        run<1>("fma(x, y_c, z_c): three distinct operand pairs", 32, threads);
        run<2>("6-FMA radix-2 butterflies, 16 points, register twiddles", 32 * 6 + 32, threads);
        run<3>("6-FMA radix-2 butterflies, 16 points, twiddles from shared memory", 32 * 6 + 32, threads);
    }
    return 0;
}
