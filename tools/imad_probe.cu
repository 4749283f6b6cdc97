// imad_probe.cu -- integer-multiply pipe rates that bound the Philox mask generator:
// 32x32->64 products as IMAD.WIDE.U32 vs separate IMAD.HI.U32 + IMAD (lo), per SM per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_probe tools/imad_probe.cu && tools/imad_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CHAINS = 8, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = seed + threadIdx.x * 977u + c * 131u + blockIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) {          // one wide multiply, both halves used
                uint64_t p = (uint64_t)x[c] * 0xD2511F53u;
                x[c] = (uint32_t)p ^ (uint32_t)(p >> 32);
            } else if (MODE == 1) {   // hi and lo as two multiplies (different constants so they cannot merge)
                uint32_t hi = __umulhi(x[c], 0xD2511F53u);
                uint32_t lo = x[c] * 0xCD9E8D57u;
                x[c] = hi ^ lo;
            } else if (MODE == 2) {   // lo only
                x[c] = x[c] * 0xCD9E8D57u + 12345u;
            } else if (MODE == 3) {   // hi only
                x[c] = __umulhi(x[c], 0xD2511F53u) + 77u;
            } else {                  // LOP3 only (alu pipe reference)
                x[c] = (x[c] ^ 0x9E3779B9u) & (x[c] >> 3 | 0xBB67AE85u);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c];
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <int MODE>
void run(const char* name, int mults_per_step) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    uint32_t* d;
    cudaMalloc(&d, 4);
    const int grid = p.multiProcessorCount * 8;
    probe<MODE><<<grid, 256>>>(d, 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MODE><<<grid, 256>>>(d, 2);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double steps = (double)grid * 256 * CHAINS * ITERS;
    const double per_sm_clk = steps / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3);
    printf("%-34s %8.3f ms  %6.1f chain-steps/clk/SM  (%5.1f multiply instr/clk/SM at max clock %d MHz)\n", name, ms,
           per_sm_clk, per_sm_clk * mults_per_step, clk_khz / 1000);
    cudaFree(d);
}

int main() {
    run<0>("IMAD.WIDE.U32 + LOP3", 1);
    run<1>("IMAD.HI.U32 + IMAD(lo) + LOP3", 2);
    run<2>("IMAD (lo) only", 1);
    run<3>("IMAD.HI.U32 + IADD", 1);
    run<4>("LOP3/SHF only", 0);
    return 0;
}
