// stream_probe.cu -- how fast can ONE SM pull an L2-resident key into shared memory / registers?
// (sizes the key stream of the small-batch blind-rotation kernel)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_probe tools/stream_probe.cu && tools/stream_probe
// modes: TMA bulk copies of `chunk` bytes with `depth` in flight (one issuing thread); LDG.128 by 256 threads;
// cp.async 16 B by 256 threads.  Reports bytes per clock per SM at grid = 1 and grid = 148.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) tma_stream(const char* src, size_t bytes, int chunk, int depth, unsigned long long* cyc, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* buf = smem + 1024;
    const int n = (int)(bytes / chunk);
    if (threadIdx.x == 0) {
        for (int d = 0; d < depth; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[d])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    float acc = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth && i < n; ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[i])), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)i * chunk)),
                         "l"(src + (size_t)i * chunk), "r"(chunk), "r"(s32(&bars[i])) : "memory");
        }
        for (int i = 0; i < n; ++i) {
            const int d = i % depth;
            const uint32_t par = (uint32_t)((i / depth) & 1);
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&bars[d])), "r"(par) : "memory");
            acc += *reinterpret_cast<float*>(buf + (size_t)d * chunk);
            const int nx = i + depth;
            if (nx < n) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[d])), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)d * chunk)),
                             "l"(src + (size_t)nx * chunk), "r"(chunk), "r"(s32(&bars[d])) : "memory");
            }
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; if (acc == 123.f) sink[0] = acc; }
}

__global__ void __launch_bounds__(256, 1) ldg_stream(const char* src, size_t bytes, int unroll_dummy, unsigned long long* cyc, float* sink) {
    const double2* p = reinterpret_cast<const double2*>(src);
    const size_t n = bytes / 16;
    const unsigned long long t0 = clock64();
    double acc = 0;
    for (size_t i = threadIdx.x; i + 7 * 256 < n; i += 8 * 256) {
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "l"(p + i + k * 256));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.0) sink[0] = (float)acc;
}

__global__ void __launch_bounds__(256, 1) cpasync_stream(const char* src, size_t bytes, int depth, unsigned long long* cyc, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    // groups of 256 threads x 6 x 16 B = 24 KB; `depth` groups in flight
    const int per = 6, group_bytes = 256 * per * 16;
    const int n = (int)(bytes / group_bytes);
    const unsigned long long t0 = clock64();
    float acc = 0.f;
    auto issue = [&](int g) {
        unsigned char* dst = smem + (size_t)(g % depth) * group_bytes;
        const char* s = src + (size_t)g * group_bytes;
#pragma unroll
        for (int k = 0; k < per; ++k) {
            const int off = (k * 256 + threadIdx.x) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + off)), "l"(s + off) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int g = 0; g < depth - 1 && g < n; ++g) issue(g);
    for (int g = 0; g < n; ++g) {
        if (g + depth - 1 < n) issue(g + depth - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (depth == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else asm volatile("cp.async.wait_group 5;" ::: "memory");
        acc += *reinterpret_cast<float*>(smem + (size_t)(g % depth) * group_bytes + threadIdx.x * 16);
        __syncthreads();
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.f) sink[0] = acc;
}

int main() {
    const size_t bytes = 72ull << 20;   // the multi-bit Fourier key is 72.9 MB: L2 resident
    char* src; unsigned long long* cyc; float* sink;
    cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    unsigned long long h[148];
    auto report = [&](const char* name, int grid) {
        cudaDeviceSynchronize();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("%-52s grid %3d  ERROR %s\n", name, grid, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-52s grid %3d  %8.1f B/clk/SM  (%llu clk)\n", name, grid, (double)bytes / (double)mx, mx);
    };
    cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(cpasync_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep)   // first pass warms L2
    for (int grid : {1, 148}) {
        char name[96];
        for (int chunk : {6144, 24576, 49152}) for (int depth : {2, 4, 6}) {
            if ((size_t)chunk * depth + 1024 > 200 * 1024) continue;
            tma_stream<<<grid, 256, 1024 + (size_t)chunk * depth>>>(src, bytes, chunk, depth, cyc, sink);
            snprintf(name, sizeof name, "TMA bulk %5d B x %d in flight", chunk, depth);
            if (rep) report(name, grid); else cudaDeviceSynchronize();
        }
        ldg_stream<<<grid, 256>>>(src, bytes, 0, cyc, sink);
        if (rep) report("LDG.128 x8 per thread, 256 threads", grid); else cudaDeviceSynchronize();
        for (int depth : {2, 4, 6}) {
            cpasync_stream<<<grid, 256, (size_t)depth * 24576>>>(src, bytes, depth, cyc, sink);
            snprintf(name, sizeof name, "cp.async 16 B, 24 KB groups x %d in flight", depth);
            if (rep) report(name, grid); else cudaDeviceSynchronize();
        }
    }
    return 0;
}
