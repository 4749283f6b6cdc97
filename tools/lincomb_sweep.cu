// tools/lincomb_sweep.cu -- development microbenchmark (not product code): times candidate
// implementations of the encrypted-dot-product kernel on one resident ciphertext set.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/lincomb_sweep tools/lincomb_sweep.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct __align__(16) u64x2 { uint64_t x, y; };
__device__ __forceinline__ u64x2 ldv(const uint64_t* p) {
    u64x2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stv(uint64_t* p, u64x2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// ---------------------------------------------------------------- V1: flat index, register prefetch
template <int U, int THREADS>
__global__ void __launch_bounds__(THREADS) v1_kernel(const uint64_t* __restrict__ ct, int d, int64_t stride,
                                                     int64_t total_vecs, const int64_t* __restrict__ W,
                                                     uint64_t* __restrict__ out) {
    extern __shared__ int64_t sW[];
    for (int i = threadIdx.x; i < d; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (g >= total_vecs) return;
    const int vecs = (int)(stride / 2);
    const int64_t b = g / vecs;
    const int v = (int)(g - b * vecs);
    const uint64_t* p = ct + (size_t)b * d * stride + 2 * v;
    uint64_t a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    u64x2 cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = ldv(p + (size_t)u * stride);
    for (int j = 0; j < d; j += U) {
        if (j + U < d) {
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = ldv(p + (size_t)(j + U + u) * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t w = (uint64_t)sW[j + u];
            a0x += w * cur[u].x; a0y += w * cur[u].y;
            a1x += cur[u].x; a1y += cur[u].y;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    }
    uint64_t* o = out + (size_t)b * 2 * stride + 2 * v;
    stv(o, u64x2{a0x, a0y});
    stv(o + stride, u64x2{a1x, a1y});
}

// ---------------------------------------------------------------- V2: TMA bulk copies into a smem ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// CTA = one (row b, column split) tile; consumers = CONS threads (one 16 B vector each); one producer warp.
template <int STAGES, int CONS>
__global__ void __launch_bounds__(CONS + 32) v2_kernel(const uint64_t* __restrict__ ct, int d, int64_t stride, int splits,
                                                       int vecs_per_split, const int64_t* __restrict__ W,
                                                       uint64_t* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tile_bytes = vecs_per_split * 16;
    const int tile_stride = (tile_bytes + 127) & ~127;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    int64_t* sW = reinterpret_cast<int64_t*>(empty + STAGES);
    unsigned char* ring = smem + ((2 * STAGES * 8 + d * 8 + 127) & ~127);
    for (int i = threadIdx.x; i < d; i += blockDim.x) sW[i] = W[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t b = blockIdx.x / splits;
    const int sp = blockIdx.x - (int)(b * splits);
    const uint64_t* base = ct + (size_t)b * d * stride + (size_t)sp * vecs_per_split * 2;
    if (threadIdx.x >= CONS) {
        if (threadIdx.x == CONS) {  // producer: one elected thread
            for (int j = 0; j < d; ++j) {
                const int s = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], tile_bytes);
                tma_load_1d(ring + (size_t)s * tile_stride, base + (size_t)j * stride, tile_bytes, &full[s]);
            }
        }
        return;
    }
    const int v = threadIdx.x;
    const bool active = v < vecs_per_split;
    uint64_t a0x = 0, a0y = 0, a1x = 0, a1y = 0;
    for (int j = 0; j < d; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(&full[s], ph);
        if (active) {
            u64x2 x = *reinterpret_cast<const u64x2*>(ring + (size_t)s * tile_stride + v * 16);
            uint64_t w = (uint64_t)sW[j];
            a0x += w * x.x; a0y += w * x.y;
            a1x += x.x; a1y += x.y;
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (active) {
        uint64_t* o = out + (size_t)b * 2 * stride + (size_t)sp * vecs_per_split * 2 + 2 * v;
        stv(o, u64x2{a0x, a0y});
        stv(o + stride, u64x2{a1x, a1y});
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1000, d = 128;
    const int64_t stride = argc > 2 ? atoi(argv[2]) : 1424;
    const size_t words = (size_t)B * d * stride;
    uint64_t *ct, *out, *ref;
    int64_t* W;
    CK(cudaMalloc(&ct, words * 8));
    CK(cudaMalloc(&out, (size_t)B * 2 * stride * 8));
    CK(cudaMalloc(&ref, (size_t)B * 2 * stride * 8));
    CK(cudaMalloc(&W, d * 8));
    std::vector<uint64_t> h(words);
    uint64_t s = 88172645463325252ULL;
    for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = s; }
    CK(cudaMemcpy(ct, h.data(), words * 8, cudaMemcpyHostToDevice));
    std::vector<int64_t> hw(d);
    for (int i = 0; i < d; ++i) hw[i] = (i * 37 % 256) - 128;
    CK(cudaMemcpy(W, hw.data(), d * 8, cudaMemcpyHostToDevice));
    const double bytes = (double)(d + 2) * stride * 8 * B;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto launch, uint64_t* dst) {
        for (int i = 0; i < 3; ++i) launch(dst);
        CK(cudaDeviceSynchronize());
        float best = 1e9, tot = 0;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0); launch(dst); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best; tot += ms;
        }
        printf("%-28s mean %.4f ms  best %.4f ms  -> %.1f GB/s (mean) %.1f GB/s (best)\n", name, tot / 10, best,
               bytes / (tot / 10 * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9);
    };
    const int64_t total_vecs = (int64_t)B * (stride / 2);
    auto check = [&](const char* name) {
        std::vector<uint64_t> a((size_t)B * 2 * stride), r((size_t)B * 2 * stride);
        CK(cudaMemcpy(a.data(), out, a.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(r.data(), ref, r.size() * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); ++i) bad += a[i] != r[i];
        printf("   check %-20s mismatches: %zu\n", name, bad);
    };
#define V1(U, T) timeit("v1 U=" #U " T=" #T, [&](uint64_t* dst) { \
        v1_kernel<U, T><<<(unsigned)((total_vecs + T - 1) / T), T, d * 8>>>(ct, d, stride, total_vecs, W, dst); }, out)
    timeit("v1 U=4 T=256 (ref)", [&](uint64_t* dst) {
        v1_kernel<4, 256><<<(unsigned)((total_vecs + 255) / 256), 256, d * 8>>>(ct, d, stride, total_vecs, W, dst); }, ref);
    V1(2, 256); V1(4, 256); V1(8, 256); V1(16, 256); V1(4, 128); V1(8, 128); V1(4, 512); V1(8, 512);
    check("v1");
#define V2(S, C, SPL) do { \
        int vps = (int)(stride / 2 / SPL); \
        if ((stride / 2) % SPL == 0 && vps <= C) { \
            size_t sm = ((2 * S * 8 + d * 8 + 127) & ~127) + (size_t)S * ((vps * 16 + 127) & ~127); \
            CK(cudaFuncSetAttribute(v2_kernel<S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            CK(cudaMemset(out, 0, (size_t)B * 2 * stride * 8)); \
            timeit("v2 S=" #S " C=" #C " split=" #SPL, [&](uint64_t* dst) { \
                v2_kernel<S, C><<<B * SPL, C + 32, sm>>>(ct, d, stride, SPL, vps, W, dst); }, out); \
            check("v2 S=" #S " split=" #SPL); \
        } } while (0)
    V2(8, 384, 2); V2(16, 384, 2); V2(32, 384, 2);
    V2(8, 192, 4); V2(16, 192, 4); V2(32, 192, 4);
    V2(8, 96, 8); V2(16, 96, 8); V2(32, 96, 8);
    V2(8, 736, 1); V2(16, 736, 1);
    CK(cudaDeviceSynchronize());
    return 0;
}
