// pbs.cu -- programmable bootstrapping: mod-switch, batched blind rotation (n CMuxes, each a
// GGSW external product over the negacyclic f64 FFT of fft.cuh) and sample extraction
// (TFHE, Chillotti et al. 2020).  Not exercised by the reference's
// compiled circuit (it has no table lookup); named by BASELINE.json's north star.
//
// Mapping (N = 2048, M = 1024 = 32 x 32):
//   * one WARP per accumulator polynomial: warp t of a ciphertext owns ACC_t (16 KB of u64 in
//     shared memory, private to that warp), produces the digits of (X^a - 1) * ACC_t, runs
//     their forward FFTs in registers, and later the inverse FFT of output column t;
//   * PBS_NCT ciphertexts per CTA walk the n key elements in lock step, so the 64 KB of
//     BSK_i they all read stays hot in L1 (v1; TMA-multicast staging is the next step);
//   * per iteration the only cross-warp traffic is the Fourier-domain digits (16 KB per
//     polynomial and level), exchanged through shared memory between two named barriers.
// Roofline: FP64 pipe and shared-memory bandwidth bind (about 176 MFLOP per PBS at the stated
// set); the 48.6 MB Fourier key is L2-resident, so HBM only sees it once per launch.
#include <cstdlib>
#include <mutex>

#include <cstdio>

#include "common.cuh"
#include "fft.cuh"
#include "kernels.h"

namespace fhe {

using nfft::cplx;

constexpr int PBS_N = nfft::NPOLY;
constexpr int PBS_M = nfft::M;
#ifndef PBS_PREFETCH
#define PBS_PREFETCH 4
#endif
constexpr int PBS_TILE = nfft::TILE_ELEMS;  // padded transpose tile / twiddle table, in complex elements

// ------------------------------------------------------------------------------- twiddle tables
constexpr int PBS_OMEGA = 128;  // two-level table of omega = exp(2*pi*i/4096): [0,64) omega^x, [64,128) omega^(64*y)
static cplx* g_tw = nullptr;  // [PBS_TILE] row-padded inter-pass twiddles + [PBS_OMEGA] (per device)
static int g_tw_device = -1;
static std::mutex g_tw_mu;

static cudaError_t get_tables(const cplx** tw) {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (g_tw == nullptr || g_tw_device != dev) {
        static cplx h[PBS_TILE + PBS_OMEGA];
        nfft::fill_twiddle_table(h);
        for (int x = 0; x < 64; ++x) {
            const long double two_pi = 6.283185307179586476925286766559005768L;
            h[PBS_TILE + x].x = (double)cosl(two_pi * x / 4096.0L);
            h[PBS_TILE + x].y = (double)sinl(two_pi * x / 4096.0L);
            h[PBS_TILE + 64 + x].x = (double)cosl(two_pi * (64 * x) / 4096.0L);
            h[PBS_TILE + 64 + x].y = (double)sinl(two_pi * (64 * x) / 4096.0L);
        }
        cplx* d = nullptr;
        if ((e = cudaMalloc(&d, sizeof(h))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        g_tw = d;  // (tables of a previous device are intentionally kept alive)
        g_tw_device = dev;
    }
    *tw = g_tw;
    return cudaSuccess;
}

// Named barrier `id` (1..8) with an IMMEDIATE barrier number: with a register operand ptxas reserves all 16 hardware
// barriers for the CTA ("used 16 barriers"), and barriers are an occupancy limit of the SM.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    switch (id) {
        case 1: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
        case 2: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
        case 3: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
        case 4: asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory"); break;
        case 5: asm volatile("bar.sync 5, %0;" ::"r"(nthreads) : "memory"); break;
        case 6: asm volatile("bar.sync 6, %0;" ::"r"(nthreads) : "memory"); break;
        case 7: asm volatile("bar.sync 7, %0;" ::"r"(nthreads) : "memory"); break;
        default: asm volatile("bar.sync 8, %0;" ::"r"(nthreads) : "memory"); break;
    }
}

__device__ __forceinline__ uint64_t f64_to_torus(double x) {
    // x mod 2^64, rounded to the nearest integer
    const double r = rint(x * 0x1p-64);
    const double y = fma(-r, 0x1p64, x);
    return (uint64_t)__double2ll_rn(y);
}

// ------------------------------------------------------------------------------- tensor memory
// TMEM (256 KB per SM, 128 lanes x 512 32-bit columns) is used here as LANE-PRIVATE storage for the
// accumulator polynomials: lane j1 of a warp owns the 64 coefficients {j1 + 32*j2} U {+1024} of its
// polynomial in both FFT directions, i.e. 128 columns of its own TMEM lane.  That takes the 16 KB
// per polynomial out of shared memory (which then fits 4 ciphertexts = 8 warps per SM instead of
// 2 = 4 warps) and moves the accumulator read-modify-write off the shared-memory pipe.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 consecutive 32-bit columns of this thread's own TMEM lane <-> 32 registers
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                 : "memory");
}
// 16-column variants (8 u64 coefficients per call): smaller register footprint per chunk
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------- key -> Fourier
// One warp per key polynomial: u64 torus coefficients (as signed) -> 1024 complex bins,
// natural order (same layout as the oracle's orc_bsk_to_fourier).
constexpr int B2F_WARPS = 4;

__global__ void __launch_bounds__(B2F_WARPS * 32)
bsk_to_fourier_kernel(const uint64_t* __restrict__ bsk, int64_t polys, const cplx* __restrict__ g_twf,
                      double* __restrict__ bskf) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* twf = reinterpret_cast<cplx*>(smem_raw);
    cplx* bufs = twf + PBS_TILE;
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) twf[i] = g_twf[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * B2F_WARPS + warp;
    if (q >= polys) return;
    cplx* buf = bufs + warp * PBS_TILE;
    const uint64_t* src = bsk + (size_t)q * PBS_N;
    double re[32], im[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        re[j2] = (double)(int64_t)src[lane + 32 * j2];
        im[j2] = (double)(int64_t)src[lane + 32 * j2 + PBS_M];
    }
    nfft::fwd_phase1(re, im, twf, buf, lane);
    __syncwarp();
    nfft::fwd_phase2(re, im, buf, lane);
    cplx* dst = reinterpret_cast<cplx*>(bskf) + (size_t)q * PBS_M;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        cplx v;
        v.x = re[p];
        v.y = im[p];
        dst[nfft::brev5(p) * 32 + lane] = v;
    }
}

cudaError_t launch_bsk_to_fourier(const fhe_b200_pbs_params& p, const uint64_t* d_bsk, double* d_bskf,
                                  cudaStream_t s) {
    const cplx* twf;
    cudaError_t e = get_tables(&twf);
    if (e != cudaSuccess) return e;
    const int64_t polys = (int64_t)p.n * (p.k + 1) * p.l_pbs * (p.k + 1);
    const size_t smem = sizeof(cplx) * PBS_TILE * (1 + B2F_WARPS);
    e = cudaFuncSetAttribute(bsk_to_fourier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bsk_to_fourier_kernel<<<(unsigned)((polys + B2F_WARPS - 1) / B2F_WARPS), B2F_WARPS * 32, smem, s>>>(
        d_bsk, polys, twf, d_bskf);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------- blind rotation
// Shared memory per CTA:
//   tw                            16 KB    inter-pass twiddles (swizzled, both directions)
//   BSK stage (L == 1 only)       64 KB    BSK_i, filled by one TMA bulk copy per iteration while
//                                          the warps run their inverse / forward FFTs
//   per ciphertext:  ACC          (K+1) x 16 KB (u64 coefficients)
//                    tile         (K+1) x 16 KB (per-warp transpose tile; with L == 1 it also
//                                                carries that warp's Fourier digits)
//                    F            (K+1) x L x 16 KB, only when L > 1
//                    a_tilde      n+1 u16  (mod-switched mask, padded)
template <int K, int L>
struct PbsSmem {
    static constexpr int POLYS = K + 1;
    static constexpr bool STAGE = (L == 1);
    static constexpr size_t tw_bytes = (size_t)PBS_TILE * 16;
    static constexpr size_t stage_bytes = STAGE ? (size_t)POLYS * L * POLYS * PBS_M * 16 : 0;
    static constexpr size_t bar_bytes = 128;
    static constexpr size_t head_bytes = tw_bytes + stage_bytes + bar_bytes;
    static constexpr size_t acc_bytes = (size_t)POLYS * PBS_N * 8;
    static constexpr size_t tile_bytes = (size_t)POLYS * PBS_TILE * 16;
    static constexpr size_t f_bytes = L > 1 ? (size_t)POLYS * L * PBS_M * 16 : 0;
    __host__ __device__ static size_t per_ct(int n) {
        return acc_bytes + tile_bytes + f_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127);
    }
    static size_t total(int n, int nct) { return head_bytes + (size_t)nct * per_ct(n); }
};

template <int K, int L, int NCT>
__global__ void __launch_bounds__(NCT*(K + 1) * 32, 1)
pbs_kernel(const cplx* __restrict__ bskf, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
           const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index, const cplx* __restrict__ g_tw,
           uint64_t* __restrict__ out) {
    using S = PbsSmem<K, L>;
    constexpr int POLYS = K + 1;
    constexpr bool STAGE = S::STAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    cplx* stage = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::tw_bytes + S::stage_bytes);
    uint64_t* bar_empty = bar_full + 1;
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) tw[i] = g_tw[i];
    if (STAGE && threadIdx.x == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_empty, NCT * POLYS);
        mbar_fence_init();
    }
    constexpr uint32_t STAGE_BYTES = (uint32_t)S::stage_bytes;
    const size_t bsk_elems = (size_t)POLYS * L * POLYS * PBS_M;  // complex elements of one BSK_i

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = warp / POLYS;        // ciphertext slot within the CTA
    const int t = warp - ctl * POLYS;    // polynomial owned by this warp
    const int64_t b = (int64_t)blockIdx.x * NCT + ctl;
    const size_t per_ct = S::per_ct(n);
    unsigned char* base = smem_raw + S::head_bytes + (size_t)ctl * per_ct;
    uint64_t* acc_all = reinterpret_cast<uint64_t*>(base);
    cplx* tile_all = reinterpret_cast<cplx*>(base + S::acc_bytes);
    cplx* f_all = reinterpret_cast<cplx*>(base + S::acc_bytes + S::tile_bytes);
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(base + S::acc_bytes + S::tile_bytes + S::f_bytes);
    uint64_t* acc = acc_all + (size_t)t * PBS_N;
    cplx* tile = tile_all + (size_t)t * PBS_TILE;
    const bool live = b < B;  // dead slots still take part in every barrier
    const int bar_id = 1 + ctl, bar_n = POLYS * 32;

    // ---- prologue: mod-switch the mask, ACC = X^(-b~) * (0,...,0,LUT)
    const uint64_t* ct = in + (size_t)(live ? b : 0) * (n + 1);
    for (int i = t * 32 + lane; i <= n; i += POLYS * 32)
        a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);   // round(a * 2N / 2^64), 2N = 4096
    named_bar_sync(bar_id, bar_n);
    {
        const uint64_t* lut = luts + (size_t)(lut_index && live ? lut_index[b] : 0) * PBS_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
        for (int x = lane; x < PBS_N; x += 32) {
            uint64_t v = 0;
            if (t == K) {
                const int src = (x - rot) & 4095;
                v = lut[src & 2047];
                if (src & 2048) v = 0 - v;
            }
            acc[x] = v;
        }
    }
    __syncthreads();  // twiddle table, mbarriers and all accumulators in place
    // BSK_i is streamed through the single stage buffer by thread 0: BSK_0 now, BSK_{i+1} as soon as
    // every warp of the CTA has finished reading BSK_i (checked between its own inverse-FFT passes,
    // when the other warps have normally arrived already, so the wait does not spin).
    if (STAGE && threadIdx.x == 0) {
        mbar_expect_tx(bar_full, STAGE_BYTES);
        tma_load_1d(stage, bskf, STAGE_BYTES, bar_full);
    }

    const uint64_t Bm = (1ULL << beta) - 1, half = 1ULL << (beta - 1);
    const int tot = L * beta;
    uint64_t offs = 0;
    for (int lev = 0; lev < L; ++lev) offs |= half << (beta * lev);
    const uint64_t rnd = 1ULL << (63 - tot);

    double re[32], im[32];
    for (int i = 0; i < n; ++i) {
        const int at = a_tilde[i];
        if (!STAGE && at == 0) continue;  // X^0 - 1 = 0 (staged kernel: every warp must consume BSK_i)
        const cplx* bk = STAGE ? stage : bskf + (size_t)i * bsk_elems;
        // ---- digits of (X^at - 1) * ACC_t, one forward FFT per level
#pragma unroll 1
        for (int lev = 0; lev < L; ++lev) {
            const int sh = beta * (L - 1 - lev);
#pragma unroll
            for (int j2 = 0; j2 < 32; ++j2) {
                const int x = lane + 32 * j2;
                const int s0 = (x - at) & 4095, s1 = (x + PBS_M - at) & 4095;
                uint64_t r0 = acc[s0 & 2047], r1 = acc[s1 & 2047];
                if (s0 & 2048) r0 = 0 - r0;
                if (s1 & 2048) r1 = 0 - r1;
                const uint64_t d0 = r0 - acc[x], d1 = r1 - acc[x + PBS_M];
                const uint64_t u0 = ((d0 + rnd) >> (64 - tot)) + offs, u1 = ((d1 + rnd) >> (64 - tot)) + offs;
                re[j2] = (double)((int32_t)((u0 >> sh) & Bm) - (int32_t)half);
                im[j2] = (double)((int32_t)((u1 >> sh) & Bm) - (int32_t)half);
            }
            nfft::fwd_phase1(re, im, tw, tile, lane);
            __syncwarp();
            nfft::fwd_phase2(re, im, tile, lane);
            if (L > 1) {
                cplx* f = f_all + (size_t)(t * L + lev) * PBS_M;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    cplx v;
                    v.x = re[p];
                    v.y = im[p];
                    f[nfft::brev5(p) * 32 + lane] = v;
                }
            }
        }
        if (L == 1) {
            // publish this warp's bins through its own tile (all lanes are done reading it)
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                cplx v;
                v.x = re[p];
                v.y = im[p];
                tile[nfft::brev5(p) * 32 + lane] = v;
            }
        }
        named_bar_sync(bar_id, bar_n);  // (A) every polynomial's Fourier digits are visible
        if (STAGE) mbar_wait(bar_full, (uint32_t)(i & 1));  // BSK_i has landed in shared memory
        // ---- output column t:  out = sum_{t',lev} F[t'][lev] * BSK_i[t'][lev][t]
        if (L == 1) {
            const cplx* bown = bk + ((size_t)(t * L) * POLYS + t) * PBS_M;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const cplx g = bown[nfft::brev5(p) * 32 + lane];
                const double a = re[p], c = im[p];
                re[p] = a * g.x - c * g.y;
                im[p] = a * g.y + c * g.x;
            }
#pragma unroll 1
            for (int tp = 0; tp < POLYS; ++tp) {
                if (tp == t) continue;
                const cplx* f = tile_all + (size_t)tp * PBS_TILE;
                const cplx* bo = bk + ((size_t)(tp * L) * POLYS + t) * PBS_M;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    const cplx v = f[nfft::brev5(p) * 32 + lane];
                    const cplx g = bo[nfft::brev5(p) * 32 + lane];
                    re[p] += v.x * g.x - v.y * g.y;
                    im[p] += v.x * g.y + v.y * g.x;
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < 32; ++p) { re[p] = 0.0; im[p] = 0.0; }
#pragma unroll 1
            for (int q = 0; q < POLYS * L; ++q) {
                const cplx* f = f_all + (size_t)q * PBS_M;
                const cplx* bo = bk + ((size_t)q * POLYS + t) * PBS_M;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    const cplx v = f[nfft::brev5(p) * 32 + lane];
                    const cplx g = bo[nfft::brev5(p) * 32 + lane];
                    re[p] += v.x * g.x - v.y * g.y;
                    im[p] += v.x * g.y + v.y * g.x;
                }
            }
        }
        if (STAGE) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty);  // this warp is done with BSK_i
        }
        named_bar_sync(bar_id, bar_n);  // (B) nobody reads the published digits any more
        // ---- inverse FFT and ACC_t += result
        nfft::inv_phase1(re, im, tw, tile, lane);
        if (STAGE && threadIdx.x == 0 && i + 1 < n) {
            mbar_wait(bar_empty, (uint32_t)(i & 1));
            mbar_expect_tx(bar_full, STAGE_BYTES);
            tma_load_1d(stage, bskf + (size_t)(i + 1) * bsk_elems, STAGE_BYTES, bar_full);
        }
        __syncwarp();
        nfft::inv_phase2(re, im, tile, lane);
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) {
            const int x = lane + 32 * j2;
            acc[x] += f64_to_torus(re[j2]);
            acc[x + PBS_M] += f64_to_torus(im[j2]);
        }
        __syncwarp();  // the next iteration reads rotated (other lanes') coefficients
    }
    // ---- sample extract coefficient 0: LWE under the flattened GLWE key
    __syncwarp();
    if (live) {
        uint64_t* o = out + (size_t)b * ((size_t)K * PBS_N + 1);
        if (t < K) {
            for (int x = lane; x < PBS_N; x += 32) o[(size_t)t * PBS_N + x] = x == 0 ? acc[0] : 0 - acc[PBS_N - x];
        } else if (lane == 0) {
            o[(size_t)K * PBS_N] = acc[0];
        }
    }
}

// ------------------------------------------------------------------------------- blind rotation, TMEM accumulators
// K = 1, L = 1.  Shared memory per CTA: tw 16 KB, BSK stage 64 KB, per ciphertext two 16 KB tiles
// (+ a_tilde): 4 ciphertexts = 8 warps per SM.  The accumulators live in TMEM; a tile carries a
// coefficient-ordered copy of its polynomial from the end of one CMux (ACC update) to the digit
// extraction of the next, where the rotation X^a needs other lanes' coefficients.
struct PbsTmemSmem {
    static constexpr size_t tw_bytes = (size_t)PBS_TILE * 16;
    static constexpr size_t stage_bytes = (size_t)4 * PBS_M * 16;
    static constexpr size_t bar_bytes = 128;
    static constexpr size_t head_bytes = tw_bytes + stage_bytes + bar_bytes;
    static constexpr size_t tile_bytes = (size_t)2 * PBS_TILE * 16;
    __host__ __device__ static size_t per_ct(int n) { return tile_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127); }
    static size_t total(int n, int nct) { return head_bytes + (size_t)nct * per_ct(n); }
};

template <int NCT>
__global__ void __launch_bounds__(NCT * 64, 1)
pbs_kernel_tmem(const cplx* __restrict__ bskf, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
                const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index,
                const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = PbsTmemSmem;
    constexpr int POLYS = 2;
    constexpr uint32_t TMEM_COLS = NCT <= 2 ? 128 : 256;  // 128 columns per warp, two warps per lane quarter
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    cplx* stage = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::tw_bytes + S::stage_bytes);
    uint64_t* bar_empty = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) tw[i] = g_tw[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_empty, NCT * POLYS);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // this warp's TMEM window: its lane quarter (warp % 4), columns (warp / 4) * 128 ..
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    constexpr uint32_t STAGE_BYTES = (uint32_t)S::stage_bytes;
    constexpr size_t bsk_elems = (size_t)4 * PBS_M;

    const int ctl = warp / POLYS, t = warp - ctl * POLYS;
    const int64_t b = (int64_t)blockIdx.x * NCT + ctl;
    unsigned char* base = smem_raw + S::head_bytes + (size_t)ctl * S::per_ct(n);
    cplx* tile_all = reinterpret_cast<cplx*>(base);
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(base + S::tile_bytes);
    cplx* tile = tile_all + (size_t)t * PBS_TILE;
    // Between CMuxes the tile carries the HIGH 32-bit words of ACC_t in coefficient order, followed by
    // their complements (~hi = high word of -A up to one unit of 2^-32): c32[(x - a) mod 4096] is the
    // rotated coefficient of X^a * ACC_t, sign included, in one 4-byte load.  The digit only needs
    // torus bits 63..41, so dropping the low words moves a rounding boundary by < 2^-32 of the torus
    // (noise +0.8 % in variance); the exact 64-bit accumulator stays in TMEM.
    uint32_t* c32 = reinterpret_cast<uint32_t*>(tile);
    const bool live = b < B;
    const int bar_id = 1 + ctl, bar_n = POLYS * 32;

    // ---- prologue: mod-switch the mask, ACC = X^(-b~) * (0, LUT) into TMEM and the tile
    const uint64_t* ct = in + (size_t)(live ? b : 0) * (n + 1);
    for (int i = t * 32 + lane; i <= n; i += POLYS * 32)
        a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);
    named_bar_sync(bar_id, bar_n);
    {
        const uint64_t* lut = luts + (size_t)(lut_index && live ? lut_index[b] : 0) * PBS_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 16 coefficients per chunk: word q = 16c+u <-> x = lane + 32*(q&31) + 1024*(q>>5)
            uint32_t rl[16], rh[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const int x = lane + 32 * (q & 31) + (q >> 5) * PBS_M;
                uint64_t v = 0;
                if (t == 1) {
                    const int src = (x - rot) & 4095;
                    v = lut[src & 2047];
                    if (src & 2048) v = 0 - v;
                }
                rl[u] = (uint32_t)v;
                rh[u] = (uint32_t)(v >> 32);
                c32[x] = rh[u];
                c32[x + PBS_N] = ~rh[u];
            }
            tmem_st_x16(tacc + 16 * c, rl);
            tmem_st_x16(tacc + 64 + 16 * c, rh);
        }
        tmem_wait_st();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_full, STAGE_BYTES);
        tma_load_1d(stage, bskf, STAGE_BYTES, bar_full);
    }

    const uint32_t rnd32 = 1u << (31 - beta);
    const int dshift = 32 - beta;
    double re[32], im[32];
    for (int i = 0; i < n; ++i) {
        const int at = a_tilde[i];
        // ---- digits of (X^at - 1) * ACC_t from the coefficient-ordered copy
        const int base0 = lane - at;  // (x - at) for j2 = 0; the +32*j2 / +1024 offsets are immediates
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            // own (unrotated) high words come from this lane's TMEM columns, the rotated ones from c32
            uint32_t own0[16], own1[16], rot0[16], rot1[16];
            tmem_ld_x16(tacc + 64 + 16 * c, own0);
            tmem_ld_x16(tacc + 96 + 16 * c, own1);
#pragma unroll
            for (int u = 0; u < 16; ++u) {  // all rotated loads of the chunk first, then the arithmetic
                const int j2 = 16 * c + u;
                rot0[u] = c32[(base0 + 32 * j2) & 4095];
                rot1[u] = c32[(base0 + 32 * j2 + PBS_M) & 4095];
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                // one level: the closest multiple of 2^(64-beta), read as a signed beta-bit integer, IS the
                // balanced digit -- an arithmetic shift of the (rounded) high word of the difference
                const int j2 = 16 * c + u;
                re[j2] = (double)((int32_t)(rot0[u] - own0[u] + rnd32) >> dshift);
                im[j2] = (double)((int32_t)(rot1[u] - own1[u] + rnd32) >> dshift);
            }
        }
        __syncwarp();  // every lane has read the ACC copy before the tile becomes the transpose buffer
        nfft::fwd_phase1(re, im, tw, tile, lane);
        __syncwarp();
        nfft::fwd_phase2(re, im, tile, lane);
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            cplx v;
            v.x = re[p];
            v.y = im[p];
            tile[nfft::brev5(p) * 32 + lane] = v;
        }
        named_bar_sync(bar_id, bar_n);                 // (A) both polynomials' Fourier digits are visible
        mbar_wait(bar_full, (uint32_t)(i & 1));        // BSK_i has landed in shared memory
        {
            const cplx* bown = stage + ((size_t)t * POLYS + t) * PBS_M;
            const cplx* f = tile_all + (size_t)(1 - t) * PBS_TILE;
            const cplx* bo = stage + ((size_t)(1 - t) * POLYS + t) * PBS_M;
            constexpr int PF = PBS_PREFETCH;
            cplx g[PF], v[PF], h[PF];
#pragma unroll
            for (int p = 0; p < PF; ++p) {
                const int bin = nfft::brev5(p) * 32 + lane;
                g[p] = bown[bin]; v[p] = f[bin]; h[p] = bo[bin];
            }
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const cplx gc = g[p % PF], vc = v[p % PF], hc = h[p % PF];
                if (p + PF < 32) {
                    const int bin = nfft::brev5(p + PF) * 32 + lane;
                    g[p % PF] = bown[bin]; v[p % PF] = f[bin]; h[p % PF] = bo[bin];
                }
                const double a = re[p], c = im[p];
                re[p] = a * gc.x - c * gc.y + (vc.x * hc.x - vc.y * hc.y);
                im[p] = a * gc.y + c * gc.x + (vc.x * hc.y + vc.y * hc.x);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty);         // this warp is done with BSK_i
        named_bar_sync(bar_id, bar_n);                 // (B) nobody reads the published digits any more
        nfft::inv_phase1(re, im, tw, tile, lane);
        if (threadIdx.x == 0 && i + 1 < n) {
            mbar_wait(bar_empty, (uint32_t)(i & 1));
            mbar_expect_tx(bar_full, STAGE_BYTES);
            tma_load_1d(stage, bskf + (size_t)(i + 1) * bsk_elems, STAGE_BYTES, bar_full);
        }
        __syncwarp();
        nfft::inv_phase2(re, im, tile, lane);
        __syncwarp();  // the tile is free again: it receives the updated accumulator
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t rl[16], rh[16];
            tmem_ld_x16(tacc + 16 * c, rl);
            tmem_ld_x16(tacc + 64 + 16 * c, rh);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const uint64_t v = (((uint64_t)rh[u] << 32) | rl[u]) + f64_to_torus(c < 2 ? re[q] : im[q - 32]);
                rl[u] = (uint32_t)v;
                rh[u] = (uint32_t)(v >> 32);
                const int x = lane + 32 * (q & 31) + (q >> 5) * PBS_M;
                c32[x] = rh[u];
                c32[x + PBS_N] = ~rh[u];
            }
            tmem_st_x16(tacc + 16 * c, rl);
            tmem_st_x16(tacc + 64 + 16 * c, rh);
        }
        tmem_wait_st();
        __syncwarp();
    }
    // ---- sample extract coefficient 0 from the exact accumulators in TMEM:
    //      o[0] = A_0[0], o[N - x] = -A_0[x] (x >= 1), o[N] = A_1[0]
    if (live) {
        uint64_t* o = out + (size_t)b * ((size_t)PBS_N + 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t rl[16], rh[16];
            tmem_ld_x16(tacc + 16 * c, rl);
            tmem_ld_x16(tacc + 64 + 16 * c, rh);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const int x = lane + 32 * (q & 31) + (q >> 5) * PBS_M;
                const uint64_t v = ((uint64_t)rh[u] << 32) | rl[u];
                if (t == 0) {
                    if (x == 0) o[0] = v;
                    else o[PBS_N - x] = 0 - v;
                } else if (x == 0) {
                    o[PBS_N] = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NCT>
static cudaError_t launch_pbs_tmem_t(const fhe_b200_pbs_params& p, const cplx* bskf, const uint64_t* d_in, int64_t B,
                                     const uint64_t* d_luts, const int32_t* d_lut_index, const cplx* tw,
                                     uint64_t* d_out, cudaStream_t s) {
    const size_t smem = PbsTmemSmem::total(p.n, NCT);
    cudaError_t e = cudaFuncSetAttribute(pbs_kernel_tmem<NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((B + NCT - 1) / NCT);
    pbs_kernel_tmem<NCT><<<grid, NCT * 64, smem, s>>>(bskf, d_in, B, p.n, p.beta_pbs, d_luts, d_lut_index, tw, d_out);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------- multi-bit blind rotation (grouping 2)
// One CMux consumes TWO key bits: ACC += sum_g (X^{e_g} - 1) * (G_g [.] ACC), g = 1..3, with
// G1 = GGSW(s_a s_b), G2 = GGSW(s_a(1-s_b)), G3 = GGSW((1-s_a)s_b) and e = (a~_a + a~_b, a~_a, a~_b).
// ACC is decomposed and transformed once per pair and the monomials are applied in the Fourier domain
// (X^e at bin k is rho_k^e = omega^((4k+1)e)), so a pair costs one forward and one inverse FFT per
// polynomial and level instead of two of each, at the price of 3 key elements per pair (L2-resident).
// The key is stored by frequency block k1 ([g][t'][lev][c][32 bins], 6*L KB per block) and streamed
// through a shared-memory ring of two-block slices by TMA while the warps walk k1 = 0..31.
//
// L decomposition levels: the external product is L independent single-level products whose outputs
// add, so a ciphertext is served by 2L warps.  Warp (t, lev) transforms the level-lev digits of ACC_t,
// accumulates sum_t' F[t'][lev] * G[t'][lev][t] and adds its inverse transform into its OWN partial
// accumulator A[t][lev] in TMEM (ACC_t = sum_lev A[t][lev]); the partials of one polynomial sit in the
// same TMEM lane quadrant (warps w and w+4), so every warp can read the full ACC_t for its digits.
#ifndef MB2_L2_K1
#define MB2_L2_K1 2
#endif
#ifndef MB2_L2_SLICES
#define MB2_L2_SLICES 3
#endif
#define MB2_SLICE_K1(L) ((L) == 1 ? 2 : MB2_L2_K1)
#ifndef MB2_L1_SLICES
#define MB2_L1_SLICES 4
#endif
#define MB2_SLICE_COUNT(L) ((L) == 1 ? MB2_L1_SLICES : MB2_L2_SLICES)
constexpr int MB2_LAG = 1;                       // refill a ring slot this many slices after warp 0 left it
template <int L>
struct PbsMb2Smem {
    static constexpr int k1 = MB2_SLICE_K1(L);                           // frequency blocks per slice
    static constexpr int slices = MB2_SLICE_COUNT(L);                    // ring slots
    static constexpr int block_elems = 3 * 2 * L * 2 * 32;              // complex elements per frequency block
    static constexpr int slice_elems = k1 * block_elems;
    static constexpr size_t tw_bytes = (size_t)PBS_TILE * 16;
    static constexpr size_t ring_bytes = (size_t)slices * slice_elems * 16;
    static constexpr size_t omega_bytes = (size_t)PBS_OMEGA * 16;
    static constexpr size_t bar_bytes = 256;
    static constexpr size_t head_bytes = tw_bytes + ring_bytes + omega_bytes + bar_bytes;
    static constexpr size_t tile_bytes = (size_t)2 * L * PBS_TILE * 16;  // one per warp (t, lev)
    __host__ __device__ static size_t per_ct(int n) { return tile_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127); }
    static size_t total(int n, int nct) { return head_bytes + (size_t)nct * per_ct(n); }
};

#ifndef MB2_L1_NCT
#define MB2_L1_NCT 4                 // ciphertexts per CTA of the large-batch kernel (A/B: 2 per CTA, one CTA per SM: 74.5 k PBS/s vs 108 k)
#endif
template <int L, int NCT>
__global__ void __launch_bounds__(NCT * 64 * L, 1)
pbs_kernel_mb2(const cplx* __restrict__ bskf2, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
               const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index,
               const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = PbsMb2Smem<L>;
    static_assert(L == 1 || (L == 2 && NCT == 2), "two levels: 2 ciphertexts x 4 warps, partials paired by TMEM quadrant");
    constexpr int WARPS = NCT * 2 * L;
    constexpr int SLICES = S::slices;
    constexpr uint32_t TMEM_COLS = WARPS <= 4 ? 128 : 256;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    cplx* ring = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes);
    cplx* omega = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes + S::ring_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::tw_bytes + S::ring_bytes + S::omega_bytes);
    uint64_t* bar_empty = bar_full + SLICES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_empty + SLICES);
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) tw[i] = g_tw[i];
    for (int i = threadIdx.x; i < PBS_OMEGA; i += blockDim.x) omega[i] = g_tw[PBS_TILE + i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int q = 0; q < SLICES; ++q) { mbar_init(&bar_full[q], 1); mbar_init(&bar_empty[q], WARPS); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // warp -> (ciphertext, polynomial t, level); its TMEM lane quadrant is warp & 3, column half warp >> 2
    const int ctl = L == 1 ? warp / 2 : (warp & 3) / 2;
    const int t = warp & 1;
    const int lev = L == 1 ? 0 : warp >> 2;
    const uint32_t tquad = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tacc = tquad + (uint32_t)((warp >> 2) * 128);   // this warp's (partial) accumulator
    constexpr uint32_t SLICE_BYTES = (uint32_t)(S::slice_elems * 16);
    constexpr int K1 = S::k1;
    constexpr int SPI = 32 / K1;  // slices per blind-rotation step
    const int pairs = n >> 1;
    const int total_slices = pairs * SPI;

    const int64_t b = (int64_t)blockIdx.x * NCT + ctl;
    unsigned char* base = smem_raw + S::head_bytes + (size_t)ctl * S::per_ct(n);
    cplx* tile_all = reinterpret_cast<cplx*>(base);
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(base + S::tile_bytes);
    cplx* tile = tile_all + (size_t)(t * L + lev) * PBS_TILE;
    const cplx* tile_other = tile_all + (size_t)((1 - t) * L + lev) * PBS_TILE;
    const bool live = b < B;
    const int bar_id = 1 + ctl, bar_n = 2 * L * 32;

    // ---- prologue: mod-switch the mask, ACC = X^(-b~) * (0, LUT) into TMEM (lo words | hi words)
    const uint64_t* ct = in + (size_t)(live ? b : 0) * (n + 1);
    for (int i = (t * L + lev) * 32 + lane; i <= n; i += 2 * L * 32)
        a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);
    named_bar_sync(bar_id, bar_n);
    {
        const uint64_t* lut = luts + (size_t)(lut_index && live ? lut_index[b] : 0) * PBS_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t rl[16], rh[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const int x = lane + 32 * (q & 31) + (q >> 5) * PBS_M;
                uint64_t v = 0;
                if (t == 1 && lev == 0) {
                    const int src = (x - rot) & 4095;
                    v = lut[src & 2047];
                    if (src & 2048) v = 0 - v;
                }
                rl[u] = (uint32_t)v;
                rh[u] = (uint32_t)(v >> 32);
            }
            tmem_st_x16(tacc + 16 * c, rl);
            tmem_st_x16(tacc + 64 + 16 * c, rh);
        }
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {  // fill the ring
        for (int q = 0; q < SLICES && q < total_slices; ++q) {
            mbar_expect_tx(&bar_full[q], SLICE_BYTES);
            tma_load_1d(ring + (size_t)q * S::slice_elems, bskf2 + (size_t)q * S::slice_elems, SLICE_BYTES, &bar_full[q]);
        }
    }

    const uint32_t rnd32 = 1u << (31 - beta);
    const int dshift = 32 - beta;
    double re[32], im[32];
    for (int i = 0; i < pairs; ++i) {
        // ---- digits of ACC_t itself (no rotation in the coefficient domain)
        if constexpr (L == 1) {  // high words of the accumulator from TMEM
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t h0[16], h1[16];
                tmem_ld_x16(tacc + 64 + 16 * c, h0);
                tmem_ld_x16(tacc + 96 + 16 * c, h1);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    re[16 * c + u] = (double)((int32_t)(h0[u] + rnd32) >> dshift);
                    im[16 * c + u] = (double)((int32_t)(h1[u] + rnd32) >> dshift);
                }
            }
        } else {  // ACC_t = A[t][0] + A[t][1] (64-bit sum of the two partials), then the level's balanced digit:
            // round to 2*beta bits, lev 1 = low beta bits sign-extended, lev 0 = the rest after the carry
            const uint32_t r_lo = 1u << (31 - 2 * beta);                            // rounding to 2*beta bits
            const uint32_t r_hi = r_lo + (1u << (31 - beta));                       // ... plus the carry out of digit 1
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t l0[16], h0[16], l1[16], h1[16];
                tmem_ld_x16(tquad + 16 * c, l0);
                tmem_ld_x16(tquad + 64 + 16 * c, h0);
                tmem_ld_x16(tquad + 128 + 16 * c, l1);
                tmem_ld_x16(tquad + 192 + 16 * c, h1);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t lo = l0[u] + l1[u];
                    const uint32_t hi = h0[u] + h1[u] + (lo < l0[u] ? 1u : 0u);
                    const int32_t dg = lev == 0 ? (int32_t)(hi + r_hi) >> dshift
                                                : (int32_t)((hi + r_lo) << beta) >> dshift;
                    if (c < 2) re[16 * c + u] = (double)dg;
                    else im[16 * (c - 2) + u] = (double)dg;
                }
            }
        }
        nfft::fwd_phase1(re, im, tw, tile, lane);
        __syncwarp();
        nfft::fwd_phase2(re, im, tile, lane);
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            cplx v;
            v.x = re[p];
            v.y = im[p];
            tile[nfft::brev5(p) * 32 + lane] = v;
        }
        // ---- monomials at this lane's bins: rho_k^e = omega^((4*lane+1)*e) * (omega^(128*e))^k1
        const int ea = a_tilde[2 * i], eb = a_tilde[2 * i + 1];
        // c_g = rho^e_g - 1 is carried directly: c' = c*r + (r - 1), all FMA chains
        double cx[3], cy[3], rx[3], ry[3], qx[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const int e = g == 0 ? ((ea + eb) & 4095) : (g == 1 ? ea : eb);
            const int E = (e * (4 * lane + 1)) & 4095;
            const cplx hi = omega[64 + (E >> 6)], lo = omega[E & 63];
            cx[g] = fma(hi.x, lo.x, fma(-hi.y, lo.y, -1.0));
            cy[g] = fma(hi.x, lo.y, hi.y * lo.x);
            const cplx r = omega[64 + (((128 * e) & 4095) >> 6)];
            rx[g] = r.x;
            ry[g] = r.y;
            qx[g] = r.x - 1.0;
        }
        named_bar_sync(bar_id, bar_n);  // (A) every warp's Fourier digits are visible
        // ---- walk the frequency blocks: out[bin] = F_t * sum_g c_g G_g[t][lev][t] + F_t' * sum_g c_g G_g[t'][lev][t],
        //      c_g = rho^e_g - 1
#pragma unroll
        for (int sl_i = 0; sl_i < SPI; ++sl_i) {
            const int sidx = i * SPI + sl_i;
            const int slot = sidx % SLICES;
            mbar_wait(&bar_full[slot], (uint32_t)((sidx / SLICES) & 1));
            const cplx* sl = ring + (size_t)slot * S::slice_elems;
#pragma unroll
            for (int kk = 0; kk < K1; ++kk) {
                const int k1 = sl_i * K1 + kk;
                const cplx* blk = sl + kk * S::block_elems;
                const int p = nfft::brev5(k1);
                const cplx fo = tile_other[k1 * 32 + lane];
                const double ax = re[p], ay = im[p];
                double kox, koy, ktx, kty;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const cplx bt = blk[(((g * 2 + t) * L + lev) * 2 + t) * 32 + lane];
                    const cplx bo = blk[(((g * 2 + (1 - t)) * L + lev) * 2 + t) * 32 + lane];
                    if (g == 0) {
                        kox = fma(cx[g], bt.x, -(cy[g] * bt.y));
                        koy = fma(cx[g], bt.y, cy[g] * bt.x);
                        ktx = fma(cx[g], bo.x, -(cy[g] * bo.y));
                        kty = fma(cx[g], bo.y, cy[g] * bo.x);
                    } else {
                        kox = fma(cx[g], bt.x, fma(-cy[g], bt.y, kox));
                        koy = fma(cx[g], bt.y, fma(cy[g], bt.x, koy));
                        ktx = fma(cx[g], bo.x, fma(-cy[g], bo.y, ktx));
                        kty = fma(cx[g], bo.y, fma(cy[g], bo.x, kty));
                    }
                    const double nx = fma(cx[g], rx[g], fma(-cy[g], ry[g], qx[g]));
                    cy[g] = fma(cx[g], ry[g], fma(cy[g], rx[g], ry[g]));
                    cx[g] = nx;
                }
                re[p] = fma(ax, kox, fma(-ay, koy, fma(fo.x, ktx, -(fo.y * kty))));
                im[p] = fma(ax, koy, fma(ay, kox, fma(fo.x, kty, fo.y * ktx)));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[slot]);
            if (threadIdx.x == 0) {  // keep the ring full: refill the slot warp 0 left MB2_LAG slices ago
                const int done = sidx - MB2_LAG;
                const int next = done + SLICES;
                if (done >= 0 && next < total_slices) {
                    const int ds = done % SLICES;
                    mbar_wait(&bar_empty[ds], (uint32_t)((done / SLICES) & 1));
                    mbar_expect_tx(&bar_full[ds], SLICE_BYTES);
                    tma_load_1d(ring + (size_t)ds * S::slice_elems, bskf2 + (size_t)next * S::slice_elems, SLICE_BYTES,
                                &bar_full[ds]);
                }
            }
        }
        named_bar_sync(bar_id, bar_n);  // (B) nobody reads the published digits any more
        nfft::inv_phase1(re, im, tw, tile, lane);
        __syncwarp();
        nfft::inv_phase2(re, im, tile, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t rl[16], rh[16];
            tmem_ld_x16(tacc + 16 * c, rl);
            tmem_ld_x16(tacc + 64 + 16 * c, rh);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const uint64_t v = (((uint64_t)rh[u] << 32) | rl[u]) + f64_to_torus(c < 2 ? re[q] : im[q - 32]);
                rl[u] = (uint32_t)v;
                rh[u] = (uint32_t)(v >> 32);
            }
            tmem_st_x16(tacc + 16 * c, rl);
            tmem_st_x16(tacc + 64 + 16 * c, rh);
        }
        tmem_wait_st();
        if constexpr (L > 1) {  // (C) the partner's partial accumulator is complete before anyone sums it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            named_bar_sync(bar_id, bar_n);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        } else {
            __syncwarp();
        }
    }
    // the tail slices (done > total - LAG) were never waited for by the producer; nothing is pending
    if (live && lev == 0) {
        uint64_t* o = out + (size_t)b * ((size_t)PBS_N + 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t rl[16], rh[16];
            tmem_ld_x16(tacc + 16 * c, rl);
            tmem_ld_x16(tacc + 64 + 16 * c, rh);
            uint32_t sl[16], sh[16];
            if constexpr (L > 1) {
                tmem_ld_x16(tquad + 128 + 16 * c, sl);
                tmem_ld_x16(tquad + 192 + 16 * c, sh);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int q = 16 * c + u;
                const int x = lane + 32 * (q & 31) + (q >> 5) * PBS_M;
                uint64_t v = ((uint64_t)rh[u] << 32) | rl[u];
                if constexpr (L > 1) v += ((uint64_t)sh[u] << 32) | sl[u];
                if (t == 0) {
                    if (x == 0) o[0] = v;
                    else o[PBS_N - x] = 0 - v;
                } else if (x == 0) {
                    o[PBS_N] = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// standard-domain bsk2 [pairs][3][2][L][2][N] -> Fourier, by frequency block: [pairs][32][3][2][L][2][32]
__global__ void __launch_bounds__(B2F_WARPS * 32)
bsk2_to_fourier_kernel(const uint64_t* __restrict__ bsk2, int64_t polys, int L, const cplx* __restrict__ g_twf,
                       cplx* __restrict__ bskf2) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* twf = reinterpret_cast<cplx*>(smem_raw);
    cplx* bufs = twf + PBS_TILE;
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) twf[i] = g_twf[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * B2F_WARPS + warp;   // q = (((i*3+g)*2+t)*L+lev)*2+c
    if (q >= polys) return;
    cplx* buf = bufs + warp * PBS_TILE;
    const uint64_t* src = bsk2 + (size_t)q * PBS_N;
    double re[32], im[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        re[j2] = (double)(int64_t)src[lane + 32 * j2];
        im[j2] = (double)(int64_t)src[lane + 32 * j2 + PBS_M];
    }
    nfft::fwd_phase1(re, im, twf, buf, lane);
    __syncwarp();
    nfft::fwd_phase2(re, im, buf, lane);
    const int per_pair = 3 * 2 * L * 2;          // polynomials per key-bit pair = entries per frequency block
    const int64_t i = q / per_pair;
    const int within = (int)(q - i * per_pair);  // ((g*2+t)*L+lev)*2+c
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const int k1 = nfft::brev5(p);
        cplx v;
        v.x = re[p];
        v.y = im[p];
        bskf2[((size_t)(i * 32 + k1) * per_pair + (size_t)within) * 32 + lane] = v;
    }
}

cudaError_t launch_bsk2_to_fourier(const fhe_b200_pbs_params& p, const uint64_t* d_bsk2, double* d_bskf2, cudaStream_t s) {
    const cplx* twf;
    cudaError_t e = get_tables(&twf);
    if (e != cudaSuccess) return e;
    if (p.k != 1 || p.l_pbs < 1 || p.l_pbs > 2 || (p.n & 1) || p.N != PBS_N) return cudaErrorInvalidValue;
    const int64_t polys = (int64_t)(p.n / 2) * 3 * 2 * p.l_pbs * 2;
    const size_t smem = sizeof(cplx) * PBS_TILE * (1 + B2F_WARPS);
    e = cudaFuncSetAttribute(bsk2_to_fourier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bsk2_to_fourier_kernel<<<(unsigned)((polys + B2F_WARPS - 1) / B2F_WARPS), B2F_WARPS * 32, smem, s>>>(
        d_bsk2, polys, p.l_pbs, twf, reinterpret_cast<cplx*>(d_bskf2));
    count_launch();
    return cudaGetLastError();
}

template <int L, int NCT>
static cudaError_t launch_pbs_mb2_t(const fhe_b200_pbs_params& p, const cplx* bskf2, const uint64_t* d_in, int64_t B,
                                    const uint64_t* d_luts, const int32_t* d_lut_index, const cplx* tw, uint64_t* d_out,
                                    cudaStream_t s) {
    const size_t smem = PbsMb2Smem<L>::total(p.n, NCT);
    cudaError_t e = cudaFuncSetAttribute(pbs_kernel_mb2<L, NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pbs_kernel_mb2<L, NCT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (getenv("FHE_B200_PBS_DEBUG")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pbs_kernel_mb2<L, NCT>, NCT * 64 * L, smem);
        fprintf(stderr, "pbs_kernel_mb2<%d,%d>: smem %zu B, %d CTA(s) per SM\n", L, NCT, smem, nb);
    }
    const unsigned grid = (unsigned)((B + NCT - 1) / NCT);
    pbs_kernel_mb2<L, NCT><<<grid, NCT * 64 * L, smem, s>>>(bskf2, d_in, B, p.n, p.beta_pbs, d_luts, d_lut_index, tw, d_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_pbs_mb2(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                           const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, int sm_count,
                           cudaStream_t s) {
    if (p.k != 1 || p.l_pbs < 1 || p.l_pbs > 2 || (p.n & 1) || p.N != PBS_N) return cudaErrorInvalidValue;
    if (p.l_pbs == 2 && (2 * p.beta_pbs > 31)) return cudaErrorInvalidValue;  // digits come from the high word
    const cplx* tw;
    cudaError_t e = get_tables(&tw);
    if (e != cudaSuccess) return e;
    const cplx* bskf2 = reinterpret_cast<const cplx*>(d_bskf2);
    if (p.l_pbs == 2) return launch_pbs_mb2_t<2, 2>(p, bskf2, d_in, B, d_luts, d_lut_index, tw, d_out, s);
    // Three kernels share a batch (measured on B200, profiles/r2_pbs_wide_times.txt):
    //   pbs_kernel_mb2<1,4>  four ciphertexts per CTA, two fat warps each: a full wave of 4 x SMs ciphertexts takes 5.47 ms
    //                        (108 k PBS/s) -- the throughput kernel;
    //   pbs_kernel_mb2_wide  one ciphertext per CTA, eight warps (pbs_wide.cu): a wave of SMs ciphertexts takes 1.55 ms
    //                        (96 k PBS/s) -- the latency kernel;
    //   pbs_kernel_mb2_pair  one ciphertext per cluster of two CTAs (pbs_wide.cu): SMs / 2 ciphertexts in 1.28 ms.
    // Full waves of 4 x SMs go to the first; what is left goes to the second while it needs at most three of its waves
    // (3 x 1.55 < 5.47), so no batch size pays for a mostly empty wave of four-ciphertext CTAs -- and to the third
    // when there are two SMs for every remaining ciphertext.
    const int64_t wave4 = (int64_t)MB2_L1_NCT * sm_count;
    int64_t full = (B / wave4) * wave4;
    int64_t rest = B - full;
    const bool wide_ok = p.beta_pbs <= 31 && !getenv("FHE_B200_PBS_NO_WIDE");
    if (!wide_ok || rest > 3 * (int64_t)sm_count) { full = B; rest = 0; }
    if (full > 0) {
        e = launch_pbs_mb2_t<1, MB2_L1_NCT>(p, bskf2, d_in, full, d_luts, d_lut_index, tw, d_out, s);
        if (e != cudaSuccess) return e;
    }
    if (rest > 0 && 2 * rest <= (int64_t)sm_count && !getenv("FHE_B200_PBS_NO_PAIR"))     // two SMs per ciphertext while they are free
        return launch_pbs_mb2_pair(p, d_bskf2, d_in + (size_t)full * (p.n + 1), rest, d_luts,
                                   d_lut_index ? d_lut_index + full : nullptr, d_out + (size_t)full * ((size_t)p.k * p.N + 1), s);
    if (rest > 0)
        return launch_pbs_mb2_wide(p, d_bskf2, d_in + (size_t)full * (p.n + 1), rest, d_luts,
                                   d_lut_index ? d_lut_index + full : nullptr, d_out + (size_t)full * ((size_t)p.k * p.N + 1), s);
    return cudaSuccess;
}

// ------------------------------------------------------------------------------- packed encrypted inner products
// out[g] = GGSW(Q) [.] GLWE_g for a whole collection of GLWE ciphertexts and ONE GGSW (the query):
// the leveled form of the both-encrypted comparison.  Each GLWE packs N/slot document vectors, Q(X) =
// sum_j x_j X^(-j), so coefficient slot*b of the product is the inner product with document b.
//
// The Fourier GGSW (2 x L x 2 polynomials, 128 KB at L = 2) is the only operand every ciphertext
// shares.  It lives in TENSOR MEMORY for the whole launch: column c of it is exactly 64 KB = one TMEM
// lane quadrant (32 lanes x 512 columns), lane k2 holding its 32 bins x 2L rows as (re, im) f64 pairs;
// warps of even quadrants produce output polynomial 0 (mask), odd quadrants polynomial 1 (body), so every
// warp finds the column it needs in its own quadrant and shared memory is left to the transforms.
// Persistent CTAs (one per SM) stream the GLWE ciphertexts from HBM: 32 KB in, 32 KB out per ciphertext.
// Warp t of a ciphertext: digits of polynomial t at both levels -> two forward FFTs, published in shared
// memory; after a named barrier, out_t = sum_{t',lev} F[t'][lev] * G[t'][lev][t] with G read from TMEM;
// inverse FFT; coefficients written back.
struct GlweDotSmem {
    static constexpr size_t tw_bytes = (size_t)PBS_TILE * 16;
    static constexpr size_t bar_bytes = 128;
    static constexpr size_t head_bytes = tw_bytes + bar_bytes;
    static constexpr size_t per_ct = (size_t)4 * PBS_TILE * 16;   // tiles (t, lev)
    static size_t total(int nct) { return head_bytes + (size_t)nct * per_ct; }
};

template <int NCT>
__global__ void __launch_bounds__(NCT * 64, 1)
glwe_dot_kernel(const cplx* __restrict__ ggswf, const uint64_t* __restrict__ in, int64_t G, int beta,
                const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = GlweDotSmem;
    constexpr int L = 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + S::tw_bytes);
    for (int i = threadIdx.x; i < PBS_TILE; i += blockDim.x) tw[i] = g_tw[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int t = warp & 1;                      // polynomial this warp decomposes AND the output column it produces
    const uint32_t tq = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // ---- GGSW column t -> this quadrant: column ((tl*32 + p)*4 + w), p = register index of bin lane + 32*brev5(p)
    if (warp < 4 && warp < NCT * 2) {
#pragma unroll 1
        for (int tl = 0; tl < 2 * L; ++tl) {
            const cplx* src = ggswf + (size_t)(tl * 2 + t) * PBS_M;
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) {
                uint32_t w[16];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const cplx v = src[lane + 32 * nfft::brev5(pc * 4 + u)];
                    const unsigned long long xr = (unsigned long long)__double_as_longlong(v.x);
                    const unsigned long long xi = (unsigned long long)__double_as_longlong(v.y);
                    w[4 * u + 0] = (uint32_t)xr;
                    w[4 * u + 1] = (uint32_t)(xr >> 32);
                    w[4 * u + 2] = (uint32_t)xi;
                    w[4 * u + 3] = (uint32_t)(xi >> 32);
                }
                tmem_st_x16(tq + (uint32_t)((tl * 32 + pc * 4) * 4), w);
            }
        }
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int ctl = warp >> 1;
    cplx* tiles = reinterpret_cast<cplx*>(smem_raw + S::head_bytes + (size_t)ctl * S::per_ct);
    const int bar_id = 1 + ctl;
    // balanced digits of the value rounded to 2*beta bits: lev 1 = low beta bits sign-extended,
    // lev 0 = the rest after the carry out of lev 1 (same digits as the oracle's decompose())
    const uint64_t rnd = 1ull << (63 - 2 * beta);
    const int s1 = 64 - 2 * beta, sx = 32 - beta;
    const uint32_t carry = 1u << (31 - beta);
    auto digit = [&](uint64_t v, int lev) -> double {
        const uint64_t r = v + rnd;   // bits [64-2*beta, 64) are the rounded value; digits are taken with 32-bit ops
        const int32_t dg = lev == 0 ? (int32_t)((uint32_t)(r >> 32) + carry) >> sx
                                    : (int32_t)((uint32_t)(r >> s1) << sx) >> sx;
        return (double)dg;
    };
    double re[32], im[32];
    for (int64_t grp = blockIdx.x; grp * NCT < G; grp += gridDim.x) {
        const int64_t g = grp * NCT + ctl;
        if (g >= G) continue;   // whole ciphertext slot idle: both of its warps skip together
        const uint64_t* src = in + ((size_t)g * 2 + t) * PBS_N;
        {   // pull the polynomial this warp will need in the NEXT iteration into L2 while this one is processed
            const int64_t gn = g + (int64_t)gridDim.x * NCT;
            if (gn < G) {
                const char* nx = reinterpret_cast<const char*>(in + ((size_t)gn * 2 + t) * PBS_N);
#pragma unroll
                for (int u = 0; u < 4; ++u) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)(lane + 32 * u) * 128));
            }
        }
#pragma unroll 1   // one copy of the forward transform in the instruction cache (ncu: no_instruction stalls)
        for (int lev = 0; lev < L; ++lev) {
            cplx* tile = tiles + (size_t)(t * L + lev) * PBS_TILE;
#pragma unroll
            for (int j2 = 0; j2 < 32; ++j2) {
                re[j2] = digit(src[lane + 32 * j2], lev);
                im[j2] = digit(src[lane + 32 * j2 + PBS_M], lev);
            }
            nfft::fwd_phase1(re, im, tw, tile, lane);
            __syncwarp();
            nfft::fwd_phase2(re, im, tile, lane);
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                cplx v;
                v.x = re[p];
                v.y = im[p];
                tile[nfft::brev5(p) * 32 + lane] = v;
            }
        }
        named_bar_sync(bar_id, 64);  // (A) all four spectra of this ciphertext are visible
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) {
            double ax[4] = {0.0, 0.0, 0.0, 0.0}, ay[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int tl = 0; tl < 2 * L; ++tl) {
                uint32_t w[16];
                tmem_ld_x16(tq + (uint32_t)((tl * 32 + pc * 4) * 4), w);
                const cplx* f = tiles + (size_t)tl * PBS_TILE;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const cplx fv = f[nfft::brev5(pc * 4 + u) * 32 + lane];
                    const double gx = __longlong_as_double((long long)(((unsigned long long)w[4 * u + 1] << 32) | w[4 * u + 0]));
                    const double gy = __longlong_as_double((long long)(((unsigned long long)w[4 * u + 3] << 32) | w[4 * u + 2]));
                    ax[u] = fma(fv.x, gx, fma(-fv.y, gy, ax[u]));
                    ay[u] = fma(fv.x, gy, fma(fv.y, gx, ay[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                re[pc * 4 + u] = ax[u];
                im[pc * 4 + u] = ay[u];
            }
        }
        named_bar_sync(bar_id, 64);  // (B) nobody reads the published spectra any more
        cplx* tile = tiles + (size_t)(t * L) * PBS_TILE;
        nfft::inv_phase1(re, im, tw, tile, lane);
        __syncwarp();
        nfft::inv_phase2(re, im, tile, lane);
        uint64_t* dst = out + ((size_t)g * 2 + t) * PBS_N;
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) {
            dst[lane + 32 * j2] = f64_to_torus(re[j2]);
            dst[lane + 32 * j2 + PBS_M] = f64_to_torus(im[j2]);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

cudaError_t launch_glwe_dot(const fhe_b200_pbs_params& p, const double* d_ggswf, const uint64_t* d_in, int64_t G,
                            uint64_t* d_out, int sm_count, cudaStream_t s) {
    if (p.k != 1 || p.l_pbs != 2 || p.N != PBS_N || 2 * p.beta_pbs > 62) return cudaErrorInvalidValue;
    if (G <= 0) return cudaSuccess;
    const cplx* tw;
    cudaError_t e = get_tables(&tw);
    if (e != cudaSuccess) return e;
    constexpr int NCT = 3;
    const size_t smem = GlweDotSmem::total(NCT);
    e = cudaFuncSetAttribute(glwe_dot_kernel<NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t groups = (G + NCT - 1) / NCT;
    const unsigned grid = (unsigned)(groups < sm_count ? groups : sm_count);
    glwe_dot_kernel<NCT><<<grid, NCT * 64, smem, s>>>(reinterpret_cast<const cplx*>(d_ggswf), d_in, G, p.beta_pbs, tw, d_out);
    count_launch();
    return cudaGetLastError();
}

bool pbs_params_supported(const fhe_b200_pbs_params& p, const char** why) {
    *why = "";
    if (p.N != PBS_N) { *why = "polynomial size N must be 2048"; return false; }
    if (p.k != 1) { *why = "GLWE dimension k must be 1"; return false; }
    if (p.l_pbs < 1 || p.l_pbs > 3) { *why = "l_pbs must be in [1,3]"; return false; }
    if (p.beta_pbs < 1 || p.beta_pbs > 31 || p.l_pbs * p.beta_pbs > 62) { *why = "beta_pbs out of range"; return false; }
    if (p.n < 1 || p.n > 4096) { *why = "n must be in [1,4096]"; return false; }
    if (p.l_ks < 1 || p.l_ks > 8 || p.beta_ks < 1 || p.l_ks * p.beta_ks > 62) { *why = "keyswitch decomposition out of range"; return false; }
    return true;
}

template <int K, int L, int NCT>
static cudaError_t launch_pbs_t(const fhe_b200_pbs_params& p, const cplx* bskf, const uint64_t* d_in, int64_t B,
                                const uint64_t* d_luts, const int32_t* d_lut_index, const cplx* tw, uint64_t* d_out,
                                cudaStream_t s) {
    const size_t smem = PbsSmem<K, L>::total(p.n, NCT);
    cudaError_t e = cudaFuncSetAttribute(pbs_kernel<K, L, NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((B + NCT - 1) / NCT);
    const unsigned threads = NCT * (K + 1) * 32;
    pbs_kernel<K, L, NCT><<<grid, threads, smem, s>>>(bskf, d_in, B, p.n, p.beta_pbs, d_luts, d_lut_index, tw, d_out);
    count_launch();
    return cudaGetLastError();
}

// 0 = TMEM accumulators (default), 1 = shared-memory accumulators; FHE_B200_PBS_VARIANT overrides
static int g_pbs_variant = [] {
    const char* v = getenv("FHE_B200_PBS_VARIANT");
    return v ? atoi(v) : 0;
}();

cudaError_t launch_pbs(const fhe_b200_pbs_params& p, const double* d_bskf, const uint64_t* d_in, int64_t B,
                       const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, int sm_count,
                       cudaStream_t s) {
    const cplx* tw;
    cudaError_t e = get_tables(&tw);
    if (e != cudaSuccess) return e;
    const cplx* bskf = reinterpret_cast<const cplx*>(d_bskf);
    const bool wide = B > (int64_t)sm_count;  // more ciphertexts than SMs: share the staged BSK_i inside a CTA
    switch (p.l_pbs) {
        case 1:
            if (g_pbs_variant == 1)  // shared-memory accumulators (kept for A/B measurements)
                return wide ? launch_pbs_t<1, 1, 2>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s)
                            : launch_pbs_t<1, 1, 1>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
            if (B <= (int64_t)sm_count) return launch_pbs_tmem_t<1>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
            if (B <= 2 * (int64_t)sm_count) return launch_pbs_tmem_t<2>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
            return launch_pbs_tmem_t<4>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
        case 2:
            return launch_pbs_t<1, 2, 1>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
        case 3:
            return launch_pbs_t<1, 3, 1>(p, bskf, d_in, B, d_luts, d_lut_index, tw, d_out, s);
        default:
            return cudaErrorInvalidValue;
    }
}

// the per-device twiddle + omega table, for kernels in other translation units (pbs_wide.cu)
cudaError_t pbs_tables(const void** tables) {
    const cplx* tw = nullptr;
    cudaError_t e = get_tables(&tw);
    *tables = tw;
    return e;
}

}  // namespace fhe
