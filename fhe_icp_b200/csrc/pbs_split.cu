// pbs_split.cu -- multi-bit blind rotation with TWO WARPS PER POLYNOMIAL (16 complex points per lane, 128 registers per
// thread, no spills): four warps per ciphertext instead of the two of pbs_kernel_mb2 (pbs.cu).
//
// Validated on a B200 (tests/test_gpu_pbs.py::test_multibit_pbs_split_kernel; same acceptance as pbs_kernel_mb2) and
// measured (profiles/r2_ncu_pbs_split_v1.txt, profiles/r2_pbs_split_times.txt): with twice the warps per ciphertext it
// is the SMALL-BATCH kernel -- one ciphertext per CTA, one or two CTAs per SM: batch 1 in 2.56 ms (pbs_kernel_mb2: 4.25),
// batch 148 at 52 k PBS/s (31 k), batch 296 at 83 k (66 k) -- and launch_pbs_mb2 dispatches it for B <= 2 x SMs.  With
// four ciphertexts per SM (16 warps) it reaches 102 k PBS/s against 107 k for pbs_kernel_mb2: the extra shared-memory
// wavefronts of the split transposes (76 % of the pipe's peak) and 30 % more issued instructions eat what the
// occupancy gains, so large batches stay on pbs_kernel_mb2.
//
// Warp w of a CTA: ciphertext w >> 2, polynomial t = (w >> 1) & 1, half h = w & 1.  Warp (t, h) owns the accumulator
// coefficients j = lane + 32(2m + h) and j + 1024 (m = 0..15; tensor memory, 64 columns of its lane quadrant) and, in
// the pointwise stage, the bins lane + 32*k1 of output polynomial t for k1 in [16h, 16h + 16).
//
// One 16.9 KB shared-memory region per polynomial is reused through the step:
//   E | O tiles  ->  P0 | P1 half-spectra  ->  exchanged pointwise halves (own half where its P was)  ->  inverse tile
// with a 64-thread named barrier (the two warps of the polynomial) at every hand-over and the two 128-thread barriers
// of the ciphertext around the pointwise stage, where the warps of one polynomial read the other's half-spectra.
//
// Key stream: the Fourier key of fhe_b200_bsk2_to_fourier is read as it is ([pair][32 frequency blocks][384 complex]);
// ring slice s of a pair is two 6 KB bulk copies, frequency blocks s and 16 + s, so that all warps walk the 16 slices
// of a step in lock step, half h using block h of each slice.
#include "common.cuh"
#include "kernels.h"
#include "pbs_split.cuh"

namespace fhe {

using nfft::cplx;

namespace {

constexpr int PS_N = nfft::NPOLY;
constexpr int PS_TILE = nfft::TILE_ELEMS;
constexpr int PS_HALF = nfft::HALF_TILE_ELEMS;
constexpr int PS_OMEGA = 128;
// ring slots per ciphertexts-per-CTA.  Deeper rings were measured and change nothing (B200, batch 1: 4 slots 2.88 ms, 8 slots
// 2.86 ms, 12 slots 2.99 ms -- a lone ciphertext's step is bound by its own dependent FP64 / shared-memory chain, not by
// the key stream), while the larger footprint stops two CTAs from sharing an SM (batch 296: 3.68 -> 6.18 ms); so 4.
#ifndef PS_SLICES_1
#define PS_SLICES_1 4
#endif
#ifndef PS_SLICES_2
#define PS_SLICES_2 4
#endif
__host__ __device__ constexpr int ps_slices(int nct) { return nct == 1 ? PS_SLICES_1 : (nct == 2 ? PS_SLICES_2 : 4); }
constexpr int PS_SLICE_ELEMS = 2 * nfft::MB2_BLOCK_ELEMS;        // two frequency blocks per slice: 768 complex = 12 KB
constexpr int PS_SPI = 16;                                       // slices per blind-rotation step
constexpr int PS_LAG = 1;

template <int NCT>
struct PsSmem {
    static constexpr int slices = ps_slices(NCT);
    static constexpr size_t tw_bytes = (size_t)PS_TILE * 16;
    static constexpr size_t ring_bytes = (size_t)slices * PS_SLICE_ELEMS * 16;
    static constexpr size_t omega_bytes = (size_t)PS_OMEGA * 16;
    static constexpr size_t bar_bytes = 256;
    static constexpr size_t head_bytes = tw_bytes + ring_bytes + omega_bytes + bar_bytes;
    static constexpr size_t region_bytes = (size_t)2 * PS_TILE * 16;   // one region per polynomial
    __host__ __device__ static size_t per_ct(int n) { return region_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127); }
    static size_t total(int n) { return head_bytes + (size_t)NCT * per_ct(n); }
};

__device__ __forceinline__ void ps_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void ps_tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ps_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ps_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ps_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void ps_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace

template <int NCT>
__global__ void __launch_bounds__(NCT * 128, 1)
pbs_kernel_mb2_split(const cplx* __restrict__ bskf2, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
                     const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index,
                     const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = PsSmem<NCT>;
    constexpr int PS_SLICES = S::slices;
    constexpr int WARPS = NCT * 4;
    constexpr uint32_t TMEM_COLS = WARPS <= 4 ? 64 : (WARPS <= 8 ? 128 : 256);   // 64 columns per warp, 4 warps share a quadrant
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    cplx* ring = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes);
    cplx* omega = reinterpret_cast<cplx*>(smem_raw + S::tw_bytes + S::ring_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::tw_bytes + S::ring_bytes + S::omega_bytes);
    uint64_t* bar_empty = bar_full + PS_SLICES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_empty + PS_SLICES);
    for (int i = threadIdx.x; i < PS_TILE; i += blockDim.x) tw[i] = g_tw[i];
    for (int i = threadIdx.x; i < PS_OMEGA; i += blockDim.x) omega[i] = g_tw[PS_TILE + i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int q = 0; q < PS_SLICES; ++q) { mbar_init(&bar_full[q], 1); mbar_init(&bar_empty[q], WARPS); }
        mbar_fence_init();
    }
    if (warp == 0) ps_tmem_alloc(tmem_slot, TMEM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int ctl = warp >> 2, t = (warp >> 1) & 1, h = warp & 1;
    // accumulator of this warp: lane quadrant warp & 3, columns (warp >> 2)*64 ..: [0,16) low words of the 16 "re"
    // coefficients c[j], [16,32) low words of the "im" coefficients c[j+1024], [32,48) / [48,64) the high words
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    constexpr uint32_t SLICE_BYTES = (uint32_t)(PS_SLICE_ELEMS * 16);
    const int pairs = n >> 1;
    const int total_slices = pairs * PS_SPI;
    // ring slot <- slice q of the whole key walk: frequency blocks s and 16 + s of pair q / 16 (s = q % 16)
    auto load_slice = [&](int slot, int q) {
        const cplx* pair = bskf2 + (size_t)(q >> 4) * 32 * nfft::MB2_BLOCK_ELEMS;
        cplx* dst = ring + (size_t)slot * PS_SLICE_ELEMS;
        mbar_expect_tx(&bar_full[slot], SLICE_BYTES);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
            tma_load_1d(dst + (size_t)hh * nfft::MB2_BLOCK_ELEMS,
                        pair + (size_t)nfft::split_ring_block(2 * (q & 15) + hh) * nfft::MB2_BLOCK_ELEMS, SLICE_BYTES / 2,
                        &bar_full[slot]);
    };

    const int64_t b = (int64_t)blockIdx.x * NCT + ctl;
    unsigned char* base = smem_raw + S::head_bytes + (size_t)ctl * S::per_ct(n);
    cplx* region = reinterpret_cast<cplx*>(base) + (size_t)t * PS_TILE;            // this polynomial's 16.9 KB
    const cplx* region_other = reinterpret_cast<cplx*>(base) + (size_t)(1 - t) * PS_TILE;
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(base + S::region_bytes);
    cplx* tile_e = region;                       // forward tiles
    cplx* tile_o = region + PS_HALF;
    cplx* half_own = region + (size_t)h * PS_HALF;        // exchange of the pointwise halves
    const cplx* half_partner = region + (size_t)(1 - h) * PS_HALF;
    const bool live = b < B;
    const int bar_ct = 1 + ctl, bar_poly = 1 + NCT + 2 * ctl + t;                  // named barriers 1 .. 3*NCT (<= 12)

    // ---- prologue: mod-switch the mask, ACC = X^(-b~) * (0, LUT) into tensor memory
    const uint64_t* ct = in + (size_t)(live ? b : 0) * (n + 1);
    for (int i = (warp & 3) * 32 + lane; i <= n; i += 128) a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);
    ps_bar(bar_ct, 128);
    {
        const uint64_t* lut = luts + (size_t)(lut_index && live ? lut_index[b] : 0) * PS_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
        uint32_t lo_re[16], lo_im[16], hi_re[16], hi_im[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int j = lane + 32 * (2 * m + h);
            uint64_t v0 = 0, v1 = 0;
            if (t == 1) {
                int src = (j - rot) & 4095;
                v0 = lut[src & 2047];
                if (src & 2048) v0 = 0 - v0;
                src = (j + 1024 - rot) & 4095;
                v1 = lut[src & 2047];
                if (src & 2048) v1 = 0 - v1;
            }
            lo_re[m] = (uint32_t)v0; hi_re[m] = (uint32_t)(v0 >> 32);
            lo_im[m] = (uint32_t)v1; hi_im[m] = (uint32_t)(v1 >> 32);
        }
        ps_tmem_st16(tacc, lo_re);
        ps_tmem_st16(tacc + 16, lo_im);
        ps_tmem_st16(tacc + 32, hi_re);
        ps_tmem_st16(tacc + 48, hi_im);
        ps_tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {  // fill the ring
        for (int q = 0; q < PS_SLICES && q < total_slices; ++q) load_slice(q, q);
    }

    double re[16], im[16];
    for (int i = 0; i < pairs; ++i) {
        // ---- digits of the owned coefficients (high words from tensor memory) -> forward pass 1 -> E | O tiles
        {
            uint32_t h0[16], h1[16];
            ps_tmem_ld16(tacc + 32, h0);
            ps_tmem_ld16(tacc + 48, h1);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                re[m] = nfft::split_digit(h0[m], beta);
                im[m] = nfft::split_digit(h1[m], beta);
            }
        }
        nfft::fwd_split_pass1(h, re, im, tile_e, tile_o, lane);
        ps_bar(bar_poly, 64);                                   // E and O are complete
        nfft::fwd_split_pass2_compute(h, re, im, tile_e, tile_o, tw, lane);
        ps_bar(bar_poly, 64);                                   // the partner has read E | O: P may overwrite them
        nfft::fwd_split_pass2_store(h, re, im, region, region + PS_HALF, lane);
        nfft::SplitMonomials mo;
        nfft::split_monomials_init(mo, omega, a_tilde[2 * i], a_tilde[2 * i + 1], lane, 16 * h);
        ps_bar(bar_ct, 128);                                    // (A) both polynomials' half-spectra are visible
        // ---- pointwise stage: bins lane + 32*(16h + s), s = 0..15; slice s of the ring holds blocks {s, 16 + s}
#pragma unroll
        for (int s = 0; s < PS_SPI; ++s) {
            const int sidx = i * PS_SPI + s;
            const int slot = sidx % PS_SLICES;
            mbar_wait(&bar_full[slot], (uint32_t)((sidx / PS_SLICES) & 1));
            const cplx* blk = ring + (size_t)slot * PS_SLICE_ELEMS + (size_t)h * nfft::MB2_BLOCK_ELEMS;
            const int k1 = 16 * h + s;
            nfft::split_pointwise_bin(t, lane, nfft::split_bin(region, region + PS_HALF, lane, k1),
                                      nfft::split_bin(region_other, region_other + PS_HALF, lane, k1), blk, mo, re[s], im[s]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[slot]);
            if (threadIdx.x == 0) {  // keep the ring full: refill the slot warp 0 left PS_LAG slices ago
                const int done = sidx - PS_LAG;
                const int next = done + PS_SLICES;
                if (done >= 0 && next < total_slices) {
                    const int ds = done % PS_SLICES;
                    mbar_wait(&bar_empty[ds], (uint32_t)((done / PS_SLICES) & 1));
                    load_slice(ds, next);
                }
            }
        }
        ps_bar(bar_ct, 128);                                    // (B) nobody reads the half-spectra any more
        // ---- exchange the pointwise halves through the bytes the half-spectra occupied
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            cplx v;
            v.x = re[p];
            v.y = im[p];
            half_own[nfft::hslot(p, lane)] = v;
        }
        ps_bar(bar_poly, 64);
        nfft::inv_split_pass1_combine(h, re, im, half_partner, lane);
        ps_bar(bar_poly, 64);                                   // the partner has read this warp's half: the tile may overwrite it
        nfft::inv_split_pass1_finish(h, re, im, tw, region, lane);
        ps_bar(bar_poly, 64);                                   // the inverse tile is complete
        nfft::inv_split_pass2(h, re, im, region, lane);
        // ---- ACC += result (64-bit wrapping adds on the (lo, hi) word pairs in tensor memory)
        {
            uint32_t lo[16], hi[16];
            ps_tmem_ld16(tacc, lo);
            ps_tmem_ld16(tacc + 32, hi);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const uint64_t v = (((uint64_t)hi[m] << 32) | lo[m]) + nfft::split_f64_to_torus(re[m]);
                lo[m] = (uint32_t)v;
                hi[m] = (uint32_t)(v >> 32);
            }
            ps_tmem_st16(tacc, lo);
            ps_tmem_st16(tacc + 32, hi);
            ps_tmem_ld16(tacc + 16, lo);
            ps_tmem_ld16(tacc + 48, hi);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const uint64_t v = (((uint64_t)hi[m] << 32) | lo[m]) + nfft::split_f64_to_torus(im[m]);
                lo[m] = (uint32_t)v;
                hi[m] = (uint32_t)(v >> 32);
            }
            ps_tmem_st16(tacc + 16, lo);
            ps_tmem_st16(tacc + 48, hi);
            ps_tmem_wait_st();
        }
        ps_bar(bar_poly, 64);                                   // the partner has read the inverse tile: next step may write E | O
    }
    // ---- sample extract coefficient 0: o[0] = A_0[0], o[N - x] = -A_0[x] (x >= 1), o[N] = A_1[0]
    if (live) {
        uint64_t* o = out + (size_t)b * ((size_t)PS_N + 1);
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            ps_tmem_ld16(tacc + 16 * part, lo);
            ps_tmem_ld16(tacc + 32 + 16 * part, hi);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int x = lane + 32 * (2 * m + h) + 1024 * part;
                const uint64_t v = ((uint64_t)hi[m] << 32) | lo[m];
                if (t == 0) {
                    if (x == 0) o[0] = v;
                    else o[PS_N - x] = 0 - v;
                } else if (x == 0) {
                    o[PS_N] = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) ps_tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NCT>
static cudaError_t launch_split_t(const fhe_b200_pbs_params& p, const cplx* bskf2, const uint64_t* d_in, int64_t B,
                                  const uint64_t* d_luts, const int32_t* d_lut_index, const cplx* tables, uint64_t* d_out,
                                  cudaStream_t s) {
    const size_t smem = PsSmem<NCT>::total(p.n);
    cudaError_t e = cudaFuncSetAttribute(pbs_kernel_mb2_split<NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((B + NCT - 1) / NCT);
    pbs_kernel_mb2_split<NCT><<<grid, NCT * 128, smem, s>>>(bskf2, d_in, B, p.n, p.beta_pbs, d_luts, d_lut_index, tables, d_out);
    count_launch();
    return cudaGetLastError();
}

// cts_per_cta: 1, 2 or 4; 0 = the best measured form for the batch (one ciphertext per CTA up to 2 x SMs, else four)
cudaError_t launch_pbs_mb2_split(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                                 const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, int sm_count,
                                 int cts_per_cta, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    if (p.k != 1 || p.l_pbs != 1 || (p.n & 1) || p.N != PS_N || p.beta_pbs < 1 || p.beta_pbs > 31) return cudaErrorInvalidValue;
    const void* tables = nullptr;
    cudaError_t e = pbs_tables(&tables);
    if (e != cudaSuccess) return e;
    const cplx* key = reinterpret_cast<const cplx*>(d_bskf2);
    const cplx* tw = reinterpret_cast<const cplx*>(tables);
    if (cts_per_cta == 0) cts_per_cta = B <= 2 * (int64_t)sm_count ? 1 : 4;
    if (cts_per_cta == 1) return launch_split_t<1>(p, key, d_in, B, d_luts, d_lut_index, tw, d_out, s);
    if (cts_per_cta == 2) return launch_split_t<2>(p, key, d_in, B, d_luts, d_lut_index, tw, d_out, s);
    if (cts_per_cta == 4) return launch_split_t<4>(p, key, d_in, B, d_luts, d_lut_index, tw, d_out, s);
    return cudaErrorInvalidValue;
}

}  // namespace fhe
