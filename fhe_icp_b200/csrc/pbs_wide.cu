// pbs_wide.cu -- multi-bit blind rotation with FOUR WARPS PER POLYNOMIAL: one ciphertext per CTA, 256 threads, 8 complex
// points per thread (pbs_wide.cuh).  The latency kernel: launch_pbs_mb2 dispatches it while there is at most one
// ciphertext per SM (the step of a lone ciphertext is a dependent chain; eight warps walk it 2.4x faster than the two of
// pbs_kernel_mb2 -- and 1.65x faster than the four warps of the round-2 "split" kernel this one replaced).
//
// Thread tid: polynomial t = tid >> 7, u = tid & 127.  It owns the accumulator coefficients j = u + 128a and j + 1024
// (a = 0..7) of polynomial t IN REGISTERS (16 torus words) and, in the Fourier domain, the bins k = u + 128 kL of
// output polynomial t.  Twiddles are per-thread constants of the launch, held in registers.
//
// Shared memory: two exchange buffers per polynomial (17 KB each) with fixed roles -- e0: forward stage 1 and inverse
// stage 3, e1: forward stage 2 and inverse stage 2:
//   stage1 -> [bar t] -> stage2 -> [bar t] -> stage3, spectrum -> [bar all] -> pointwise, inverse stage3 -> [bar all] ->
//   inverse stage2 -> [bar t] -> inverse stage1, accumulate
// so a buffer is rewritten only after a barrier that every one of its readers has passed.  The spectrum itself -- the one
// thing the two polynomials exchange -- goes through TENSOR MEMORY (a lane-aligned mailbox, see the kernel) between the
// two 256-thread barriers.  tests/emul/pbs_wide_emul.cpp replays this plan in several thread orders.
//
// Key stream: the Fourier key of fhe_b200_bsk2_to_fourier as it is ([pair][32 frequency blocks][384 complex]).  Ring
// slice q of a step is the eight consecutive blocks 8q .. 8q + 7 (48 KB, ONE bulk copy: a bulk copy takes ~1200 clocks
// from L2 whatever its size up to 48 KB, tools/stream_probe.cu, so few large copies in flight beat many small ones);
// three slots.  A slot is refilled by the LAST warp that leaves it (a shared counter per slot), i.e. at the earliest
// possible moment and without anybody spinning; the wait on a slice's full-barrier is issued one turn early so that
// the ~90 clocks a try_wait takes even on a completed barrier overlap the arithmetic.
//
// Memory safety: compute-sanitizer is closed on this pool; the plan above is emulated on the CPU and the
// kernel is covered by the acceptance of the other blind-rotation kernels on ragged batches.
#include <cstdio>

#include "common.cuh"
#include "kernels.h"
#include "pbs_wide.cuh"

namespace fhe {

using nfft::cplx;

namespace {

constexpr int PW_N = nfft::NPOLY;
constexpr int PW_TILE = nfft::TILE_ELEMS;     // offset of the omega table inside pbs_tables()
constexpr int PW_OMEGA = 128;
#ifndef WIDE_SLOTS
#define WIDE_SLOTS 3
#endif
#ifndef WIDE_SKEW
#define WIDE_SKEW 200       // clocks polynomial 1 is held back per step (see the kernel)
#endif
constexpr int PW_SLOTS = WIDE_SLOTS;
constexpr int PW_THREADS = 2 * wfft::WT;
constexpr unsigned PW_WARPS = PW_THREADS / 32;

struct PwSmem {
    static constexpr size_t xbuf_bytes = (size_t)wfft::XBUF_ELEMS * 16;                 // 17,408
    static constexpr size_t x_bytes = 4 * xbuf_bytes;                                   // [t][2]
    static constexpr size_t ring_bytes = (size_t)PW_SLOTS * wfft::SLICE_ELEMS * 16;
    static constexpr size_t omega_bytes = (size_t)PW_OMEGA * 16;
    static constexpr size_t bar_bytes = 256;
    static size_t total(int n) { return ring_bytes + x_bytes + omega_bytes + bar_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127); }
};

// tensor memory as a lane-aligned mailbox between the two polynomials (warps w and w + 4 share a lane quadrant)
__device__ __forceinline__ void pw_tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pw_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void pw_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void pw_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void pw_bar_poly(int t) {
    if (t == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

}  // namespace

__global__ void __launch_bounds__(PW_THREADS, 1)
pbs_kernel_mb2_wide(const cplx* __restrict__ bskf2, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
                    const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index,
                    const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = PwSmem;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* ring = reinterpret_cast<cplx*>(smem_raw);
    cplx* xbufs = reinterpret_cast<cplx*>(smem_raw + S::ring_bytes);
    cplx* omega = reinterpret_cast<cplx*>(smem_raw + S::ring_bytes + S::x_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::ring_bytes + S::x_bytes + S::omega_bytes);
    unsigned* left = reinterpret_cast<unsigned*>(bar_full + PW_SLOTS);      // warps that have left each slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(left + PW_SLOTS + 1);
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(smem_raw + S::ring_bytes + S::x_bytes + S::omega_bytes + S::bar_bytes);

    const int tid = threadIdx.x, t = tid >> 7, u = tid & 127, lane = tid & 31, wp = u >> 5;
    const int64_t b = blockIdx.x;
    if (b >= B) return;
    for (int i = tid; i < PW_OMEGA; i += PW_THREADS) omega[i] = g_tw[PW_TILE + i];
    if (tid == 0) {
        for (int q = 0; q < PW_SLOTS; ++q) { mbar_init(&bar_full[q], 1); left[q] = 0; }
        mbar_fence_init();
    }
    const uint64_t* ct = in + (size_t)b * (n + 1);
    for (int i = tid; i <= n; i += PW_THREADS) a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);
    if (tid < 32) pw_tmem_alloc(tmem_slot, 64);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // The spectrum mailbox: thread (t, u) and thread (1 - t, u) are the same lane of two warps that share a TMEM lane
    // quadrant (warps w and w + 4), and what one publishes -- its 8 bins u + 128 kL -- is exactly what the other needs.
    // 32 columns per polynomial: tcgen05.st / tcgen05.ld instead of 8 shared-memory stores + 8 loads per thread and step
    // (the shared-memory pipe is this kernel's busiest).
    const uint32_t tmail = *tmem_slot + ((uint32_t)((tid >> 5 & 3) * 32) << 16);
    const uint32_t tmail_own = tmail + 32u * (uint32_t)t, tmail_oth = tmail + 32u * (uint32_t)(1 - t);

    const int pairs = n >> 1;
    const int total_slices = pairs * wfft::SLICES_PER_STEP;
    constexpr uint32_t SLICE_BYTES = (uint32_t)(wfft::SLICE_ELEMS * 16);
    auto load_slice = [&](int slot, int q) {     // slice q of the whole key walk is contiguous: blocks 8(q%4).. of pair q/4
        mbar_expect_tx(&bar_full[slot], SLICE_BYTES);
        tma_load_1d(ring + (size_t)slot * wfft::SLICE_ELEMS, bskf2 + (size_t)q * wfft::SLICE_ELEMS, SLICE_BYTES, &bar_full[slot]);
    };
    if (tid == 0)
        for (int q = 0; q < PW_SLOTS && q < total_slices; ++q) load_slice(q, q);

    wfft::Twiddles tw;
    wfft::twiddles_init(tw, u);

    // ---- ACC = X^(-b~) * (0, LUT), in registers
    uint64_t acc_re[8], acc_im[8];
    {
        const uint64_t* lut = luts + (size_t)(lut_index ? lut_index[b] : 0) * PW_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int j = u + 128 * a;
            uint64_t v0 = 0, v1 = 0;
            if (t == 1) {
                int src = (j - rot) & 4095;
                v0 = lut[src & 2047];
                if (src & 2048) v0 = 0 - v0;
                src = (j + 1024 - rot) & 4095;
                v1 = lut[src & 2047];
                if (src & 2048) v1 = 0 - v1;
            }
            acc_re[a] = v0;
            acc_im[a] = v1;
        }
    }

    // this polynomial's two exchange buffers: e0 carries forward stage 1 and inverse stage 3, e1 forward stage 2 and inverse
    // stage 2 -- four exchanges per step, every rewrite behind a barrier all readers of the buffer have passed
    cplx* const e0 = xbufs + (size_t)(2 * t) * wfft::XBUF_ELEMS;
    cplx* const e1 = e0 + wfft::XBUF_ELEMS;

#ifdef WIDE_TIMING      // debug builds: clocks per phase of a step, printed by one warp of each polynomial of CTA 0
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define WIDE_TICK(k) { const unsigned long long now = clock64(); tacc[k] += now - tprev; tprev = now; }
#else
#define WIDE_TICK(k)
#endif
    double re[8], im[8];
    for (int i = 0; i < pairs; ++i) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            re[a] = wfft::top_digit((uint32_t)(acc_re[a] >> 32), beta);
            im[a] = wfft::top_digit((uint32_t)(acc_im[a] >> 32), beta);
        }
        wfft::fwd_stage1(re, im, tw, u, e0);
        pw_bar_poly(t);
        WIDE_TICK(0)
        wfft::fwd_stage2(tw, u, e0, e1);
        pw_bar_poly(t);
        WIDE_TICK(1)
        wfft::fwd_stage3(u, e1, re, im);
        {   // publish the spectrum: the other polynomial's pointwise stage reads it (tensor-memory mailbox)
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                uint32_t w[16];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const unsigned long long xr = (unsigned long long)__double_as_longlong(re[4 * hlf + k4]);
                    const unsigned long long xi = (unsigned long long)__double_as_longlong(im[4 * hlf + k4]);
                    w[4 * k4 + 0] = (uint32_t)xr;
                    w[4 * k4 + 1] = (uint32_t)(xr >> 32);
                    w[4 * k4 + 2] = (uint32_t)xi;
                    w[4 * k4 + 3] = (uint32_t)(xi >> 32);
                }
                pw_tmem_st16(tmail_own + 16u * (uint32_t)hlf, w);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        wfft::Monomials mo;
        wfft::monomials_init(mo, omega, a_tilde[2 * i], a_tilde[2 * i + 1], u);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WIDE_TICK(2)
        // ---- pointwise stage
        double gre[8], gim[8];
        uint32_t ready = mbar_try_wait(&bar_full[(i * wfft::SLICES_PER_STEP) % PW_SLOTS],
                                       (uint32_t)(((i * wfft::SLICES_PER_STEP) / PW_SLOTS) & 1));
#pragma unroll
        for (int q = 0; q < wfft::SLICES_PER_STEP; ++q) {     // two bins (one ring slice) per turn
            const int sidx = i * wfft::SLICES_PER_STEP + q;
            const int slot = sidx % PW_SLOTS;
            cplx fo0, fo1;
            {
                uint32_t w[8];
                pw_tmem_ld8(tmail_oth + 8u * (uint32_t)q, w);      // the other polynomial's bins 2q, 2q + 1 of this thread
                fo0.x = __longlong_as_double((long long)(((unsigned long long)w[1] << 32) | w[0]));
                fo0.y = __longlong_as_double((long long)(((unsigned long long)w[3] << 32) | w[2]));
                fo1.x = __longlong_as_double((long long)(((unsigned long long)w[5] << 32) | w[4]));
                fo1.y = __longlong_as_double((long long)(((unsigned long long)w[7] << 32) | w[6]));
            }
            cplx fa0, fa1;
            fa0.x = re[2 * q]; fa0.y = im[2 * q];
            fa1.x = re[2 * q + 1]; fa1.y = im[2 * q + 1];
#ifndef WIDE_NOSTREAM      // (timing experiment: WIDE_NOSTREAM computes on whatever the ring holds -- invalid results)
            if (!ready) mbar_wait(&bar_full[slot], (uint32_t)((sidx / PW_SLOTS) & 1));
            if (q + 1 < wfft::SLICES_PER_STEP)      // the next turn's wait, issued now: its latency hides under this turn
                ready = mbar_try_wait(&bar_full[(sidx + 1) % PW_SLOTS], (uint32_t)(((sidx + 1) / PW_SLOTS) & 1));
#endif
            const cplx* blk0 = ring + (size_t)slot * wfft::SLICE_ELEMS + (size_t)wp * wfft::MB2_BLOCK_ELEMS;
            const cplx* blk1 = blk0 + 4 * wfft::MB2_BLOCK_ELEMS;
            wfft::pointwise_bin(t, lane, fa0, fo0, blk0, mo, gre[2 * q], gim[2 * q]);
            wfft::pointwise_bin(t, lane, fa1, fo1, blk1, mo, gre[2 * q + 1], gim[2 * q + 1]);
#ifndef WIDE_NOSTREAM
            __syncwarp();
            if (lane == 0) {
                // The last warp to leave the slot refills it.  No fence: a warp issues in order and every load from the
                // slot has returned before the FMAs above could issue, so when this atomic issues the warp's reads of
                // the slot are complete (a __threadfence_block here is a MEMBAR.SC per turn: ~6 % of the step).
                if (atomicAdd(&left[slot], 1u) == PW_WARPS - 1) {
                    left[slot] = 0;
                    const int next = sidx + PW_SLOTS;
                    if (next < total_slices) load_slice(slot, next);
                }
            }
#endif
        }
        WIDE_TICK(3)
        wfft::inv_stage3(u, gre, gim, e0);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WIDE_TICK(4)                               // also: nobody reads the published spectra any more
        // Phase offset between the two polynomials.  Their warps run every phase of a step in lock step (load burst,
        // arithmetic, store burst, barrier), so the shared-memory pipe and the FP64 pipe take turns.  Holding polynomial 1
        // back by a fraction of a phase after this barrier makes its bursts fall under polynomial 0's arithmetic for the
        // five phases up to the next CTA-wide barrier (where polynomial 0 then waits the same time): batch 148 measured
        // 1.780 ms without, 1.709 / 1.697 / 1.611 / 1.598 / 1.641 ms at 50 / 100 / 150 / 200 / 250 clocks.  (Letting the
        // polynomials run free of each other -- flags instead of CTA-wide barriers -- settles at an offset of a WHOLE
        // phase and gains nothing: 1.78 ms; profiles/r2_pbs_wide_times.txt.)
        if (WIDE_SKEW > 0 && t == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < WIDE_SKEW) { }
        }
        wfft::inv_stage2(tw, u, e0, e1);
        pw_bar_poly(t);
        WIDE_TICK(5)
        wfft::inv_stage1(tw, u, e1, re, im);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            acc_re[a] += wfft::f64_to_torus_u64(re[a]);
            acc_im[a] += wfft::f64_to_torus_u64(im[a]);
        }
        WIDE_TICK(6)
    }
#ifdef WIDE_TIMING
    if (blockIdx.x == 0 && (tid & 31) == 0 && (tid >> 5) % 4 == 0)
        printf("warp %d clocks per step: stage1+bar %llu | stage2+bar %llu | stage3+publish+bar %llu | pointwise %llu | inv3+bar %llu | inv2+bar %llu | inv1+acc+digits %llu\n",
               tid >> 5, tacc[0] / pairs, tacc[1] / pairs, tacc[2] / pairs, tacc[3] / pairs, tacc[4] / pairs, tacc[5] / pairs, tacc[6] / pairs);
#endif
    // ---- sample extract coefficient 0: o[0] = A_0[0], o[N - x] = -A_0[x] (x >= 1), o[N] = A_1[0]
    uint64_t* o = out + (size_t)b * ((size_t)PW_N + 1);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int x = u + 128 * a + 1024 * part;
            const uint64_t v = part ? acc_im[a] : acc_re[a];
            if (t == 0) {
                if (x == 0) o[0] = v;
                else o[PW_N - x] = 0 - v;
            } else if (x == 0) {
                o[PW_N] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) pw_tmem_dealloc(*tmem_slot, 64);
}

// ------------------------------------------------------------------------------------------------------------------
// pbs_kernel_mb2_pair -- ONE ciphertext on a CLUSTER OF TWO CTAs (two SMs): CTA t owns polynomial t (128 threads, the same
// four warps and the same 8 x 8 x 16 transform as above), streams only COLUMN t of the Fourier key (half the bytes, from
// the column layout of bsk2_column_split_kernel) and hands its spectrum to the other CTA through distributed shared
// memory: after the forward transform it stores the 1024 bins in a send buffer and one thread issues a 16 KB bulk copy
// shared::cta -> shared::cluster that completes on an mbarrier IN THE PEER's shared memory.  The pointwise stage first
// forms, per bin, S_own = sum c_g K_g[t][t], S_oth = sum c_g K_g[1-t][t] and F_t * S_own -- none of which needs the peer --
// and only then waits for the peer's spectrum to add F_(1-t) * S_oth, so the ~1 us the exchange takes hides under the
// stage.  Send and receive buffers alternate with the step's parity: the two CTAs can never be more than one exchange
// apart (each needs the other's spectrum every step), so a buffer is rewritten two exchanges after it was last read.
// Shared-memory exchange buffers: A for forward stage 1 / inverse stage 3, B for forward stage 2 / inverse stage 2,
// four CTA-wide barriers per step.  A cluster barrier after the mbarrier initialisation (nobody sends into a barrier
// that does not exist yet) and one before exit (nobody's shared memory disappears under an in-flight copy).
namespace {

constexpr int PP_THREADS = wfft::WT;
constexpr unsigned PP_WARPS = PP_THREADS / 32;
constexpr int PP_SLOTS = 3;
constexpr uint32_t PP_SPEC_BYTES = 1024 * 16;

struct PpSmem {
    static constexpr size_t ring_bytes = (size_t)PP_SLOTS * wfft::COL_SLICE_ELEMS * 16;      // 72 KB
    static constexpr size_t x_bytes = 2 * (size_t)wfft::XBUF_ELEMS * 16;                     // A, B
    static constexpr size_t spec_bytes = 4 * (size_t)PP_SPEC_BYTES;                          // send[2], recv[2]
    static constexpr size_t omega_bytes = (size_t)PW_OMEGA * 16;
    static constexpr size_t bar_bytes = 256;
    static size_t total(int n) { return ring_bytes + x_bytes + spec_bytes + omega_bytes + bar_bytes + (((size_t)(n + 1) * 2 + 127) & ~(size_t)127); }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_peer(const void* local_smem, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 KB of this CTA's shared memory -> the peer's, completing (bytes) on the peer's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t peer_dst, const void* local_src, uint32_t bytes, uint32_t peer_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(peer_dst),
                 "r"(smem_u32(local_src)), "r"(bytes), "r"(peer_bar) : "memory");
}

// keycol[((pair*2 + c)*32 + k1)*6 + (g*2 + t')][lane] = bskf2[((pair*32 + k1)*12 + ((g*2 + t')*2 + c))][lane]
__global__ void bsk2_column_split_kernel(const cplx* __restrict__ bskf2, int64_t rows /* pairs*32*12 */, cplx* __restrict__ keycol) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = idx >> 5;
    const int lane = (int)(idx & 31);
    if (row >= rows) return;
    const int within = (int)(row % 12);          // (g*2 + t')*2 + c
    const int64_t pk = row / 12;                 // pair*32 + k1
    const int c = within & 1, gt = within >> 1;
    const int64_t pair = pk >> 5;
    const int k1 = (int)(pk & 31);
    keycol[((((pair * 2 + c) * 32 + k1) * 6) + gt) * 32 + lane] = bskf2[row * 32 + lane];
}

}  // namespace

__global__ void __launch_bounds__(PP_THREADS, 1)
pbs_kernel_mb2_pair(const cplx* __restrict__ keycol, const uint64_t* __restrict__ in, int64_t B, int n, int beta,
                    const uint64_t* __restrict__ luts, const int32_t* __restrict__ lut_index,
                    const cplx* __restrict__ g_tw, uint64_t* __restrict__ out) {
    using S = PpSmem;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* ring = reinterpret_cast<cplx*>(smem_raw);
    cplx* xa = reinterpret_cast<cplx*>(smem_raw + S::ring_bytes);
    cplx* xb = xa + wfft::XBUF_ELEMS;
    cplx* send = reinterpret_cast<cplx*>(smem_raw + S::ring_bytes + S::x_bytes);          // [2][1024]
    cplx* recv = send + 2 * 1024;                                                          // [2][1024]
    cplx* omega = reinterpret_cast<cplx*>(smem_raw + S::ring_bytes + S::x_bytes + S::spec_bytes);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + S::ring_bytes + S::x_bytes + S::spec_bytes + S::omega_bytes);
    uint64_t* bar_recv = bar_full + PP_SLOTS;                                              // [2]
    unsigned* left = reinterpret_cast<unsigned*>(bar_recv + 2);
    uint16_t* a_tilde = reinterpret_cast<uint16_t*>(smem_raw + S::ring_bytes + S::x_bytes + S::spec_bytes + S::omega_bytes + S::bar_bytes);

    const int u = threadIdx.x, lane = u & 31, wp = u >> 5;
    const int t = (int)cluster_ctarank();
    const uint32_t peer = (uint32_t)(1 - t);
    const int64_t b = blockIdx.x >> 1;
    for (int i = u; i < PW_OMEGA; i += PP_THREADS) omega[i] = g_tw[PW_TILE + i];
    if (u == 0) {
        for (int q = 0; q < PP_SLOTS; ++q) { mbar_init(&bar_full[q], 1); left[q] = 0; }
        mbar_init(&bar_recv[0], 1);
        mbar_init(&bar_recv[1], 1);
        mbar_fence_init();
    }
    const uint64_t* ct = in + (size_t)b * (n + 1);
    for (int i = u; i <= n; i += PP_THREADS) a_tilde[i] = (uint16_t)((((ct[i] >> 51) + 1) >> 1) & 4095);
    __syncthreads();
    cluster_sync_all();

    const int pairs = n >> 1;
    const int total_slices = pairs * wfft::SLICES_PER_STEP;
    constexpr uint32_t SLICE_BYTES = (uint32_t)(wfft::COL_SLICE_ELEMS * 16);
    auto load_slice = [&](int slot, int sidx) {   // slice sidx = 4*pair + q of column t: blocks 8q .. 8q + 7, contiguous
        const int pair = sidx >> 2, q = sidx & 3;
        mbar_expect_tx(&bar_full[slot], SLICE_BYTES);
        tma_load_1d(ring + (size_t)slot * wfft::COL_SLICE_ELEMS,
                    keycol + ((size_t)(pair * 2 + t) * 32 + 8 * q) * wfft::COL_BLOCK_ELEMS, SLICE_BYTES, &bar_full[slot]);
    };
    if (u == 0) {
        for (int q = 0; q < PP_SLOTS && q < total_slices; ++q) load_slice(q, q);
        mbar_expect_tx(&bar_recv[0], PP_SPEC_BYTES);          // the peer's spectra of steps 0 and 1
        if (pairs > 1) mbar_expect_tx(&bar_recv[1], PP_SPEC_BYTES);
    }
    const uint32_t peer_recv = map_to_peer(recv, peer), peer_bar = map_to_peer(bar_recv, peer);

    wfft::Twiddles tw;
    wfft::twiddles_init(tw, u);
    uint64_t acc_re[8], acc_im[8];
    {
        const uint64_t* lut = luts + (size_t)(lut_index ? lut_index[b] : 0) * PW_N;
        const int rot = (4096 - (int)a_tilde[n]) & 4095;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int j = u + 128 * a;
            uint64_t v0 = 0, v1 = 0;
            if (t == 1) {
                int src = (j - rot) & 4095;
                v0 = lut[src & 2047];
                if (src & 2048) v0 = 0 - v0;
                src = (j + 1024 - rot) & 4095;
                v1 = lut[src & 2047];
                if (src & 2048) v1 = 0 - v1;
            }
            acc_re[a] = v0;
            acc_im[a] = v1;
        }
    }

    double re[8], im[8];
    for (int i = 0; i < pairs; ++i) {
        const int par = i & 1;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            re[a] = wfft::top_digit((uint32_t)(acc_re[a] >> 32), beta);
            im[a] = wfft::top_digit((uint32_t)(acc_im[a] >> 32), beta);
        }
        wfft::fwd_stage1(re, im, tw, u, xa);
        __syncthreads();
        wfft::fwd_stage2(tw, u, xa, xb);
        __syncthreads();
        wfft::fwd_stage3(u, xb, re, im);
        // ---- hand the spectrum to the peer
        cplx* sb = send + (size_t)par * 1024;
#pragma unroll
        for (int kL = 0; kL < 8; ++kL) {
            cplx v;
            v.x = re[kL];
            v.y = im[kL];
            sb[kL * wfft::WT + u] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the bulk copy
        __syncthreads();
        if (u == 0) dsmem_bulk_copy(peer_recv + (uint32_t)par * PP_SPEC_BYTES, sb, PP_SPEC_BYTES, peer_bar + (uint32_t)par * 8);
        wfft::Monomials mo;
        wfft::monomials_init(mo, omega, a_tilde[2 * i], a_tilde[2 * i + 1], u);
        // ---- pointwise, part 1: everything that needs only this CTA's spectrum
        double pre[8], pim[8], sre[8], sim[8];
        const int s0 = i * wfft::SLICES_PER_STEP;
        uint32_t ready = mbar_try_wait(&bar_full[s0 % PP_SLOTS], (uint32_t)((s0 / PP_SLOTS) & 1));
#pragma unroll
        for (int q = 0; q < wfft::SLICES_PER_STEP; ++q) {
            const int sidx = s0 + q, slot = sidx % PP_SLOTS;
            if (!ready) mbar_wait(&bar_full[slot], (uint32_t)((sidx / PP_SLOTS) & 1));
            if (q + 1 < wfft::SLICES_PER_STEP)
                ready = mbar_try_wait(&bar_full[(sidx + 1) % PP_SLOTS], (uint32_t)(((sidx + 1) / PP_SLOTS) & 1));
            const cplx* blk0 = ring + (size_t)slot * wfft::COL_SLICE_ELEMS + (size_t)wp * wfft::COL_BLOCK_ELEMS;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int kL = 2 * q + d;
                cplx so, st;
                wfft::pointwise_sums(t, lane, blk0 + (size_t)d * 4 * wfft::COL_BLOCK_ELEMS, mo, so, st);
                pre[kL] = fma(re[kL], so.x, -(im[kL] * so.y));
                pim[kL] = fma(re[kL], so.y, im[kL] * so.x);
                sre[kL] = st.x;
                sim[kL] = st.y;
            }
            __syncwarp();
            if (lane == 0) {       // the last warp to leave the slot refills it (see pbs_kernel_mb2_wide)
                if (atomicAdd(&left[slot], 1u) == PP_WARPS - 1) {
                    left[slot] = 0;
                    const int next = sidx + PP_SLOTS;
                    if (next < total_slices) load_slice(slot, next);
                }
            }
        }
        // ---- part 2: the peer's spectrum
        mbar_wait(&bar_recv[par], (uint32_t)((i >> 1) & 1));
        if (u == 0 && i + 2 < pairs) mbar_expect_tx(&bar_recv[par], PP_SPEC_BYTES);    // arm the phase of step i + 2
        const cplx* rb = recv + (size_t)par * 1024;
#pragma unroll
        for (int kL = 0; kL < 8; ++kL) {
            const cplx fo = rb[kL * wfft::WT + u];
            re[kL] = fma(fo.x, sre[kL], fma(-fo.y, sim[kL], pre[kL]));
            im[kL] = fma(fo.x, sim[kL], fma(fo.y, sre[kL], pim[kL]));
        }
        wfft::inv_stage3(u, re, im, xa);
        __syncthreads();
        wfft::inv_stage2(tw, u, xa, xb);
        __syncthreads();
        wfft::inv_stage1(tw, u, xb, re, im);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            acc_re[a] += wfft::f64_to_torus_u64(re[a]);
            acc_im[a] += wfft::f64_to_torus_u64(im[a]);
        }
    }
    uint64_t* o = out + (size_t)b * ((size_t)PW_N + 1);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int x = u + 128 * a + 1024 * part;
            const uint64_t v = part ? acc_im[a] : acc_re[a];
            if (t == 0) {
                if (x == 0) o[0] = v;
                else o[PW_N - x] = 0 - v;
            } else if (x == 0) {
                o[PW_N] = v;
            }
        }
    }
    cluster_sync_all();     // the peer's last copy out of / into this CTA's shared memory has long completed; be explicit
}

cudaError_t launch_pbs_mb2_pair(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                                const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    if (p.k != 1 || p.l_pbs != 1 || (p.n & 1) || p.N != PW_N || p.beta_pbs < 1 || p.beta_pbs > 31) return cudaErrorInvalidValue;
    if (B > 0x3fffffffLL) return cudaErrorInvalidValue;
    const void* tables = nullptr;
    cudaError_t e = pbs_tables(&tables);
    if (e != cudaSuccess) return e;
    // the key by output column (73 MB at the stated set, a 20 us permutation): scratch from the stream-ordered pool
    const int64_t rows = (int64_t)(p.n / 2) * 32 * 12;
    cplx* keycol = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void**>(&keycol), (size_t)rows * 32 * sizeof(cplx), s);
    if (e != cudaSuccess) return e;
    bsk2_column_split_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const cplx*>(d_bskf2), rows, keycol);
    count_launch();
    const size_t smem = PpSmem::total(p.n);
    e = cudaFuncSetAttribute(pbs_kernel_mb2_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * B));
        cfg.blockDim = dim3(PP_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cplx* kc = keycol;
        const cplx* tb = reinterpret_cast<const cplx*>(tables);
        int n = p.n, beta = p.beta_pbs;
        e = cudaLaunchKernelEx(&cfg, pbs_kernel_mb2_pair, kc, d_in, B, n, beta, d_luts, d_lut_index, tb, d_out);
        count_launch();
    }
    cudaError_t e2 = cudaFreeAsync(keycol, s);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_pbs_mb2_wide(const fhe_b200_pbs_params& p, const double* d_bskf2, const uint64_t* d_in, int64_t B,
                                const uint64_t* d_luts, const int32_t* d_lut_index, uint64_t* d_out, cudaStream_t s) {
    if (B <= 0) return cudaSuccess;
    if (p.k != 1 || p.l_pbs != 1 || (p.n & 1) || p.N != PW_N || p.beta_pbs < 1 || p.beta_pbs > 31) return cudaErrorInvalidValue;
    if (B > 0x7fffffffLL) return cudaErrorInvalidValue;
    const void* tables = nullptr;
    cudaError_t e = pbs_tables(&tables);
    if (e != cudaSuccess) return e;
    const size_t smem = PwSmem::total(p.n);
    e = cudaFuncSetAttribute(pbs_kernel_mb2_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pbs_kernel_mb2_wide<<<(unsigned)B, PW_THREADS, smem, s>>>(reinterpret_cast<const cplx*>(d_bskf2), d_in, B, p.n, p.beta_pbs,
                                                             d_luts, d_lut_index, reinterpret_cast<const cplx*>(tables), d_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fhe
