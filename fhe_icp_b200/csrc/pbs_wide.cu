// pbs_wide.cu -- multi-bit blind rotation with FOUR WARPS PER POLYNOMIAL: one ciphertext per CTA, 256 threads, 8 complex
// points per thread (pbs_wide.cuh).  The latency kernel: launch_pbs_mb2 dispatches it while there is at most one
// ciphertext per SM (the step of a lone ciphertext is a dependent chain; eight warps walk it 2.4x faster than the two of
// pbs_kernel_mb2 -- and 1.65x faster than the four warps of the round-2 "split" kernel this one replaced).
//
// Thread tid: polynomial t = tid >> 7, u = tid & 127.  It owns the accumulator coefficients j = u + 128a and j + 1024
// (a = 0..7) of polynomial t IN REGISTERS (16 torus words) and, in the Fourier domain, the bins k = u + 128 kL of
// output polynomial t.  Twiddles are per-thread constants of the launch, held in registers.
//
// Shared memory: two exchange buffers per polynomial (17 KB each), written alternately -- a step has five exchanges
//   stage1 -> [bar t] -> stage2 -> [bar t] -> stage3, spectrum -> [bar all] -> pointwise, inverse stage3 -> [bar all] ->
//   inverse stage2 -> [bar t] -> inverse stage1, accumulate
// and because five is odd the alternation simply continues into the next step: a buffer is rewritten only after a
// barrier that every one of its readers has passed (the two 256-thread barriers are the ones around the pointwise
// stage, where each polynomial reads the other's spectrum).  tests/emul/pbs_wide_emul.cpp replays this plan in several
// thread orders.
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
    __syncthreads();

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

    cplx* xa = xbufs + (size_t)(2 * t) * wfft::XBUF_ELEMS;              // this polynomial's two exchange buffers
    cplx* xb = xa + wfft::XBUF_ELEMS;
    const cplx* oa = xbufs + (size_t)(2 * (1 - t)) * wfft::XBUF_ELEMS;  // the other polynomial's
    const cplx* ob = oa + wfft::XBUF_ELEMS;

#ifdef WIDE_TIMING      // debug builds: clocks per phase of a step, printed by one warp of each polynomial of CTA 0
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define WIDE_TICK(k) { const unsigned long long now = clock64(); tacc[k] += now - tprev; tprev = now; }
#else
#define WIDE_TICK(k)
#endif
    double re[8], im[8];
    for (int i = 0; i < pairs; ++i) {
        // exchange buffers of this step in write order: w0 r0 w1 r1 w2(spectrum) ...; five per step, so the roles swap
        cplx* e0 = (i & 1) ? xb : xa;
        cplx* e1 = (i & 1) ? xa : xb;
        const cplx* o0 = (i & 1) ? ob : oa;
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
#pragma unroll
        for (int kL = 0; kL < 8; ++kL) {          // publish the spectrum: the other polynomial's pointwise stage reads it
            cplx v;
            v.x = re[kL];
            v.y = im[kL];
            e0[kL * wfft::WT + u] = v;
        }
        wfft::Monomials mo;
        wfft::monomials_init(mo, omega, a_tilde[2 * i], a_tilde[2 * i + 1], u);
        __syncthreads();
        WIDE_TICK(2)
        // ---- pointwise stage
        double gre[8], gim[8];
        uint32_t ready = mbar_try_wait(&bar_full[(i * wfft::SLICES_PER_STEP) % PW_SLOTS],
                                       (uint32_t)(((i * wfft::SLICES_PER_STEP) / PW_SLOTS) & 1));
#pragma unroll
        for (int q = 0; q < wfft::SLICES_PER_STEP; ++q) {     // two bins (one ring slice) per turn
            const int sidx = i * wfft::SLICES_PER_STEP + q;
            const int slot = sidx % PW_SLOTS;
            const cplx fo0 = o0[(2 * q) * wfft::WT + u], fo1 = o0[(2 * q + 1) * wfft::WT + u];
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
        wfft::inv_stage3(u, gre, gim, e1);
        __syncthreads();
        WIDE_TICK(4)                               // also: nobody reads the published spectra any more
        wfft::inv_stage2(tw, u, e1, e0);
        pw_bar_poly(t);
        WIDE_TICK(5)
        wfft::inv_stage1(tw, u, e0, re, im);
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
