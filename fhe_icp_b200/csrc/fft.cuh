// fft.cuh -- negacyclic transform of a degree-2048 real polynomial as a 1024-point complex
// FFT (fold + twist), organised as 32 x 32 so that ONE WARP transforms one polynomial:
//
//   j = j1 + 32*j2 (time),  k = k2 + 32*k1 (frequency),  W = exp(2*pi*i/1024)
//   pass 1  lane j1 : 32-point DFT over j2 held in registers           -> index k2
//           multiply by W^(j1*k2) * omega^j1   (omega = exp(2*pi*i/4096), the twist)
//   transpose through a 16 KB XOR-swizzled shared-memory tile (the only exchange)
//   pass 2  lane k2 : 32-point DFT over j1 held in registers           -> index k1
//
// All passes are radix-2 DIT butterflies (fused with their twiddle product); the forward passes
// run them on a bit-reversed view of the registers (natural in, bit-reversed out), the inverse
// passes directly (bit-reversed in, natural out), so no permutation is ever executed: "bit
// reversal" is a compile-time renaming of registers.  Bins come out in natural order k = lane + 32*k1,
// which is also the layout of the Fourier-domain bootstrapping key.
//
// The per-lane code is __host__ __device__ so that tests/ can emulate a warp on the CPU
// and check the index algebra against the oracle without a GPU.
#pragma once
#include <stdint.h>

#include <cmath>

#if defined(__CUDACC__)
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif

#ifndef FHE_FFT_PREFETCH
#define FHE_FFT_PREFETCH 4  // twiddle loads run this many elements ahead of their use
#endif

namespace fhe {
namespace nfft {

constexpr int NPOLY = 2048;  // polynomial size N
constexpr int M = 1024;      // complex FFT size N/2
constexpr int R = 32;        // radix of each in-register pass = lanes per warp

struct cplx { double x, y; };

#define FHE_W32_RE_LIST                                                                                   \
    0x1.0000000000000p+0, 0x1.f6297cff75cb0p-1, 0x1.d906bcf328d46p-1, 0x1.a9b66290ea1a3p-1,              \
    0x1.6a09e667f3bcdp-1, 0x1.1c73b39ae68c8p-1, 0x1.87de2a6aea963p-2, 0x1.8f8b83c69a60bp-3, 0.0,         \
    -0x1.8f8b83c69a60bp-3, -0x1.87de2a6aea963p-2, -0x1.1c73b39ae68c8p-1, -0x1.6a09e667f3bcdp-1,          \
    -0x1.a9b66290ea1a3p-1, -0x1.d906bcf328d46p-1, -0x1.f6297cff75cb0p-1
#define FHE_W32_IM_LIST                                                                                   \
    0.0, 0x1.8f8b83c69a60bp-3, 0x1.87de2a6aea963p-2, 0x1.1c73b39ae68c8p-1, 0x1.6a09e667f3bcdp-1,         \
    0x1.a9b66290ea1a3p-1, 0x1.d906bcf328d46p-1, 0x1.f6297cff75cb0p-1, 1.0, 0x1.f6297cff75cb0p-1,         \
    0x1.d906bcf328d46p-1, 0x1.a9b66290ea1a3p-1, 0x1.6a09e667f3bcdp-1, 0x1.1c73b39ae68c8p-1,              \
    0x1.87de2a6aea963p-2, 0x1.8f8b83c69a60bp-3
// exp(2*pi*i*j2/128), j2 = 0..31: the j2-dependent factor omega^(32*j2) of the twist
#define FHE_C128_RE_LIST                                                                                  \
    0x1.0000000000000p+0, 0x1.ff621e3796d7ep-1, 0x1.fd88da3d12526p-1, 0x1.fa7557f08a517p-1,              \
    0x1.f6297cff75cb0p-1, 0x1.f0a7efb9230d7p-1, 0x1.e9f4156c62ddap-1, 0x1.e212104f686e5p-1,              \
    0x1.d906bcf328d46p-1, 0x1.ced7af43cc773p-1, 0x1.c38b2f180bdb1p-1, 0x1.b728345196e3ep-1,              \
    0x1.a9b66290ea1a3p-1, 0x1.9b3e047f38741p-1, 0x1.8bc806b151741p-1, 0x1.7b5df226aafafp-1,              \
    0x1.6a09e667f3bcdp-1, 0x1.57d69348ceca0p-1, 0x1.44cf325091dd6p-1, 0x1.30ff7fce17035p-1,              \
    0x1.1c73b39ae68c8p-1, 0x1.073879922ffeep-1, 0x1.e2b5d3806f63bp-2, 0x1.b5d1009e15cc0p-2,              \
    0x1.87de2a6aea963p-2, 0x1.58f9a75ab1fddp-2, 0x1.294062ed59f06p-2, 0x1.f19f97b215f1bp-3,              \
    0x1.8f8b83c69a60bp-3, 0x1.2c8106e8e613ap-3, 0x1.917a6bc29b42cp-4, 0x1.91f65f10dd814p-5
#define FHE_C128_IM_LIST                                                                                  \
    0.0, 0x1.91f65f10dd814p-5, 0x1.917a6bc29b42cp-4, 0x1.2c8106e8e613ap-3, 0x1.8f8b83c69a60bp-3,         \
    0x1.f19f97b215f1bp-3, 0x1.294062ed59f06p-2, 0x1.58f9a75ab1fddp-2, 0x1.87de2a6aea963p-2,              \
    0x1.b5d1009e15cc0p-2, 0x1.e2b5d3806f63bp-2, 0x1.073879922ffeep-1, 0x1.1c73b39ae68c8p-1,              \
    0x1.30ff7fce17035p-1, 0x1.44cf325091dd6p-1, 0x1.57d69348ceca0p-1, 0x1.6a09e667f3bcdp-1,              \
    0x1.7b5df226aafafp-1, 0x1.8bc806b151741p-1, 0x1.9b3e047f38741p-1, 0x1.a9b66290ea1a3p-1,              \
    0x1.b728345196e3ep-1, 0x1.c38b2f180bdb1p-1, 0x1.ced7af43cc773p-1, 0x1.d906bcf328d46p-1,              \
    0x1.e212104f686e5p-1, 0x1.e9f4156c62ddap-1, 0x1.f0a7efb9230d7p-1, 0x1.f6297cff75cb0p-1,              \
    0x1.fa7557f08a517p-1, 0x1.fd88da3d12526p-1, 0x1.ff621e3796d7ep-1

#if defined(__CUDACC__)
__constant__ double d_W32_RE[16] = {FHE_W32_RE_LIST};
__constant__ double d_W32_IM[16] = {FHE_W32_IM_LIST};
__constant__ double d_C128_RE[32] = {FHE_C128_RE_LIST};
__constant__ double d_C128_IM[32] = {FHE_C128_IM_LIST};
#endif
static const double h_W32_RE[16] = {FHE_W32_RE_LIST};
static const double h_W32_IM[16] = {FHE_W32_IM_LIST};
static const double h_C128_RE[32] = {FHE_C128_RE_LIST};
static const double h_C128_IM[32] = {FHE_C128_IM_LIST};

#if defined(__CUDA_ARCH__)
#define FHE_W32_RE(i) d_W32_RE[i]
#define FHE_W32_IM(i) d_W32_IM[i]
#define FHE_C128_RE(i) d_C128_RE[i]
#define FHE_C128_IM(i) d_C128_IM[i]
#else
#define FHE_W32_RE(i) h_W32_RE[i]
#define FHE_W32_IM(i) h_W32_IM[i]
#define FHE_C128_RE(i) h_C128_RE[i]
#define FHE_C128_IM(i) h_C128_IM[i]
#endif

FHE_HD constexpr int brev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// slot of element (row, col) of the 32x32 transpose tile of 16-byte elements.  Rows are padded to
// 33 elements: a warp writing one row or reading one column touches every bank group exactly once,
// and -- unlike an XOR swizzle -- the address is base(lane) + compile-time offset, so the unrolled
// code needs no per-element address registers (the XOR version spilled them, ncu r1_ncu_pbs_v3).
constexpr int TILE_PITCH = 33;
constexpr int TILE_ELEMS = 32 * TILE_PITCH;  // 1056 elements = 16,896 bytes
FHE_HD constexpr int slot(int row, int col) { return row * TILE_PITCH + col; }

// 32-point DFT, radix-2 decimation in time, fully unrolled in registers.  The input for index q
// sits at LOGICAL position brev5(q), the output for index f at logical position f.  With
// PERM = true logical position i lives in register brev5(i): natural-order input registers, output
// f in register brev5(f) (what a decimation-in-frequency pass would leave) -- so both transform
// directions use DIT butterflies and "bit reversal" stays a compile-time renaming of registers.
// DIT is chosen because its butterfly fuses with the twiddle product: y0 = a + w*b costs 4 FMAs and
// y1 = 2a - y0 two more (6 FP64 instructions instead of 8 for multiply-then-add/sub).
template <int SIGN, bool PERM>
FHE_HD void dit32(double (&re)[32], double (&im)[32]) {
#pragma unroll
    for (int half = 1; half <= 16; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int a = PERM ? brev5(base + j) : base + j;
                const int b = PERM ? brev5(base + j + half) : base + j + half;
                const int tw = j * (16 / half);
                const double ar = re[a], ai = im[a], br = re[b], bi = im[b];
                if (tw == 0) {
                    re[a] = ar + br;
                    im[a] = ai + bi;
                    re[b] = ar - br;
                    im[b] = ai - bi;
                } else if (tw == 8) {  // w = +-i
                    if (SIGN > 0) {
                        re[a] = ar - bi; im[a] = ai + br;
                        re[b] = ar + bi; im[b] = ai - br;
                    } else {
                        re[a] = ar + bi; im[a] = ai - br;
                        re[b] = ar - bi; im[b] = ai + br;
                    }
                } else {
                    const double wr = FHE_W32_RE(tw), wi = SIGN > 0 ? FHE_W32_IM(tw) : -FHE_W32_IM(tw);
                    const double y0r = fma(-bi, wi, fma(br, wr, ar));
                    const double y0i = fma(bi, wr, fma(br, wi, ai));
                    re[a] = y0r;
                    im[a] = y0i;
                    re[b] = fma(2.0, ar, -y0r);
                    im[b] = fma(2.0, ai, -y0i);
                }
            }
        }
    }
}

// One 16 KB twiddle table serves both directions: tw[slot(k2, j1)] = W^(j1*k2) * omega^j1.
// The forward pass reads row k2 across lanes j1, the inverse pass reads, per lane k2, the
// conjugates along j1; the row padding makes both patterns bank-conflict free.
//
// ---- forward: registers hold z[j2] = (c[j] + i*c[j+1024]) for j = lane + 32*j2 (the caller
// has NOT applied any twist).  After phase 2 register brev5(k1) holds bin k = lane + 32*k1.
FHE_HD void fwd_phase1(double (&re)[32], double (&im)[32], const cplx* tw, cplx* buf, int lane) {
#pragma unroll
    for (int j2 = 1; j2 < 32; ++j2) {  // twist factor omega^(32*j2)
        const double cr = FHE_C128_RE(j2), ci = FHE_C128_IM(j2);
        const double a = re[j2], b = im[j2];
        re[j2] = a * cr - b * ci;
        im[j2] = a * ci + b * cr;
    }
    dit32<+1, true>(re, im);
    // twiddle loads run PF elements ahead of their use (explicit software prefetch: with one or two
    // warps per scheduler the shared-memory latency is otherwise exposed at every multiply)
    constexpr int PF = FHE_FFT_PREFETCH;
    cplx w[PF];
#pragma unroll
    for (int p = 0; p < PF; ++p) w[p] = tw[slot(brev5(p), lane)];
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const int k2 = brev5(p);
        const cplx wc = w[p % PF];
        if (p + PF < 32) w[p % PF] = tw[slot(brev5(p + PF), lane)];
        cplx v;
        v.x = re[p] * wc.x - im[p] * wc.y;
        v.y = re[p] * wc.y + im[p] * wc.x;
        buf[slot(k2, lane)] = v;
    }
}
FHE_HD void fwd_phase2(double (&re)[32], double (&im)[32], const cplx* buf, int lane) {
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) {
        const cplx v = buf[slot(lane, j1)];
        re[j1] = v.x;
        im[j1] = v.y;
    }
    dit32<+1, true>(re, im);
}

// ---- inverse: register brev5(k1) holds bin k = lane + 32*k1.  After phase 2 register j2 holds
// (c[j] + i*c[j+1024]) for j = lane + 32*j2, fully untwisted and scaled by 1/1024.
FHE_HD void inv_phase1(double (&re)[32], double (&im)[32], const cplx* tw, cplx* buf, int lane) {
    dit32<-1, false>(re, im);
    constexpr int PF = FHE_FFT_PREFETCH;
    cplx w[PF];
#pragma unroll
    for (int j1 = 0; j1 < PF; ++j1) w[j1] = tw[slot(lane, j1)];
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) {
        const cplx wc = w[j1 % PF];  // multiply by conj(w)
        if (j1 + PF < 32) w[j1 % PF] = tw[slot(lane, j1 + PF)];
        cplx v;
        v.x = re[j1] * wc.x + im[j1] * wc.y;
        v.y = im[j1] * wc.x - re[j1] * wc.y;
        buf[slot(j1, lane)] = v;
    }
}
FHE_HD void inv_phase2(double (&re)[32], double (&im)[32], const cplx* buf, int lane) {
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const cplx v = buf[slot(lane, brev5(p))];
        re[p] = v.x;
        im[p] = v.y;
    }
    dit32<-1, false>(re, im);
    re[0] *= 0x1p-10;
    im[0] *= 0x1p-10;
#pragma unroll
    for (int j2 = 1; j2 < 32; ++j2) {  // conj(omega^(32*j2)) / 1024 (power-of-two scale: exact)
        const double cr = FHE_C128_RE(j2) * 0x1p-10, ci = -FHE_C128_IM(j2) * 0x1p-10;
        const double a = re[j2], b = im[j2];
        re[j2] = a * cr - b * ci;
        im[j2] = a * ci + b * cr;
    }
}

// host: the inter-pass twiddle table (computed in long double, rounded once)
inline void fill_twiddle_table(cplx* tw) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k2 = 0; k2 < 32; ++k2) {
        for (int j1 = 0; j1 < 32; ++j1) {
            long double ang = two_pi * ((long double)(j1 * k2) / 1024.0L + (long double)j1 / 4096.0L);
            tw[slot(k2, j1)].x = (double)cosl(ang);
            tw[slot(k2, j1)].y = (double)sinl(ang);
        }
    }
}

}  // namespace nfft
}  // namespace fhe
