// pbs_split.cuh -- per-lane arithmetic of one multi-bit blind-rotation step in the TWO-WARPS-PER-POLYNOMIAL form
// (fft_split.cuh): digit extraction for the coefficients a warp owns, the pointwise stage on half-spectra, and the
// accumulator update.  One decomposition level (l_pbs = 1), k = 1, N = 2048 -- the stated 4-bit set.
//
// Used by pbs_kernel_mb2_split (pbs_split.cu).  Everything here is __host__ __device__; tests/test_pbs_split_emul.py
// runs a complete blind rotation through these functions on the CPU (four emulated warps per ciphertext) and checks it
// against the oracle's multi-bit PBS.
//
// Ownership.  Warp (t, h) of a ciphertext: polynomial t of the accumulator, coefficients j = lane + 32*(2m + h)
// and j + 1024 for m = 0..15 (register m of re / im).  In the pointwise stage the same warp produces the bins
// k = lane + 32*k1 of output polynomial t for k1 in [16h, 16h + 16).
//
// Key block of frequency block k1 (what the TMA ring delivers, layout of bsk2_to_fourier_kernel):
//   blk[(((g*2 + t')*1 + 0)*2 + c)*32 + k2]   g = 0..2 (s_a s_b, s_a(1-s_b), (1-s_a)s_b), t' = decomposed polynomial, c = output column
#pragma once
#include "fft_split.cuh"

namespace fhe {
namespace nfft {

constexpr int MB2_BLOCK_ELEMS = 3 * 2 * 1 * 2 * 32;   // complex elements per frequency block (l_pbs = 1)

// Ring order of the key: position pos = 2*s + hh of a pair (slice s = 0..15, block hh = 0..1 of the slice) holds
// frequency block k1 = 16*hh + s, so that half h of every polynomial finds ITS block of step s in slice s.
FHE_HD constexpr int split_ring_block(int pos) { return 16 * (pos & 1) + (pos >> 1); }

// x mod 2^64, rounded to the nearest integer (pbs.cu's f64_to_torus)
FHE_HD uint64_t split_f64_to_torus(double x) {
    const double r = rint(x * 0x1p-64);
    const double y = fma(-r, 0x1p64, x);
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ll_rn(y);
#else
    return (uint64_t)(int64_t)llrint(y);
#endif
}

// balanced beta-bit digit of the top of a torus word, from its HIGH 32 bits (pbs_kernel_mb2, L = 1)
FHE_HD double split_digit(uint32_t hi, int beta) {
    const uint32_t rnd32 = 1u << (31 - beta);
    return (double)((int32_t)(hi + rnd32) >> (32 - beta));
}

// monomial factors c_g = rho_k^{e_g} - 1 at bin k = lane + 32*k1 and the per-frequency-block step r_g = omega^(128 e_g):
// rho_k^e = omega^((4k+1) e) = omega^((4*lane+1) e) * (omega^(128 e))^k1.  c' = c*r + (r - 1) walks k1 -> k1 + 1.
// omega: two-level table of exp(2*pi*i/4096): [0,64) omega^x, [64,128) omega^(64*y)   (pbs.cu's PBS_OMEGA table)
struct SplitMonomials {
    double cx[3], cy[3], rx[3], ry[3], qx[3];
};
FHE_HD void split_monomials_init(SplitMonomials& mo, const cplx* omega, int ea, int eb, int lane, int k1_start) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int e = g == 0 ? ((ea + eb) & 4095) : (g == 1 ? ea : eb);
        const int E = (e * (4 * lane + 1) + 128 * e * k1_start) & 4095;
        const cplx hi = omega[64 + (E >> 6)], lo = omega[E & 63];
        mo.cx[g] = fma(hi.x, lo.x, fma(-hi.y, lo.y, -1.0));
        mo.cy[g] = fma(hi.x, lo.y, hi.y * lo.x);
        const cplx r = omega[64 + (((128 * e) & 4095) >> 6)];
        mo.rx[g] = r.x;
        mo.ry[g] = r.y;
        mo.qx[g] = r.x - 1.0;
    }
}

// one bin of the pointwise stage: G = F_t * sum_g c_g K_g[t][t] + F_t' * sum_g c_g K_g[t'][t] with the key block `blk`
// of this bin's frequency block; advances the monomial factors to the next frequency block.
FHE_HD void split_pointwise_bin(int t, int lane, const cplx fa, const cplx fo, const cplx* blk, SplitMonomials& mo,
                                double& out_re, double& out_im) {
    double kox = 0, koy = 0, ktx = 0, kty = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const cplx bt = blk[((g * 2 + t) * 2 + t) * 32 + lane];
        const cplx bo = blk[((g * 2 + (1 - t)) * 2 + t) * 32 + lane];
        kox = fma(mo.cx[g], bt.x, fma(-mo.cy[g], bt.y, kox));
        koy = fma(mo.cx[g], bt.y, fma(mo.cy[g], bt.x, koy));
        ktx = fma(mo.cx[g], bo.x, fma(-mo.cy[g], bo.y, ktx));
        kty = fma(mo.cx[g], bo.y, fma(mo.cy[g], bo.x, kty));
        const double nx = fma(mo.cx[g], mo.rx[g], fma(-mo.cy[g], mo.ry[g], mo.qx[g]));
        mo.cy[g] = fma(mo.cx[g], mo.ry[g], fma(mo.cy[g], mo.rx[g], mo.ry[g]));
        mo.cx[g] = nx;
    }
    out_re = fma(fa.x, kox, fma(-fa.y, koy, fma(fo.x, ktx, -(fo.y * kty))));
    out_im = fma(fa.x, koy, fma(fa.y, kox, fma(fo.x, kty, fo.y * ktx)));
}

// pointwise stage of warp (t, h): out[kk] = G[lane + 32*(16h + kk)] of output polynomial t, F from the published
// half-spectra of both polynomials.  key_pair: the 32 frequency blocks of this pair of key bits, in k1 order.
FHE_HD void split_pointwise(int t, int h, int lane, const cplx* own0, const cplx* own1, const cplx* oth0, const cplx* oth1,
                            const cplx* key_pair, SplitMonomials& mo, double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        const int k1 = 16 * h + kk;
        split_pointwise_bin(t, lane, split_bin(own0, own1, lane, k1), split_bin(oth0, oth1, lane, k1),
                            key_pair + (size_t)k1 * MB2_BLOCK_ELEMS, mo, re[kk], im[kk]);
    }
}

}  // namespace nfft
}  // namespace fhe
