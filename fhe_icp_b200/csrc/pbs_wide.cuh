// pbs_wide.cuh -- negacyclic transform and blind-rotation step with FOUR WARPS PER POLYNOMIAL (128 threads, 8 complex
// points per thread): the per-thread arithmetic of pbs_kernel_mb2_wide (pbs_wide.cu), the latency kernel for small
// batches -- one ciphertext per CTA, 256 threads.  A lone ciphertext's blind-rotation step is a dependent chain
// (digits -> forward transform -> pointwise -> inverse transform -> accumulate); the chain gets shorter with the
// number of threads that share a transform, until the exchanges through shared memory cost more than the butterflies
// (4 points per thread: five exchanges per transform, shared-memory bound).  8 points per thread need two.
//
// 1024-point complex FFT as 8 x 8 x 16, decimation in frequency, W = exp(2*pi*i/1024), omega = exp(2*pi*i/4096):
//   j = 128a + 16b + c (time),  k = ka + 8kb + 64kc (frequency),  kc = kH + 2kL,  c = cL + 8cH
//   F[k] = sum_c W16^(c kc) W128^(c kb) W1024^(c ka)  sum_b W8^(b kb) W64^(b ka)  sum_a W8^(a ka) z_j omega^j
//   stage 1  thread u = 16b + c : z_(u+128a) W32^a (constants) -> DFT8 over a -> * omega^(u(4ka+1))   [twist folded in]
//   exchange 1 (natural layout [ka][u])
//   stage 2  thread 16ka + c    : DFT8 over b -> * W128^(c kb)
//   exchange 2 (rows p = ka + 8kb of 16 elements, pitch 17)
//   stage 3  thread p + 64kH    : reads the 16 elements of row p, radix-2 over cH for ITS kH, * W16^(cL kH), DFT8 over cL
// Thread v of stage 3 ends with the bins k = v + 128 kL, kL = 0..7 -- the same shape as the time-domain ownership
// j = u + 128a, so the pointwise stage and the inverse transform (the mirror image, conjugate twiddles, 1/1024 and the
// untwist folded into the last stage) need no further permutation.
//
// Everything is __host__ __device__: tests/emul/pbs_wide_emul.cpp runs the transforms and a whole blind rotation on the
// CPU, phase by phase in several thread orders (a missing barrier would make the result order dependent).
#pragma once
#include "fft.cuh"

namespace fhe {
namespace wfft {

using nfft::cplx;
using nfft::h_W32_IM;
using nfft::h_W32_RE;
#if defined(__CUDACC__)
using nfft::d_W32_IM;
using nfft::d_W32_RE;
#endif

constexpr int WT = 128;                    // threads per polynomial
constexpr int MB2_BLOCK_ELEMS = 3 * 2 * 1 * 2 * 32;   // complex elements per frequency block (l_pbs = 1)

// x mod 2^64, rounded to the nearest integer (pbs.cu's f64_to_torus)
FHE_HD uint64_t f64_to_torus_u64(double x) {
    const double r = rint(x * 0x1p-64);
    const double y = fma(-r, 0x1p64, x);
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ll_rn(y);
#else
    return (uint64_t)(int64_t)llrint(y);
#endif
}

// balanced beta-bit digit of the top of a torus word, from its HIGH 32 bits (pbs_kernel_mb2, L = 1)
FHE_HD double top_digit(uint32_t hi, int beta) {
    const uint32_t rnd32 = 1u << (31 - beta);
    return (double)((int32_t)(hi + rnd32) >> (32 - beta));
}

// monomial factors c_g = rho_k^(e_g) - 1 of a bin and the factor r_g that takes them to the thread's next bin:
// rho_(k+128)^e = rho_k^e * omega^(512 e), so c' = c*r + (r - 1).
struct Monomials {
    double cx[3], cy[3], rx[3], ry[3], qx[3];
};

// one bin of the pointwise stage: G = F_t * sum_g c_g K_g[t][t] + F_t' * sum_g c_g K_g[t'][t] with the key block `blk`
// of this bin's frequency block (layout of bsk2_to_fourier_kernel: blk[((g*2 + t')*2 + c)*32 + lane], g = 0..2 for
// s_a s_b, s_a(1-s_b), (1-s_a)s_b, t' = decomposed polynomial, c = output column); advances the monomial factors.
FHE_HD void pointwise_bin(int t, int lane, const cplx fa, const cplx fo, const cplx* blk, Monomials& mo, double& out_re,
                          double& out_im) {
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

constexpr int PITCH = 17;                  // row pitch of the exchange-2 layout (16 elements + 1: conflict-free columns)
constexpr int XBUF_ELEMS = 64 * PITCH;     // one exchange buffer: 1088 elements = 17,408 B (the natural layout uses 1024)
constexpr double RSQRT2 = 0x1.6a09e667f3bcdp-1;

// y[k] = sum_a x[a] exp(SIGN * 2*pi*i * a*k / 8), natural order in and out (radix-2 decimation in frequency; the
// reordering is a compile-time renaming of registers)
template <int SIGN>
FHE_HD void dft8(double (&re)[8], double (&im)[8]) {
    double sr[4], si[4], dr[4], di[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        sr[a] = re[a] + re[a + 4];
        si[a] = im[a] + im[a + 4];
        dr[a] = re[a] - re[a + 4];
        di[a] = im[a] - im[a + 4];
    }
    {   // d[a] *= w8^a, w8 = exp(SIGN*2*pi*i/8)
        const double x1 = dr[1], y1 = di[1], x2 = dr[2], y2 = di[2], x3 = dr[3], y3 = di[3];
        if (SIGN > 0) {
            dr[1] = (x1 - y1) * RSQRT2; di[1] = (x1 + y1) * RSQRT2;
            dr[2] = -y2;                di[2] = x2;
            dr[3] = (-x3 - y3) * RSQRT2; di[3] = (x3 - y3) * RSQRT2;
        } else {
            dr[1] = (x1 + y1) * RSQRT2; di[1] = (y1 - x1) * RSQRT2;
            dr[2] = y2;                 di[2] = -x2;
            dr[3] = (y3 - x3) * RSQRT2; di[3] = (-x3 - y3) * RSQRT2;
        }
    }
    // two 4-point transforms: sums -> even outputs, rotated differences -> odd outputs
#pragma unroll
    for (int hsel = 0; hsel < 2; ++hsel) {
        double (&ur)[4] = hsel == 0 ? sr : dr;
        double (&ui)[4] = hsel == 0 ? si : di;
        const double s0r = ur[0] + ur[2], s0i = ui[0] + ui[2], s1r = ur[1] + ur[3], s1i = ui[1] + ui[3];
        const double d0r = ur[0] - ur[2], d0i = ui[0] - ui[2];
        double d1r = ur[1] - ur[3], d1i = ui[1] - ui[3];
        {   // d1 *= w4 = SIGN * i
            const double x = d1r, y = d1i;
            d1r = SIGN > 0 ? -y : y;
            d1i = SIGN > 0 ? x : -x;
        }
        // 4-point outputs k' = 0, 1, 2, 3  ->  8-point outputs 2k' + hsel
        re[0 + hsel] = s0r + s1r; im[0 + hsel] = s0i + s1i;
        re[2 + hsel] = d0r + d1r; im[2 + hsel] = d0i + d1i;
        re[4 + hsel] = s0r - s1r; im[4 + hsel] = s0i - s1i;
        re[6 + hsel] = d0r - d1r; im[6 + hsel] = d0i - d1i;
    }
}

FHE_HD void cmul(double& x, double& y, double wr, double wi) {
    const double a = x, b = y;
    x = fma(a, wr, -(b * wi));
    y = fma(a, wi, b * wr);
}

// exp(2*pi*i * num / den): the per-thread twiddles, computed once per launch
FHE_HD void unit_root(int num, int den, double& x, double& y) {
#if defined(__CUDA_ARCH__)
    sincospi(2.0 * (double)num / (double)den, &y, &x);
#else
    const long double ang = 6.283185307179586476925286766559005768L * (long double)num / (long double)den;
    x = (double)cosl(ang);
    y = (double)sinl(ang);
#endif
}

// loop-invariant twiddles of thread u: tw1[ka] = omega^(u(4ka+1)) (stage 1 / inverse stage 1, u = 16b + c),
// tw2[kb] = W128^(c kb) (stage 2 / inverse stage 2, c = u & 15)
struct Twiddles {
    double t1x[8], t1y[8], t2x[8], t2y[8];
};
FHE_HD void twiddles_init(Twiddles& tw, int u) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unit_root((u * (4 * q + 1)) & 4095, 4096, tw.t1x[q], tw.t1y[q]);
        unit_root(((u & 15) * q) & 127, 128, tw.t2x[q], tw.t2y[q]);
    }
}

// ---- forward.  re/im[a] = z_(u+128a) = (c[j] + i c[j+1024]), j = u + 128a, untwisted.
FHE_HD void fwd_stage1(double (&re)[8], double (&im)[8], const Twiddles& tw, int u, cplx* x1) {
#pragma unroll
    for (int a = 1; a < 8; ++a) cmul(re[a], im[a], FHE_W32_RE(a), FHE_W32_IM(a));     // omega^(128a) = W32^a
    dft8<+1>(re, im);
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
        cmul(re[ka], im[ka], tw.t1x[ka], tw.t1y[ka]);
        cplx v;
        v.x = re[ka];
        v.y = im[ka];
        x1[ka * WT + u] = v;
    }
}
FHE_HD void fwd_stage2(const Twiddles& tw, int v, const cplx* x1, cplx* x2) {
    const int ka = v >> 4, c = v & 15;
    double re[8], im[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const cplx e = x1[ka * WT + b * 16 + c];
        re[b] = e.x;
        im[b] = e.y;
    }
    dft8<+1>(re, im);
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        if (kb) cmul(re[kb], im[kb], tw.t2x[kb], tw.t2y[kb]);
        cplx e;
        e.x = re[kb];
        e.y = im[kb];
        x2[(ka + 8 * kb) * PITCH + c] = e;
    }
}
// thread v = p + 64 kH; on return re/im[kL] = bin v + 128 kL
FHE_HD void fwd_stage3(int v, const cplx* x2, double (&re)[8], double (&im)[8]) {
    const int p = v & 63, kH = v >> 6;
#pragma unroll
    for (int cL = 0; cL < 8; ++cL) {
        const cplx lo = x2[p * PITCH + cL], hi = x2[p * PITCH + cL + 8];
        if (kH == 0) {
            re[cL] = lo.x + hi.x;
            im[cL] = lo.y + hi.y;
        } else {
            re[cL] = lo.x - hi.x;
            im[cL] = lo.y - hi.y;
            if (cL) cmul(re[cL], im[cL], FHE_W32_RE(2 * cL), FHE_W32_IM(2 * cL));       // W16^cL
        }
    }
    dft8<+1>(re, im);
}

// ---- inverse.  re/im[kL] = bin v + 128 kL of thread v = p + 64 kH.
FHE_HD void inv_stage3(int v, double (&re)[8], double (&im)[8], cplx* x2) {
    const int p = v & 63, kH = v >> 6;
    dft8<-1>(re, im);
#pragma unroll
    for (int cL = 0; cL < 8; ++cL) {
        if (kH && cL) cmul(re[cL], im[cL], FHE_W32_RE(2 * cL), -FHE_W32_IM(2 * cL));
        cplx e;
        e.x = re[cL];
        e.y = im[cL];
        x2[p * PITCH + cL + 8 * kH] = e;
    }
}
FHE_HD void inv_stage2(const Twiddles& tw, int v, const cplx* x2, cplx* x1) {
    const int ka = v >> 4, c = v & 15, cL = c & 7, cH = c >> 3;
    double re[8], im[8];
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        const cplx t0 = x2[(ka + 8 * kb) * PITCH + cL], t1 = x2[(ka + 8 * kb) * PITCH + cL + 8];
        re[kb] = cH ? t0.x - t1.x : t0.x + t1.x;
        im[kb] = cH ? t0.y - t1.y : t0.y + t1.y;
        if (kb) cmul(re[kb], im[kb], tw.t2x[kb], -tw.t2y[kb]);
    }
    dft8<-1>(re, im);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        cplx e;
        e.x = re[b];
        e.y = im[b];
        x1[ka * WT + b * 16 + c] = e;
    }
}
// on return re[a] / im[a] = coefficients j = u + 128a / j + 1024, untwisted and scaled by 1/1024
FHE_HD void inv_stage1(const Twiddles& tw, int u, const cplx* x1, double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
        const cplx e = x1[ka * WT + u];
        re[ka] = e.x;
        im[ka] = e.y;
        cmul(re[ka], im[ka], tw.t1x[ka], -tw.t1y[ka]);
    }
    dft8<-1>(re, im);
    re[0] *= 0x1p-10;
    im[0] *= 0x1p-10;
#pragma unroll
    for (int a = 1; a < 8; ++a) cmul(re[a], im[a], FHE_W32_RE(a) * 0x1p-10, -FHE_W32_IM(a) * 0x1p-10);
}

// ---- pointwise stage of thread v of output polynomial t: bins k = v + 128 kL.  The key of bin k sits in frequency
// block k >> 5 = (v >> 5) + 4 kL at lane v & 31 (layout of bsk2_to_fourier_kernel); ring slice q
// holds the eight consecutive blocks 8q .. 8q + 7, i.e. the bins kL = 2q and 2q + 1 of every thread.
constexpr int SLICE_BLOCKS = 8;                                      // two bins per thread (kL = 2q, 2q + 1)
constexpr int SLICE_ELEMS = SLICE_BLOCKS * MB2_BLOCK_ELEMS;    // 3072 complex = 48 KB
constexpr int SLICES_PER_STEP = 4;

// monomial factors c_g = rho_k^(e_g) - 1 at k = v (kL = 0) and the step r_g = omega^(512 e_g) = exp(2*pi*i*e_g/8)
// from one slice to the next: rho_k^e = omega^((4k+1)e).  omega: pbs.cu's two-level table ([0,64) omega^x, [64,128)
// omega^(64y)).
FHE_HD void monomials_init(Monomials& mo, const cplx* omega, int ea, int eb, int v) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int e = g == 0 ? ((ea + eb) & 4095) : (g == 1 ? ea : eb);
        const int E = (e * (4 * v + 1)) & 4095;
        const cplx hi = omega[64 + (E >> 6)], lo = omega[E & 63];
        mo.cx[g] = fma(hi.x, lo.x, fma(-hi.y, lo.y, -1.0));
        mo.cy[g] = fma(hi.x, lo.y, hi.y * lo.x);
        const cplx r = omega[64 + (((512 * e) & 4095) >> 6)];
        mo.rx[g] = r.x;
        mo.ry[g] = r.y;
        mo.qx[g] = r.x - 1.0;
    }
}

// ---- the same stage for ONE output column held apart (pbs_kernel_mb2_pair: a cluster of two CTAs per ciphertext, CTA t
// owns polynomial t and streams only column t of the key).  Column layout, written by bsk2_column_split_kernel:
//   keycol[((pair*2 + c)*32 + k1)*6 + (g*2 + t')][32 lanes]
// so that the bins of slice q (frequency blocks 8q .. 8q + 7) of column c are 24 KB contiguous.
constexpr int COL_BLOCK_ELEMS = 3 * 2 * 32;                          // complex elements per (column, frequency block)
constexpr int COL_SLICE_ELEMS = SLICE_BLOCKS * COL_BLOCK_ELEMS;      // 1536 complex = 24 KB

// the two key combinations of a bin, S_own = sum_g c_g K_g[t][t] and S_oth = sum_g c_g K_g[1-t][t]; advances the
// monomials.  The output bin is F_t * S_own + F_(1-t) * S_oth: the first product needs nothing from the other CTA.
FHE_HD void pointwise_sums(int t, int lane, const cplx* blkc, Monomials& mo, cplx& s_own, cplx& s_oth) {
    double kox = 0, koy = 0, ktx = 0, kty = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const cplx bt = blkc[(g * 2 + t) * 32 + lane];
        const cplx bo = blkc[(g * 2 + (1 - t)) * 32 + lane];
        kox = fma(mo.cx[g], bt.x, fma(-mo.cy[g], bt.y, kox));
        koy = fma(mo.cx[g], bt.y, fma(mo.cy[g], bt.x, koy));
        ktx = fma(mo.cx[g], bo.x, fma(-mo.cy[g], bo.y, ktx));
        kty = fma(mo.cx[g], bo.y, fma(mo.cy[g], bo.x, kty));
        const double nx = fma(mo.cx[g], mo.rx[g], fma(-mo.cy[g], mo.ry[g], mo.qx[g]));
        mo.cy[g] = fma(mo.cx[g], mo.ry[g], fma(mo.cy[g], mo.rx[g], mo.ry[g]));
        mo.cx[g] = nx;
    }
    s_own.x = kox; s_own.y = koy;
    s_oth.x = ktx; s_oth.y = kty;
}

}  // namespace wfft
}  // namespace fhe
