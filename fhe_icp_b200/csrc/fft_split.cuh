// fft_split.cuh -- the 32x32 negacyclic FFT of fft.cuh with every 32-point in-register DFT split by one radix-2
// step between TWO warps, so that a polynomial is transformed by 64 lanes holding 16 complex points each
// (~64 data registers per thread instead of 128: four warps per scheduler instead of two).
//
// Used by pbs_kernel_mb2_split (pbs_split.cu).  __host__ __device__: tests/test_fft_split_emul.py emulates the two
// warps on the CPU against numpy and against the one-warp transform.
//
//   j = j1 + 32*j2 (time),  k = k2 + 32*k1 (frequency),  W = exp(2*pi*i/1024), w32 = W^32, omega = exp(2*pi*i/4096)
//
// forward (decimation in time at both levels; warp h owns the samples / rows of parity h):
//   pass 1, warp h, lane j1 : a_m = z[j1 + 32(2m+h)] * omega^(32(2m+h)),  Y_h[q] = sum_m a_m w16^(mq)       (16 points)
//                             tile E[q][j1] = Y_0[q]   |   tile O[q][j1] = w32^q * Y_1[q]
//                             -- the closing butterfly X[q + 16s] = E[q] +- O[q] is NOT executed here --
//   pass 2, warp h', lane k2 = q + 16s : for j1 = 2m + h':  x = E[q][j1] +- O[q][j1]   (the deferred butterfly: one
//                             add per element, the sign is per lane),  b_m = x * W^(j1*k2) * omega^j1,
//                             Z_h'[p] = sum_m b_m w16^(mp),  published half-spectra P0[p][k2] = Z_0[p],  P1[p][k2] = w32^p * Z_1[p]
//   bins: F[k2 + 32p] = P0[p][k2] + P1[p][k2],   F[k2 + 32(p+16)] = P0[p][k2] - P1[p][k2]    (taken by the consumer)
//
// inverse (decimation in frequency at both levels; warp h produces the outputs of parity h):
//   pass 1, warp h, lane k2 : S_p = (G[k2+32p] +- G[k2+32(p+16)]) * w32^(-hp),  u[2m+h] = sum_p S_p w16^(-mp),
//                             tile T[j1][k2] = u[j1] * conj(W^(j1*k2) * omega^j1)          (rows of parity h)
//   pass 2, warp h, lane j1 : S_q = (T[j1][q] +- T[j1][q+16]) * w32^(-hq),  v[2m+h] = sum_q S_q w16^(-mq),
//                             z[j1 + 32(2m+h)] = v[2m+h] * conj(omega^(32(2m+h))) / 1024
//   so warp h ends with exactly the coefficients it started the forward transform from: the accumulator of a
//   polynomial splits between its two warps by the parity of j2, 32 coefficients (+32 of the upper half) per lane.
//
// Shared-memory traffic per polynomial and transform pair, in complex elements: tiles 1024 written + 2048 read
// (forward), half-spectra 1024 written, 1024 + 2048 (inverse) -- against 1024 + 1024 per direction for the one-warp
// form; the FP64 instruction count is unchanged up to the 2 x 1024 extra complex adds of the deferred butterflies.
#pragma once
#include "fft.cuh"

namespace fhe {
namespace nfft {

FHE_HD constexpr int brev4(int v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3); }

constexpr int HALF_ROWS = 16;                           // rows of an E / O / P tile
constexpr int HALF_TILE_ELEMS = HALF_ROWS * TILE_PITCH;  // 528 elements = 8,448 bytes

// 16-point DFT, radix-2 decimation in time, same conventions as dit32 (fft.cuh): input index q at logical
// position brev4(q), output index f at logical position f; PERM = true keeps logical position i in register
// brev4(i) (natural-order input registers, output f in register brev4(f)).
template <int SIGN, bool PERM>
FHE_HD void dit16(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int half = 1; half <= 8; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 16; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int a = PERM ? brev4(base + j) : base + j;
                const int b = PERM ? brev4(base + j + half) : base + j + half;
                const int tw = j * (16 / half);  // index into the w32 table: w16^(j*8/half) = w32^(j*16/half)
                const double ar = re[a], ai = im[a], br = re[b], bi = im[b];
                if (tw == 0) {
                    re[a] = ar + br; im[a] = ai + bi;
                    re[b] = ar - br; im[b] = ai - bi;
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

FHE_HD constexpr int hslot(int row, int col) { return row * TILE_PITCH + col; }  // row < 16

// ---- forward pass 1.  Warp h, lane j1: registers m = 0..15 hold z[j1 + 32(2m+h)] = c[j] + i*c[j+1024]
// (no twist applied).  Writes Y_0 to the E tile (h = 0) or w32^q * Y_1 to the O tile (h = 1), row q, column j1.
FHE_HD void fwd_split_pass1(int h, double (&re)[16], double (&im)[16], cplx* tile_e, cplx* tile_o, int lane) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {  // twist factor omega^(32*j2), j2 = 2m + h
        const int j2 = 2 * m + h;
        if (j2 == 0) continue;
        const double cr = FHE_C128_RE(j2), ci = FHE_C128_IM(j2);
        const double a = re[m], b = im[m];
        re[m] = a * cr - b * ci;
        im[m] = a * ci + b * cr;
    }
    dit16<+1, true>(re, im);
    cplx* dst = h ? tile_o : tile_e;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        cplx v;
        v.x = re[brev4(q)];
        v.y = im[brev4(q)];
        if (h && q) {  // O[q] = w32^q * Y_1[q]
            const double wr = FHE_W32_RE(q), wi = FHE_W32_IM(q);
            const double a = v.x, b = v.y;
            v.x = a * wr - b * wi;
            v.y = a * wi + b * wr;
        }
        dst[hslot(q, lane)] = v;
    }
}

// ---- forward pass 2.  Warp h2, lane k2 = q + 16s: deferred butterfly, inter-pass twiddle tw[slot(k2, j1)] =
// W^(j1*k2) * omega^j1 (fft.cuh's table, read along the lane's own row), 16-point DFT over the j1 of parity h2.
// Publishes P0[p][k2] = Z_0[p] (h2 = 0) or P1[p][k2] = w32^p * Z_1[p] (h2 = 1).
// compute part: results stay in registers (index p in register brev4(p), w32^p already applied for h2 = 1) so that
// a kernel can put a barrier between the last read of the E / O tiles and the store over the same bytes.
FHE_HD void fwd_split_pass2_compute(int h2, double (&re)[16], double (&im)[16], const cplx* tile_e, const cplx* tile_o,
                                    const cplx* tw, int lane) {
    const int q = lane & 15;
    const double sg = lane >= 16 ? -1.0 : 1.0;   // the deferred butterfly as one FMA: X[q + 16s] = E + (-1)^s O
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int j1 = 2 * m + h2;
        const cplx e = tile_e[hslot(q, j1)], o = tile_o[hslot(q, j1)];
        const double xr = fma(sg, o.x, e.x);
        const double xi = fma(sg, o.y, e.y);
        const cplx w = tw[slot(lane, j1)];
        re[m] = xr * w.x - xi * w.y;
        im[m] = xr * w.y + xi * w.x;
    }
    dit16<+1, true>(re, im);
    if (h2) {
#pragma unroll
        for (int p = 1; p < 16; ++p) {
            const double wr = FHE_W32_RE(p), wi = FHE_W32_IM(p);
            const double a = re[brev4(p)], b = im[brev4(p)];
            re[brev4(p)] = a * wr - b * wi;
            im[brev4(p)] = a * wi + b * wr;
        }
    }
}
FHE_HD void fwd_split_pass2_store(int h2, const double (&re)[16], const double (&im)[16], cplx* pub0, cplx* pub1, int lane) {
    cplx* dst = h2 ? pub1 : pub0;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        cplx v;
        v.x = re[brev4(p)];
        v.y = im[brev4(p)];
        dst[hslot(p, lane)] = v;
    }
}
FHE_HD void fwd_split_pass2(int h2, double (&re)[16], double (&im)[16], const cplx* tile_e, const cplx* tile_o,
                            const cplx* tw, cplx* pub0, cplx* pub1, int lane) {
    fwd_split_pass2_compute(h2, re, im, tile_e, tile_o, tw, lane);
    fwd_split_pass2_store(h2, re, im, pub0, pub1, lane);
}

// bin k = k2 + 32*k1 from the published half-spectra (what the pointwise stage reads)
FHE_HD cplx split_bin(const cplx* pub0, const cplx* pub1, int k2, int k1) {
    const cplx a = pub0[hslot(k1 & 15, k2)], b = pub1[hslot(k1 & 15, k2)];
    const double sg = k1 < 16 ? 1.0 : -1.0;
    cplx f;
    f.x = fma(sg, b.x, a.x);
    f.y = fma(sg, b.y, a.y);
    return f;
}

// ---- inverse pass 1.  Warp h, lane k2: g_lo[p] = G[k2 + 32p], g_hi[p] = G[k2 + 32(p+16)] (one half computed by this
// warp's pointwise stage, the other half read from its partner).  Writes rows j1 = 2m + h of the 32-row tile.
FHE_HD void inv_split_pass1(int h, const cplx (&g_lo)[16], const cplx (&g_hi)[16], double (&re)[16], double (&im)[16],
                            const cplx* tw, cplx* tile, int lane) {
    const double sg = h ? -1.0 : 1.0;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        double sr = fma(sg, g_hi[p].x, g_lo[p].x);
        double si = fma(sg, g_hi[p].y, g_lo[p].y);
        if (h && p) {  // times w32^(-p)
            const double wr = FHE_W32_RE(p), wi = -FHE_W32_IM(p);
            const double a = sr, b = si;
            sr = a * wr - b * wi;
            si = a * wi + b * wr;
        }
        re[brev4(p)] = sr;  // bit-reversed in, natural out
        im[brev4(p)] = si;
    }
    dit16<-1, false>(re, im);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int j1 = 2 * m + h;
        const cplx w = tw[slot(lane, j1)];  // multiply by conj(w)
        cplx v;
        v.x = re[m] * w.x + im[m] * w.y;
        v.y = im[m] * w.x - re[m] * w.y;
        tile[slot(j1, lane)] = v;
    }
}
// The same in three parts, for a kernel whose warp holds its OWN half of the pointwise output in registers (index p in
// register p: G[k2 + 32(16h + p)]) and finds its partner's half in shared memory as exch[hslot(p, lane)]:
//   combine  : S_p from own registers and the partner's half, in place (own registers become the DFT input)
//   (barrier : the partner has read this warp's half, the bytes may be overwritten by the tile)
//   finish   : 16-point DFT, conj twiddle, tile rows of parity h
FHE_HD void inv_split_pass1_combine(int h, double (&re)[16], double (&im)[16], const cplx* partner_half, int lane) {
    double sr[16], si[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const cplx o = partner_half[hslot(p, lane)];
        // h = 0: own = G[p] (low half), partner = G[p+16];  h = 1: own = G[p+16], partner = G[p]:  S = G_lo + (-1)^h G_hi
        sr[p] = h ? o.x - re[p] : re[p] + o.x;
        si[p] = h ? o.y - im[p] : im[p] + o.y;
        if (h && p) {
            const double wr = FHE_W32_RE(p), wi = -FHE_W32_IM(p);
            const double a = sr[p], b = si[p];
            sr[p] = a * wr - b * wi;
            si[p] = a * wi + b * wr;
        }
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        re[brev4(p)] = sr[p];
        im[brev4(p)] = si[p];
    }
}
FHE_HD void inv_split_pass1_finish(int h, double (&re)[16], double (&im)[16], const cplx* tw, cplx* tile, int lane) {
    dit16<-1, false>(re, im);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int j1 = 2 * m + h;
        const cplx w = tw[slot(lane, j1)];
        cplx v;
        v.x = re[m] * w.x + im[m] * w.y;
        v.y = im[m] * w.x - re[m] * w.y;
        tile[slot(j1, lane)] = v;
    }
}

// ---- inverse pass 2.  Warp h, lane j1: reads its own row of the tile; registers m end with
// (c[j] + i*c[j+1024]) for j = j1 + 32(2m+h), untwisted and scaled by 1/1024.
FHE_HD void inv_split_pass2(int h, double (&re)[16], double (&im)[16], const cplx* tile, int lane) {
    const double sg = h ? -1.0 : 1.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const cplx a = tile[slot(lane, q)], b = tile[slot(lane, q + 16)];
        double sr = fma(sg, b.x, a.x);
        double si = fma(sg, b.y, a.y);
        if (h && q) {
            const double wr = FHE_W32_RE(q), wi = -FHE_W32_IM(q);
            const double x = sr, y = si;
            sr = x * wr - y * wi;
            si = x * wi + y * wr;
        }
        re[brev4(q)] = sr;
        im[brev4(q)] = si;
    }
    dit16<-1, false>(re, im);
#pragma unroll
    for (int m = 0; m < 16; ++m) {  // conj(omega^(32*j2)) / 1024 (power-of-two scale: exact)
        const int j2 = 2 * m + h;
        const double cr = FHE_C128_RE(j2) * 0x1p-10, ci = -FHE_C128_IM(j2) * 0x1p-10;
        const double a = re[m], b = im[m];
        re[m] = a * cr - b * ci;
        im[m] = a * ci + b * cr;
    }
}

}  // namespace nfft
}  // namespace fhe
