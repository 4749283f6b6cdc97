// CPU emulation of a complete multi-bit blind rotation in the two-warps-per-polynomial form
// (fhe_icp_b200/csrc/pbs_split.cuh + fft_split.cuh): four emulated warps per ciphertext, run one after the other,
// shared memory as plain arrays, barriers as loop boundaries.  Built by tests/test_pbs_split_emul.py.
#include "../../fhe_icp_b200/csrc/pbs_split.cuh"
#include <cmath>
#include <cstring>
#include <vector>
using namespace fhe::nfft;

extern "C" int emul_pbs_mb2_split(const double* key_blocks /* [pairs][32][3][2][1][2][32][2] */, const uint64_t* ct, int n,
                                  int beta, const uint64_t* lut, uint64_t* out /* N + 1 */) {
    if (n % 2) return 1;
    std::vector<cplx> tw(TILE_ELEMS), omega(128);
    fill_twiddle_table(tw.data());
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int x = 0; x < 64; ++x) {
        omega[x].x = (double)cosl(two_pi * x / 4096.0L);        omega[x].y = (double)sinl(two_pi * x / 4096.0L);
        omega[64 + x].x = (double)cosl(two_pi * (64 * x) / 4096.0L); omega[64 + x].y = (double)sinl(two_pi * (64 * x) / 4096.0L);
    }
    std::vector<int> a_tilde(n + 1);
    for (int i = 0; i <= n; ++i) a_tilde[i] = (int)((((ct[i] >> 51) + 1) >> 1) & 4095);
    // ACC = X^(-b~) * (0, LUT)
    std::vector<uint64_t> acc(2 * NPOLY, 0);
    const int rot = (4096 - a_tilde[n]) & 4095;
    for (int x = 0; x < NPOLY; ++x) {
        const int src = (x - rot) & 4095;
        uint64_t v = lut[src & 2047];
        if (src & 2048) v = 0 - v;
        acc[NPOLY + x] = v;
    }
    std::vector<cplx> te[2], to[2], p0[2], p1[2], tile[2];
    for (int t = 0; t < 2; ++t) { te[t].resize(HALF_TILE_ELEMS); to[t].resize(HALF_TILE_ELEMS); p0[t].resize(HALF_TILE_ELEMS);
                                  p1[t].resize(HALF_TILE_ELEMS); tile[t].resize(TILE_ELEMS); }
    static double g_re[2][2][32][16], g_im[2][2][32][16];   // pointwise halves [t][h][lane][kk]
    double re[16], im[16];
    for (int i = 0; i < n / 2; ++i) {
        const cplx* key_pair = reinterpret_cast<const cplx*>(key_blocks) + (size_t)i * 32 * MB2_BLOCK_ELEMS;
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) {   // digits of the owned coefficients -> forward pass 1
                    for (int m = 0; m < 16; ++m) {
                        const int j = lane + 32 * (2 * m + h);
                        re[m] = split_digit((uint32_t)(acc[t * NPOLY + j] >> 32), beta);
                        im[m] = split_digit((uint32_t)(acc[t * NPOLY + j + 1024] >> 32), beta);
                    }
                    fwd_split_pass1(h, re, im, te[t].data(), to[t].data(), lane);
                }
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane)
                {   // compute, (barrier), store -- the kernel's order
                    fwd_split_pass2_compute(h, re, im, te[t].data(), to[t].data(), tw.data(), lane);
                    fwd_split_pass2_store(h, re, im, p0[t].data(), p1[t].data(), lane);
                }
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) {   // pointwise on half-spectra
                    SplitMonomials mo;
                    split_monomials_init(mo, omega.data(), a_tilde[2 * i], a_tilde[2 * i + 1], lane, 16 * h);
                    split_pointwise(t, h, lane, p0[t].data(), p1[t].data(), p0[1 - t].data(), p1[1 - t].data(), key_pair, mo,
                                    g_re[t][h][lane], g_im[t][h][lane]);
                }
        // exchange of the halves as the kernel will do it: every warp stores its own half where the half-spectra were
        // (after the barrier that ends the pointwise stage), the partner combines it with the half it holds in registers
        static cplx exch[2][2][HALF_TILE_ELEMS];
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane)
                    for (int p = 0; p < 16; ++p) { exch[t][h][hslot(p, lane)].x = g_re[t][h][lane][p]; exch[t][h][hslot(p, lane)].y = g_im[t][h][lane][p]; }
        static double s_re[2][2][32][16], s_im[2][2][32][16];
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) {
                    std::memcpy(s_re[t][h][lane], g_re[t][h][lane], sizeof(double) * 16);
                    std::memcpy(s_im[t][h][lane], g_im[t][h][lane], sizeof(double) * 16);
                    inv_split_pass1_combine(h, s_re[t][h][lane], s_im[t][h][lane], exch[t][1 - h], lane);
                }
        for (int t = 0; t < 2; ++t)      // (barrier) then the tile overwrites the exchange area
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) inv_split_pass1_finish(h, s_re[t][h][lane], s_im[t][h][lane], tw.data(), tile[t].data(), lane);
        for (int t = 0; t < 2; ++t)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) {   // inverse pass 2 -> accumulator update
                    inv_split_pass2(h, re, im, tile[t].data(), lane);
                    for (int m = 0; m < 16; ++m) {
                        const int j = lane + 32 * (2 * m + h);
                        acc[t * NPOLY + j] += split_f64_to_torus(re[m]);
                        acc[t * NPOLY + j + 1024] += split_f64_to_torus(im[m]);
                    }
                }
    }
    // sample extract coefficient 0: o[0] = A_0[0], o[N - x] = -A_0[x], o[N] = A_1[0]
    out[0] = acc[0];
    for (int x = 1; x < NPOLY; ++x) out[NPOLY - x] = 0 - acc[x];
    out[NPOLY] = acc[NPOLY];
    return 0;
}

// ---- the same blind rotation with the KERNEL's memory plan (csrc/pbs_split.cu): ONE 1056-element region per
// polynomial reused as E|O tiles -> P0|P1 half-spectra -> exchanged pointwise halves -> inverse tile, every warp's
// registers living across the phases, and the phases separated exactly where the kernel has its named barriers.
// `order` permutes the order in which the warps (and lanes) of a phase run: 0 = ascending, 1 = descending, 2 = odd
// warps first.  A hand-over that is missing a barrier shows up as a result that depends on the order.
extern "C" int emul_pbs_mb2_split_aliased(const double* key_blocks, const uint64_t* ct, int n, int beta, const uint64_t* lut,
                                          int order, uint64_t* out) {
    if (n % 2) return 1;
    std::vector<cplx> tw(TILE_ELEMS), omega(128);
    fill_twiddle_table(tw.data());
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int x = 0; x < 64; ++x) {
        omega[x].x = (double)cosl(two_pi * x / 4096.0L);        omega[x].y = (double)sinl(two_pi * x / 4096.0L);
        omega[64 + x].x = (double)cosl(two_pi * (64 * x) / 4096.0L); omega[64 + x].y = (double)sinl(two_pi * (64 * x) / 4096.0L);
    }
    std::vector<int> a_tilde(n + 1);
    for (int i = 0; i <= n; ++i) a_tilde[i] = (int)((((ct[i] >> 51) + 1) >> 1) & 4095);
    // accumulators as the kernel keeps them: per warp (t, h) and lane, 16 "re" and 16 "im" coefficients
    static uint64_t acc_re[2][2][32][16], acc_im[2][2][32][16];
    const int rot = (4096 - a_tilde[n]) & 4095;
    for (int t = 0; t < 2; ++t) for (int h = 0; h < 2; ++h) for (int lane = 0; lane < 32; ++lane) for (int m = 0; m < 16; ++m) {
        const int j = lane + 32 * (2 * m + h);
        uint64_t v0 = 0, v1 = 0;
        if (t == 1) {
            int src = (j - rot) & 4095; v0 = lut[src & 2047]; if (src & 2048) v0 = 0 - v0;
            src = (j + 1024 - rot) & 4095; v1 = lut[src & 2047]; if (src & 2048) v1 = 0 - v1;
        }
        acc_re[t][h][lane][m] = v0; acc_im[t][h][lane][m] = v1;
    }
    std::vector<cplx> region[2] = {std::vector<cplx>(TILE_ELEMS), std::vector<cplx>(TILE_ELEMS)};
    static double re[2][2][32][16], im[2][2][32][16];
    static SplitMonomials mo[2][2][32];
    int warps[4] = {0, 1, 2, 3};                      // warp = 2t + h
    if (order == 1) { warps[0] = 3; warps[1] = 2; warps[2] = 1; warps[3] = 0; }
    if (order == 2) { warps[0] = 1; warps[1] = 3; warps[2] = 0; warps[3] = 2; }
    auto lane_at = [&](int i) { return order == 1 ? 31 - i : (order == 2 ? (i * 7 + 3) & 31 : i); };
#define FOR_WARPS for (int wi = 0; wi < 4; ++wi) for (int li = 0; li < 32; ++li) { const int t = warps[wi] >> 1, h = warps[wi] & 1, lane = lane_at(li); \
        double (&R)[16] = re[t][h][lane]; double (&I)[16] = im[t][h][lane]; cplx* reg = region[t].data(); (void)reg; (void)R; (void)I;
#define END_WARPS }
    for (int i = 0; i < n / 2; ++i) {
        const cplx* key_pair = reinterpret_cast<const cplx*>(key_blocks) + (size_t)i * 32 * MB2_BLOCK_ELEMS;
        FOR_WARPS   // phase 1: digits -> forward pass 1 -> E | O
            for (int m = 0; m < 16; ++m) {
                R[m] = split_digit((uint32_t)(acc_re[t][h][lane][m] >> 32), beta);
                I[m] = split_digit((uint32_t)(acc_im[t][h][lane][m] >> 32), beta);
            }
            fwd_split_pass1(h, R, I, reg, reg + HALF_TILE_ELEMS, lane);
        END_WARPS   // bar_poly
        FOR_WARPS fwd_split_pass2_compute(h, R, I, reg, reg + HALF_TILE_ELEMS, tw.data(), lane); END_WARPS   // bar_poly
        FOR_WARPS   // phase 3: half-spectra over E | O, monomials
            fwd_split_pass2_store(h, R, I, reg, reg + HALF_TILE_ELEMS, lane);
            split_monomials_init(mo[t][h][lane], omega.data(), a_tilde[2 * i], a_tilde[2 * i + 1], lane, 16 * h);
        END_WARPS   // bar_ct (A)
        FOR_WARPS   // phase 4: pointwise, ring slice s = blocks {s, 16 + s}
            const cplx* oth = region[1 - t].data();
            for (int s2 = 0; s2 < 16; ++s2) {
                const int k1 = 16 * h + s2;
                split_pointwise_bin(t, lane, split_bin(reg, reg + HALF_TILE_ELEMS, lane, k1), split_bin(oth, oth + HALF_TILE_ELEMS, lane, k1),
                                    key_pair + (size_t)k1 * MB2_BLOCK_ELEMS, mo[t][h][lane], R[s2], I[s2]);
            }
        END_WARPS   // bar_ct (B)
        FOR_WARPS   // phase 5: own half where its half-spectrum was
            for (int p = 0; p < 16; ++p) { cplx v; v.x = R[p]; v.y = I[p]; reg[h * HALF_TILE_ELEMS + hslot(p, lane)] = v; }
        END_WARPS   // bar_poly
        FOR_WARPS inv_split_pass1_combine(h, R, I, reg + (1 - h) * HALF_TILE_ELEMS, lane); END_WARPS   // bar_poly
        FOR_WARPS inv_split_pass1_finish(h, R, I, tw.data(), reg, lane); END_WARPS                      // bar_poly
        FOR_WARPS   // phase 8: inverse pass 2 -> accumulator
            inv_split_pass2(h, R, I, reg, lane);
            for (int m = 0; m < 16; ++m) {
                acc_re[t][h][lane][m] += split_f64_to_torus(R[m]);
                acc_im[t][h][lane][m] += split_f64_to_torus(I[m]);
            }
        END_WARPS   // bar_poly
    }
#undef FOR_WARPS
#undef END_WARPS
    for (int t = 0; t < 2; ++t) for (int h = 0; h < 2; ++h) for (int lane = 0; lane < 32; ++lane) for (int m = 0; m < 16; ++m)
        for (int part = 0; part < 2; ++part) {
            const int x = lane + 32 * (2 * m + h) + 1024 * part;
            const uint64_t v = part ? acc_im[t][h][lane][m] : acc_re[t][h][lane][m];
            if (t == 0) { if (x == 0) out[0] = v; else out[NPOLY - x] = 0 - v; }
            else if (x == 0) out[NPOLY] = v;
        }
    return 0;
}

// ring position -> frequency block (the re-slicing of fhe_b200_bsk2_fourier_split)
extern "C" int emul_split_ring_block(int pos) { return split_ring_block(pos); }
