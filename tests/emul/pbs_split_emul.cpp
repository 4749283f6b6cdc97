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
