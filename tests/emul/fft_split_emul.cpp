// CPU emulation of the two-warp ("split") form of the 32x32 negacyclic FFT (fhe_icp_b200/csrc/fft_split.cuh):
// warps h = 0, 1 are run one after the other, lanes in order; shared memory is a plain array and the barrier
// between passes is the end of a loop.  Built by tests/test_fft_split_emul.py with g++.
#include "../../fhe_icp_b200/csrc/fft_split.cuh"
#include <vector>
using namespace fhe::nfft;

extern "C" {
// coef: 2048 doubles -> bins: 1024 complex (re,im interleaved), natural order k = k2 + 32*k1
void emul_split_forward(const double* coef, double* bins) {
    std::vector<cplx> tw(TILE_ELEMS), te(HALF_TILE_ELEMS), to(HALF_TILE_ELEMS), p0(HALF_TILE_ELEMS), p1(HALF_TILE_ELEMS);
    fill_twiddle_table(tw.data());
    double re[16], im[16];
    for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) {
            for (int m = 0; m < 16; ++m) {
                const int j = lane + 32 * (2 * m + h);
                re[m] = coef[j];
                im[m] = coef[j + 1024];
            }
            fwd_split_pass1(h, re, im, te.data(), to.data(), lane);
        }
    for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) fwd_split_pass2(h, re, im, te.data(), to.data(), tw.data(), p0.data(), p1.data(), lane);
    for (int k2 = 0; k2 < 32; ++k2)
        for (int k1 = 0; k1 < 32; ++k1) {
            const cplx f = split_bin(p0.data(), p1.data(), k2, k1);
            bins[2 * (k2 + 32 * k1)] = f.x;
            bins[2 * (k2 + 32 * k1) + 1] = f.y;
        }
}
void emul_split_inverse(const double* bins, double* coef) {
    std::vector<cplx> tw(TILE_ELEMS), tile(TILE_ELEMS);
    fill_twiddle_table(tw.data());
    double re[16], im[16];
    for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) {
            cplx lo[16], hi[16];
            for (int p = 0; p < 16; ++p) {
                lo[p].x = bins[2 * (lane + 32 * p)];        lo[p].y = bins[2 * (lane + 32 * p) + 1];
                hi[p].x = bins[2 * (lane + 32 * (p + 16))]; hi[p].y = bins[2 * (lane + 32 * (p + 16)) + 1];
            }
            inv_split_pass1(h, lo, hi, re, im, tw.data(), tile.data(), lane);
        }
    for (int h = 0; h < 2; ++h)
        for (int lane = 0; lane < 32; ++lane) {
            inv_split_pass2(h, re, im, tile.data(), lane);
            for (int m = 0; m < 16; ++m) {
                const int j = lane + 32 * (2 * m + h);
                coef[j] = re[m];
                coef[j + 1024] = im[m];
            }
        }
}
}
